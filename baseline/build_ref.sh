#!/usr/bin/env bash
# Builds the UNMODIFIED reference CUDA extensions (diff_surfel_rasterization, simple_knn) for sm_100a
# and installs them, together with the reference's own Python glue, into baseline/_ref/ (git-ignored;
# it travels to the GPU box with gpurun).  Nothing from /root/reference is copied into tracked files.
# The only deviation from the reference's own setup.py invocation is a force-included <cstdint>
# (gcc 13 no longer pulls it in transitively; rasterizer_impl.h uses uint32_t without including it).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "reference tree $REF not present; keeping existing $OUT"; exit 0; fi
TMP=$(mktemp -d /tmp/isr_ref_build.XXXXXX)
mkdir -p "$OUT"
export TORCH_CUDA_ARCH_LIST="10.0a" NVCC_APPEND_FLAGS="-include cstdint" MAX_JOBS=${MAX_JOBS:-8}
if [ ! -f "$OUT/diff_surfel_rasterization/_C.so" ] && ! ls "$OUT"/diff_surfel_rasterization/_C*.so >/dev/null 2>&1; then
  cp -r "$REF/submodules/diff-surfel-rasterization" "$TMP/dsr"
  (cd "$TMP/dsr" && python setup.py build_ext --inplace >"$TMP/dsr_build.log" 2>&1) || { tail -50 "$TMP/dsr_build.log"; exit 1; }
  mkdir -p "$OUT/diff_surfel_rasterization"
  cp "$TMP"/dsr/diff_surfel_rasterization/__init__.py "$TMP"/dsr/diff_surfel_rasterization/_C*.so "$OUT/diff_surfel_rasterization/"
fi
if ! ls "$OUT"/simple_knn/_C*.so >/dev/null 2>&1; then
  cp -r "$REF/submodules/simple-knn" "$TMP/knn"
  (cd "$TMP/knn" && python setup.py build_ext >"$TMP/knn_build.log" 2>&1) || { tail -50 "$TMP/knn_build.log"; exit 1; }
  mkdir -p "$OUT/simple_knn"
  touch "$OUT/simple_knn/__init__.py"
  find "$TMP/knn/build" -name "_C*.so" -exec cp {} "$OUT/simple_knn/" \;
fi
# the reference's Python glue on the hot path (render(), contrastive_loss and what they import) and the tracker
# extraction that consumes the pair list (spatial_track/modules/init_tracker.py, golden generation only)
for d in gaussian_renderer utils scene arguments spatial_track; do
  rm -rf "$OUT/$d"; cp -r "$REF/$d" "$OUT/$d"
done
rm -rf "$TMP"
echo "reference installed into $OUT"
