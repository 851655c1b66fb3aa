"""CPU tests of the host-side logic: the C-ABI library loads and exports every declared symbol (no compute calls
without a GPU), the Python mirrors keep the reference's interface, the oracle satisfies domain properties, and the
data-parallel plumbing works over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_loads_and_exports_every_declared_symbol():
    from instascene_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "isr.h")).read()
    declared = set(re.findall(r"\b(isr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/isr.h but not exported by libisr.so"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert L.isr_version() == 200
    assert L.isr_status_string(-3).decode() == "workspace too small"
    assert ctypes.sizeof(_lib.IsrForwardArgs) % 8 == 0 and ctypes.sizeof(_lib.IsrBackwardArgs) % 8 == 0


def test_boundary_cabi_compiled_translation_unit(tmp_path):
    """A C++ translation unit compiled against include/isr.h and linked to libisr.so (the INTEGRATION.md level-3
    binding): every struct field sits at the offset the ctypes mirror assumes, the sizes agree, and calls that need no
    GPU go through the linked symbols."""
    import shutil
    import subprocess
    from instascene_b200 import _lib
    _lib.lib()
    cxx = shutil.which("g++")
    assert cxx, "g++ is part of the image"
    lines = ['#include <cstddef>', '#include <cstdio>', '#include "isr.h"', 'int main() {']
    for sname, st in (("IsrForwardArgs", _lib.IsrForwardArgs), ("IsrBackwardArgs", _lib.IsrBackwardArgs)):
        lines.append(f'  std::printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in st._fields_:
            lines.append(f'  std::printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    lines += ['  std::printf("version %d\\n", isr_version());',
              '  std::printf("geom %zu\\n", isr_geom_bytes(1000));',
              '  std::printf("status %s\\n", isr_status_string(ISR_ERR_WORKSPACE));',
              '  IsrForwardArgs a = IsrForwardArgs();', '  a.P = 10; a.W = 64; a.H = 64; a.F = ISR_MAX_EXTRA_DIMS + 1;',
              '  std::printf("unsupported %d\\n", isr_forward_geometry(&a, nullptr));',
              '  std::printf("sparse %d\\n", isr_backward_extra_sparse(-1, 0, 1, 1, nullptr, nullptr, nullptr, nullptr, 0, 0, '
              'nullptr, nullptr, nullptr, ISR_FLAG_SPEC_ARITH, nullptr));',
              '  return 0;', '}']
    src = tmp_path / "tu.cpp"
    src.write_text("\n".join(lines))
    exe = tmp_path / "tu"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run([cxx, "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lisr", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.rsplit(" ", 1) for line in r.stdout.strip().splitlines())
    for sname, st in (("IsrForwardArgs", _lib.IsrForwardArgs), ("IsrBackwardArgs", _lib.IsrBackwardArgs)):
        assert int(got[sname]) == ctypes.sizeof(st), sname
        for fname, _ in st._fields_:
            assert int(got[f"{sname}.{fname}"]) == getattr(st, fname).offset, f"{sname}.{fname}"
    assert got["version"] == "200" and int(got["geom"]) == _lib.lib().isr_geom_bytes(1000)
    assert got["status workspace too"] == "small" or "status" in r.stdout
    assert got["unsupported"] == "-2" and got["sparse"] == "-1"
    # every field of the header's structs is mirrored (no field silently missing from the ctypes side)
    header = open(os.path.join(ROOT, "include", "isr.h")).read()
    for sname, st in (("IsrForwardArgs", _lib.IsrForwardArgs), ("IsrBackwardArgs", _lib.IsrBackwardArgs)):
        body = header[header.index(f"typedef struct {sname} {{"):header.index(f"}} {sname};")]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:,|;)", body)
        assert set(names) == {f for f, _ in st._fields_}, (sname, set(names) ^ {f for f, _ in st._fields_})


def test_argument_validation_without_gpu():
    from instascene_b200 import _lib
    L = _lib.lib()
    a = _lib.IsrForwardArgs()
    a.P, a.W, a.H, a.F = 10, 64, 64, 40
    assert L.isr_forward_geometry(ctypes.byref(a), None) == -2  # F > ISR_MAX_EXTRA_DIMS
    a.F = 0
    assert L.isr_forward_geometry(ctypes.byref(a), None) == -1  # null pointers
    assert L.isr_forward_geometry(None, None) == -1
    assert L.isr_knn_mean_dist2(-1, None, None, None, 0, None) == -1


def test_product_has_no_cpu_fallback():
    import torch
    import instascene_b200 as isr
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = isr.GaussianRasterizationSettings(8, 8, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                          torch.zeros(3), False, False)
    r = isr.GaussianRasterizer(s)
    z = torch.zeros((4, 3))
    with pytest.raises(Exception):
        r(z, z, torch.zeros((4, 1)), colors_precomp=z, scales=torch.ones((4, 2)), rotations=torch.ones((4, 4)))
    with pytest.raises(Exception):
        isr.distCUDA2(torch.zeros((8, 3)))
    # product sources never import the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "instascene_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("oracle/isr_oracle.c", "").replace(
                    "the oracle", "").replace("oracle:", "").replace("CPU oracle", ""), f


def test_reference_interface_mirrors():
    import inspect
    import instascene_b200 as isr
    assert isr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    sig = inspect.signature(isr.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp", "extra_attrs"]
    sig = inspect.signature(isr.render)
    assert list(sig.parameters)[:7] == ["viewpoint_camera", "pc", "pipe", "bg_color", "scaling_modifier",
                                        "override_color", "norm_seg_feat"]
    sig = inspect.signature(isr.contrastive_loss)
    assert list(sig.parameters)[:6] == ["features", "masks", "predef_u_list", "min_pixnum", "temp_lambda",
                                        "consider_negative"]


def test_synth_is_deterministic_and_matches_reference_conventions():
    from instascene_b200 import synth
    a, b = synth.synth_scene(1000, F=8, seed=5), synth.synth_scene(1000, F=8, seed=5)
    assert np.array_equal(a.xyz, b.xyz) and np.array_equal(a.seg_feature_raw, b.seg_feature_raw)
    cam = synth.ring_cameras(8, 640, 360)[3]
    # camera centre maps to the view-space origin; the cloud centre is in front of the camera (+z)
    c = np.append(cam.camera_center, 1.0) @ cam.world_view_transform
    assert np.abs(c[:3]).max() < 1e-5
    o = np.array([0, 0, 0, 1.0]) @ cam.world_view_transform
    assert abs(o[2] - 4.0) < 1e-4
    # full_proj: clip w equals view z
    p = np.array([0.3, -0.2, 0.1, 1.0])
    assert abs((p @ cam.full_proj_transform)[3] - (p @ cam.world_view_transform)[2]) < 1e-5


def _small(oracle, P=800, F=4, W=48, H=32, seed=3):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_forward, scene_inputs
    inp = scene_inputs(P, F, W, H, seed, scale_mult=1.5)
    return inp, oracle_forward(oracle, inp)


def test_oracle_forward_properties(oracle):
    inp, o = _small(oracle)
    # sortedness: the instance list is ordered by (tile, depth bits, id), ranges partition it
    keys = o["keys"]
    assert np.all(keys[1:] >= keys[:-1])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    same = tiles[1:] == tiles[:-1]
    d = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    tie = same & (d[1:] == d[:-1])
    assert np.all(o["point_list"][1:][tie] > o["point_list"][:-1][tie])  # stable: ties by ascending id
    r = o["ranges"].astype(np.int64)
    assert int((r[:, 1] - r[:, 0]).sum()) == o["num_rendered"] == int(o["tiles_touched"].sum())
    # alpha = 1 - T_final, weights sum to alpha, <= 9 pairs per pixel
    assert np.allclose(o["others"][1], 1.0 - o["final_T"][0], atol=1e-6)
    assert np.all(o["final_T"][0] >= 1e-4 - 1e-9)
    cnt = np.bincount(o["pairs"][:, 1], minlength=inp["W"] * inp["H"])
    assert cnt.max() <= 9
    # idempotence
    o2 = _small(oracle)[1]
    assert all(np.array_equal(o[k], o2[k]) for k in ("color", "others", "extra", "n_contrib", "point_list"))


def test_oracle_backward_is_linear_in_cotangents_and_matches_finite_differences(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_backward, oracle_forward
    inp, o = _small(oracle, P=300, F=4, W=32, H=32, seed=9)
    W, H, F = inp["W"], inp["H"], 4
    rng = np.random.default_rng(0)
    c1 = [rng.standard_normal(s).astype(np.float32) for s in ((3, H, W), (7, H, W), (F, H, W))]
    c2 = [rng.standard_normal(s).astype(np.float32) for s in ((3, H, W), (7, H, W), (F, H, W))]
    g1, g2 = oracle_backward(oracle, inp, o, *c1), oracle_backward(oracle, inp, o, *c2)
    g12 = oracle_backward(oracle, inp, o, *[a + 2 * b for a, b in zip(c1, c2)])
    for k in ("dL_dextra", "dL_dopacity", "dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dsh"):
        want = g1[k] + 2 * g2[k]
        assert np.abs(g12[k] - want).max() <= 2e-4 * (np.abs(want).max() + 1e-12), k
    # finite differences on the feature channels (the map is linear in the features: exact up to rounding)
    cot = c1[2]
    base = float((o["extra"].astype(np.float64) * cot).sum())
    g = oracle_backward(oracle, inp, o, np.zeros_like(c1[0]), np.zeros_like(c1[1]), cot)
    vis = np.flatnonzero(o["radii"] > 0)[:5]
    for gi in vis:
        for ch in (0, 3):
            inp2 = dict(inp)
            ea = inp["extra_attrs"].copy()
            ea[gi, ch] += 0.25
            inp2["extra_attrs"] = ea
            o2 = oracle_forward(oracle, inp2)
            fd = (float((o2["extra"].astype(np.float64) * cot).sum()) - base) / 0.25
            assert abs(fd - g["dL_dextra"][gi, ch]) <= 1e-3 * (abs(fd) + 1e-3), (gi, ch, fd, g["dL_dextra"][gi, ch])
    # finite differences on opacity (smooth away from the thresholds; loose tolerance)
    cotc = c1[0]
    basec = float((o["color"].astype(np.float64) * cotc).sum())
    gc = oracle_backward(oracle, inp, o, cotc, np.zeros_like(c1[1]), np.zeros_like(c1[2]))
    ok = 0
    gmax = float(np.abs(gc["dL_dopacity"]).max())
    for gi in np.argsort(-np.abs(gc["dL_dopacity"][:, 0]))[:6]:  # largest gradients: fp32 forward noise is negligible
        inp2 = dict(inp)
        op = inp["opacities"].copy()
        op[gi] += 1e-3
        inp2["opacities"] = op
        o2 = oracle_forward(oracle, inp2)
        if not np.array_equal(o2["n_contrib"], o["n_contrib"]):
            continue  # crossed a threshold: not differentiable there
        fd = (float((o2["color"].astype(np.float64) * cotc).sum()) - basec) / 1e-3
        assert abs(fd - gc["dL_dopacity"][gi, 0]) <= 5e-2 * abs(fd) + 2e-2 * gmax, (gi, fd, gc["dL_dopacity"][gi, 0])
        ok += 1
    assert ok >= 1


def test_contrastive_oracle_matches_closed_form():
    import torch
    from oracle.contrastive_ref import contrastive_loss_ref
    torch.manual_seed(0)
    f = torch.randn(64, 8, dtype=torch.float64)
    lab = torch.randint(1, 5, (64,))
    loss = contrastive_loss_ref(f, lab)
    fh = f / f.norm(dim=1, keepdim=True)
    ids = torch.unique(lab)
    u = torch.stack([fh[lab == i].mean(0) for i in ids])
    n = torch.tensor([(lab == i).sum() for i in ids], dtype=torch.float64)
    y = torch.searchsorted(ids, lab)
    phi = torch.stack([(fh[lab == i] - u[j]).norm(dim=1).sum() for j, i in enumerate(ids)]) / (n * torch.log(n + 1000))
    phi = torch.clip(phi * 10, 0.5, 1.0)
    z = fh @ u.T / phi
    want = (torch.logsumexp(z, 1) - z[torch.arange(64), y]).sum()
    assert abs(float(loss) - float(want)) < 1e-6 * abs(float(want))


def test_view_sharding():
    from instascene_b200 import dist as idist
    for world in (1, 2, 8):
        all_views = sorted(v for r in range(world) for v in idist.shard_views(200, r, world))
        assert all_views == list(range(200))
        step0 = [idist.step_views(0, 200, r, world) for r in range(world)]
        assert len(set(step0)) == world


_GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
from instascene_b200 import dist as idist
rank, local_rank, world = idist.init("gloo")
assert world == 2
torch.manual_seed(0)
P, F = 1000, 16
per_view = [torch.randn(P, F) for _ in range(4)]          # "gradient of view v" (same on both ranks)
mine = [v for v in range(4) if v %% world == rank]
g = torch.zeros(P, F)
for v in mine:
    g += per_view[v]
idist.allreduce_grads([g], world)
want = sum(per_view)
assert torch.allclose(g, want, atol=1e-5), (g - want).abs().max()
t = idist.max_over_ranks(float(rank + 1), world, "cpu")
assert t == 2.0
idist.barrier(world)
print("rank", rank, "ok")
"""


def test_gloo_world2_gradient_allreduce(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    port = 29500 + (os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


_GLOO_SHARDED_WORKER = """
import os, sys, torch
sys.path.insert(0, %r)
from instascene_b200 import dist as idist
rank, local_rank, world = idist.init("gloo")
assert world == 2
torch.manual_seed(0)
P, F = 960, 8

def adam_update(p, g, m, v, step, cfg, lr=0.025, b1=0.9, b2=0.999, eps=1e-15):
    assert cfg is None
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr / (1 - b1 ** step) * m / (v.sqrt() / (1 - b2 ** step) ** 0.5 + eps))

p0 = torch.randn(P, F)
param = p0.clone()
opt = idist.ShardedAdam(param, world, rank, lr=0.025, eps=1e-15, chunks=4, update_fn=adam_update)
assert opt.exp_avg.numel() * world == P * F          # moments are sharded
ref = p0.clone().requires_grad_(True)
ref_opt = torch.optim.Adam([ref], lr=0.025, eps=1e-15)
for it in range(3):
    grads = [torch.randn(P, F) for _ in range(world)]   # gradient of each rank's view (same list on both ranks)
    opt.step(grads[rank].clone())
    ref.grad = sum(grads)
    ref_opt.step()
    assert torch.allclose(param, ref.detach(), atol=1e-6), (it, (param - ref.detach()).abs().max())
print("rank", rank, "ok")
"""


def test_gloo_world2_sharded_adam_equals_full_adam_on_summed_gradients(tmp_path):
    """reduce-scatter -> Adam on the rank's rows -> all-gather (instascene_b200.dist.ShardedAdam, world_size 2 over gloo)
    equals torch.optim.Adam on the sum of the ranks' gradients, for several steps, with the moments sharded."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_SHARDED_WORKER % ROOT)
    port = 31500 + (os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_cfg1_plumbing_config_on_the_cpu_port(oracle):
    """BASELINE.json configs[0]: 50k random Gaussians, one 512x512 camera, RGB-only forward on the CPU restatement
    (the reference has no CPU path; this is the plumbing config that needs no GPU).  Size-independent properties."""
    from helpers import oracle_forward, scene_inputs
    P, W, H = 50_000, 512, 512
    inp = scene_inputs(P, 0, W, H, 1001, view=0, n_views=8)
    o = oracle_forward(oracle, inp, want_pairs=False)
    vis = o["radii"] > 0
    assert 0.3 * P < vis.sum() <= P and o["num_rendered"] == int(o["tiles_touched"].sum()) > P
    assert o["color"].shape == (3, H, W) and np.isfinite(o["color"]).all() and np.isfinite(o["others"]).all()
    r = o["ranges"].astype(np.int64)
    assert r.shape[0] == (W // 16) * (H // 16) and int((r[:, 1] - r[:, 0]).sum()) == o["num_rendered"]
    keys = o["keys"]
    assert np.all(keys[1:] >= keys[:-1])                                        # (tile, depth) sortedness
    T = o["final_T"][0]
    assert np.all(T >= 1e-4 - 1e-9) and np.all(T <= 1.0) and np.allclose(o["others"][1], 1.0 - T, atol=1e-6)
    # colour = blended colour + T * background: with a black background the image is bounded by alpha * max colour
    inp_black = dict(inp, bg=np.zeros(3, np.float32))
    ob = oracle_forward(oracle, inp_black, want_pairs=False)
    assert np.allclose(o["color"] - ob["color"], T[None] * inp["bg"][:, None, None], atol=1e-6)   # linear in the background
    assert np.array_equal(o["n_contrib"], ob["n_contrib"]) and np.array_equal(o["others"], ob["others"])
