"""Full-size parity report against the UNMODIFIED reference CUDA rasterizer (baseline/_ref) on a B200.

    gpurun -- 'python tools/parity_at_scale.py [cfg2 cfg3 cfg5]'   ->  gpurun_out/parity_scale.json  (copied to profiles/)

The same comparison runs as driver-visible tests in tests/test_reference_scale_gpu.py; this tool only writes the
full report.  cfg5 (F=32): the reference stops at 24 feature dims (MAX_EXTRA_DIMS); its feature map / feature gradient
are assembled from two 16-channel reference passes.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_compare as rc  # noqa: E402

CASES = {"cfg2": (500_000, 0, 1920, 1080, 1002), "cfg3": (2_000_000, 16, 1920, 1080, 1003),
         "cfg5": (5_000_000, 32, 1600, 1200, 1005)}


def run_case(name, ref_C):
    P, F, W, H, seed = CASES[name]
    inp = rc.make_inputs(P, F, W, H, seed)
    mfw, mbw = rc.run_mine(inp)
    torch.cuda.synchronize()
    if F <= 24:
        rfw, rbw = rc.run_ref(ref_C, inp)
        r_extra, r_dextra = None, None
    else:
        h = F // 2
        rfw, rbw = rc.run_ref(ref_C, inp, extra=inp["extra"][:, :h].contiguous(), dextra=inp["dextra"][:h].contiguous())
        rfw2, rbw2 = rc.run_ref(ref_C, inp, extra=inp["extra"][:, h:].contiguous(), dextra=inp["dextra"][h:].contiguous(),
                                dcolor=torch.zeros_like(inp["dcolor"]), dothers=torch.zeros_like(inp["dothers"]))
        r_extra = torch.cat([rfw[4], rfw2[4]], 0)
        r_dextra = torch.cat([rbw[8], rbw2[8]], 1)
        rbw = tuple(a + b for a, b in zip(rbw[:8], rbw2[:8])) + (r_dextra,)
    torch.cuda.synchronize()
    rep = rc.compare_forward(inp, mfw, rfw, r_extra)
    rep["grads"] = rc.compare_backward(inp, mbw, rbw, r_dextra)
    print(name, json.dumps(rep), flush=True)
    return rep


def main():
    ref_C = rc.reference_C()
    assert ref_C is not None, "baseline/_ref is missing (baseline/build_ref.sh)"
    names = [a for a in sys.argv[1:] if a in CASES] or list(CASES)
    out = {n: run_case(n, ref_C) for n in names}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_scale.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
