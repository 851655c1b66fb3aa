"""Sampled-pixel rendering vs dense rendering of the semantic-training terms on the GPU box (development aid; bench.py is
the contract and keeps rendering whole views).  cfg3 scene: 2M Gaussians, F=16, 1080p, 32768 samples per term.
    single-view term : render() + sample_pixels + ProtoNCE + backward      vs  single_view_loss_sampled + backward
    multi-view term  : 5 x render() + multiview_loss + backward            vs  multiview_loss_sampled + backward"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import instascene_b200 as isr  # noqa: E402
from instascene_b200 import semantic_step as sstep, synth  # noqa: E402
from quick_bench import Pipe, make_cam, make_pc  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    P, F, W, H, V = int(os.environ.get("QB_P", 2000000)), 16, 1920, 1080, 5
    dev = "cuda:0"
    sc = synth.synth_scene(P, F=F, seed=1003)
    cams = [make_cam(c, dev) for c in synth.ring_cameras(200, W, H)[:8]]
    pc = make_pc(sc, dev)
    for name in ("get_xyz", "get_opacity", "get_scaling", "get_rotation", "get_features"):
        assert not getattr(pc, name).requires_grad
    bg = torch.zeros(3, device=dev)
    labs = [torch.from_numpy(synth.label_map(W, H, 1005 + v)).to(dev).reshape(-1) for v in range(V)]
    opt = sstep.SemanticOpt()
    state = {"i": 0}

    def step(loss_fn):
        def run():
            state["i"] += 1
            loss = loss_fn(state["i"])
            loss.backward()
            pc._seg_feature.grad = None
        return run

    out = {"P": P, "F": F, "image": [W, H], "samples": opt.sample_batchsize}
    out["single_dense_ms"] = timed(step(lambda i: sstep.single_view_loss(
        isr.render(cams[i % 8], pc, Pipe, bg, want_pairs=False)["seg_feature"], [labs[0]], None, opt, num_labels=64)), 8)
    out["single_sampled_ms"] = timed(step(lambda i: sstep.single_view_loss_sampled(
        cams[i % 8], pc, Pipe, bg, [labs[0]], None, opt, num_labels=64)[0]), 8)
    out["multi5_dense_ms"] = timed(step(lambda i: sstep.multiview_loss(
        [isr.render(cams[(i + v) % 8], pc, Pipe, bg, want_pairs=False)["seg_feature"] for v in range(V)], labs, None, opt,
        num_labels=64)), 4)
    out["multi5_sampled_ms"] = timed(step(lambda i: sstep.multiview_loss_sampled(
        [cams[(i + v) % 8] for v in range(V)], pc, Pipe, bg, labs, None, opt, num_labels=64)), 4)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sampled_bench.json"), "w"))


if __name__ == "__main__":
    main()
