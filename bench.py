#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: views/sec fwd+bwd @1080p, N Gaussians x feat_dim).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5]

Workloads (synthetic, seeded -- SURVEY.md §8d):
  cfg3 (default, BASELINE.json configs[2], the one the north-star target is quoted on):
        2M Gaussians, F=16, 1920x1080; one step per rank = render() of one view + 32768 sampled pixels +
        ProtoNCE contrastive loss + backward to the raw seg-feature parameter + (N>1) NCCL all-reduce of
        dL/d_seg_feature + fused Adam step.   Views shard over ranks (weak scaling: one view per rank per step).
  cfg2 (BASELINE.json configs[1]): 500k Gaussians, F=0, 1080p, RGB+depth+normal forward + backward of ALL
        gradients with seeded random cotangents.
  cfg5 (BASELINE.json configs[4]): 5M Gaussians, F=32, 1600x1200, the --gram_feat_3d step (two single-view terms, fixed
        class prototypes, 3D term).  The reference CUDA leg is unavailable there (its rasterizer stops at F=24).
One JSON line on stdout (rank 0).  See the task contract for the keys.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Native libraries write to fd 1 behind Python's back (NCCL prints its
# "NCCL version ..." banner there at NCCL_DEBUG=VERSION and ignores NCCL_DEBUG_FILE at that level), so fd 1 points at
# stderr while the bench runs and is restored only to emit the result line.
_REAL_STDOUT_FD = None


def _quiet_stdout():
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)
    if _REAL_STDOUT_FD is not None:
        os.dup2(2, 1)


WORKLOADS = {
    "cfg3": dict(P=2_000_000, F=16, W=1920, H=1080, seed=1003, n_views=200, samples=32768, labels=64,
                 desc="cfg3: 2M Gaussians x 16-dim features @1920x1080, render + ProtoNCE(32768 px) + backward + Adam"),
    "cfg2": dict(P=500_000, F=0, W=1920, H=1080, seed=1002, n_views=200, samples=0, labels=0,
                 desc="cfg2: 500k Gaussians @1920x1080, RGB+depth+normal forward + backward of all gradients"),
    # BASELINE.json configs[4]: the --gram_feat_3d step of train_semantic.py:102-205 -- two label maps per view (cluster
    # means / fixed Gram-Schmidt class prototypes) + the 3D term over visible labelled Gaussians
    "cfg5": dict(P=5_000_000, F=32, W=1600, H=1200, seed=1005, n_views=200, samples=32768, labels=64,
                 desc="cfg5: 5M Gaussians x 32-dim features @1600x1200, --gram_feat_3d step: render + 2 single-view "
                      "ProtoNCE terms (32768 px each) + 3D ProtoNCE (32768 Gaussians) + backward + Adam"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch comes from the committed `ncu --set full`
# capture of THIS round, stamped with the kernel it was captured on (profiles/r2_traffic.json); a workload without a
# capture reports null.
def ncu_traffic(workload: str, kernel_name: str):
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(workload)
    except Exception:
        return None, None
    if not rec or rec.get("kernel") != kernel_name:
        return None, None
    return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), rec.get("source")


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / max clock / power / throttle reasons of one GPU WHILE the timed region runs -- the fields of the
    B200_PROFILING.md `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.*` line, read
    through NVML in-process (the library nvidia-smi itself uses) every few milliseconds.  Not one nvidia-smi process per
    sample: each start-up enumerates every GPU of the box under a driver lock (it stalled the kernel launches of all
    ranks for tens of ms), and a piped `nvidia-smi -lms` block-buffers its output, so its lines carry no usable time.
    NVML is initialised BEFORE the warm-up.  Falls back to two one-shot nvidia-smi queries (before / after the timed
    region) when the NVML binding is missing."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_s: float = 0.010, enabled: bool = True):
        self.gpu, self.period_s = gpu_index, period_s
        self.enabled = enabled and os.environ.get("ISR_BENCH_NO_CLOCKS", "0") != "1"
        self.samples, self._stop, self._t, self._nvml, self._h, self._active = [], threading.Event(), None, None, None, False

    def start(self):
        if not self.enabled:
            return self
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            try:  # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the CUDA device
                import torch
                bus = int(torch.cuda.get_device_properties(self.gpu).pci_bus_id)
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == bus:
                        self._h = h
                        break
            except Exception:
                pass
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self._nvml = None
        return self

    def _query(self):
        n, h = self._nvml, self._h
        sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        return sm, mask, None   # (the power query is slow -- tens of ms -- and is read once after the region)

    def _run(self):
        while not self._stop.is_set():
            if self._active:
                try:
                    self.samples.append(self._query())
                except Exception:
                    pass
                self._stop.wait(self.period_s)
            else:
                self._stop.wait(0.002)

    def _smi_once(self):
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                               capture_output=True, text=True, timeout=10)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            mask = sum(bit for (bit, _), v in zip(self.REASONS, f[4:8]) if v.lower().startswith("active"))
            self._max = float(f[2])
            self.samples.append((float(f[1]), mask, float(f[3])))
        except Exception:
            pass

    def __enter__(self):  # the timed region starts
        if self.enabled and self._nvml is None:
            self._smi_once()
        self._active = True
        return self

    def __exit__(self, *a):
        self._active = False
        if self.enabled and self._nvml is None:
            self._smi_once()
        elif self.enabled and self._nvml is not None:
            try:
                self._power_after = self._nvml.nvmlDeviceGetPowerUsage(self._h) / 1e3
            except Exception:
                self._power_after = None

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        mask = 0
        for s in self.samples:
            mask |= s[1]
        power = [s[2] for s in self.samples if s[2] is not None]
        if getattr(self, "_power_after", None) is not None:
            power.append(self._power_after)
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": self._max,
                "reasons": sorted(name for bit, name in self.REASONS if mask & bit), "samples": len(self.samples),
                "power_w_max": max(power) if power else None,
                "source": "nvml, sampled during the timed region" if self._nvml is not None else "nvidia-smi before/after the timed region"}


# ----------------------------------------------------------------------------------------------------------------
def build_workload(name: str, n_views_needed: int):
    """Seeded synthetic scene + the ring of views.  The views travel through a synthetic COLMAP model on disk
    (cameras.bin / images.bin, BASELINE.json configs[3] "200 synthetic COLMAP views") and are read back the way the
    reference derives its cameras from those files (scene/colmap_loader.py:180-241, scene/dataset_readers.py:68-101)."""
    import tempfile
    from instascene_b200 import io as isr_io, synth
    w = WORKLOADS[name]
    scene = synth.synth_scene(w["P"], F=w["F"], seed=w["seed"])
    with tempfile.TemporaryDirectory(prefix="isr_colmap_") as d:
        isr_io.write_synthetic_colmap(os.path.join(d, "sparse", "0"), synth.ring_cameras(w["n_views"], w["W"], w["H"]))
        cams = isr_io.load_colmap_cameras(os.path.join(d, "sparse", "0"))
    assert len(cams) == w["n_views"]
    return w, scene, cams


class _Pipe:
    compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 1.0


class _PipeDeferred(_Pipe):
    # the chain rule of the two seg-feature normalisations is applied inside the Adam kernel (instascene_b200.FusedAdam /
    # dist.ShardedAdam) instead of a separate rownorm-backward pass
    defer_seg_feature_grad = True


def _make_pc(scene, dev):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3

        def __init__(self):
            self.get_xyz = t(scene.xyz)
            self.get_opacity = t(scene.opacities()).reshape(-1, 1)
            self.get_scaling = t(scene.scales())
            self.get_rotation = t(scene.rotations())
            self.get_features = t(scene.shs())
            self._seg_feature = t(scene.seg_feature_raw).requires_grad_(True) if scene.seg_feature_raw is not None else None

        @property
        def get_seg_feature(self):  # scene/gaussian_model.py:121-125
            if self._seg_feature is None:
                return None
            return self._seg_feature / (torch.norm(self._seg_feature, p=2, dim=1, keepdim=True) + 1e-6)

    return PC()


class _Cam:
    znear, zfar = 0.01, 100.0

    def __init__(self, cam, wvt, fpt, center):
        self.FoVx, self.FoVy, self.image_width, self.image_height = cam.FoVx, cam.FoVy, cam.image_width, cam.image_height
        self.world_view_transform, self.full_proj_transform, self.camera_center = wvt, fpt, center


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import instascene_b200 as isr
    from instascene_b200 import dist as idist, synth
    rank, local_rank, world = idist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    wl, scene, cams = build_workload(args.workload, args.steps + args.warmup)
    P, F, W, H = wl["P"], wl["F"], wl["W"], wl["H"]
    pc = _make_pc(scene, dev)
    bg = torch.zeros(3, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    total_steps = args.steps + args.warmup
    my_views = [idist.step_views(s, len(cams), rank, world) for s in range(total_steps)]

    # ---- per-view host data (pinned) and its device copies ------------------------------------------------------
    host, devdata = {}, {}
    for v in sorted(set(my_views)):
        c = cams[v]
        h = dict(wvt=torch.from_numpy(c.world_view_transform).pin_memory(),
                 fpt=torch.from_numpy(c.full_proj_transform).pin_memory(),
                 center=torch.from_numpy(c.camera_center).pin_memory())
        if args.workload in ("cfg3", "cfg5"):
            lab = synth.label_map(W, H, wl["seed"] + 2 + v)
            h["labels"] = torch.from_numpy(lab.reshape(-1).astype(np.int16)).pin_memory()  # the view's segmap
            if args.workload == "cfg5":  # the view's sorted_segmap (globally consistent ids): a second labelling
                lab2 = synth.label_map(W, H, wl["seed"] + 5000 + v, grid=6)
                h["labels2"] = torch.from_numpy(lab2.reshape(-1).astype(np.int16)).pin_memory()
        else:
            rng = np.random.default_rng(wl["seed"] + 1 + v)
            h["dcolor"] = torch.from_numpy(rng.standard_normal((3, H, W)).astype(np.float32)).pin_memory()
            h["dothers"] = torch.from_numpy(rng.standard_normal((7, H, W)).astype(np.float32)).pin_memory()
        host[v] = h
        devdata[v] = {k: x.to(dev) for k, x in h.items()}
    torch.cuda.synchronize()

    opt = sharded = None
    sem_opt, class_feat, labels3d = None, None, None
    # cfg3: the raw parameter reaches the loss only through the rasterizer's normalisation -> deferred chain rule;
    # cfg5's 3D term also reads the parameter directly (ordinary gradient), kept on the plain path
    pipe = _PipeDeferred if args.workload == "cfg3" else _Pipe
    if args.workload in ("cfg3", "cfg5"):
        from instascene_b200 import semantic_step as sstep
        sem_opt = sstep.SemanticOpt(sample_batchsize=wl["samples"])
        opt = isr.FusedAdam([pc._seg_feature], lr=0.025, eps=1e-15)  # scene/gaussian_model.py:217-249
        if world > 1:  # reduce-scatter -> Adam on this rank's rows -> all-gather (moments sharded)
            sharded = idist.ShardedAdam(pc._seg_feature, world, rank, lr=0.025, eps=1e-15, chunks=4)
    if args.workload == "cfg5":
        class_feat = t(synth.gram_schmidt_prototypes(wl["labels"], F, wl["seed"] + 7))   # gaussian_model.py:158-176
        labels3d = t(synth.morton_labels(scene.xyz, wl["labels"]))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    import collections
    held = collections.deque(maxlen=2)
    geo_params = []
    if args.workload == "cfg2":
        for name in ("get_xyz", "get_scaling", "get_rotation", "get_opacity", "get_features"):
            p = getattr(pc, name).detach().clone().requires_grad_(True)
            setattr(pc, name, p)
            geo_params.append(p)

    # Semantic-feature training optimises `_seg_feature` only (train_semantic.py), so the geometry phase of the NEXT
    # view (projection, depth sort, instance count -- none of it reads the features) can be started on a side stream
    # while the loss / backward / optimizer tail of the current view runs: isr.prefetch_geometry.  Not valid for cfg2
    # (RGB training moves the geometry every step).  ISR_BENCH_PREFETCH=0 disables it.
    use_prefetch = args.workload in ("cfg3", "cfg5") and os.environ.get("ISR_BENCH_PREFETCH", "1") != "0"
    prefetched = {}

    def prefetch(nx):
        vn, dn = nx
        if vn not in prefetched:
            prefetched[vn] = isr.prefetch_geometry(_Cam(cams[vn], dn["wvt"], dn["fpt"], dn["center"]), pc, pipe, bg)

    def step(v, data, nxt=None, nxt2=None):
        cam = _Cam(cams[v], data["wvt"], data["fpt"], data["center"])
        if args.workload in ("cfg3", "cfg5"):
            pkg = isr.render(cam, pc, pipe, bg, prefetched=prefetched.pop(v, None))
            if use_prefetch and nxt is not None:
                prefetch(nxt)
            segmaps = [data["labels"]] if class_feat is None else [data["labels"], data["labels2"]]
            loss = sstep.single_view_loss(pkg["seg_feature"], segmaps, class_feat, sem_opt, generator=gen,
                                          num_labels=wl["labels"])
            if class_feat is not None:
                loss = loss + sstep.contrastive_3d_loss(pc._seg_feature, labels3d, pkg["radii"], class_feat, sem_opt,
                                                        generator=gen)
            loss.backward()
            if world > 1:
                # gradient reduce-scatter + sharded Adam + parameter all-gather on a side stream; the next render() launches
                # (or has prefetched) its geometry phase and binning first and waits for this event only before it reads
                # the features (renderer.py)
                seg = pc._seg_feature
                deferred = isr.optim.deferred_grad(seg)
                g, cfg = deferred if deferred is not None else (seg.grad, None)
                comm.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(comm):
                    sharded.step(g, cfg)
                    ev = torch.cuda.Event()
                    ev.record(comm)
                seg.grad = None
                seg._isr_deferred_dy = None
                pc._isr_param_ready_event = ev
                # keep the gradient buffer alive until the main stream has waited for `ev` (next render) instead of
                # record_stream(): its deferred frees made the caching allocator grow for dozens of steps
                held.append(g)
                # The next blend has to wait for the all-gathered parameter (~0.4 ms during which the main stream is
                # idle and the next view's geometry is long finished): start the geometry + binning of the view AFTER
                # the next one now, so that the SMs have work during the collective (prefetch depth 2 at N > 1).
                if use_prefetch and nxt2 is not None:
                    prefetch(nxt2)
            else:
                opt.step()
                opt.zero_grad(set_to_none=True)
            return loss
        else:
            pkg = isr.render(cam, pc, _Pipe, bg)
            # cotangent pattern of train.py:76-104 replaced by seeded random cotangents on the three maps it uses
            loss = (pkg["render"] * data["dcolor"]).sum() + (pkg["rend_normal"] * data["dothers"][2:5]).sum() \
                + (pkg["rend_dist"] * data["dothers"][6:7]).sum() + (pkg["surf_depth"] * data["dothers"][0:1]).sum() \
                + (pkg["rend_alpha"] * data["dothers"][1:2]).sum()
            loss.backward()
            idist.allreduce_grads([p.grad for p in geo_params], world)
            for p in geo_params:
                p.grad = None
            return loss

    import gc
    from instascene_b200 import _lib as _isr_lib
    launch_count = _isr_lib.lib().isr_kernel_launch_count

    def timed(n_steps, first, e2e, wrap=False, report=False):
        """Times exactly n_steps steps: barrier + synchronize on both sides, one CUDA event per step boundary on the
        launching stream; returns the max over ranks of the region time plus per-step statistics.  The cyclic GC is
        switched off inside the region (a generation-2 collection over a heap with hundreds of live tensors is a
        multi-millisecond host stall that a 50-ms region cannot absorb)."""
        gc.collect()
        gc.disable()
        idist.barrier(world)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
        h2d = d2h = 0
        n_launch0 = launch_count()
        evs[0].record()
        last = None
        view_of = lambda s: my_views[s % len(my_views)] if wrap else my_views[s]
        cam_keys = ("wvt", "fpt", "center")
        upload = lambda v, keys: {k: host[v][k].to(dev, non_blocking=True) for k in keys}
        cam_ahead = {}  # e2e: camera matrices are uploaded up to two steps ahead (they feed the prefetch)
        prefetched.clear()
        last_s = first + n_steps - 1
        for n, s in enumerate(range(first, first + n_steps)):
            v = view_of(s)
            nxt = nxt2 = None
            if e2e:
                data = cam_ahead.pop(s, None) or upload(v, cam_keys)
                data.update(upload(v, [k for k in host[v] if k not in cam_keys]))
                if s == first:
                    h2d = sum(x.numel() * x.element_size() for x in host[v].values())
                for d in (1, 2):
                    if s + d <= last_s and (s + d) not in cam_ahead:
                        cam_ahead[s + d] = upload(view_of(s + d), cam_keys)
                if s + 1 <= last_s:
                    nxt = (view_of(s + 1), cam_ahead[s + 1])
                if s + 2 <= last_s:
                    nxt2 = (view_of(s + 2), cam_ahead[s + 2])
            else:
                data = devdata[v]
                if s + 1 <= last_s:
                    nxt = (view_of(s + 1), devdata[view_of(s + 1)])
                if s + 2 <= last_s:
                    nxt2 = (view_of(s + 2), devdata[view_of(s + 2)])
            loss = step(v, data, nxt, nxt2)
            if e2e:
                last = float(loss.detach().to("cpu", non_blocking=False))  # device -> host read of the step's result
                d2h = 4
            if n + 1 < n_steps:
                evs[n + 1].record()
        if comm is not None:  # the last step's all-reduce + optimizer step belong to the timed region
            torch.cuda.current_stream().wait_stream(comm)
        evs[n_steps].record()
        torch.cuda.synchronize()
        n_launch = launch_count() - n_launch0
        idist.barrier(world)
        gc.enable()
        total = evs[0].elapsed_time(evs[n_steps])
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)] or [total]
        ms = idist.max_over_ranks(total, world, dev)
        stats = step_stats(per_step, total, rank, world, dev, warn=report)
        return ms, h2d, d2h, last, stats, n_launch

    def step_stats(per_step, total, rank, world, dev, warn=True):
        """p50 / p90 / max of the per-step device times over all ranks, and which rank owned the slowest step."""
        mine = torch.tensor(per_step + [total], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as tdist
            allr = [torch.empty_like(mine) for _ in range(world)]
            tdist.all_gather(allr, mine)
            allr = torch.stack(allr).cpu().numpy()
        else:
            allr = mine.cpu().numpy()[None, :]
        steps_ms = allr[:, :-1]
        worst = np.unravel_index(int(np.argmax(steps_ms)), steps_ms.shape) if steps_ms.size else (0, 0)
        out = {"p50": float(np.percentile(steps_ms, 50)), "p90": float(np.percentile(steps_ms, 90)), "max": float(steps_ms.max()),
               "max_rank": int(worst[0]), "max_step": int(worst[1]), "region_ms_per_rank": [float(x) for x in allr[:, -1]]}
        if warn and out["max"] > 3.0 * out["p50"] and rank == 0:
            print(f"[bench] WARNING: slowest step {out['max']:.3f} ms (rank {out['max_rank']}, step {out['max_step']}) is more than "
                  f"3x the median {out['p50']:.3f} ms -- the timed region swallowed a stall", file=sys.stderr)
        return out

    cs = ClockSampler(local_rank, enabled=(rank == 0)).start()
    timed(args.warmup, 0, False)  # warm-up (untimed)
    # Pre-heat (untimed, not part of W or K): the process has just spent ~10 s building the synthetic scene on the CPU with
    # the GPU idle, and W = 3 steps are ~10 ms of GPU work -- not enough for a cold GPU (and, at N > 1, NCCL and the
    # caching allocator) to reach steady state: measured on a fresh 4-GPU box the first timed region ran 4.2-4.5 ms/step,
    # every later one 2.86.  ~1 s of the same steps first; the count derives from the all-reduced warm-up time, so every
    # rank runs the same number of collectives.
    preheat = 0
    if os.environ.get("ISR_BENCH_PREHEAT", "1") != "0":
        ms_probe = timed(10, 0, False, wrap=True)[0]   # the W warm-up steps include first-call overheads
        preheat = 10 + int(min(600, max(1, 1000.0 / max(ms_probe / 10.0, 0.5))))
        timed(preheat - 10, 0, False, wrap=True)
    with cs:
        ms, _, _, _, stats, n_launch = timed(args.steps, args.warmup, False, report=True)
    timed(max(3, args.warmup), 0, True, wrap=True)  # untimed: the end-to-end path's own allocations (upload buffers) warm
    ms_e2e, h2d, d2h, _, stats_e2e, _ = timed(args.steps, args.warmup, True, report=True)
    cs.stop()
    clocks = cs.summary()
    views = args.steps * world
    value = views / (ms / 1e3)
    e2e_value = views / (ms_e2e / 1e3)

    line = {"metric": "views/sec fwd+bwd @1080p", "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "step_ms_p50": stats["p50"], "step_ms_p90": stats["p90"], "step_ms_max": stats["max"],
            "step_ms_max_owner": {"rank": stats["max_rank"], "step": stats["max_step"]},
            "config": {"workload": wl["desc"], "gaussians": P, "feat_dim": F, "image": [W, H],
                       "views_per_step_per_gpu": 1, "views": f"{len(cams)} synthetic COLMAP views (cameras.bin/images.bin round trip)",
                       "preheat_steps_untimed": preheat,
                       "geometry_prefetch": ("next view's projection + depth sort + binning overlap the current step's loss/backward/Adam "
                                             "(isr.prefetch_geometry)" + ("; at N > 1 one view further ahead, so that it runs under the "
                                                                          "parameter all-gather" if world > 1 else "")) if use_prefetch else "off", "parallelism": f"dp{world} (views sharded; gradient reduce-scatter, Adam on 1/N of the rows, parameter all-gather)" if world > 1 else "dp1",
                       "trainable": "_seg_feature only; geometry frozen, as GaussianModel.training_setup does for semantic "
                                    "training (scene/gaussian_model.py:226-232)" if opt is not None else "all geometry / appearance tensors",
                       "l2": "inputs larger than L2 (Gaussian state %.0f MB >> 126 MB), distinct view every step" % (P * (232 + 4 * F) / 1e6),
                       "optimizer": "Adam(lr=0.025, eps=1e-15) on _seg_feature (isr.FusedAdam)" if opt is not None else "none"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "step_ms_p50": stats_e2e["p50"], "step_ms_max": stats_e2e["max"]},
            # kernels of libisr.so launched by THIS rank inside the timed region (isr_kernel_launch_count; library
            # kernels -- CUB sort/scan, torch glue, NCCL -- are not included)
            "gpu_launches": int(n_launch)}

    # Supplementary (NOT the headline, which renders whole views like the reference): the same training step with the
    # pixels drawn first and only they composited (semantic_step.single_view_loss_sampled) -- what a train_semantic.py
    # iteration needs, since its losses read nothing but the sampled rows of seg_feature.
    if args.workload == "cfg3" and world == 1 and os.environ.get("ISR_BENCH_SAMPLED", "1") != "0":
        def sampled_step(v, data, nxt, handles):
            cam = _Cam(cams[v], data["wvt"], data["fpt"], data["center"])
            loss, _ = sstep.single_view_loss_sampled(cam, pc, pipe, bg, [data["labels"]], None, sem_opt, generator=gen,
                                                     num_labels=wl["labels"], prefetched=handles.pop(v, None))
            if use_prefetch and nxt is not None:
                vn, dn = nxt
                handles[vn] = isr.prefetch_geometry(_Cam(cams[vn], dn["wvt"], dn["fpt"], dn["center"]), pc, pipe, bg,
                                                    want_pairs=False)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss

        def run_sampled(n_steps):
            handles = {}
            gc.collect(); gc.disable()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(n_steps):
                v, vn = my_views[s % len(my_views)], my_views[(s + 1) % len(my_views)]
                sampled_step(v, devdata[v], (vn, devdata[vn]) if s + 1 < n_steps else None, handles)
            e1.record()
            torch.cuda.synchronize()
            gc.enable()
            return e0.elapsed_time(e1)

        run_sampled(10)
        ms_s = run_sampled(args.steps)
        line["sampled_step"] = {
            "value": args.steps / (ms_s / 1e3), "unit": "training steps/s", "ms_per_step": ms_s / args.steps, "steps": args.steps,
            "note": "supplementary, same workload and optimizer: labelled pixels are drawn before rendering and only those "
                    "32768 pixels are composited (isr.render_sampled: projection + binning per view, one warp per sample, "
                    "bit-identical features); the headline `value` composites the whole 1080p view like the reference"}

        # Supplementary: the same step when the caller leaves requires_grad=True on the geometry (what the reference's
        # GaussianModel does during train_semantic.py although its optimizer only holds _seg_feature; round-1 advisor
        # finding): render() then cannot use the prefetch and runs the dense backward of every gradient.
        if os.environ.get("ISR_BENCH_GEOM_TRAINABLE", "1") != "0":
            names = ("get_xyz", "get_scaling", "get_rotation", "get_opacity", "get_features")
            saved = {n: getattr(pc, n) for n in names}
            for n in names:
                setattr(pc, n, saved[n].detach().clone().requires_grad_(True))

            def run_trainable(n_steps):
                gc.collect(); gc.disable()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for s in range(n_steps):
                    v = my_views[s % len(my_views)]
                    data = devdata[v]
                    pkg = isr.render(_Cam(cams[v], data["wvt"], data["fpt"], data["center"]), pc, pipe, bg)
                    loss = sstep.single_view_loss(pkg["seg_feature"], [data["labels"]], None, sem_opt, generator=gen,
                                                  num_labels=wl["labels"])
                    loss.backward()
                    opt.step()
                    opt.zero_grad(set_to_none=True)
                    for n in names:
                        getattr(pc, n).grad = None
                e1.record()
                torch.cuda.synchronize()
                gc.enable()
                return e0.elapsed_time(e1)

            run_trainable(3)
            k = max(4, args.steps // 2)
            ms_t = run_trainable(k)
            for n in names:
                setattr(pc, n, saved[n])
            line["geom_trainable_step"] = {
                "value": k / (ms_t / 1e3), "unit": "views/s", "ms_per_step": ms_t / k, "steps": k,
                "note": "supplementary: geometry / appearance tensors left with requires_grad=True (the reference's GaussianModel "
                        "state during train_semantic.py): no prefetch, dense backward of every gradient; freeze them "
                        "(requires_grad_(False)) to get the headline path"}

    if rank == 0:
        line["roofline"] = measure_roofline(args, wl, pc, cams, devdata, my_views, dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=25.0)
        if world == 1 and not args.no_ref_cuda:
            rc = ref_cuda_leg(args, wl, pc, cams, devdata, my_views, dev)
            if rc is not None:
                line["ref_cuda"] = rc
        emit_line(line)
    idist.barrier(world)


# ----------------------------------------------------------------------------------------------------------------
def measure_roofline(args, wl, pc, cams, devdata, my_views, dev):
    """Dominant kernel = blend_fwd_kernel.  Re-launches ONLY that kernel (ISR_FLAG_SKIP_BINNING) on an already
    binned view and times it with CUDA events on the launching (= torch current) stream."""
    import ctypes as C
    import torch
    from instascene_b200 import _lib
    from instascene_b200.rasterizer import c_rasterize_gaussians
    P, F, W, H = wl["P"], wl["F"], wl["W"], wl["H"]
    v = my_views[-1]
    d = devdata[v]
    e = torch.empty(0, device=dev)
    with torch.no_grad():
        seg = e
        if F:
            seg = pc.get_seg_feature
            seg = (seg / (seg.norm(dim=-1, keepdim=True) + 1e-9)).contiguous()
        cam = cams[v]
        res = c_rasterize_gaussians(torch.zeros(3, device=dev), pc.get_xyz.detach(), e, pc.get_opacity.detach(),
                                    pc.get_scaling.detach(), pc.get_rotation.detach(), 1.0, e, seg, F, d["wvt"], d["fpt"],
                                    cam.tanfovx, cam.tanfovy, H, W, pc.get_features.detach(), 3, d["center"], False, False,
                                    want_pairs=True, return_args=True)
    (R, color, others, radii, extra, geom, binning, img, pairs, pidx, a) = res
    G = int(pidx.item()) + 1
    V = int((radii > 0).sum().item())
    L = _lib.lib()
    a.flags = a.flags | _lib.FLAG_SKIP_BINNING
    stream = torch.cuda.current_stream().cuda_stream
    times = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.isr_forward_render(C.byref(a), a._n_inst, stream), "isr_forward_render(blend only)")
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    HW = W * H
    # sorted id (4) + instance record gather (64 + 12 + 4F) per instance, outputs 4*(3+7+F) and
    # saved per-pixel state (20) per pixel, 8 bytes per emitted pair.  The kernel walks the footprint-culled list
    # (R_emitted instances), so THAT count defines the bytes it has to move; the figure with the reference's instance
    # count R (what a kernel without the culling would have to gather) is reported beside it.
    n_inst = int(a._n_inst)
    per_inst, per_px = 4 + 64 + 12 + 4 * F, 4 * (3 + 7 + F) + 20
    alg = n_inst * per_inst + HW * per_px + 8 * G
    alg_ref = R * per_inst + HW * per_px + 8 * G
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg / (ms / 1e3) / 1e9
    kname = f"blend_fwd_kernel<F={F}>"
    traffic, traffic_src = ncu_traffic(args.workload, kname)
    return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": "measured" if peaks else "fallback",
            "traffic": traffic, "traffic_source": traffic_src,
            "kernel_ms": ms, "algorithmic_bytes": alg, "R": R, "R_emitted": n_inst, "V": V, "pairs": G,
            "frac_with_reference_instance_count": alg_ref / (ms / 1e3) / 1e9 / peak,
            "note": "instruction-issue bound (~80% of issue slots busy, DRAM 3% of peak): every pixel walks its tile list; "
                    "algorithmic bytes count one record gather per emitted (tile, Gaussian) instance, most of which hit L2"}


# ----------------------------------------------------------------------------------------------------------------
class CpuWorkload:
    """The same step as the GPU arm, on the reference's algorithm restated in C (oracle/, OpenMP on every host core; the
    reference ships no CPU rasterizer, DSR/rasterize_points.cu:27-28), LIKE FOR LIKE:
      cfg3 / cfg5: forward of one view (colour, 7 aux maps, F feature channels, pair list) + sampled-pixel ProtoNCE
                   (oracle/contrastive_ref.py) + backward of dL/d(features) over exactly the sampled pixels + the double
                   normalisation backward + Adam on [P,F] (numpy) -- only `_seg_feature` is trainable
                   (scene/gaussian_model.py:226-232), which is also all the GPU arm differentiates;
      cfg2:        forward + dense backward of every gradient with seeded random cotangents + K8.
    `tile_stride` > 1 runs the per-tile blends on every S-th tile only (bounded sample; projection, binning and the
    per-Gaussian work always cover all Gaussians)."""

    def __init__(self, workload: str):
        from instascene_b200 import synth
        from oracle import oracle as orc
        self.orc, self.wl, self.name = orc, WORKLOADS[workload], workload
        wl = self.wl
        P, F, W, H = wl["P"], wl["F"], wl["W"], wl["H"]
        orc.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
        self.cores = orc.num_threads()
        self.scene = scene = synth.synth_scene(P, F=F, seed=wl["seed"])
        self.cams = synth.ring_cameras(wl["n_views"], W, H)
        self.raw = scene.seg_feature_raw.astype(np.float32).copy() if F else None
        self.kw = dict(scales=scene.scales(), rotations=scene.rotations(), shs=scene.shs(), sh_degree=3)
        self.adam = None
        rng = np.random.default_rng(1)
        if not F:
            self.dcolor = rng.standard_normal((3, H, W)).astype(np.float32)
            self.dothers = rng.standard_normal((7, H, W)).astype(np.float32)
        self.rng = rng
        self.synth = synth

    def _features(self):
        x = self.raw
        n1 = np.linalg.norm(x, axis=1, keepdims=True) + 1e-6
        y = x / n1
        n2 = np.linalg.norm(y, axis=1, keepdims=True) + 1e-9
        return (y / n2).astype(np.float32)

    def preprocess_only(self, view=0):
        cam, sc, wl = self.cams[view], self.scene, self.wl
        return self.orc.forward(sc.xyz, sc.opacities(), cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                wl["W"], wl["H"], np.zeros(3, np.float32), **self.kw, want_pairs=False, blend=False)

    def step(self, view: int, tile_stride: int = 1):
        import torch
        cam, sc, wl, orc = self.cams[view % len(self.cams)], self.scene, self.wl, self.orc
        P, F, W, H = wl["P"], wl["F"], wl["W"], wl["H"]
        base = (sc.xyz, sc.opacities(), cam.world_view_transform, cam.full_proj_transform, cam.camera_center, W, H,
                np.zeros(3, np.float32))
        bw_args = (sc.xyz, cam.world_view_transform, cam.full_proj_transform, cam.camera_center, W, H, np.zeros(3, np.float32),
                   cam.tanfovx, cam.tanfovy)
        t = {}
        t0 = time.time()
        if not F:
            fwd = orc.forward(*base, **self.kw, want_pairs=True, tile_stride=tile_stride)
            t["fwd"] = time.time() - t0
            t0 = time.time()
            orc.backward(fwd, *bw_args, self.dcolor, self.dothers, None, **self.kw, tile_stride=tile_stride)
            t["bwd"] = time.time() - t0
            return t
        feats = self._features()
        fwd = orc.forward(*base, **self.kw, extra_attrs=feats, want_pairs=True, tile_stride=tile_stride)
        t["fwd"] = time.time() - t0
        t0 = time.time()
        from oracle.contrastive_ref import contrastive_loss_ref
        n_terms = 2 if self.name == "cfg5" else 1
        dextra = np.zeros((F, H, W), np.float32)
        mask = np.zeros((H, W), np.uint8)
        for k in range(n_terms):
            lab = self.synth.label_map(W, H, wl["seed"] + (2 if k == 0 else 5000) + view, **({} if k == 0 else {"grid": 6})).reshape(-1)
            valid = np.flatnonzero(lab > 0)
            pix = valid[self.rng.integers(0, len(valid), size=wl["samples"])]
            f = torch.tensor(fwd["extra"].reshape(F, -1)[:, pix].T.astype(np.float64), requires_grad=True)
            loss = contrastive_loss_ref(f, torch.tensor(lab[pix].astype(np.int64))) * (1e-6 * (0.5 if k == 0 else 1.0))
            loss.backward()
            np.add.at(dextra.reshape(F, -1).T, pix, f.grad.numpy().astype(np.float32))
            mask.reshape(-1)[pix] = 1
        t["loss"] = time.time() - t0
        t0 = time.time()
        g = orc.backward(fwd, *bw_args, np.zeros((3, H, W), np.float32), np.zeros((7, H, W), np.float32), dextra, **self.kw,
                         extra_attrs=feats, tile_stride=tile_stride, pixel_mask=mask, preprocess=False)
        t["bwd"] = time.time() - t0
        t0 = time.time()
        # backward of the two row normalisations (gaussian_model.py:124 then gaussian_renderer/__init__.py:61-62) + Adam
        x, dy = self.raw, g["dL_dextra"]
        n1 = np.linalg.norm(x, axis=1, keepdims=True)
        y = x / (n1 + 1e-6)
        n2 = np.linalg.norm(y, axis=1, keepdims=True)
        z = y / (n2 + 1e-9)
        dyy = dy / (n2 + 1e-9) - y * ((dy * z).sum(1, keepdims=True) / (n2 + 1e-9) / np.maximum(n2, 1e-30) * (n2 > 0))
        dx = dyy / (n1 + 1e-6) - x * ((dyy * y).sum(1, keepdims=True) / (n1 + 1e-6) / np.maximum(n1, 1e-30) * (n1 > 0))
        if self.adam is None:
            self.adam = [np.zeros_like(x), np.zeros_like(x), 0]
        m, v, n = self.adam
        n += 1
        m *= 0.9; m += 0.1 * dx
        v *= 0.999; v += 0.001 * dx * dx
        self.raw = (x - 0.025 / (1 - 0.9 ** n) * m / (np.sqrt(v) / math.sqrt(1 - 0.999 ** n) + 1e-15)).astype(np.float32)
        self.adam[2] = n
        t["optim"] = time.time() - t0
        return t


def cpu_baseline(workload: str, budget_s: float = 25.0):
    """cpu_baseline leg of the GPU arm (rank 0, N = 1): ONE step of CpuWorkload.  The whole view when that fits the budget
    (no extrapolation); otherwise the per-tile blends on every S-th tile, scaled by S, and the line says so."""
    cw = CpuWorkload(workload)
    wl = cw.wl
    t0 = time.time()
    cw.preprocess_only()
    t_pre = time.time() - t0
    # only the per-tile blends (forward minus projection/binning, backward) scale with the tile stride
    scaled = lambda parts: max(parts["fwd"] - t_pre, 0.0) + parts["bwd"]
    fixed = lambda parts: t_pre + parts.get("loss", 0.0) + parts.get("optim", 0.0)
    cal = cw.step(0, tile_stride=64)
    est_full = fixed(cal) + scaled(cal) * 64
    stride = 1 if est_full <= budget_s * 1.6 else int(min(256, math.ceil(scaled(cal) * 64 / max(budget_s - fixed(cal), 1.0))))
    t0 = time.time()
    parts = cw.step(1, tile_stride=stride)
    t_step = time.time() - t0
    n_tiles = ((wl["W"] + 15) // 16) * ((wl["H"] + 15) // 16)
    if stride == 1:
        t_view, how = t_step, f"the whole view ({n_tiles} tiles), nothing extrapolated"
    else:
        t_view = fixed(parts) + scaled(parts) * stride
        how = f"per-tile blends on every {stride}-th of {n_tiles} tiles, that part scaled x{stride}"
    like = ("forward + sampled-pixel ProtoNCE + feature-only backward on the sampled pixels + Adam" if wl["F"] else
            "forward + dense backward of all gradients")
    return {"value": 1.0 / t_view, "unit": "views/s", "cores": cw.cores, "kind": "port",
            "sample": f"1 step of {wl['desc'].split(':')[0]} ({like}; like for like with the GPU step): all {wl['P']} Gaussians "
                      f"projected + sorted, {how}",
            "seconds_per_view": t_view, "seconds_by_part": parts}


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference ships no CPU path).
    Whole views only -- nothing is sampled or extrapolated here.  Every step is one full view of the GPU arm's workload;
    if K steps do not fit the time budget, fewer are run and the line reports the number actually timed."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    budget_s = float(os.environ.get("ISR_REF_BUDGET_S", 420.0))
    t_start = time.time()
    cw = CpuWorkload(args.workload)
    times, parts = [], None
    n_warm = 0
    for s in range(args.warmup + args.steps):
        warm = s < args.warmup
        if warm and time.time() - t_start > 0.25 * budget_s:
            continue                      # warm-up steps are skipped once they would eat the budget
        if not warm and times and (time.time() - t_start) + float(np.mean(times)) > budget_s:
            break
        t0 = time.time()
        parts = cw.step(s)
        if warm:
            n_warm += 1
        else:
            times.append(time.time() - t0)
    t_view = float(np.mean(times))
    world = int(os.environ.get("WORLD_SIZE", 1))
    cb = {"value": 1.0 / t_view, "unit": "views/s", "cores": cw.cores, "kind": "port",
          "sample": f"{len(times)} whole-view step(s) of {wl['desc'].split(':')[0]}, like for like with the GPU step, "
                    f"nothing extrapolated ({n_warm} warm-up)"}
    line = {"impl": "reference", "metric": "views/sec fwd+bwd @1080p", "value": cb["value"], "unit": "views/s",
            "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps, "warmup": n_warm,
            "ms_per_step": 1e3 * t_view, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "gaussians": wl["P"], "feat_dim": wl["F"], "image": [wl["W"], wl["H"]]},
            "cpu_baseline": cb, "seconds_by_part": parts,
            "e2e": {"value": cb["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t_start, "world_size_env": world}
    emit_line(line)


# ----------------------------------------------------------------------------------------------------------------
def ref_cuda_leg(args, wl, pc, cams, devdata, my_views, dev, n=4):
    """Extra baseline: the UNMODIFIED reference CUDA rasterizer (baseline/_ref) driven through the reference's own
    render() / contrastive_loss on the same inputs, timed in the same run with CUDA events."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_loader
        if not ref_loader.available():
            return None
        rrender, rloss, _ = ref_loader.load()
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": repr(ex)[:200]}
    F, W, H = wl["F"], wl["W"], wl["H"]
    if F > 24:
        return {"unavailable": "the reference rasterizer is compiled for at most 24 feature dims (auxiliary.h:20)"}
    bg = torch.zeros(3, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    ref_class_feat = ref_labels3d = None

    def step(v):
        d = devdata[v]
        cam = _Cam(cams[v], d["wvt"], d["fpt"], d["center"])
        pkg = rrender(cam, pc, _Pipe, bg)
        if args.workload in ("cfg3", "cfg5"):
            loss = 0
            maps = [d["labels"]] if args.workload == "cfg3" else [d["labels"], d["labels2"]]
            for k, lab in enumerate(maps):                        # train_semantic.py:108-143
                segmap = lab.reshape(H, W)
                mask = segmap > 0
                valid_feat = pkg["seg_feature"][:, mask]
                valid_lab = segmap[mask]
                idx = torch.randint(0, len(valid_lab), size=(wl["samples"],), device=dev, generator=gen)
                loss = loss + rloss(valid_feat[:, idx].T, valid_lab[idx].long(),
                                    predef_u_list=ref_class_feat if k == 1 else None) * (1e-6 * (1.0 if k == 1 else 0.5))
            if args.workload == "cfg5":                           # train_semantic.py:175-197
                vis = pkg["visibility_filter"]
                vfeat, vlab = pc.get_seg_feature[vis], ref_labels3d[vis]
                m3 = vlab > 0
                vfeat, vlab = vfeat[m3], vlab[m3]
                idx = torch.randint(0, len(vlab), size=(wl["samples"],), device=dev, generator=gen)
                loss = loss + rloss(vfeat[idx], vlab[idx], predef_u_list=ref_class_feat) * 2.5e-6
            loss.backward()
            pc._seg_feature.grad = None
        else:
            loss = (pkg["render"] * d["dcolor"]).sum() + (pkg["rend_normal"] * d["dothers"][2:5]).sum() \
                + (pkg["rend_dist"] * d["dothers"][6:7]).sum() + (pkg["surf_depth"] * d["dothers"][0:1]).sum() \
                + (pkg["rend_alpha"] * d["dothers"][1:2]).sum()
            loss.backward()
            for name in ("get_xyz", "get_scaling", "get_rotation", "get_opacity", "get_features"):
                getattr(pc, name).grad = None

    views = [my_views[i % len(my_views)] for i in range(n + 2)]
    for v in views[:2]:
        step(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for v in views[2:]:
        step(v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": 1e3 / ms, "unit": "views/s", "ms_per_step": ms, "steps": n,
            "what": "unmodified reference CUDA (diff_surfel_rasterization built for sm_100a) via its own render()/contrastive_loss, no optimizer step"}


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS))
    ap.add_argument("--gaussians", type=int, default=0, help="override the workload's Gaussian count (smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.gaussians > 0:
        WORKLOADS[args.workload] = dict(WORKLOADS[args.workload], P=args.gaussians,
                                        desc=WORKLOADS[args.workload]["desc"] + f" [--gaussians {args.gaussians}]")
    _quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
