"""SURVEY.md §8 row f-2 / BASELINE.json configs[4]: one --gram_feat_3d iteration of the semantic training loop
(train_semantic.py:102-205) through instascene_b200.semantic_step -- two single-view ProtoNCE terms on the same render
(cluster means, weight 0.5 / fixed class prototypes, weight 1), the 3D term over visible labelled Gaussians, backward
to the raw seg-feature parameter -- against the CPU oracle rasterizer composed with the restated loss in torch-CPU
float64.  F = 32 (beyond the reference's 24).  Tolerance 2e-4 norm-wise on the parameter gradient, 1e-4 on the loss."""
import numpy as np
import pytest

from helpers import oracle_backward, oracle_forward, rel_err, scene_inputs

pytestmark = pytest.mark.gpu


def _scene_objects(inp, W, H, dev):
    import torch
    sc, cam = inp["scene"], inp["cam"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        active_sh_degree, max_sh_degree = 3, 3
        get_xyz, get_opacity = t(sc.xyz), t(sc.opacities()).reshape(-1, 1)
        get_scaling, get_rotation, get_features = t(sc.scales()), t(sc.rotations()), t(sc.shs())
        _seg_feature = t(sc.seg_feature_raw).requires_grad_(True)

        @property
        def get_seg_feature(self):
            return self._seg_feature / (torch.norm(self._seg_feature, p=2, dim=1, keepdim=True) + 1e-6)

    class Cam:
        FoVx, FoVy, image_width, image_height = cam.FoVx, cam.FoVy, W, H
        world_view_transform, full_proj_transform, camera_center = t(cam.world_view_transform), t(cam.full_proj_transform), t(cam.camera_center)
        znear, zfar = 0.01, 100.0

    class Pipe:
        compute_cov3D_python, convert_SHs_python, depth_ratio = False, False, 1.0

    return PC(), Cam(), Pipe(), t


def test_gram_feat_3d_step_matches_oracle(oracle, monkeypatch):
    import torch
    import instascene_b200 as isr
    from instascene_b200 import semantic_step as sstep, synth
    from oracle.contrastive_ref import contrastive_loss_ref
    P, F, W, H, seed, n = 5000, 32, 128, 80, 71, 4096
    inp = scene_inputs(P, F, W, H, seed)
    pc, cam, pipe, t = _scene_objects(inp, W, H, "cuda:0")
    sc = inp["scene"]
    lab_a = synth.label_map(W, H, seed + 2)              # view.segmap        (8x8 grid, ids 1..64)
    lab_b = synth.label_map(W, H, seed + 3, grid=6)      # view.sorted_segmap (6x6 grid, ids 1..36)
    class_feat = synth.gram_schmidt_prototypes(40, F, seed + 4)
    labels3d = synth.morton_labels(sc.xyz, 36)
    opt = sstep.SemanticOpt(sample_batchsize=n)

    pkg = isr.render(cam, pc, pipe, t(inp["bg"]))
    radii = pkg["radii"].cpu().numpy()
    # fixed samples, shared with the oracle: the three draws of the step, in call order
    rng = np.random.default_rng(seed + 5)
    va, vb = np.flatnonzero(lab_a.reshape(-1) > 0), np.flatnonzero(lab_b.reshape(-1) > 0)
    v3 = np.flatnonzero((radii > 0) & (labels3d > 0))
    draws = [va[rng.integers(0, va.size, n)], vb[rng.integers(0, vb.size, n)], v3[rng.integers(0, v3.size, n)]]
    queue = list(draws)

    def fixed_sampler(labels_flat, count, generator=None):
        ids = torch.from_numpy(queue.pop(0)).to(labels_flat.device)
        assert count == n and bool((labels_flat[ids] > 0).all())
        return ids, labels_flat[ids]

    monkeypatch.setattr(sstep, "sample_labelled_pixels", fixed_sampler)
    cf = t(class_feat)
    loss = sstep.single_view_loss(pkg["seg_feature"], [t(lab_a.reshape(-1)), t(lab_b.reshape(-1))], cf, opt)
    loss = loss + sstep.contrastive_3d_loss(pc._seg_feature, t(labels3d), pkg["radii"], cf, opt)
    loss.backward()
    assert not queue
    got = pc._seg_feature.grad.cpu().numpy()

    # ---- oracle ------------------------------------------------------------------------------------------------
    with torch.no_grad():
        segn = isr.normalize_rows(pc._seg_feature.detach(), 1e-6, 1e-9, stages=2)
    inp["extra_attrs"] = segn.cpu().numpy()
    o = oracle_forward(oracle, inp)
    assert np.array_equal(pkg["seg_feature"].detach().cpu().numpy().view(np.uint32), o["extra"].view(np.uint32))
    emap = torch.tensor(o["extra"], dtype=torch.float64, requires_grad=True)
    raw = torch.tensor(sc.seg_feature_raw, dtype=torch.float64, requires_grad=True)
    cf64 = torch.tensor(class_feat, dtype=torch.float64)
    fa = emap.reshape(F, -1)[:, torch.tensor(draws[0])].t()
    fb = emap.reshape(F, -1)[:, torch.tensor(draws[1])].t()
    la = torch.tensor(lab_a.reshape(-1)[draws[0]].astype(np.int64))
    lb = torch.tensor(lab_b.reshape(-1)[draws[1]].astype(np.int64))
    l_ref = contrastive_loss_ref(fa, la) * (1e-6 * 0.5) + contrastive_loss_ref(fb, lb, cf64) * (1e-6 * 1.0)   # :133-143
    act = raw / (torch.norm(raw, p=2, dim=1, keepdim=True) + 1e-6)                     # get_seg_feature
    l3 = contrastive_loss_ref(act[torch.tensor(draws[2])], torch.tensor(labels3d[draws[2]]), cf64) * 2.5e-6     # :191-194
    (l_ref + l3).backward()
    assert abs(float(loss) - float(l_ref + l3)) / abs(float(l_ref + l3)) < 1e-4
    og = oracle_backward(oracle, inp, o, np.zeros((3, H, W), np.float32), np.zeros((7, H, W), np.float32),
                         emap.grad.numpy().astype(np.float32))
    raw2 = torch.tensor(sc.seg_feature_raw, dtype=torch.float64, requires_grad=True)
    a = raw2 / (torch.norm(raw2, p=2, dim=1, keepdim=True) + 1e-6)
    a = a / (a.norm(dim=-1, keepdim=True) + 1e-9)
    a.backward(torch.tensor(og["dL_dextra"], dtype=torch.float64))
    want = raw.grad.numpy() + raw2.grad.numpy()
    assert rel_err(got, want) < 2e-4


def test_multiview_loss_equals_stacked_reference_sampling(monkeypatch):
    """multiview_loss (train_semantic.py:146-173) gathers per view instead of stacking [V,F,H,W]; same value as the
    reference's stacked formulation for the same draw."""
    import torch
    from instascene_b200 import semantic_step as sstep, synth
    from oracle.contrastive_ref import contrastive_loss_ref
    V, F, W, H, n = 3, 8, 40, 24, 2048
    gen = torch.Generator(device="cuda").manual_seed(3)
    maps = [torch.randn((F, H, W), device="cuda", generator=gen, requires_grad=True) for _ in range(V)]
    labs = [torch.from_numpy(synth.label_map(W, H, 90 + v, grid=3)).cuda().reshape(-1) for v in range(V)]
    flat = torch.cat(labs)
    valid = torch.nonzero(flat > 0).reshape(-1)
    draw = valid[torch.randint(0, valid.numel(), (n,), device="cuda", generator=gen)]
    monkeypatch.setattr(sstep, "sample_labelled_pixels", lambda lf, c, generator=None: (draw, lf[draw]))
    cf = torch.from_numpy(synth.gram_schmidt_prototypes(12, F, 5)).cuda()
    loss = sstep.multiview_loss(maps, labs, cf, sstep.SemanticOpt(sample_batchsize=n))
    loss.backward()
    stacked = torch.stack([m.detach().double().cpu() for m in maps], 0).requires_grad_(True)   # [V,F,H,W]
    feats = stacked.permute(1, 0, 2, 3).reshape(F, -1)[:, draw.cpu()].t()                       # :163-167
    want = contrastive_loss_ref(feats, flat[draw].cpu().long(), cf.double().cpu()) * 1e-6
    want.backward()
    assert abs(float(loss) - float(want)) / abs(float(want)) < 1e-4
    got = torch.stack([m.grad for m in maps], 0).cpu().numpy()
    assert rel_err(got, stacked.grad.numpy()) < 1e-4


def test_prefetched_geometry_render_is_bitwise_the_inline_render():
    """isr.prefetch_geometry starts phase A of a later view on a side stream; render(prefetched=...) must return exactly
    what an inline render returns (outputs, pair list, gradient of the features), also when several views are in
    flight, and must ignore a handle made for a different camera."""
    import torch
    import instascene_b200 as isr
    from instascene_b200 import synth
    P, F, W, H, seed = 6000, 16, 160, 96, 81
    inp = scene_inputs(P, F, W, H, seed)
    pc, cam0, pipe, t = _scene_objects(inp, W, H, "cuda:0")
    cams = []
    for c in synth.ring_cameras(4, W, H):
        class Cam:
            FoVx, FoVy, image_width, image_height = c.FoVx, c.FoVy, W, H
            world_view_transform, full_proj_transform, camera_center = t(c.world_view_transform), t(c.full_proj_transform), t(c.camera_center)
            znear, zfar = 0.01, 100.0
        cams.append(Cam())
    bg = t(inp["bg"])

    def run(cam, handle=None):
        pc._seg_feature.grad = None
        pkg = isr.render(cam, pc, pipe, bg, prefetched=handle)
        (pkg["seg_feature"] * pkg["seg_feature"]).sum().backward()
        pairs = pkg["gau_related_pixels"]
        return (pkg["render"].detach().clone(), pkg["seg_feature"].detach().clone(), pkg["radii"].clone(),
                pairs[torch.argsort(pairs[:, 0].long() * (W * H) + pairs[:, 1].long())].clone(), pc._seg_feature.grad.clone())

    want = [run(c) for c in cams]
    handles = [isr.prefetch_geometry(c, pc, pipe, bg) for c in cams[:3]]       # three views in flight
    got = [run(cams[1], handles[1]), run(cams[0], handles[0]), run(cams[2], handles[2])]
    for g, w in zip(got, [want[1], want[0], want[2]]):
        for a, b in zip(g[:4], w[:4]):
            assert torch.equal(a, b)
        assert rel_err(g[4].cpu().numpy(), w[4].cpu().numpy()) < 1e-5        # float atomics: summation order differs run to run
    stale = isr.prefetch_geometry(cams[0], pc, pipe, bg)
    g = run(cams[3], stale)                                                    # wrong view: handle ignored
    for a, b in zip(g[:4], want[3][:4]):
        assert torch.equal(a, b)
    # The prefetch also queues the binning, into a workspace sized from the instance-count high-water mark of earlier
    # views (the inline renders above set it).  A workspace that turns out too small must be detected and the binning
    # repeated inline: force a capacity of 100 instances.
    from instascene_b200 import renderer as isr_renderer
    assert handles[0].state.binning is not None and handles[0].state.bin_capacity > 0
    monkey = isr_renderer.binning_capacity_hint
    isr_renderer.binning_capacity_hint = lambda P_, W_, H_: 100
    try:
        small = isr.prefetch_geometry(cams[2], pc, pipe, bg)
        assert small.state.bin_capacity == 100
        g = run(cams[2], small)
    finally:
        isr_renderer.binning_capacity_hint = monkey
    for a, b in zip(g[:4], want[2][:4]):
        assert torch.equal(a, b)
    no_bin = isr.prefetch_geometry(cams[1], pc, pipe, bg, bin_ahead=False)    # phase A only
    assert no_bin.state.binning is None
    g = run(cams[1], no_bin)
    for a, b in zip(g[:4], want[1][:4]):
        assert torch.equal(a, b)
    torch.cuda.synchronize()


def _ring_cams(t, W, H, n):
    from instascene_b200 import synth
    cams = []
    for c in synth.ring_cameras(max(n, 4), W, H)[:n]:
        class Cam:
            FoVx, FoVy, image_width, image_height = c.FoVx, c.FoVy, W, H
            world_view_transform, full_proj_transform, camera_center = t(c.world_view_transform), t(c.full_proj_transform), t(c.camera_center)
            znear, zfar = 0.01, 100.0
        cams.append(Cam())
    return cams


@pytest.mark.parametrize("F", [16, 7, 32])
def test_render_sampled_is_bitwise_the_dense_render_at_the_samples(F):
    """render_sampled composites only the sampled pixels, the samples of all views in ONE launch: the rows must be
    bit-identical to the dense render's seg_feature map at those pixels (same arithmetic, same order), and the gradient
    of the raw parameter must equal the one obtained through render() + sample_pixels()."""
    import torch
    import instascene_b200 as isr
    P, W, H, seed, V, n = 6000, 160, 96, 83, 3, 5000
    inp = scene_inputs(P, F, W, H, seed)
    pc, _, pipe, t = _scene_objects(inp, W, H, "cuda:0")
    cams = _ring_cams(t, W, H, V)
    bg = t(inp["bg"])
    gen = torch.Generator(device="cuda").manual_seed(5)
    pix = torch.randint(0, W * H, (n,), device="cuda", generator=gen)
    view = torch.randint(0, V, (n,), device="cuda", generator=gen)
    wgt = torch.randn((n, F), device="cuda", generator=gen)

    # dense path: V renders, gather, weighted sum
    pc._seg_feature.grad = None
    rows, radii = torch.zeros((n, F), device="cuda"), []
    loss = 0.0
    for v, cam in enumerate(cams):
        pkg = isr.render(cam, pc, pipe, bg, want_pairs=False)
        sel = torch.nonzero(view == v).reshape(-1)
        r = isr.sample_pixels(pkg["seg_feature"], pix[sel])
        rows[sel] = r.detach()
        radii.append(pkg["radii"].clone())
        loss = loss + (r * wgt[sel]).sum()
    loss.backward()
    want_grad = pc._seg_feature.grad.clone()

    pc._seg_feature.grad = None
    launches0 = isr._lib.lib().isr_kernel_launch_count()
    out = isr.render_sampled(cams, pc, pipe, bg, pix, view)
    assert torch.equal(out["features"], rows)
    assert torch.equal(out["radii"], torch.stack(radii))
    (out["features"] * wgt).sum().backward()
    assert rel_err(pc._seg_feature.grad.cpu().numpy(), want_grad.cpu().numpy()) < 1e-5   # float atomics: summation order
    # a single view without view ids, through a prefetched handle
    pc._seg_feature.grad = None
    h = isr.prefetch_geometry(cams[1], pc, pipe, bg, want_pairs=False)
    sel = torch.nonzero(view == 1).reshape(-1)
    one = isr.render_sampled(cams[1], pc, pipe, bg, pix[sel], None, prefetched=[h])
    assert torch.equal(one["features"], rows[sel])
    # out-of-range samples render zeros and receive no gradient
    bad = isr.render_sampled(cams[:2], pc, pipe, bg, torch.tensor([0, W * H, 5], device="cuda"), torch.tensor([0, 1, 7], device="cuda"))
    assert torch.equal(bad["features"][1:], torch.zeros((2, F), device="cuda"))
    torch.cuda.synchronize()
    assert isr._lib.lib().isr_kernel_launch_count() > launches0


def test_sampled_losses_equal_the_rendered_ones():
    """multiview_loss_sampled / single_view_loss_sampled (pixels drawn first, only they are composited) against
    multiview_loss / single_view_loss on dense renders: same random stream -> same draw -> same loss and gradient."""
    import torch
    import instascene_b200 as isr
    from instascene_b200 import semantic_step as sstep, synth
    P, F, W, H, seed, V, n = 6000, 16, 160, 96, 85, 4, 4096
    inp = scene_inputs(P, F, W, H, seed)
    pc, _, pipe, t = _scene_objects(inp, W, H, "cuda:0")
    cams = _ring_cams(t, W, H, V)
    bg = t(inp["bg"])
    labs = [torch.from_numpy(synth.label_map(W, H, 120 + v, grid=4)).cuda().reshape(-1) for v in range(V)]
    cf = torch.from_numpy(synth.gram_schmidt_prototypes(20, F, 7)).cuda()
    opt = sstep.SemanticOpt(sample_batchsize=n)

    def grads(fn):
        pc._seg_feature.grad = None
        loss = fn()
        loss.backward()
        return float(loss), pc._seg_feature.grad.clone().cpu().numpy()

    mk = lambda: torch.Generator(device="cuda").manual_seed(11)
    l_d, g_d = grads(lambda: sstep.multiview_loss([isr.render(c, pc, pipe, bg, want_pairs=False)["seg_feature"] for c in cams],
                                                  labs, cf, opt, generator=mk()))
    l_s, g_s = grads(lambda: sstep.multiview_loss_sampled(cams, pc, pipe, bg, labs, cf, opt, generator=mk()))
    assert abs(l_d - l_s) <= 1e-5 * abs(l_d) and rel_err(g_s, g_d) < 1e-5

    segmaps = [labs[0], labs[1]]
    l_d, g_d = grads(lambda: sstep.single_view_loss(isr.render(cams[2], pc, pipe, bg, want_pairs=False)["seg_feature"], segmaps,
                                                    cf, opt, generator=mk()))
    l_s, g_s = grads(lambda: sstep.single_view_loss_sampled(cams[2], pc, pipe, bg, segmaps, cf, opt, generator=mk())[0])
    assert abs(l_d - l_s) <= 1e-5 * abs(l_d) and rel_err(g_s, g_d) < 1e-5
