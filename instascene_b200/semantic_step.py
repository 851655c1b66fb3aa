"""One iteration of the semantic-feature training loop (train_semantic.py:102-205) on top of the B200 path
(SURVEY.md §8 row f-2): the same loss terms with the same weights, without the reference's host round trips --

* single-view term(s) (:108-143): per label map, `sample_batchsize` labelled pixels drawn uniformly (device-side scan +
  searchsorted instead of the boolean-mask gather of the whole [F,H,W] map), ProtoNCE on the sampled rows; the first map
  (`segmap`) uses cluster means and weight 0.5, the second (`sorted_segmap`, only with --gram_feat_3d class prototypes)
  uses the fixed prototypes `class_feat` and weight 1;
* 3D term (:175-197): `sample_batchsize` visible Gaussians with a 3D label > 0, ProtoNCE on their (single-normalised)
  features against `class_feat`;
* the gradient reaches the raw `_seg_feature` through the sampled-pixel sparse backward of the rasterizer and, for the
  3D term, through a row gather.

The multi-view term (:146-173, every 10th iteration, 5 extra renders): `multiview_loss` composes it from rendered maps;
`multiview_loss_sampled` draws the pixels FIRST and composites only them -- the samples of all views in one launch
(renderer.render_sampled), no [V,F,H,W] stack, no dense blend.  `single_view_loss_sampled` is the same for the
per-iteration term: the losses of train_semantic.py never read anything but the sampled rows of `seg_feature`."""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import torch

from .contrastive import contrastive_loss
from .rasterizer import normalize_rows, sample_labelled_pixels, sample_pixels
from .renderer import render_sampled


class SemanticOpt(NamedTuple):  # arguments/__init__.py:103-119
    sample_batchsize: int = 32 * 1024
    lambda_singview_contras: float = 1e-6
    lambda_multiview_contras: float = 1e-6
    lambda_3D_contras: float = 2.5e-6
    consider_negative_labels: bool = False


def single_view_loss(seg_feature_map: torch.Tensor, segmaps: Sequence[torch.Tensor], class_feat: Optional[torch.Tensor],
                     opt: SemanticOpt = SemanticOpt(), generator=None, num_labels: Optional[int] = None):
    """train_semantic.py:108-143.  segmaps: flat [H*W] integer label maps -- [segmap] or, when `class_feat` is given,
    [segmap, sorted_segmap].  Returns the weighted sum of the terms (0-dim tensor)."""
    total = None
    for k, gt in enumerate(segmaps):
        gt = gt.reshape(-1)
        negative = k == 0 and opt.consider_negative_labels
        if negative:  # :121-122 every pixel is a candidate, label 0 is a cluster of its own
            pix = torch.randint(0, gt.numel(), (opt.sample_batchsize,), device=gt.device, generator=generator)
            labels = gt[pix]
        else:
            pix, labels = sample_labelled_pixels(gt, opt.sample_batchsize, generator=generator)
        feats = sample_pixels(seg_feature_map, pix)
        weight = 1.0 if k == 1 else 0.5                                   # :133 "mv with larger weight"
        term = contrastive_loss(feats, labels, predef_u_list=class_feat if k == 1 else None, consider_negative=negative,
                                num_labels=None if (k == 1 and class_feat is not None) else num_labels)
        term = term * (opt.lambda_singview_contras * weight)
        total = term if total is None else total + term
    return total


def contrastive_3d_loss(seg_feature_raw: torch.Tensor, labels3d: torch.Tensor, radii: torch.Tensor,
                        class_feat: Optional[torch.Tensor], opt: SemanticOpt = SemanticOpt(), generator=None,
                        num_labels: Optional[int] = None):
    """train_semantic.py:175-197: ProtoNCE over `sample_batchsize` Gaussians that are visible in this view (radii > 0)
    and carry a 3D label > 0.  `seg_feature_raw` is the trainable [P,F] parameter; get_seg_feature's normalisation
    (scene/gaussian_model.py:121-125) is applied to the sampled rows only."""
    lab = torch.where(radii > 0, labels3d.to(radii.device), torch.zeros_like(labels3d))
    ids, labels = sample_labelled_pixels(lab.reshape(-1), opt.sample_batchsize, generator=generator)
    rows = seg_feature_raw[ids]                                            # gather; backward = index_add into [P,F]
    feats = normalize_rows(rows, 1e-6)                                     # == get_seg_feature[ids]
    return contrastive_loss(feats, labels, predef_u_list=class_feat, num_labels=num_labels) * opt.lambda_3D_contras


def multiview_loss(seg_feature_maps: Sequence[torch.Tensor], sorted_segmaps: Sequence[torch.Tensor],
                   class_feat: Optional[torch.Tensor], opt: SemanticOpt = SemanticOpt(), generator=None,
                   num_labels: Optional[int] = None):
    """train_semantic.py:146-173: one ProtoNCE over `sample_batchsize` labelled pixels drawn uniformly from the UNION of
    the labelled pixels of several views.  Equivalent sampling without stacking the [V,F,H,W] maps: one draw over the
    concatenated label maps, then each view gathers its own share."""
    flat = torch.cat([m.reshape(-1) for m in sorted_segmaps])
    pix, labels = sample_labelled_pixels(flat, opt.sample_batchsize, generator=generator)
    hw = sorted_segmaps[0].numel()
    view = torch.div(pix, hw, rounding_mode="floor")
    order = torch.argsort(view, stable=True)
    counts = torch.bincount(view, minlength=len(seg_feature_maps)).tolist()   # one host read of V integers
    feats, start = [], 0
    for v, fmap in enumerate(seg_feature_maps):
        sel = order[start:start + counts[v]]
        start += counts[v]
        feats.append(sample_pixels(fmap, pix[sel] - v * hw))
    feats = torch.cat(feats)
    return contrastive_loss(feats, labels[order], predef_u_list=class_feat, num_labels=num_labels) * opt.lambda_multiview_contras


def multiview_loss_sampled(viewpoint_cameras, pc, pipe, bg_color, sorted_segmaps: Sequence[torch.Tensor],
                           class_feat: Optional[torch.Tensor], opt: SemanticOpt = SemanticOpt(), generator=None,
                           num_labels: Optional[int] = None, prefetched=None):
    """train_semantic.py:146-173 without rendering the views: the same draw as `multiview_loss` (uniform over the union
    of the labelled pixels of the V views, same random stream), then the sampled pixels of ALL views are composited by
    one kernel launch.  Same value and gradient as multiview_loss([render(v)["seg_feature"] for v in views], ...)."""
    flat = torch.cat([m.reshape(-1) for m in sorted_segmaps])
    pix, labels = sample_labelled_pixels(flat, opt.sample_batchsize, generator=generator)
    hw = sorted_segmaps[0].numel()
    view = torch.div(pix, hw, rounding_mode="floor")
    out = render_sampled(viewpoint_cameras, pc, pipe, bg_color, pix - view * hw, view, prefetched=prefetched)
    return contrastive_loss(out["features"], labels, predef_u_list=class_feat, num_labels=num_labels) * opt.lambda_multiview_contras


def single_view_loss_sampled(viewpoint_camera, pc, pipe, bg_color, segmaps: Sequence[torch.Tensor],
                             class_feat: Optional[torch.Tensor], opt: SemanticOpt = SemanticOpt(), generator=None,
                             num_labels: Optional[int] = None, prefetched=None):
    """train_semantic.py:102-143 with the pixels drawn before the view is rendered: every label map contributes
    `sample_batchsize` samples, all of them composited by one launch.  Same value and gradient as
    single_view_loss(render(view)["seg_feature"], ...).  Returns (loss, radii [P]) -- radii feed the 3D term."""
    pix_l, lab_l, neg_l = [], [], []
    for k, gt in enumerate(segmaps):
        gt = gt.reshape(-1)
        negative = k == 0 and opt.consider_negative_labels
        if negative:
            pix = torch.randint(0, gt.numel(), (opt.sample_batchsize,), device=gt.device, generator=generator)
            labels = gt[pix]
        else:
            pix, labels = sample_labelled_pixels(gt, opt.sample_batchsize, generator=generator)
        pix_l.append(pix); lab_l.append(labels); neg_l.append(negative)
    out = render_sampled([viewpoint_camera], pc, pipe, bg_color, torch.cat(pix_l), None,
                         prefetched=None if prefetched is None else [prefetched])
    feats = out["features"]
    total, start = None, 0
    for k in range(len(segmaps)):
        n = int(pix_l[k].numel())
        weight = 1.0 if k == 1 else 0.5
        term = contrastive_loss(feats[start:start + n], lab_l[k], predef_u_list=class_feat if k == 1 else None,
                                consider_negative=neg_l[k],
                                num_labels=None if (k == 1 and class_feat is not None) else num_labels)
        start += n
        term = term * (opt.lambda_singview_contras * weight)
        total = term if total is None else total + term
    return total, out["radii"][0]
