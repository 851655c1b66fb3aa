"""CPU ORACLE (test infrastructure) for the sampled-pixel ProtoNCE loss.

Restates the algorithm of the reference's `contrastive_loss` (utils/contrastive_utils.py:18-73) in closed form with
dense one-hot algebra instead of the reference's unique / LUT-remap / scatter_add sequence; device agnostic (the
reference hard-codes .cuda()), differentiable through torch autograd, usable in float64 as ground truth.

Semantics being restated (line numbers of the reference):
  * samples with label <= 0 are ignored unless `consider_negative` (:28-31); labels then shift down by one (:38-39)
  * features are divided by (their L2 norm + 1e-9) with the norm DETACHED (:41)
  * clusters = the distinct labels present (:43); cluster centre u_k = predef_u_list[k] if given (:44-45), else the
    mean of the normalised member features, through which gradients flow (:54-58)
  * temperature phi_k = clip(10 * sum_{i in k} |f_i - u_k| / (n_k * log(n_k + temp_lambda)), 0.5, 1), detached (:60-66)
  * loss = - sum_i log( exp(f_i.u_{y_i} / phi_{y_i}) / (sum_k exp(f_i.u_k / phi_k) + 1e-9) )      (:68-71, a SUM)
  * `min_pixnum` drops clusters with <= min_pixnum samples before anything else (:33-35)
"""
import torch


def contrastive_loss_ref(features, masks, predef_u_list=None, min_pixnum=0, temp_lambda=1000, consider_negative=False):
    lab = masks.to(torch.int64)
    keep = torch.ones_like(lab, dtype=torch.bool) if consider_negative else lab > 0
    if min_pixnum > 0:
        ids, cnt = torch.unique(lab, return_counts=True)
        keep &= torch.isin(lab, ids[cnt > min_pixnum])
    y = lab[keep] - (0 if consider_negative else 1)
    f = features[keep]
    f = f / (f.norm(dim=1, keepdim=True) + 1e-9).detach()

    present = torch.unique(y)                                   # ascending cluster ids, K = len(present)
    onehot = (y[:, None] == present[None, :]).to(f.dtype)       # [N, K] membership
    n_k = onehot.sum(0)                                          # [K]
    if predef_u_list is None:
        centres = (onehot.t() @ f) / n_k[:, None]                # cluster means (differentiable)
    else:
        centres = predef_u_list[present].to(f.dtype)
    spread = (f - onehot @ centres).norm(dim=1)                  # |f_i - u_{y_i}|
    phi = (onehot.t() @ spread) / (n_k * torch.log(n_k + temp_lambda))
    phi = torch.clamp(10.0 * phi, 0.5, 1.0).detach()
    e = torch.exp((f @ centres.t()) / phi[None, :])              # [N, K]
    own = (e * onehot).sum(1)
    return -(torch.log(own / (e.sum(1) + 1e-9))).sum()
