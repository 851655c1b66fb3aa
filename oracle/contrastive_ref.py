"""CPU ORACLE (test infrastructure) for the ProtoNCE contrastive loss: a device-agnostic torch restatement of
utils/contrastive_utils.py:18-73, line for line (the reference hard-codes .cuda()); autograd gives the gradient.
Run it in float64 for a ground truth or float32 for the reference's own precision."""
import torch


def contrastive_loss_ref(features, masks, predef_u_list=None, min_pixnum=0, temp_lambda=1000, consider_negative=False):
    dev = features.device
    if not consider_negative:
        valid_semantic_idx = masks > 0                                                    # :28-29
    else:
        valid_semantic_idx = torch.ones_like(masks, dtype=torch.bool)                     # :30-31
    mask_ids, mask_nums = torch.unique(masks, return_counts=True)                         # :33
    valid_mask_ids = mask_ids[mask_nums > min_pixnum]                                     # :34
    valid_semantic_idx = valid_semantic_idx & torch.isin(masks, valid_mask_ids)           # :35
    masks = masks[valid_semantic_idx].type(torch.int64)                                   # :37
    if not consider_negative:
        masks = masks - 1                                                                 # :38-39
    features = features[valid_semantic_idx, :]                                            # :40
    features = features / (torch.norm(features, dim=-1, keepdim=True) + 1e-9).detach()    # :41
    mask_ids, mask_nums = torch.unique(masks, return_counts=True)                         # :43
    if predef_u_list is not None:
        u_list = predef_u_list[mask_ids]                                                  # :44-45
    label_mapping = torch.zeros(int(mask_ids.max()) + 1, dtype=torch.long, device=dev)    # :47
    label_mapping[mask_ids] = torch.arange(len(mask_ids), device=dev)                     # :48
    masks = label_mapping[masks]                                                          # :49
    mask_ids, mask_nums = torch.unique(masks, return_counts=True)                         # :50
    mask_num = mask_ids.shape[0]
    if predef_u_list is None:
        u_list_sum = torch.zeros(mask_num, features.shape[1], dtype=features.dtype, device=dev)
        u_list_sum = u_list_sum.scatter_add(0, masks.unsqueeze(1).expand(-1, features.shape[1]), features)  # :56-57
        u_list = u_list_sum / mask_nums[:, None]                                          # :58
    cluster_diff = features - u_list[masks]                                               # :60
    cluster_diff_norm = torch.norm(cluster_diff, dim=1, keepdim=True)                     # :61
    phi_list_sum = torch.zeros(mask_num, 1, dtype=features.dtype, device=dev)
    phi_list_sum = phi_list_sum.scatter_add(0, masks.unsqueeze(1), cluster_diff_norm)     # :62-63
    phi_list = phi_list_sum / (mask_nums.unsqueeze(1) * torch.log(mask_nums.unsqueeze(1) + temp_lambda))  # :64
    phi_list = torch.clip(phi_list * 10, min=0.5, max=1.0).detach()                       # :65-66
    dist = torch.exp(torch.matmul(features, u_list.T) / phi_list.T)                       # :68
    dist_sum = dist.sum(dim=1, keepdim=True)                                              # :69
    return -torch.sum(torch.log(dist[torch.arange(features.shape[0]), masks].unsqueeze(1) / (dist_sum + 1e-9)))  # :71
