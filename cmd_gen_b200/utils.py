"""The two mask helpers of DiffPhar/utils.py the hot path keeps calling
(utils.py:122-145), restated; the rest of that file (rdkit/Bio I/O) is out of scope."""
import torch


def num_nodes_to_batch_mask(n_samples, num_nodes, device):
    assert isinstance(num_nodes, int) or len(num_nodes) == n_samples
    if isinstance(num_nodes, torch.Tensor):
        num_nodes = num_nodes.to(device)
    return torch.repeat_interleave(torch.arange(n_samples, device=device), num_nodes)


def batch_to_list(data, batch_mask):
    order = torch.argsort(batch_mask)          # stable enough: masks are sorted already
    batch_mask, data = batch_mask[order], data[order]
    sizes = torch.unique(batch_mask, return_counts=True)[1].tolist()
    return torch.split(data, sizes)


def scatter_add(src, index, dim_size=None):
    """index-sum over dim 0 (torch_scatter.scatter_add semantics) — host-side glue only."""
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def scatter_mean(src, index, dim_size=None):
    tot = scatter_add(src, index, dim_size)
    cnt = scatter_add(torch.ones(index.shape[0], dtype=src.dtype, device=src.device), index, tot.shape[0])
    return tot / cnt.clamp(min=1).view((-1,) + (1,) * (src.dim() - 1))
