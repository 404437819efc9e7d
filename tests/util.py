import numpy as np
import torch


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'max rel err' of BASELINE.json's north_star)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def cuda(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def randomize_bn_(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def nbr_to_pairs(nbr, n_out):
    """(kvol, cap) neighbour table -> list of (in_rows, out_rows) per offset (sorted by out row)."""
    out = []
    for k in range(nbr.shape[0]):
        col = nbr[k, :n_out]
        o = np.nonzero(col >= 0)[0].astype(np.int32)
        out.append((col[o].astype(np.int32), o))
    return out
