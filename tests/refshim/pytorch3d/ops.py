"""Stand-in for pytorch3d.ops (pytorch3d==0.4.0 is pinned by the reference's
requirements.txt:56 but is neither vendored nor installable here).

TEST INFRASTRUCTURE: only put on sys.path by tests/refharness.py so that the
UNMODIFIED reference code under /root/reference can be imported on CPU.
`knn_points` restates the published brute-force kernel (squared L2, fma chain,
lowest index wins ties) through the oracle's C routine; see oracle/geom.c.
"""
import numpy as np
import torch

from oracle import clib


def knn_points(p1, p2, K=1, return_nn=False, **kw):
    assert K == 1, "the reference only uses K=1 on the render path"
    B = p1.shape[0]
    idxs, dists = [], []
    for b in range(B):
        q = p1[b].detach().cpu().numpy().astype(np.float32)
        c = p2[b].detach().cpu().numpy().astype(np.float32)
        idx, d2 = clib.nearest(q, c, want_d2=True)
        idxs.append(torch.from_numpy(idx.astype(np.int64))[:, None])
        dists.append(torch.from_numpy(d2)[:, None])
    idx = torch.stack(idxs).to(p1.device)
    dist = torch.stack(dists).to(p1.device)
    nn = knn_gather(p2, idx) if return_nn else None
    return dist, idx, nn


def knn_gather(x, idx, lengths=None):
    B, P, K = idx.shape
    D = x.shape[-1]
    return torch.gather(x[:, :, None].expand(-1, -1, K, -1), 1, idx[..., None].expand(-1, -1, -1, D))
