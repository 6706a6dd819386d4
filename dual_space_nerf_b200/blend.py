"""Drop-in for the reference's `utils/blend_utils.py:ppts_to_pts` (inverse linear-blend skinning, :72-81) on the B200 path.

Same signature and shapes as the reference function; the work is one launch of `lbs_inverse_kernel` per batch element
through the C ABI (`dsnerf_ppts_to_pts`).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import lib

_ctx = {}


def _context(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx:
        _ctx[idx] = lib.Context(idx)
    return _ctx[idx]


def ppts_to_pts(pts, bw, A):
    """transform points from the pose space to the t pose: pts (B,P,3), bw (B,24,P), A (B,24,4,4) -> (B,P,3)."""
    if pts.dim() != 3 or bw.dim() != 3 or bw.shape[1] != 24 or A.shape[-3:] != (24, 4, 4):
        raise ValueError("expected pts (B,P,3), bw (B,24,P), A (B,24,4,4)")
    if not pts.is_cuda:
        raise RuntimeError("dual_space_nerf_b200.blend.ppts_to_pts needs CUDA tensors (there is no CPU path)")
    dev = pts.device
    ctx = _context(dev)
    B, P = pts.shape[0], pts.shape[1]
    out = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for b in range(B):
        x = pts[b].to(dev, torch.float32).contiguous()
        w = bw[b].to(dev, torch.float32).contiguous()
        a = A[b].to(dev, torch.float32).contiguous()
        ctx.check(ctx.L.dsnerf_ppts_to_pts(ctx.h, p(x), p(w), p(a), P, p(out[b]), st))
    return out
