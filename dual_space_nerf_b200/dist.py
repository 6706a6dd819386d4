"""Multi-GPU plumbing: one process per GPU (torchrun), rays are independent units.

The reference has no distributed code.  Rays -- and whole frames --
are independent given the per-frame constants, so the path shards with no
data-path collective; the only exchange is one all-gather of the 6 floats per
ray (rgb, depth, acc, disp) that reassembles the frame.

Works on CUDA tensors over NCCL and on CPU tensors over gloo (the latter is what
the CPU tests exercise with world_size 2).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

OUT_CHANNELS = 6  # rgb(3) depth acc disp


def shard_range(n_rays: int, rank: int, world: int):
    """Contiguous ray range [lo, hi) of `rank`; every rank gets ceil(n/world) rays except the tail."""
    per = (n_rays + world - 1) // world
    lo = min(rank * per, n_rays)
    return lo, min(lo + per, n_rays)


def pack_outputs(out):
    """dict(color (R,3), depth_map, acc_map, disp_map (R,)) -> (R,6) tensor."""
    return torch.cat([out["color"], out["depth_map"][:, None], out["acc_map"][:, None], out["disp_map"][:, None]], 1).contiguous()


def unpack_outputs(t):
    return {"color": t[:, :3], "depth_map": t[:, 3], "acc_map": t[:, 4], "disp_map": t[:, 5]}


def gather_rays(local, n_rays: int, group=None):
    """All-gather per-rank ray shards (made by `shard_range`) into the full (n_rays, C) tensor on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_rays + world - 1) // world
    C = local.shape[1]
    if local.shape[0] < per:  # tail rank: pad so that every contribution has the same size
        local = torch.cat([local, local.new_zeros(per - local.shape[0], C)], 0)
    full = local.new_empty(world * per, C)
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full[:n_rays]


def gather_frames(local, group=None):
    """All-gather one (R, C) frame per rank into (world, R, C) (config 5: one novel-pose frame per GPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[None]
    world = dist.get_world_size(group)
    flat = local.contiguous().reshape(-1)
    full = flat.new_empty(world * flat.numel())
    dist.all_gather_into_tensor(full, flat, group=group)
    return full.reshape((world,) + tuple(local.shape))


def render_sharded(renderer, batch, group=None):
    """Render one frame with its rays split across the ranks (config 4) and reassemble it everywhere.

    `renderer` is a dual_space_nerf_b200.Renderer (or anything with the same ``render``)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    R = batch["ray_o"].shape[1]
    lo, hi = shard_range(R, rank, world)
    sub = dict(batch)
    for k in ("ray_o", "ray_d", "near", "far"):
        sub[k] = batch[k][:, lo:hi]
    out = renderer.render(sub)["coarse"]
    return unpack_outputs(gather_rays(pack_outputs(out), R, group))
