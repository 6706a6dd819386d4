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


def interleaved_indices(n_rays: int, rank: int, world: int, block: int):
    """Ray indices of `rank` when blocks of `block` consecutive rays are dealt round-robin to the ranks (block = one image
    row: rank r renders rows r, r + world, ...).  Hit and miss rays differ in cost by two orders of magnitude (transparent
    samples are skipped), and a body fills the middle rows of a frame: contiguous ranges leave the first and last ranks idle,
    interleaved rows give every rank the same mix.  Every rank gets the same count when block * world divides
    n_rays; otherwise the tail blocks go to the low ranks and `gather_interleaved` pads."""
    idx = torch.arange(n_rays)
    return idx[(idx // block) % world == rank]


def gather_interleaved(local, n_rays: int, block: int, group=None):
    """Inverse of `interleaved_indices`: all-gather the per-rank (n_local, C) outputs and restore ray order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    C = local.shape[1]
    counts = [int(((torch.arange(n_rays) // block) % world == r).sum()) for r in range(world)]
    per = max(counts)
    if local.shape[0] < per:
        local = torch.cat([local, local.new_zeros(per - local.shape[0], C)], 0)
    full = local.new_empty(world, per, C)
    dist.all_gather_into_tensor(full.view(world * per, C), local.contiguous(), group=group)
    if n_rays % (block * world) == 0:  # regular case: (world, n_blocks, block, C) -> (n_blocks, world, block, C)
        return full.view(world, per // block, block, C).transpose(0, 1).reshape(n_rays, C)
    out = local.new_empty(n_rays, C)
    for r in range(world):
        out[interleaved_indices(n_rays, r, world, block).to(out.device)] = full[r, : counts[r]]
    return out


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


def unpack_blocks(blocks, rays_per_rank: int):
    """(world, 6 * R) blocks [rgb (R,3) | depth | acc | disp] of `FrameExchange` / `dsnerf_render_gather` -> (world * R, 6)."""
    world, R = blocks.shape[0], rays_per_rank
    return torch.cat([blocks[:, : 3 * R].reshape(world, R, 3), blocks[:, 3 * R: 4 * R, None], blocks[:, 4 * R: 5 * R, None],
                      blocks[:, 5 * R:, None]], 2).reshape(world * R, 6)


def render_sharded(renderer, batch, group=None, interleave=0, exchange=None, slot=0):
    """Render one frame with its rays split across the ranks (config 4) and reassemble it everywhere.

    `renderer` is a dual_space_nerf_b200.Renderer (or anything with the same ``render``).  ``interleave`` = 0: contiguous
    ray ranges; > 0: blocks of that many rays (one image row) dealt round-robin, which balances hit and miss rays.
    ``exchange`` (a `FrameExchange` sized for n_rays / world rays; needs interleave > 0 and n_rays divisible by
    interleave * world): the reassembly is fused into the render kernel (`Renderer.render_gather`) instead of an NCCL call."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    R = batch["ray_o"].shape[1]
    sub = dict(batch)
    if interleave > 0:
        sel = interleaved_indices(R, rank, world, interleave).to(batch["ray_o"].device)
        for k in ("ray_o", "ray_d", "near", "far"):
            sub[k] = batch[k][:, sel]
        if exchange is not None:
            if R % (interleave * world):
                raise ValueError("fused reassembly needs n_rays divisible by interleave * world")
            per = R // world
            flat = unpack_blocks(renderer.render_gather(sub, exchange, slot), per)  # rank-major, rows interleaved
            full = flat.view(world, per // interleave, interleave, OUT_CHANNELS).transpose(0, 1).reshape(R, OUT_CHANNELS)
            return unpack_outputs(full)
        out = renderer.render(sub)["coarse"]
        return unpack_outputs(gather_interleaved(pack_outputs(out), R, interleave, group))
    lo, hi = shard_range(R, rank, world)
    for k in ("ray_o", "ray_d", "near", "far"):
        sub[k] = batch[k][:, lo:hi]
    out = renderer.render(sub)["coarse"]
    return unpack_outputs(gather_rays(pack_outputs(out), R, group))


class FrameExchange:
    """Frame buffers of all ranks in CUDA symmetric memory (torch.distributed._symmetric_memory): every GPU maps every other
    GPU's buffer over NVLink / NVSwitch, so the compositor kernel of `dsnerf_render_gather` stores its per-ray outputs
    straight into the buffers of all GPUs -- the all-gather of SURVEY.md 8e happens inside the render kernel, tile by tile,
    instead of as a separate NCCL collective after it.

    Layout of each GPU's buffer (float32): ``n_slots`` frames of ``world`` blocks of ``6 * rays_per_rank`` floats,
    block = [rgb (R,3) | depth (R) | acc (R) | disp (R)] of one rank.  ``n_slots`` = 2 lets frame k + 1 be rendered while
    frame k is still being read."""

    def __init__(self, rays_per_rank: int, device, group=None, n_slots: int = 2, multicast: bool = False):
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("FrameExchange works inside one NVSwitch domain (at most 8 GPUs)")
        self.R = int(rays_per_rank)
        self.block = 6 * self.R
        self.n_slots = n_slots
        self.buf = symm_mem.empty(n_slots * self.world * self.block, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.mc = 0
        if multicast:  # NVSwitch multicast (NVLS) address of the buffer; 0 / None where the fabric or driver has none
            try:
                self.mc = int(self.hdl.multicast_ptr or 0)
            except Exception:
                self.mc = 0
        self.buf.zero_()
        self.barrier()

    def offsets(self, slot: int, rank=None):
        rank = self.rank if rank is None else rank
        return (slot * self.world + rank) * self.block

    def targets(self, slot: int):
        """(own block pointer, ctypes array of the peers' pointers to this rank's block, n_peers, multicast pointer or None)."""
        import ctypes

        off = 4 * self.offsets(slot)
        peers = [self.ptrs[p] + off for p in range(self.world) if p != self.rank]
        arr = (ctypes.c_void_p * max(len(peers), 1))(*peers)
        return ctypes.c_void_p(self.ptrs[self.rank] + off), arr, len(peers), (ctypes.c_void_p(self.mc + off) if self.mc else None)

    def frames(self, slot: int):
        """(world, 6 * R) view of all ranks' blocks of a slot in THIS GPU's buffer (valid after `barrier`)."""
        lo = slot * self.world * self.block
        return self.buf[lo: lo + self.world * self.block].view(self.world, self.block)

    def barrier(self):
        """Group barrier on the current stream (signal pads in symmetric memory): every rank's stores of the frames
        enqueued before it are visible to every rank after it."""
        self.hdl.barrier(channel=0)
