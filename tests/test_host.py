"""CPU tests of the host-side logic: C-ABI surface, weight container, scene, sharding (gloo)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dual_space_nerf_b200 import lib

    hdr = open(os.path.join(ROOT, "include", "dsnerf.h")).read()
    declared = sorted(set(re.findall(r"\b(dsnerf_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 18
    assert os.path.exists(lib.LIB_PATH), "libdsnerf.so not built (run __graft_entry__.build())"
    L = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/dsnerf.h but not exported"
    assert sorted(lib.ENTRY_POINTS) == declared
    assert L.dsnerf_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_no_cpu_fallback():
    from dual_space_nerf_b200 import lib

    with pytest.raises(lib.DsnerfError):
        lib.Context(0)
    from dual_space_nerf_b200 import net as N

    with pytest.raises(RuntimeError):
        N.synthetic_net(0)(torch.zeros(4, 6), torch.zeros(4, 6))


def test_weight_container_matches_reference_layout():
    from dual_space_nerf_b200 import net as N

    net = N.synthetic_net(0)
    sd = net.state_dict()
    assert list(sd.keys()) == N.STATE_DICT_ORDER and len(sd) == 33
    assert sum(v.numel() for v in sd.values()) == 500021
    assert sd["nerf.stage1.0.weight"].shape == (256, 87) and sd["nerf.stage2.0.weight"].shape == (256, 319)
    v0 = net._weights_version
    net.load_state_dict(sd)
    assert net._weights_version == v0 + 1
    net.set_light_center(torch.tensor([0.1, 0.2, 0.3]))
    assert net.light_center.dtype == torch.float32


def test_scene_is_deterministic_and_smpl_sized():
    from dual_space_nerf_b200 import scene as S

    a, b = S.make_scene(64, 64), S.make_scene(64, 64)
    assert a["canonical"].shape == (6890, 3) and a["faces"].shape == (13776, 3)
    for k in ("canonical", "posed", "ray_d", "near", "far", "poses"):
        assert np.array_equal(a[k], b[k])
    assert int(a["hit_box"].sum()) == 931
    # closed genus-0 mesh: every edge shared by exactly two faces
    f = a["faces"]
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert np.all(counts == 2)
    assert not np.array_equal(S.make_scene(64, 64, pose_seed=1)["poses"], a["poses"])


def test_shard_ranges_cover_everything():
    from dual_space_nerf_b200 import dist as D

    for n in (0, 1, 7, 262144, 1048576 + 3):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_interleaved_shards_partition_the_rays():
    from dual_space_nerf_b200 import dist as D

    for n, block in ((0, 4), (10, 3), (4096, 64), (1024 * 1024, 1024), (1000, 7)):
        for world in (1, 2, 8):
            parts = [D.interleaved_indices(n, r, world, block) for r in range(world)]
            allidx = torch.cat(parts)
            assert allidx.numel() == n and torch.equal(torch.sort(allidx)[0], torch.arange(n))
            if n and n % (block * world) == 0:
                assert len({p.numel() for p in parts}) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ray_counts, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from dual_space_nerf_b200 import dist as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeRenderer:  # per-ray function of the ray itself, like the real renderer
        def render(self, batch):
            d = batch["ray_d"][0]
            n = batch["near"][0]
            return {"coarse": {"color": d * 2.0, "depth_map": n + 1.0, "acc_map": d[:, 0] * n, "disp_map": 1.0 / (n + 1.0)}}

    ok = True
    for n_rays in ray_counts:
        g = torch.Generator().manual_seed(0)
        batch = {"ray_o": torch.zeros(1, n_rays, 3), "ray_d": torch.rand(1, n_rays, 3, generator=g),
                 "near": torch.rand(1, n_rays, generator=g), "far": torch.ones(1, n_rays)}
        full = FakeRenderer().render(batch)["coarse"]
        got = D.render_sharded(FakeRenderer(), batch)
        ok = ok and all(torch.equal(got[k], full[k]) for k in full)
        for block in (1, 3, 16):  # rows dealt round-robin (config 4's cost-balanced split), regular and ragged tails
            got = D.render_sharded(FakeRenderer(), batch, interleave=block)
            ok = ok and all(torch.equal(got[k], full[k]) for k in full)
    frames = D.gather_frames(torch.full((5, 6), float(rank)))
    ok = ok and frames.shape == (world, 5, 6) and all(float(frames[r, 0, 0]) == r for r in range(world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_render_equals_unsharded_gloo_world2():
    """Rays are independent: a frame rendered in two shards and all-gathered is bit-identical."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, (10, 4097, 1, 96), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_candidate_filter_keeps_every_possible_winner():
    """The lookup table's candidate lists (csrc/geom.cuh: can_beat, settle_cell, long_list) rest on one inequality: with c0 the
    nearest centroid of a cube's centre x and a its half edge, a centroid c can be the nearest one of SOME point of the cube
    only if |x - c|^2 - |x - c0|^2 <= 2 a |c - c0|_1 (the difference of the two squared distances is linear in the point).
    Checked here with the kernel's fp32 slack on the synthetic mesh's centroids (dense clusters at the poles included): the
    brute-force nearest centroid of random points of random cubes is always in the filtered list, and the 1-NN over the list
    (strict '<', lowest index on ties, the rule of pytorch3d's knn_points as called at utils/render_utils.py:95) equals the
    1-NN over all centroids."""
    from dual_space_nerf_b200 import scene as S

    sc = S.make_scene(8, 8)
    cent = sc["canonical"][sc["faces"]].mean(1).astype(np.float32)
    rng = np.random.RandomState(3)
    lo, hi = cent.min(0), cent.max(0)
    a = np.float32(0.0125)  # half edge of a 2.5 cm table cell
    worst = 0
    for trial in range(60):
        # cube centres: near the surface, near the poles (dense clusters) and inside the body (rings of near-equidistant centroids)
        if trial % 3 == 0:
            x = cent[rng.randint(len(cent))] + rng.uniform(-0.05, 0.05, 3)
        elif trial % 3 == 1:
            x = cent[np.argmax(cent[:, trial % 3 + 1] * (1 if trial % 2 else -1))] + rng.uniform(-0.03, 0.03, 3)
        else:
            x = rng.uniform(lo, hi)
        x = x.astype(np.float32)
        d = ((cent - x) ** 2).sum(1, dtype=np.float32)
        i0 = int(np.argmin(d))
        l1 = np.abs(cent - cent[i0]).sum(1, dtype=np.float32)
        keep = d - d[i0] <= 2.0 * a * l1 * np.float32(1.0002) + np.float32(1e-7) + np.float32(2e-6) * d  # can_beat
        worst = max(worst, int(keep.sum()))
        pts = (x + rng.uniform(-a, a, (400, 3))).astype(np.float32)
        dd = ((pts[:, None, :] - cent[None, :, :]) ** 2).sum(-1, dtype=np.float32)
        nn = dd.argmin(1)            # first minimum = lowest index on ties
        assert keep[nn].all(), (trial, int((~keep[nn]).sum()))
        idx = np.flatnonzero(keep)
        assert np.array_equal(idx[dd[:, idx].argmin(1)], nn)
    assert worst > 256  # the sample holds cubes whose list exceeds the builder's scratch list (the long_list path)


def test_pe_sincos_algorithm_accuracy():
    """The tensor-core kernel's positional encoding does not call sincosf: csrc/mlp_tc.cuh:pe_sincos reduces 2^k * x / 2pi
    exactly (two-float) and evaluates sin / cos (2 pi r), |r| <= 1/8, by polynomials.  Restated here operation by operation
    in fp32 (fma emulated through float64) and held to 1e-7 absolute against float64 over the encoding's whole input range."""
    f32 = np.float32

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

    def mul(a, b):
        return (a.astype(np.float64) * b.astype(np.float64)).astype(f32)

    def add(a, b):
        return (a.astype(np.float64) + b.astype(np.float64)).astype(f32)

    def full(x, v):
        return np.full_like(x, f32(v))

    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-1.5, 1.5, 200000), rng.uniform(-1e-3, 1e-3, 20000), rng.uniform(-30, 30, 30000),
                        [0.0, 1.0, -1.0, 0.125, 0.25]]).astype(f32)
    chi, clo = f32(0.15915493667125702), f32(6.4206382432985265e-09)
    yh = mul(x, full(x, chi))
    yl = fma(x, full(x, clo), fma(x, full(x, chi), -yh))
    for k in range(10):
        sc = f32(2.0 ** k)
        th, tl = mul(yh, full(x, sc)), mul(yl, full(x, sc))
        fh = add(th, -np.rint(th).astype(f32))
        q = np.rint(mul(fh, full(x, 4.0))).astype(f32)
        rh = fma(q, full(x, -0.25), fh)
        s = add(rh, tl)
        e = add(tl, -add(s, -rh))
        u = mul(s, s)
        ps = full(x, 41.46822738647461)
        for c in (-76.69773864746094, 81.6052017211914, -41.34170150756836, -1.7484555e-07):
            ps = fma(ps, u, full(x, c))
        pc = full(x, 59.41782760620117)
        for c in (-85.44869995117188, 64.93936920166016, -19.739208221435547):
            pc = fma(pc, u, full(x, c))
        s0 = fma(s, ps, mul(s, full(x, 6.2831854820251465)))
        c0 = fma(pc, u, full(x, 1.0))
        e2 = mul(e, full(x, 6.2831854820251465))
        s1, c1 = fma(e2, c0, s0), fma(-e2, s0, c0)
        qi = q.astype(np.int32) & 3
        a = np.where(qi & 1, c1, s1)
        b = np.where(qi & 1, s1, c1)
        sn = np.where(qi & 2, -a, a)
        cs = np.where((qi + 1) & 2, -b, b)
        arg = x.astype(np.float64) * 2.0 ** k
        assert np.abs(sn - np.sin(arg)).max() < 1e-7 and np.abs(cs - np.cos(arg)).max() < 1e-7, k


def test_load_checkpoint_file(tmp_path):
    """Checkpoint files of the reference (utils/checkpoint.py:102-125: {"model": state_dict, ...}) load into the container."""
    import torch

    from dual_space_nerf_b200 import net as N

    src = N.synthetic_net(0)
    path = tmp_path / "model_epoch_0000010.pth"
    torch.save({"model": {"module." + k: v for k, v in src.state_dict().items()}, "epoch": 10}, path)
    dst = N.DualSpaceNeRF(None)
    data = dst.load_checkpoint(str(path))
    assert data["epoch"] == 10
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
