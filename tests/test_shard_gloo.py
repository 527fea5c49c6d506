"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: contiguous sharding of views /
frames / points and the all-reduce of the 21-double normal-equation block and the 4 error sums.
The per-rank numbers come from the oracle here (no GPU in this tier); on the GPU box the same
collective runs over NCCL inside cameracalibrations_b200.reproj_jtj / calculate_errors."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(nviews=11, seed=7):
    from oracle import oracle_c as oc
    rng = np.random.default_rng(seed)
    intr = (2800.0, 2800.0, 1080.0, 1920.0, -0.12, 1.0)
    obj = np.array([[a, b, 0.0] for b in range(14) for a in range(20)], dtype=np.float64)
    rv = rng.normal(0, 0.3, (nviews, 3))
    tv = np.array([-10.0, -7.0, 40.0]) + rng.normal(0, 2.0, (nviews, 3))
    views = [(rv[i], tv[i]) for i in range(nviews)]
    img = np.empty((nviews, 280, 2))
    for i in range(nviews):
        r, q = oc.world2img(oc.chain(intr, rv[i], tv[i]), obj)
        img[i, :, 0], img[i, :, 1] = r, q
    img += rng.normal(0, 0.25, img.shape)
    return intr, views, obj, img


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c as oc
    from cameracalibrations_b200.shard import shard_range, shard_frames
    intr, views, obj, img = _case()
    lo, hi = shard_range(len(views), rank, world)
    pv, sh, _ = oc.reproj_jtj(intr, 1.0, views[lo:hi], obj, img[lo:hi])
    t = torch.from_numpy(sh.copy())
    dist.all_reduce(t)                      # the exchange step of the path (NCCL on the GPU box)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, pv))
    if rank == 0:
        out.put((t.numpy(), gathered, shard_frames(10, 0, world, ring=4), shard_frames(10, 1, world, ring=4)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_of_normal_equation_block():
    from oracle import oracle_c as oc
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    shared, gathered, fr0, fr1 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    intr, views, obj, img = _case()
    pv_full, sh_full, _ = oc.reproj_jtj(intr, 1.0, views, obj, img)
    np.testing.assert_allclose(shared, sh_full, rtol=1e-12, atol=1e-6)
    # ranges tile [0, n) in rank order and the per-view blocks are the same numbers
    assert [g[:2] for g in gathered] == [(0, 6), (6, 11)]
    np.testing.assert_array_equal(np.concatenate([g[2] for g in gathered]), pv_full)
    assert fr0 == [(0, 4), (4, 5)] and fr1 == [(5, 9), (9, 10)]


def test_shard_range_properties():
    from cameracalibrations_b200.shard import shard_range
    for n in (0, 1, 7, 64, 4096, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


# ---------------------------------------------------------------------------------------------
# LM step with views sharded over ranks: per-rank Schur shares (what lm_schur_kernel emits),
# all-reduced, must give the step of the unsharded dense system (tests/test_lm.py::_dense_step).
# The per-rank arithmetic is numpy here (checker), the exchange is the real collective.
def _schur_share(pv, lam):
    S, s, Y, Z = np.zeros((4, 4)), np.zeros(4), [], []
    for b in pv:
        A = b[:36].reshape(6, 6).copy()
        A[np.diag_indices(6)] *= 1.0 + lam
        B, g = b[36:60].reshape(6, 4), b[60:66]
        y, z = np.linalg.solve(A, B), np.linalg.solve(A, g)
        S += B.T @ y
        s += B.T @ z
        Y.append(y)
        Z.append(z)
    return np.concatenate([S.ravel(), s, [0.0]]), Y, Z


def _lm_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c as oc
    from cameracalibrations_b200.shard import shard_range
    intr, views, obj, img = _case(nviews=9)
    lam = 1e-3
    lo, hi = shard_range(len(views), rank, world)
    pv, sh, _ = oc.reproj_jtj(intr, 1.0, views[lo:hi], obj, img[lo:hi])
    sh_t = torch.from_numpy(sh.copy())
    dist.all_reduce(sh_t)                                   # reproj_jtj's exchange
    share, Y, Z = _schur_share(pv, lam)
    sc_t = torch.from_numpy(share.copy())
    dist.all_reduce(sc_t)                                   # lm_fit's exchange after cc_lm_schur_f64
    sh_all, sc = sh_t.numpy(), sc_t.numpy()
    M = sh_all[:16].reshape(4, 4) - sc[:16].reshape(4, 4)
    M[np.diag_indices(4)] += lam * np.diag(sh_all[:16].reshape(4, 4))
    di = np.linalg.solve(M, -(sh_all[16:20] - sc[16:20]))   # cc_lm_update_f64, every rank the same
    de = np.array([-(Z[i] + Y[i] @ di) for i in range(hi - lo)]).reshape(hi - lo, 6)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, di, de))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_lm_step_equals_unsharded_dense_solve():
    from oracle import oracle_c as oc
    from test_lm import _dense_step
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lm_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    intr, views, obj, img = _case(nviews=9)
    pv, sh, _ = oc.reproj_jtj(intr, 1.0, views, obj, img)
    start = np.array([np.concatenate(v) for v in views])
    di_ref, cand_ref = _dense_step(pv, sh, 1e-3, start)
    assert np.array_equal(gathered[0][2], gathered[1][2])               # identical decision inputs
    np.testing.assert_allclose(gathered[0][2], di_ref, rtol=1e-8, atol=1e-14)
    de = np.concatenate([g[3] for g in gathered])
    np.testing.assert_allclose(start + de, cand_ref, rtol=1e-9, atol=1e-12)
