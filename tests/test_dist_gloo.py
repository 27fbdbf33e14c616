"""CPU, world_size 2 over gloo: the N>1 path of bench.py — index-addressed shards, per-rank summary
statistics, ONE all-gather of CRB_STATS_LEN doubles, max-over-ranks timing — with the oracle standing in
for the GPU kernels (this tests the host logic, not the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stats_numpy(v, status, iters, i0):
    ok = np.isfinite(v)
    idx = i0 + np.arange(v.size)
    return np.array([v[ok].astype(np.float64).sum(), v[ok].min(), v[ok].max(), (~ok).sum(),
                     (status == 0).sum(), iters.sum(),
                     (v[ok].astype(np.float64) * ((idx[ok] % 251) + 1)).sum(), v.size], np.float64)


def _worker(rank, world, port, n_per, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from cpprobotics_b200 import synth
    from oracle import oracle as O
    T = 6
    course = synth.mpc_course()
    st, pind = synth.mpc_states(n_per, i0=rank * n_per, course=course)
    xref, _ = synth.mpc_xref_numpy(st, pind, T, course=course)
    r = O.mpc_solve_batched(st, xref, T, nthreads=1)
    stats = torch.from_numpy(_stats_numpy(r["cost"], r["status"], r["iters"], rank * n_per))
    g = bench.gather_stats(stats, world)                    # the only collective of the data path
    ms = bench.max_over_ranks_cpu(10.0 + rank, world)
    out_q.put((rank, g.numpy().copy(), ms, r["cost"].copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_reproduce_the_single_process_run():
    world, n_per = 2, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _q
    import time as _t
    t_end = _t.time() + 240
    while len(res) < world and _t.time() < t_end:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            assert all(p.exitcode in (None, 0) for p in procs), "a rank died"
    assert len(res) == world
    res.sort(key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from cpprobotics_b200 import synth
    from oracle import oracle as O
    T = 6
    course = synth.mpc_course()
    st, pind = synth.mpc_states(world * n_per, course=course)
    xref, _ = synth.mpc_xref_numpy(st, pind, T, course=course)
    one = O.mpc_solve_batched(st, xref, T, nthreads=1)
    # per-agent outputs are bit-identical whatever the shard count
    assert np.array_equal(np.concatenate([res[0][3], res[1][3]]), one["cost"])
    g0, g1 = res[0][1], res[1][1]
    assert np.array_equal(g0, g1) and g0.shape == (2, 8)    # every rank holds every rank's row
    whole = _stats_numpy(one["cost"], one["status"], one["iters"], 0)
    assert abs(g0[:, 0].sum() - whole[0]) < 1e-9 * abs(whole[0])
    assert g0[:, 1].min() == whole[1] and g0[:, 2].max() == whole[2]
    assert g0[:, 4].sum() == whole[4] and g0[:, 5].sum() == whole[5] and g0[:, 7].sum() == whole[7]
    assert abs(g0[:, 6].sum() - whole[6]) < 1e-9 * abs(whole[6])   # checksum is shard-invariant
    assert res[0][2] == res[1][2] == 11.0                   # max over ranks
