"""Probe: staged host pipeline (crb_*_host) versus launching the resident kernels directly on pinned,
device-mapped host memory (zero-copy over PCIe).  Also prints the CPU oracle's thread scaling on this box.
Run on a GPU box: python scripts/zerocopy_probe.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpprobotics_b200 import synth  # noqa: E402
from cpprobotics_b200.engine import Engine, mpc_default_params  # noqa: E402


def wall(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def main():
    eng = Engine()
    lib = eng.lib
    # ---------------- EKF
    n = 1 << 20
    host = synth.ekf_inputs(n)
    a = [pin(h.copy()) for h in host]
    b = [pin(h.copy()) for h in host]
    eng.ekf_estimation_host(*a)
    eng._ekf(lib.crb_ekf_step_batched, None, *b, None, 1)
    eng.sync()
    print("EKF zero-copy == staged:", all(torch.equal(p, q) for p, q in zip(a, b)), flush=True)
    t_st = wall(lambda: eng.ekf_estimation_host(*a), 10)
    t_zc = wall(lambda: (eng._ekf(lib.crb_ekf_step_batched, None, *b, None, 1), eng.sync()), 10)
    print("EKF staged %.3f ms %.1f M/s | zero-copy %.3f ms %.1f M/s (%.1f GB/s duplex total)" % (
        t_st * 1e3, n / t_st / 1e6, t_zc * 1e3, n / t_zc / 1e6, 176 * n / t_zc / 1e9), flush=True)
    # ---------------- PF
    lm = synth.pf_landmarks()
    px, pw, noise = synth.pf_inputs(n)
    a = [pin(px.copy()), pin(pw.copy()), pin(noise.copy())]
    b = [pin(px.copy()), pin(pw.copy()), pin(noise.copy())]
    eng.pf_predict_weight_host(a[0], a[1], a[2], lm)
    eng._pf(lib.crb_pf_predict_weight_batched, None, b[0], b[1], b[2], lm, None, 0)
    eng.sync()
    print("PF zero-copy == staged:", all(torch.equal(p, q) for p, q in zip(a, b)), flush=True)
    t_st = wall(lambda: eng.pf_predict_weight_host(a[0], a[1], a[2], lm), 10)
    t_zc = wall(lambda: (eng._pf(lib.crb_pf_predict_weight_batched, None, b[0], b[1], b[2], lm, None, 0),
                         eng.sync()), 10)
    print("PF staged %.3f ms %.1f M/s | zero-copy %.3f ms %.1f M/s" % (
        t_st * 1e3, n / t_st / 1e6, t_zc * 1e3, n / t_zc / 1e6), flush=True)
    # ---------------- MPC
    T, m = 20, 65536
    course = synth.mpc_course()
    st, pind = synth.mpc_states(m, course=course)
    xref, _ = synth.mpc_xref_numpy(st, pind, T, course=course)
    prm = mpc_default_params()
    nsol = 4 * T + 2 * (T - 1)

    def outs():
        return dict(sol=torch.empty((nsol, m), dtype=torch.float32).pin_memory(),
                    u0=torch.empty((2, m), dtype=torch.float32).pin_memory(),
                    cost=torch.empty(m, dtype=torch.float32).pin_memory(),
                    status=torch.empty(m, dtype=torch.int32).pin_memory(),
                    iters=torch.empty(m, dtype=torch.int32).pin_memory())
    hst, hxr = pin(st), pin(xref)
    oa, ob = outs(), outs()
    eng.mpc_solve_host(hst, hxr, T, prm, **oa)
    eng._mpc(lib.crb_mpc_solve_batched, None, hst, hxr, T, prm, None, ob["sol"], ob["u0"], ob["cost"],
             ob["status"], ob["iters"])
    eng.sync()
    print("MPC zero-copy == staged:", all(torch.equal(oa[k], ob[k]) for k in oa), flush=True)
    t_st = wall(lambda: eng.mpc_solve_host(hst, hxr, T, prm, **oa), 5, 1)
    t_zc = wall(lambda: (eng._mpc(lib.crb_mpc_solve_batched, None, hst, hxr, T, prm, None, ob["sol"],
                                  ob["u0"], ob["cost"], ob["status"], ob["iters"]), eng.sync()), 5, 1)
    print("MPC staged %.3f ms %.2f M/s | zero-copy %.3f ms %.2f M/s" % (
        t_st * 1e3, m / t_st / 1e6, t_zc * 1e3, m / t_zc / 1e6), flush=True)
    # ---------------- CPU oracle thread scaling
    from oracle import oracle as O
    x, P, z, u = (h.copy() for h in host)
    cores = len(os.sched_getaffinity(0))
    print("host cpus in affinity mask:", cores, "loadavg", open("/proc/loadavg").read().strip(), flush=True)
    thr = 1
    while thr <= cores:
        O.ekf_step_batched(x, P, z, u, nthreads=thr, inplace=True)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            O.ekf_step_batched(x, P, z, u, nthreads=thr, inplace=True)
            ts.append(time.perf_counter() - t0)
        print("oracle EKF %4d threads: best %.1f ms  %.1f M updates/s" % (thr, min(ts) * 1e3,
                                                                         n / min(ts) / 1e6), flush=True)
        thr *= 2


if __name__ == "__main__":
    main()
