"""Kernel-only timing of the MPC solve for A/B runs (env knobs are read once per process):
   CRB_MPC_VARIANT=0|1, CRB_MPC_WARPS, CRB_MPC_SLOTS.  Prints one line per batch size.
   Checks the first 2048 problems against the oracle bit for bit (status, iters, u0, cost)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpprobotics_b200 import Engine, mpc_default_params, synth  # noqa: E402

T = 20
sizes = [int(s) for s in (sys.argv[1:] or ["65536", "1048576"])]
eng = Engine(0)
tag = " ".join(f"{k}={os.environ[k]}" for k in ("CRB_MPC_VARIANT", "CRB_MPC_WARPS", "CRB_MPC_SLOTS") if k in os.environ)
for n in sizes:
    course = synth.mpc_course()
    st, pind = synth.mpc_states(n, course=course)
    xref, _ = synth.mpc_xref_numpy(st, pind, T, course=course)
    prm = mpc_default_params()
    nsol = 4 * T + 2 * (T - 1)
    sets = [(torch.from_numpy(st).cuda(), torch.from_numpy(xref).cuda()) for _ in range(3)]
    sol = torch.empty((nsol, n), dtype=torch.float32, device="cuda")
    u0 = torch.empty((2, n), dtype=torch.float32, device="cuda")
    cost = torch.empty(n, dtype=torch.float32, device="cuda")
    status = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    iters = torch.empty(n, dtype=torch.int32, device="cuda")
    eng.bind_current_stream()
    reps = 20 if n <= 1 << 17 else 5
    for k in range(3):
        eng.mpc_solve(*sets[k % 3], T, prm, sol=sol, u0=u0, cost=cost, status=status, iters=iters)
    torch.cuda.synchronize()
    times = []
    for k in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.mpc_solve(*sets[k % 3], T, prm, sol=sol, u0=u0, cost=cost, status=status, iters=iters)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times = np.array(times)
    ok = "unchecked"
    if os.environ.get("MPC_PROBE_CHECK", "1") == "1":
        from oracle import oracle as O
        m = min(n, 2048)
        want = O.mpc_solve_batched(st[:, :m], xref[:, :m], T, O.mpc_params())
        got = dict(status=status[:m].cpu().numpy(), iters=iters[:m].cpu().numpy(), u0=u0[:, :m].cpu().numpy(),
                   cost=cost[:m].cpu().numpy(), sol=sol[:, :m].cpu().numpy())
        ok = "bit-exact" if all(np.array_equal(got[k], want[k]) for k in got) else "MISMATCH"
    st_all = status.cpu().numpy()
    print(f"[{tag}] n={n} median {np.median(times):.4f} ms min {times.min():.4f} ms  "
          f"{n / np.median(times) * 1e3:.4e} solves/s (min-time {n / times.min() * 1e3:.4e})  "
          f"mean_iters {iters.float().mean().item():.3f} converged {(st_all == 0).mean():.6f} "
          f"unsolved {(st_all == -7).sum()} vs oracle: {ok}", flush=True)
eng.close()
