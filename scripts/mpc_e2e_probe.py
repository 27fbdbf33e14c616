"""A/B for the MPC host entry: CRB_MPC_CHUNK x CRB_HOST_ZEROCOPY (both read once per process).
Usage on a GPU box:  for c in 8192 16384 32768 65536; do for z in 0 1; do
  CRB_MPC_CHUNK=$c CRB_HOST_ZEROCOPY=$z python scripts/mpc_e2e_probe.py; done; done"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpprobotics_b200 import synth  # noqa: E402
from cpprobotics_b200.engine import Engine, mpc_default_params  # noqa: E402

T, m = 20, 65536
eng = Engine()
course = synth.mpc_course()
st, pind = synth.mpc_states(m, course=course)
xref, _ = synth.mpc_xref_numpy(st, pind, T, course=course)
prm = mpc_default_params()
nsol = 4 * T + 2 * (T - 1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
hst, hxr = pin(st), pin(xref)
o = dict(sol=torch.empty((nsol, m), dtype=torch.float32).pin_memory(),
         u0=torch.empty((2, m), dtype=torch.float32).pin_memory(),
         cost=torch.empty(m, dtype=torch.float32).pin_memory(),
         status=torch.empty(m, dtype=torch.int32).pin_memory(),
         iters=torch.empty(m, dtype=torch.int32).pin_memory())
for _ in range(2):
    eng.mpc_solve_host(hst, hxr, T, prm, **o)
ts = []
for _ in range(8):
    t0 = time.perf_counter()
    eng.mpc_solve_host(hst, hxr, T, prm, **o)
    ts.append(time.perf_counter() - t0)
print("CRB_MPC_CHUNK=%s CRB_HOST_ZEROCOPY=%s  median %.3f ms  %.2f M solves/s  (min %.3f ms) checksum %.6f" % (
    os.environ.get("CRB_MPC_CHUNK", "default"), os.environ.get("CRB_HOST_ZEROCOPY", "default"),
    np.median(ts) * 1e3, m / np.median(ts) / 1e6, min(ts) * 1e3, float(o["cost"].double().sum())))
