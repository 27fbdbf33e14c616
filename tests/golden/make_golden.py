"""Regenerates tests/golden/*.npz.

The reference ships no golden vectors and cannot be built as-is, so the fixtures
are produced here from (a) the CPU restatement (oracle/), (b) independent float64 numpy statements
(tests/test_oracle_ekf.py::ekf_numpy_f64, tests/ref_mpc.py) stored beside it, and - when oracle/_ref was
built - cross-checked against the reference's own sources compiled with header shims
(tests/test_oracle_vs_ref.py does that check; this script only writes the files).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import ref_mpc as M  # noqa: E402
from cpprobotics_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_oracle_ekf import ekf_numpy_f64  # noqa: E402


def main():
    # EKF: 256 agents, 3 steps
    n, steps = 256, 3
    x, P, z, u = synth.ekf_inputs(n, seed=0xA11CE, n_steps=steps)
    xo, Po = O.ekf_step_batched(x, P, z, u, n_steps=steps, nthreads=1)
    dt, Q, R = O.ekf_constants()
    x64, P64 = np.zeros_like(xo, np.float64), np.zeros_like(Po, np.float64)
    for i in range(n):
        xi, Pi = x[:, i].astype(float), P[:, i].reshape(4, 4).T.astype(float)
        for s in range(steps):
            xi, Pi = ekf_numpy_f64(xi, Pi, z[2 * s:2 * s + 2, i].astype(float), u[2 * s:2 * s + 2, i].astype(float),
                                   Q.reshape(4, 4).T.astype(float), R.reshape(2, 2).T.astype(float), dt)
        x64[:, i], P64[:, i] = xi, Pi.T.reshape(-1)
    np.savez_compressed(os.path.join(HERE, "ekf_golden.npz"), x=x, P=P, z=z, u=u, n_steps=steps,
                        x_out=xo, P_out=Po, x_f64=x64, P_f64=P64)
    # PF: 512 particles, 8 landmarks
    px, pw, noise = synth.pf_inputs(512, seed=0xA11CE)
    lm = synth.pf_landmarks(8, seed=0xA11CE)
    pxo, pwo = O.pf_predict_weight_batched(px, pw, noise, lm, nthreads=1)
    np.savez_compressed(os.path.join(HERE, "pf_golden.npz"), px=px, pw=pw, noise=noise, lm=lm, px_out=pxo,
                        pw_out=pwo)
    # MPC: 48 agents, T = 20
    T = 20
    course = synth.mpc_course()
    st, pind = synth.mpc_states(48, seed=0xA11CE, course=course)
    xref, _ = synth.mpc_xref_numpy(st, pind, T, course=course)
    r = O.mpc_solve_batched(st, xref, T, nthreads=1)
    u64, c64 = np.zeros((2, 48)), np.zeros(48)
    for i in range(48):
        ref = M.box_ilqr(st[:, i].astype(float), xref[:, i].reshape(T, 4).T.astype(float), dict(j_tol=0.0))
        u64[:, i] = ref["U"][1, 0], ref["U"][0, 0]
        c64[i] = ref["cost"]
    np.savez_compressed(os.path.join(HERE, "mpc_golden.npz"), x0=st, xref=xref, T=T, sol=r["sol"], u0=r["u0"],
                        cost=r["cost"], status=r["status"], iters=r["iters"], u0_f64=u64, cost_f64=c64)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
