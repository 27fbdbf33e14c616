"""CPU: the oracle against tests/golden/ref_golden.npz - outputs of the reference's OWN source text
(oracle/_ref, see tests/golden/make_ref_golden.py) committed as fixtures, so this pin holds without
/root/reference or a compiler.  Bit-exact wherever the oracle restates the reference arithmetic; the
documented deviations (double accumulation in the PF estimate) are compared with a tolerance."""
import os

import numpy as np

import ref_mpc as M
from cpprobotics_b200 import synth
from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))


def test_ekf_step_is_the_reference_bit_for_bit():
    xo, Po = O.ekf_step_batched(G["ekf_x"], G["ekf_P"], G["ekf_z"], G["ekf_u"], nthreads=1)
    assert np.array_equal(xo, G["ekf_x_out"]) and np.array_equal(Po, G["ekf_P_out"])
    # and the committed inputs are what the generator in this repo produces today
    x, P, z, u = synth.ekf_inputs(512, seed=0xBEEF)
    assert np.array_equal(x, G["ekf_x"]) and np.array_equal(P, G["ekf_P"])


def test_pf_localization_is_the_reference():
    c = O.pf_constants()
    lib = O.lib()
    for case in range(3):
        px, pw, lm, draws = (G[f"pf{case}_{k}"] for k in ("px", "pw", "lm", "draws"))
        NP = px.shape[1]
        pxo, pwo = np.zeros((NP, 4), np.float32), np.zeros(NP, np.float32)
        for ip in range(NP):       # the reference's draws (two doubles per particle, :87-88) are part of the fixture
            xx, ww = np.ascontiguousarray(px[:, ip]).copy(), np.array([pw[ip]], np.float32)
            lib.crb_oracle_pf_particle(xx, ww, np.ascontiguousarray(draws[2 * ip:2 * ip + 2]), c["u"],
                                       c["rsim_diag"], np.ascontiguousarray(lm.reshape(-1)), len(lm),
                                       float(c["Q"]), c["dt"], c["pi"])
            pxo[ip], pwo[ip] = xx, ww[0]
        assert np.array_equal(pxo.T, G[f"pf{case}_px_out"])                      # predict: bit for bit
        s = np.float32(0.0)
        for w in pwo:                                                            # Eigen's float sum (:104)
            s = np.float32(s + w)
        assert np.array_equal((pwo / s).astype(np.float32), G[f"pf{case}_pw_out"])
        _, xe, Pe, _ = O.pf_estimate(np.ascontiguousarray(pxo.T), pwo)           # double accumulation: tolerance
        assert np.abs(xe - G[f"pf{case}_xEst"]).max() < 1e-5
        assert np.abs(Pe.T.reshape(-1) - G[f"pf{case}_PEst"]).max() < 1e-5


def test_resampling_is_the_reference_bit_for_bit():
    for case in range(3):
        px, pw = G[f"pf{case}_px_out"], G[f"pf{case}_pw_out"]
        pxo, pwo, did, _ = O.pf_resample(px, pw, G[f"rs{case}_draws"], reference_mode=True)
        assert bool(G[f"rs{case}_did"]) == did
        assert np.array_equal(pxo, G[f"rs{case}_px_out"]) and np.array_equal(pwo, G[f"rs{case}_pw_out"])


def test_plant_update_and_ref_trajectory_are_the_reference_bit_for_bit():
    st, a, d = G["upd_state"], G["upd_a"], G["upd_delta"]
    for i in range(st.shape[1]):
        assert np.array_equal(O.plant_update(st[:, i], a[i], d[i]), G["upd_out"][:, i])
    cx, cy, cyaw, sp = synth.mpc_course()
    T = int(G["crt_T"])
    for i in range(G["crt_state"].shape[1]):
        xo, to = O.calc_ref_trajectory(np.ascontiguousarray(G["crt_state"][:, i]), cx, cy, cyaw, sp, 1.0, T,
                                       int(G["crt_pind"][i]))
        assert to == G["crt_tind"][i] and np.array_equal(xo.reshape(-1), G["crt_xref"][:, i])


def test_nlp_cost_is_fg_eval():
    T = int(G["crt_T"])
    p = dict(M.DEFAULTS)
    for i in range(G["fg_vars"].shape[0]):
        v, fg = G["fg_vars"][i], G["fg_out"][i]
        X, U = v[:4 * T].reshape(4, T), v[4 * T:].reshape(2, T - 1)
        xr = G["crt_xref"][:, i].reshape(T, 4).T.astype(float)
        assert abs(fg[0] - M.nlp_cost(X, U, xr, p)) <= 1e-9 * abs(fg[0])


def test_dlqr_is_the_reference_bit_for_bit():
    for nx, nu in ((4, 1), (5, 2)):
        r = O.dlqr_batched(G[f"lqr{nx}_A"], G[f"lqr{nx}_B"], G[f"lqr{nx}_Q"], G[f"lqr{nx}_R"], nx, nu)
        assert np.array_equal(r["K"], G[f"lqr{nx}_K"]) and np.array_equal(r["X"], G[f"lqr{nx}_X"])
