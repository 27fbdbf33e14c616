"""CPU: size-independent properties of the oracle (hypothesis drives the seeds).  These are the properties
the GPU tests rely on at the BASELINE sizes, checked here where every case is cheap."""
import numpy as np
from hypothesis import given, settings, strategies as st

from cpprobotics_b200 import synth
from oracle import oracle as O

SEEDS = st.integers(min_value=1, max_value=2**31 - 1)
FAST = settings(max_examples=int(__import__("os").environ.get("CRB_HYP_EXAMPLES", 8)), deadline=None, derandomize=True, database=None)


@FAST
@given(SEEDS)
def test_ekf_multi_step_is_composition_and_batch_order_is_irrelevant(seed):
    n, k = 257, 4
    x, P, z, u = synth.ekf_inputs(n, seed=seed, n_steps=k)
    xa, Pa = O.ekf_step_batched(x, P, z, u, n_steps=k, nthreads=2)
    xb, Pb = x.copy(), P.copy()
    for s in range(k):
        xb, Pb = O.ekf_step_batched(xb, Pb, z[2 * s:2 * s + 2], u[2 * s:2 * s + 2], nthreads=1)
    assert np.array_equal(xa, xb) and np.array_equal(Pa, Pb)
    perm = np.random.default_rng(seed).permutation(n)
    xc, Pc = O.ekf_step_batched(x[:, perm], P[:, perm], z[:, perm], u[:, perm], n_steps=k, nthreads=3)
    assert np.array_equal(xc, xa[:, perm]) and np.array_equal(Pc, Pa[:, perm])
    assert np.isfinite(Pa).all()
    # covariance stays (numerically) symmetric positive on the diagonal
    Pm = Pa.T.reshape(n, 4, 4)
    assert (np.einsum("nii->ni", Pm) > 0).all()
    assert np.abs(Pm - Pm.transpose(0, 2, 1)).max() <= 1e-4 * np.abs(Pm).max()


@FAST
@given(SEEDS)
def test_pf_weights_only_shrink_by_at_most_the_prefactor_and_no_landmark_is_identity(seed):
    n = 300
    px, pw, noise = synth.pf_inputs(n, seed=seed)
    lm = synth.pf_landmarks(5, seed=seed)
    px1, pw1 = O.pf_predict_weight_batched(px, pw, noise, lm)
    px0, pw0 = O.pf_predict_weight_batched(px, pw, noise, np.zeros((0, 3), np.float32))
    assert np.array_equal(px0, px1)                       # the motion step does not depend on the landmarks
    assert np.array_equal(pw0, pw)                        # no observation: weights untouched (:92 loop empty)
    pre = 1.0 / np.sqrt(2.0 * 3.141592653 * 0.01)         # gauss_likelihood(0, 0.1), the largest factor
    assert (pw1 <= pw * np.float32(pre ** 5 * (1 + 1e-5))).all() and (pw1 >= 0).all()


@FAST
@given(SEEDS)
def test_resample_is_monotone_and_counts_follow_the_weights(seed):
    n = 4096
    rng = np.random.default_rng(seed)
    w = rng.gamma(0.3, 1.0, n)
    w = (w / w.sum()).astype(np.float32)
    px = np.stack([np.arange(n, dtype=np.float32)] + [rng.standard_normal(n).astype(np.float32) for _ in range(3)])
    u = (1.0 + rng.random(n)).astype(np.float64)
    pxo, pwo, did, neff = O.pf_resample(px, w, u, nth=float(n))
    assert did and 1.0 <= neff <= n
    src = pxo[0].astype(np.int64)
    assert (np.diff(src) >= 0).all() and np.array_equal(pxo[1:], px[1:, src])
    counts = np.bincount(src, minlength=n)
    assert np.abs(counts - n * w.astype(np.float64)).max() <= 2.0 + 1e-3 * n * w.max()
    assert np.array_equal(pwo, np.full(n, np.float32(1.0 / n)))


@FAST
@given(SEEDS)
def test_mpc_solution_is_feasible_consistent_and_a_fixed_point(seed):
    n, T = 48, 20
    course = synth.mpc_course()
    stt, pind = synth.mpc_states(n, seed=seed, course=course)
    xref, _ = synth.mpc_xref_numpy(stt, pind, T, course=course)
    p = O.mpc_params()
    r = O.mpc_solve_batched(stt, xref, T, p)
    ok = r["status"] == 0
    assert ok.mean() > 0.9
    N = T - 1
    sol = r["sol"]
    X = sol[:4 * T].reshape(4, T, n)
    delta, acc = sol[4 * T:4 * T + N], sol[4 * T + N:]
    assert (np.abs(delta) <= p.max_steer * (1 + 1e-6)).all() and (np.abs(acc) <= p.max_accel * (1 + 1e-6)).all()
    assert np.array_equal(r["u0"][0], acc[0]) and np.array_equal(r["u0"][1], delta[0])
    assert np.array_equal(X[:, 0, :], stt)                        # :213-216 initial-state constraint
    v = X[3]
    assert (v[1:] <= p.max_speed + 1e-4).all() and (v[1:] >= p.min_speed - 1e-4).all()
    # dynamics constraints :242-245 hold along the returned trajectory (float32 roll-out)
    x_next = X[0, :-1] + v[:-1] * np.cos(X[2, :-1]) * p.dt
    yaw_next = X[2, :-1] + v[:-1] * np.tan(delta) / p.wb * p.dt
    assert np.abs(x_next - X[0, 1:]).max() <= 2e-4 and np.abs(yaw_next - X[2, 1:]).max() <= 2e-5
    # idempotence: restarting from the converged controls stops almost immediately at the same cost
    u_init = np.concatenate([delta, acc]).astype(np.float32)
    r2 = O.mpc_solve_batched(stt, xref, T, p, u_init=u_init)
    assert (r2["iters"][ok] <= 3).all()
    assert np.abs(r2["cost"][ok] - r["cost"][ok]).max() <= 2e-5 * np.abs(r["cost"][ok]).max()


@FAST
@given(SEEDS)
def test_dlqr_gain_stabilises_the_plant(seed):
    for nx, nu in ((4, 1), (5, 2)):
        A, B, Q, R = synth.lqr_inputs(16, nx, seed=seed)
        r = O.dlqr_batched(A, B, Q, R, nx, nu)
        for i in range(16):
            a = A[:, i].reshape(nx, nx).T.astype(float)
            b = B[:, i].reshape(nu, nx).T.astype(float)
            k = r["K"][:, i].reshape(nx, nu).T.astype(float)
            assert np.abs(np.linalg.eigvals(a - b @ k)).max() < 1.0
