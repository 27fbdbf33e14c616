"""Independent float64 numpy statement of the batched-MPC algorithm, and the exact reference NLP.

Two things live here, both test-only:

1. `nlp_cost`, `nlp_solve_scipy`: the reference's NLP exactly as FG_EVAL defines it
   (src/model_predictive_control.cpp:199-252, bounds :283-301), solved with SciPy SLSQP in float64.
   This is the point IPOPT converges to when it is not cut off by max_cpu_time (:328).

2. `box_ilqr`: a dense, textbook float64 implementation of the algorithm libcrb / the C oracle use in
   place of IPOPT (linearise along the rolled-out trajectory -> LQ model with the reference weights ->
   Riccati recursion (with the exact second derivatives of the dynamics, i.e. DDP) on the (state,
   previous input) augmented system -> exact per-stage box QP ->
   clamped non-linear roll-out with step halving).  It is written with generic 6x6 / 6x2 matrices and
   np.linalg, sharing no code and no operation order with oracle/crb_oracle_mpc.c, so agreement
   between the two is evidence about the algorithm, not about a shared bug.
"""
from __future__ import annotations

import numpy as np

REG_EPS = 1.0e-3
DEFAULTS = dict(dt=0.2, wb=2.5, max_steer=np.deg2rad(45.0), max_accel=1.0, max_speed=55.0 / 3.6,
                min_speed=-20.0 / 3.6, w_a=0.01, w_delta=0.01, w_da=0.01, w_ddelta=1.0, w_x=1.0,
                w_y=1.0, w_yaw=0.5, w_v=0.5, max_iter=50, du_th=1e-4, max_ls=4, j_tol=1e-6)


def f_dyn(x, u, p):
    """One step of the dynamics constraint, src/model_predictive_control.cpp:242-245. u = (delta, a)."""
    px, py, yaw, v = x
    d, a = u
    return np.array([px + v * np.cos(yaw) * p["dt"], py + v * np.sin(yaw) * p["dt"],
                     yaw + v * np.tan(d) / p["wb"] * p["dt"], v + a * p["dt"]])


def nlp_cost(X, U, xref, p):
    """fg[0] of FG_EVAL (:199-250).  X [4,T], U [2,T-1] rows (delta, a), xref [4,T]."""
    T = X.shape[1]
    J = 0.0
    for i in range(T - 1):
        J += p["w_a"] * U[1, i] ** 2 + p["w_delta"] * U[0, i] ** 2            # :203-204
    for i in range(T - 2):
        J += p["w_da"] * (U[1, i + 1] - U[1, i]) ** 2                         # :208
        J += p["w_ddelta"] * (U[0, i + 1] - U[0, i]) ** 2                     # :209
    for i in range(T - 1):
        xn = f_dyn(X[:, i], U[:, i], p)                                       # :247-250
        e = xref[:, i + 1] - xn
        J += p["w_x"] * e[0] ** 2 + p["w_y"] * e[1] ** 2 + p["w_yaw"] * e[2] ** 2 + p["w_v"] * e[3] ** 2
    return J


def a_bounds(v, p):
    """Acceleration interval that keeps v + a*dt inside [min_speed, max_speed] (:298-301) and
    |a| <= max_accel (:293-296).  Returns (lo, hi, lo_is_speed, hi_is_speed)."""
    lo_v = (p["min_speed"] - v) / p["dt"]
    hi_v = (p["max_speed"] - v) / p["dt"]
    lo = max(-p["max_accel"], min(lo_v, p["max_accel"]))
    hi = min(p["max_accel"], max(hi_v, -p["max_accel"]))
    return lo, hi, lo_v > -p["max_accel"], hi_v < p["max_accel"]


def rollout(x0, U, p):
    """Clamped non-linear roll-out; returns X [4,T], clamped U."""
    N = U.shape[1]
    X = np.zeros((4, N + 1))
    Uc = U.copy()
    X[:, 0] = x0
    for t in range(N):
        Uc[0, t] = min(max(Uc[0, t], -p["max_steer"]), p["max_steer"])
        lo, hi, _, _ = a_bounds(X[3, t], p)
        Uc[1, t] = min(max(Uc[1, t], lo), hi)
        X[:, t + 1] = f_dyn(X[:, t], Uc[:, t], p)
    return X, Uc


def box_qp2(H, g, lo, hi):
    """Projected-Newton step for min 0.5 u'Hu + g'u on a 2-D box that contains 0 (H may be indefinite).
    Inputs that sit on a bound with the gradient pushing outward are fixed first (Bertsekas' strongly
    active set); the Hessian of the remaining inputs is shifted to be positive definite, and the convex
    box QP in those inputs is solved exactly.  Returns (u, clamped[2], H_reg)."""
    sa = np.zeros(2, bool)
    u = np.zeros(2)
    for i in (0, 1):
        if lo[i] >= 0.0 and g[i] > 0.0:
            sa[i], u[i] = True, lo[i]
        elif hi[i] <= 0.0 and g[i] < 0.0:
            sa[i], u[i] = True, hi[i]
    Hr = H.copy()
    if sa.all():
        return u, np.array([True, True]), Hr
    if sa.any():
        i = int(np.argmax(sa)); j = 1 - i
        Hr[j, j] = max(abs(H[j, j]), REG_EPS)          # curvature magnitude (saddle-free Newton)
        uj = -(g[j] + H[j, i] * u[i]) / Hr[j, j]
        cj = False
        if uj <= lo[j]:
            uj, cj = lo[j], True
        elif uj >= hi[j]:
            uj, cj = hi[j], True
        u[j] = uj
        cl = np.array([True, True]); cl[j] = cj
        return u, cl, Hr
    lam = np.linalg.eigvalsh(H)[0]
    if lam < REG_EPS:                                   # smallest eigenvalue -> max(|lam|, eps)
        Hr = H + (max(-lam, REG_EPS) - lam) * np.eye(2)
    u = -np.linalg.solve(Hr, g)
    if np.all(u >= lo) and np.all(u <= hi):
        return u, np.array([False, False]), Hr
    best, bu, bc = np.inf, None, None
    for i in (0, 1):
        j = 1 - i
        for b in (lo[i], hi[i]):
            uj = -(g[j] + Hr[j, i] * b) / Hr[j, j]
            cj = False
            if uj <= lo[j]:
                uj, cj = lo[j], True
            elif uj >= hi[j]:
                uj, cj = hi[j], True
            cand = np.zeros(2)
            cand[i], cand[j] = b, uj
            val = 0.5 * cand @ Hr @ cand + g @ cand
            if val < best:
                cl = np.zeros(2, bool)
                cl[i], cl[j] = True, cj
                best, bu, bc = val, cand, cl
    return bu, bc, Hr


def backward(X, U, xref, p, gauss_newton=False):
    """Riccati recursion on z = (x, w) with w = previous input.  Returns k [2,N], K [2,6,N], dV1, dV2."""
    N = U.shape[1]
    dt, wb = p["dt"], p["wb"]
    R2 = 2.0 * np.diag([p["w_delta"], p["w_a"]])
    Rd2 = 2.0 * np.diag([p["w_ddelta"], p["w_da"]])
    Q2 = 2.0 * np.diag([p["w_x"], p["w_y"], p["w_yaw"], p["w_v"]])
    P = np.zeros((6, 6))
    pv = np.zeros(6)
    P[:4, :4] = Q2
    pv[:4] = Q2 @ (X[:, N] - xref[:, N])
    ks = np.zeros((2, N))
    Ks = np.zeros((2, 6, N))
    dV1 = dV2 = 0.0
    for t in range(N - 1, -1, -1):
        yaw, v = X[2, t], X[3, t]
        d = U[0, t]
        A = np.eye(4)
        A[0, 2] = -v * np.sin(yaw) * dt
        A[0, 3] = np.cos(yaw) * dt
        A[1, 2] = v * np.cos(yaw) * dt
        A[1, 3] = np.sin(yaw) * dt
        A[2, 3] = np.tan(d) / wb * dt
        B = np.zeros((4, 2))
        B[2, 0] = v * dt / (wb * np.cos(d) ** 2)
        B[3, 1] = dt
        Az = np.zeros((6, 6)); Az[:4, :4] = A            # w_{t+1} = u_t does not depend on z_t
        Bz = np.zeros((6, 2)); Bz[:4] = B; Bz[4:] = np.eye(2)
        has_rate = t >= 1
        lz = np.zeros(6); lu = R2 @ U[:, t]
        Lzz = np.zeros((6, 6)); Luu = R2.copy(); Luz = np.zeros((2, 6))
        if has_rate:
            du = U[:, t] - U[:, t - 1]
            lz[:4] = Q2 @ (X[:, t] - xref[:, t]); Lzz[:4, :4] = Q2
            lz[4:] = -Rd2 @ du; Lzz[4:, 4:] = Rd2
            lu = lu + Rd2 @ du; Luu = Luu + Rd2; Luz[:, 4:] = -Rd2
        qz = lz + Az.T @ pv
        qu = lu + Bz.T @ pv
        Qzz = Lzz + Az.T @ P @ Az
        Quz = Luz + Bz.T @ P @ Az
        Quu = Luu + Bz.T @ P @ Bz
        # second derivatives of the dynamics (:242-245) contracted with the costate of x_{t+1}
        if not gauss_newton:
            p0, p1, p2 = pv[0], pv[1], pv[2]
            Qzz[2, 2] += p0 * (-v * np.cos(yaw) * dt) + p1 * (-v * np.sin(yaw) * dt)
            Qzz[2, 3] += p0 * (-np.sin(yaw) * dt) + p1 * (np.cos(yaw) * dt)
            Qzz[3, 2] += p0 * (-np.sin(yaw) * dt) + p1 * (np.cos(yaw) * dt)
            Quz[0, 3] += p2 * dt / (wb * np.cos(d) ** 2)
            Quu[0, 0] += p2 * v * dt * 2.0 * np.tan(d) / (wb * np.cos(d) ** 2)
        lo_a, hi_a, lo_sp, hi_sp = a_bounds(v, p)
        lo = np.array([-p["max_steer"], lo_a]) - U[:, t]
        hi = np.array([p["max_steer"], hi_a]) - U[:, t]
        k, cl, Hr = box_qp2(Quu, qu, lo, hi)
        K = np.zeros((2, 6))
        # a speed-induced bound on a moves with v: a = (bound - v)/dt  =>  da/dv = -1/dt
        if cl[1]:
            at_lo = k[1] <= lo[1]
            if (at_lo and lo_sp) or ((not at_lo) and hi_sp):
                K[1, 3] = -1.0 / dt
        fr = ~cl
        if fr.all():
            K = -np.linalg.solve(Hr, Quz)
        elif fr.any():
            j = int(np.argmax(fr)); i = 1 - j
            K[j] = -(Quz[j] + Hr[j, i] * K[i]) / Hr[j, j]
        ks[:, t] = k; Ks[:, :, t] = K
        dV1 += k @ qu
        dV2 += 0.5 * k @ Quu @ k
        pv = qz + K.T @ Quu @ k + K.T @ qu + Quz.T @ k
        P = Qzz + K.T @ Quu @ K + K.T @ Quz + Quz.T @ K
        P = 0.5 * (P + P.T)
    return ks, Ks, dV1, dV2


def forward(x0, X, U, ks, Ks, alpha, p):
    N = U.shape[1]
    Xn = np.zeros_like(X); Un = np.zeros_like(U)
    Xn[:, 0] = x0
    for t in range(N):
        dz = np.zeros(6)
        dz[:4] = Xn[:, t] - X[:, t]
        if t >= 1:
            dz[4:] = Un[:, t - 1] - U[:, t - 1]
        u = U[:, t] + alpha * ks[:, t] + Ks[:, :, t] @ dz
        u[0] = min(max(u[0], -p["max_steer"]), p["max_steer"])
        lo, hi, _, _ = a_bounds(Xn[3, t], p)
        u[1] = min(max(u[1], lo), hi)
        Un[:, t] = u
        Xn[:, t + 1] = f_dyn(Xn[:, t], u, p)
    return Xn, Un


def box_ilqr(x0, xref, p=None, U_init=None):
    """Returns dict(X [4,T], U [2,T-1] rows (delta,a), cost, status, iters)."""
    p = dict(DEFAULTS, **(p or {}))
    T = xref.shape[1]
    U = np.zeros((2, T - 1)) if U_init is None else np.array(U_init, float)
    X, U = rollout(np.asarray(x0, float), U, p)
    J = nlp_cost(X, U, xref, p)
    status, iters = 1, 0
    gn = False
    while iters < p["max_iter"]:
        ks, Ks, dV1, dV2 = backward(X, U, xref, p, gauss_newton=gn)
        iters += 1
        accepted = tiny = False
        jacc = 0
        for j in range(p["max_ls"] + 1):
            Xn, Un = forward(x0, X, U, ks, Ks, 0.5 ** j, p)
            Jn = nlp_cost(Xn, Un, xref, p)
            du = np.abs(Un - U).sum()
            if j == 0:   # full step below the input tolerance, or below the cost resolution
                tiny = du <= p["du_th"] or abs(Jn - J) <= p["j_tol"] * abs(J)
            if Jn < J:
                accepted, jacc = True, j
                break
            if tiny:
                break
        if not accepted:
            if tiny:
                status = 0
                break
            if not gn:          # Newton step found no descent: retry this iterate with Gauss-Newton
                gn = True
                continue
            status = 2
            break
        gn = False
        X, U, J = Xn, Un, Jn
        if (jacc == 0 and tiny) or du <= p["du_th"]:
            status = 0
            break
    return dict(X=X, U=U, cost=J, status=status, iters=iters)


def nlp_solve_scipy(x0, xref, p=None, maxiter=500):
    """The reference NLP in single-shooting form (states eliminated through :242-245), SLSQP, float64.
    Box bounds on delta, a; speed bounds as inequality constraints on the rolled-out v."""
    from scipy.optimize import minimize
    p = dict(DEFAULTS, **(p or {}))
    T = xref.shape[1]
    N = T - 1

    def unpack(w):
        return np.vstack([w[:N], w[N:]])

    def roll(w):
        U = unpack(w)
        X = np.zeros((4, T)); X[:, 0] = x0
        for t in range(N):
            X[:, t + 1] = f_dyn(X[:, t], U[:, t], p)
        return X, U

    def obj(w):
        X, U = roll(w)
        return nlp_cost(X, U, xref, p)

    def vcon(w):
        X, _ = roll(w)
        return np.concatenate([X[3, 1:] - p["min_speed"], p["max_speed"] - X[3, 1:]])

    bnds = [(-p["max_steer"], p["max_steer"])] * N + [(-p["max_accel"], p["max_accel"])] * N
    best = None
    for w0 in (np.zeros(2 * N),):
        r = minimize(obj, w0, method="SLSQP", bounds=bnds, constraints=[dict(type="ineq", fun=vcon)],
                     options=dict(maxiter=maxiter, ftol=1e-12))
        if best is None or r.fun < best.fun:
            best = r
    X, U = roll(best.x)
    return dict(X=X, U=U, cost=best.fun, success=best.success, nit=best.nit)
