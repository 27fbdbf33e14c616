"""The independent float64 numpy statement of the MPC solver lives in oracle/ref_mpc_f64.py (bench.py's
cpu_baseline leg uses it too, to report the GPU solver's distance to the float64 optimum); tests keep importing
it under this name."""
from oracle.ref_mpc_f64 import *  # noqa: F401,F403
from oracle.ref_mpc_f64 import box_ilqr, nlp_solve_scipy  # noqa: F401
