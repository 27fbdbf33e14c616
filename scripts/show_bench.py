"""Prints the figures of one bench.py JSON line that matter at a glance.  usage: python scripts/show_bench.py file.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])


def g(o, *k):
    for x in k:
        o = o.get(x, {}) if isinstance(o, dict) else {}
    return o


print("n_gpus", d.get("n_gpus"))
print("EKF", d["value"], "frac", d["roofline"]["frac"], "timing", d["config"].get("timing"))
print("EKF e2e", d["e2e"]["value"], "resident", g(d, "e2e", "resident_state", "value"))
print("MPC", d["config"].get("mpc_solves_per_s"), "kernel-only", g(d, "roofline", "mpc", "solves_per_s_kernel_only"),
      "frac", g(d, "roofline", "mpc", "frac"), "peak", g(d, "roofline", "mpc", "peak"), "traffic", g(d, "roofline", "mpc", "traffic"))
print("MPC e2e", g(d, "e2e", "mpc", "value"), "cpu", g(d, "cpu_baseline", "mpc", "value"), "acc",
      d["config"].get("mpc_accuracy_vs_float64_optimum"))
print("MPC hinted", g(d, "roofline", "mpc", "receding_horizon"))
print("MPC config5", d["config"].get("mpc_config5"))
print("MPC config5 hinted", g(d, "roofline", "mpc_config5", "receding_horizon"))
print("NUMA", d["config"].get("host_buffers_numa"))
print("PF", d["config"].get("pf_particles_per_s"), "frac", g(d, "roofline", "pf", "frac"), "full iter ms",
      g(d, "extra", "pf_full_iteration", "ms_per_step"))
print("cpu ekf", g(d, "cpu_baseline", "value"), g(d, "cpu_baseline", "cores"), g(d, "cpu_baseline", "spread"),
      g(d, "cpu_baseline", "pinning"))
print("LQR", g(d, "extra", "lqr", "value"), g(d, "extra", "lqr", "roofline", "frac"))
print("clocks", d.get("clocks"), "stale", d["config"].get("traffic_stale"))
