"""Digest of an ncu report for one kernel: headline metrics + the instructions with the most stall samples.
   usage: python scripts/ncu_digest.py gpurun_out/prof.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "launch__shared_mem_per_block_dynamic"]
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, h in enumerate(hdr):
        if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if "issue_stalled" in h and v < 0.05:
                continue
            print(f"  {h:90s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print(f"total samples {tot:.0f}; instructions executed {sum(f(r, 'Instructions Executed') for r in data):.0f}")
agg = {}
for r in data:
    for k in h:
        if k.startswith("stall_") and "Not" not in k:
            agg[k] = agg.get(k, 0) + f(r, k)
print("stall samples:", ", ".join(f"{k[6:]} {v / tot * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / tot > 0.005))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    st = {k: f(r, k) for k in h if k.startswith("stall_") and "Not" not in k}
    big = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {r[0][-5:]} {r[1][:58]:58s} smp {f(r, '# Samples'):6.0f} exec {f(r, 'Instructions Executed'):9.0f} thr {f(r, 'Avg. Threads Executed'):4.1f} "
          + " ".join(f"{k[6:]}={v:.0f}" for k, v in big))
