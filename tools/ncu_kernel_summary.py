"""Per-kernel summary of an `ncu --set full` report: duration, DRAM traffic and achieved GB/s, issue-slot
utilisation, shared-memory bank conflicts, the top warp-stall reasons, registers and shared memory.

    python tools/ncu_kernel_summary.py gpurun_out/prof.ncu-rep [peak_GBps] > profiles/<name>.txt
"""
import csv
import json
import os
import re
import subprocess
import sys

rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default


def scale(name, v, want):
    """convert v from the column's unit to `want` (ns / byte based)"""
    u = units[col[name]] if name in col else ""
    m = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    u = u.split("/")[0]
    return v * m.get(u if u in m else u.lower(), 1.0)


stall_cols = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)]
print("# %s — peak used for 'of peak': %.0f GB/s (measured copy bandwidth)" % (os.path.basename(rep), peak))
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", name).replace("<unnamed>::", "")
    dur_ns = scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum"), "ns")
    rd = scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum"), "byte")
    wr = scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum"), "byte")
    gbs = (rd + wr) / dur_ns if dur_ns else 0.0
    stalls = sorted(((f(r, c), re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active", c).group(1)) for c in stall_cols), reverse=True)
    top = ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4] if v > 0.005)
    grid = r[col["Grid Size"]] if "Grid Size" in col else "?"
    blk = r[col["Block Size"]] if "Block Size" in col else "?"
    print("\n%s   grid %s block %s" % (short, grid, blk))
    print("  duration %.1f us | DRAM read %.1f MB + write %.1f MB = %.1f GB/s (%.1f %% of peak; ncu dram throughput %.1f %%)"
          % (dur_ns / 1e3, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak, f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
    print("  issue slots busy %.1f %% | warps active %.1f %% of peak | regs/thread %d | smem/block %d B static + %d B dynamic"
          % (f(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
             int(f(r, "launch__registers_per_thread")), int(scale("launch__shared_mem_per_block_static", f(r, "launch__shared_mem_per_block_static"), "byte")),
             int(scale("launch__shared_mem_per_block_dynamic", f(r, "launch__shared_mem_per_block_dynamic"), "byte"))))
    print("  shared-memory bank conflicts %d (ld %d, st %d, atom %d)"
          % (int(f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")), int(f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum")),
             int(f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum")), int(f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum"))))
    print("  warp stalls per issue: %s" % top)
