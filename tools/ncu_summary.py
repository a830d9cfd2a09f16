#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (`ncu -i X --page raw --csv` piped in or given as a csv path):
duration, DRAM bytes, issue-slot / pipe utilisation, warp-stall breakdown (cycles stalled per issued instruction)."""
import csv
import re
import subprocess
import sys

COLS = [("dur_us", "gpu__time_duration.sum"), ("rd_MB", "dram__bytes_read.sum"), ("wr_MB", "dram__bytes_write.sum"),
        ("regs", "launch__registers_per_thread"), ("smemKB", "launch__shared_mem_per_block_allocated"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed"),
        ("lsu_sh%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("Minst", "smsp__inst_executed.sum"), ("bankconf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "wait", "not_selected", "sleeping", "membar", "dispatch_stall", "no_instruction", "branch_resolving", "misc"]


def short(n):
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return re.sub(r"\(.*", "", n)


def main(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
    else:
        rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout, lineterminator="\n")      # kernel names contain commas (template arguments): quoted
    w.writerow(["kernel", "grid", "block"] + [c for c, _ in COLS] + ["stalls(cycles per issued inst)"])
    for r in data:
        out = [short(r[ix["Kernel Name"]]), r[ix["Grid Size"]].replace(",", " "), r[ix["Block Size"]].replace(",", " ")]
        for c, k in COLS:
            v = r[ix[k]].replace(",", "") if k in ix else ""
            try:
                f = float(v)
                if c == "Minst":
                    f /= 1e6
                if c in ("rd_MB", "wr_MB"):
                    u = units[ix[k]]
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if c == "dur_us":
                    u = units[ix[k]]
                    f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                v = "%.2f" % f
            except ValueError:
                pass
            out.append(v)
        st = []
        for s in STALLS:
            k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if k in ix:
                try:
                    f = float(r[ix[k]])
                except ValueError:
                    continue
                if f >= 0.3:
                    st.append("%s=%.1f" % (s, f))
        out.append(" ".join(st))
        w.writerow(out)


if __name__ == "__main__":
    main(sys.argv[1])
