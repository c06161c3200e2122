"""Per-launch summary table of an `ncu --set full` report (read here with `ncu -i <rep> --page raw --csv`).

usage: python tools/ncu_summary.py <report.ncu-rep> [--json out.json] > profiles/<name>.txt
Columns: duration, DRAM bytes read / written (traffic = their sum), DRAM throughput % of peak, tensor-pipe cycles active
(tcgen05 and mma.sync both run on it), issue-slot utilisation, shared-memory pipe utilisation, L2 hit rate, L2 -> SM bytes,
registers per thread.  Numbers taken under the profiler are cold-cache and serialised: they explain a kernel, they are
never a bench value.
"""
import csv
import io
import json
import re
import subprocess
import sys

COLS = [
    ("ms", "gpu__time_duration.sum", "ms"),
    ("dram_rd_GB", "dram__bytes_read.sum", "Gbyte"),
    ("dram_wr_GB", "dram__bytes_write.sum", "Gbyte"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("tensor_%", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", None),
    ("hmma_inst_%", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", None),
    ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    ("smem_pipe_%", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", None),
    ("l2_hit_%", "lts__t_sector_hit_rate.pct", None),
    ("l2_to_sm_GB", "l1tex__m_xbar2l1tex_read_bytes.sum", "Gbyte"),
    ("regs", "launch__registers_per_thread", None),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
]
SCALE = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    rep = sys.argv[1]
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]

    def find(name):
        for i, h in enumerate(head):
            if h == name or h.endswith("." + name):
                return i
        return -1

    ik, ig, ib = find("Kernel Name"), find("Grid Size"), find("Block Size")
    idx = [(lab, find(name), want) for lab, name, want in COLS]
    print("# %s" % rep)
    print("%-46s %-14s %5s " % ("kernel", "grid", "block") + " ".join("%11s" % lab for lab, _, _ in idx))
    recs = []
    for r in body:
        name = r[ik].split("(")[0].split("::")[-1]
        if "<" in r[ik] and ">" in r[ik].split("(")[0]:
            name = r[ik].split("(")[0].split("::")[-1]
        grid = r[ig].replace(" ", "")
        block = r[ib].split(",")[0].strip("( ")
        vals = []
        for lab, i, want in idx:
            if i < 0 or not re.match(r"^-?[0-9.,]+(e[-+]?[0-9]+)?$", r[i]):
                vals.append(None)
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if want and u in SCALE and want in SCALE:
                v = v * SCALE[u] / SCALE[want]
            vals.append(v)
        print("%-46s %-14s %5s " % (name[:46], grid, block) + " ".join("%11s" % ("-" if v is None else "%.3f" % v) for v in vals))
        recs.append(dict(kernel=name, grid=grid, block=block, **{lab: v for (lab, _, _), v in zip(idx, vals)}))
    if out_json:
        json.dump(recs, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
