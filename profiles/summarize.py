"""Turn gpurun_out/*.ncu-rep / launch-list CSVs into the small text summaries committed under profiles/.
usage: python profiles/summarize.py launches <csv> <out.md> | full <ncu-rep> <out.md> | traffic <ncu-rep> <log_n> [<ncu-rep> <log_n> ...]
`traffic` rewrites profiles/ncu_traffic.json (dram bytes per launch of every captured kernel), the file bench.py reads roofline.traffic from."""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = OrderedDict()
    total = 0.0
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        ns = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1; a[1] += ns
        total += ns
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` -- cold-cache, serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | avg ms | share | grid | block |\n|---|---|---|---|---|---|---|\n")
        for name, (cnt, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {name} | {cnt} | {ns / 1e6:.3f} | {ns / 1e6 / cnt:.3f} | {100 * ns / total:.1f} % | {g} | {b} |\n")
        f.write(f"\ntotal {total / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for r in rows[2:]:
            f.write(f"## {r[ki][:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def traffic(args):
    import json
    import os
    import re

    out = {"source": "ncu --set full --clock-control none captures: " + ", ".join(args[0::2]) + " (profiles/summarize.py traffic)", "kernels": []}
    for path, log_n in zip(args[0::2], args[1::2]):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h, units = rows[0], rows[1]
        ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        # the capture covers the launches of ONE MSM: a kernel that is launched once per scatter range is summed over its launches
        seen = {}
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[ki]).split("(")[0]
            if name not in seen:
                seen[name] = {"kernel": name, "log_n": int(log_n), "launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0}
                out["kernels"].append(seen[name])
            seen[name]["launches"] += 1
            seen[name]["dram_bytes_read"] += to_bytes(r[ri], units[ri])
            seen[name]["dram_bytes_write"] += to_bytes(r[wi], units[wi])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
