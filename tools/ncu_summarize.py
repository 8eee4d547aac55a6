"""Condense `ncu --page raw --csv` exports (tools/profile_round.sh) into profiles/<tag>_ncu_summary.md and
profiles/<tag>_ncu_traffic.json (DRAM bytes per unit of each kernel, read by bench.py for roofline.traffic).

usage: python tools/ncu_summarize.py <tag> gpurun_out/<tag>_prof_single.raw.csv [more.csv ...] --units k_decode=N ...
The first launch of each kernel name found in the files is used."""
import csv
import json
import os
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
]


def read(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = {}
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        k = d["Kernel Name"].split("(")[0].split("::")[-1]
        k = "k_equation_vargen" if k == "k_equation<1>" else k.split("<")[0]   # k_equation<0>: fixed base, <1>: var-generator
        if k not in out:
            out[k] = (d, dict(zip(names, units)))
    return out


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    tag = sys.argv[1]
    files = [a for a in sys.argv[2:] if not a.startswith("--") and "=" not in a]
    units = dict(a.split("=") for a in sys.argv[2:] if "=" in a and not a.startswith("--"))
    kernels = {}
    for f in files:
        for k, v in read(f).items():
            kernels.setdefault(k, v)
    order = [k for k in ("k_decode", "k_challenge", "k_equation", "k_equation_vargen", "k_rtest", "k_agg_coeffs", "k_aggregate") if k in kernels]
    lines = [f"# {tag} ncu summary", "",
             "ncu --set full --clock-control none --import-source on (tools/profile_round.sh); one launch per kernel after warm-up. "
             "Reports stay in gpurun_out/; the launch list of the default bench command is `" + tag + "_launches.csv`.", "",
             "| metric | unit | " + " | ".join(order) + " |", "|---|---|" + "---|" * len(order)]
    for m in METRICS:
        if not any(m in kernels[k][0] for k in order):
            continue
        unit = next((kernels[k][1].get(m, "") for k in order if m in kernels[k][0]), "")
        lines.append(f"| {m} | {unit} | " + " | ".join(kernels[k][0].get(m, "") for k in order) + " |")
    os.makedirs("profiles", exist_ok=True)
    open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(lines) + "\n")
    traffic = {"source": f"profiles/{tag}_ncu_summary.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch)", "kernels": {}}
    for k in order:
        d, u = kernels[k]
        b = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        n = int(units.get(k, 0))
        traffic["kernels"][k] = {"dram_bytes_in_profiled_launch": b, "units_in_profiled_launch": n, "dram_bytes_per_unit": b / n if n else None}
    json.dump(traffic, open(f"profiles/{tag}_ncu_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
