#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch, --set full) as text for profiles/.
usage: ncu_summary.py report.ncu-rep [units_per_launch alg_bytes_per_unit] > profiles/xxx.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_tma_ld.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print(f"# {rep}")
    print(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}")
    for k in KEYS:
        if k in d:
            print(f"{k:72s} {d[k][1]:>18s} {d[k][0]}")
    rd = float(d["dram__bytes_read.sum"][1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][0]]
    wr = float(d["dram__bytes_write.sum"][1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][0]]
    t = float(d["gpu__time_duration.sum"][1]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[d["gpu__time_duration.sum"][0]]
    print(f"{'traffic = dram read + write per launch':72s} {(rd + wr) / 1e6:18.1f} MB   ({(rd + wr) / t / 1e9:.0f} GB/s under the profiler)")
    if len(sys.argv) >= 4:
        units_n, alg = float(sys.argv[2]), float(sys.argv[3])
        print(f"{'algorithmic bytes per launch':72s} {units_n * alg / 1e6:18.1f} MB   (traffic / algorithmic = {(rd + wr) / (units_n * alg):.2f})")
        ins = float(d["smsp__inst_executed.sum"][1])
        print(f"{'warp instructions x 32 / unit':72s} {ins * 32 / units_n:18.2f}")
    print("\nwarp issue-stall reasons (warps stalled per issue-active cycle):")
    st = []
    for k, (u, v) in d.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
        if m:
            st.append((float(v), m.group(1)))
    for v, name in sorted(st, reverse=True):
        if v >= 0.02:
            print(f"  {name:28s} {v:6.2f}")
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    ops = collections.Counter()
    tot = 0
    for r in data:
        src = re.sub(r"^@!?U?P\d\s+", "", r[ci["Source"]].strip())
        if not src:
            continue
        n = int(r[ci["Instructions Executed"]] or 0)
        ops[src.split()[0].split(".")[0]] += n
        tot += n
    print(f"\nexecuted warp instructions by opcode (total {tot}):")
    for op, n in ops.most_common(24):
        print(f"  {op:10s} {n:12d}  {100.0 * n / tot:5.1f} %")
    evidence = [op for op in ops if op in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCHMMA", "LDTM", "HMMA", "FADD2", "FFMA2", "FMUL2", "LDGSTS")]
    print("\nBlackwell-specific / async-copy opcodes present (FADD2 = packed fp32x2, UBLKCP = cp.async.bulk, LDGSTS = cp.async):", ", ".join(sorted(evidence)) or "none")


if __name__ == "__main__":
    main()
