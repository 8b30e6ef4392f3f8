#!/usr/bin/env python3
"""Turns an .ncu-rep (ncu --set full --import-source on) into the small text summary kept under profiles/:
key raw metrics per kernel, warp-stall breakdown and the hottest SASS instructions (from the source page).
usage: ncu_summary.py report.ncu-rep [kernel-regex] > profiles/xyz.txt"""
import csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]

def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout

def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    sel = ["-k", "regex:" + rx] if rx else []
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"] + sel))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ki = hdr.index("Kernel Name")
    print("# ncu summary of %s" % rep.split("/")[-1])
    for r in rows:
        print("\n## %s" % r[ki])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-72s %s %s" % (k, r[i], units[i]))
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "-k", "regex:" + re.escape(r[ki].split("(")[0].split("<")[0].split()[-1])]))))
        if len(src) < 3:
            continue
        h = src[1]
        idx = {n: i for i, n in enumerate(h)}
        seen, stalls, insts, total_exec = set(), {}, [], 0
        for row in src[2:]:
            if len(row) < len(h) or row[0] in seen or not row[idx["# Samples"]].isdigit():
                continue
            seen.add(row[0])
            for n in h:
                if n.startswith("stall_") and "Not Issued" not in n and row[idx[n]].isdigit():
                    stalls[n] = stalls.get(n, 0) + int(row[idx[n]])
            ex = row[idx["Instructions Executed"]]
            total_exec += int(ex) if ex.isdigit() else 0
            insts.append((int(row[idx["# Samples"]]), row[idx["Source"]].strip(), ex, row[idx["Avg. Threads Executed"]]))
        tot = sum(stalls.values()) or 1
        print("  SASS instructions: %d   warp-instructions executed: %d" % (len(seen), total_exec))
        print("  warp stall samples:")
        for n, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            print("    %-26s %6.1f %%" % (n, 100.0 * v / tot))
        print("  hottest instructions (samples | executed | avg threads | SASS):")
        for s_, txt, ex, thr in sorted(insts, key=lambda t: -t[0])[:12]:
            print("    %7d | %11s | %3s | %s" % (s_, ex, thr, txt[:90]))

if __name__ == "__main__":
    main()
