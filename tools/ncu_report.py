"""Summarise an `ncu --set full` report (exported with `ncu -i x.ncu-rep --page raw --csv > x.csv`) as a markdown table
and, with --traffic KEY, write the measured DRAM bytes per step into profiles/r02_traffic.json.
usage: python tools/ncu_report.py raw.csv [--traffic lfa_pool|pp_dense] [--source "text"]"""
import csv
import json
import os
import re
import sys

METRICS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("launch__registers_per_thread", "regs"),
           ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
           ("lts__t_sector_hit_rate.pct", "L2 hit %"),
           ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
           ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
           ("smsp__inst_executed.sum", "warp instr"),
           ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
           ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
           ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
           ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
           ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
           ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle")]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    hdr, units, rows = load(sys.argv[1])
    ki = hdr.index("Kernel Name")
    names = [re.sub(r"\(.*", "", r[ki]).replace("void ", "") for r in rows]
    print("| metric | " + " | ".join("`%s`" % n for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for m, label in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        vals = []
        for r in rows:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(("%.4g" % v) + (" " + units[i] if units[i] not in ("", "%", "inst", "register/thread") else ""))
            except ValueError:
                vals.append(r[i])
        print("| %s | " % label + " | ".join(vals) + " |")
    if "--traffic" in sys.argv:
        key = sys.argv[sys.argv.index("--traffic") + 1]
        src = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else os.path.basename(sys.argv[1])
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        tot = sum(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for r in rows)
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
        d = json.load(open(out)) if os.path.exists(out) else {}
        d[key] = dict(dram_bytes_per_step=tot, launches=len(rows), source=src,
                      per_launch={n: to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for n, r in zip(names, rows)})
        json.dump(d, open(out, "w"), indent=1)
        print("\nDRAM read + write of these %d launches: %.1f MB -> %s[%s]" % (len(rows), tot / 1e6, out, key))


if __name__ == "__main__":
    main()
