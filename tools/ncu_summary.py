#!/usr/bin/env python3
"""Markdown summary of an ncu report: one column per captured kernel launch, the metrics the roofline discussion uses.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [--traffic-key KEY --traffic-out profiles/traffic.json --kernel k_accumulate<true>]"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active (%)"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe (% of peak, elapsed)"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy inst (% of peak)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu pipe (%)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__inst_executed_op_shared_atom.sum", "shared atomics (warp inst)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per warp instruction"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=None, help="substring of the kernel name whose DRAM traffic goes to --traffic-out")
    ap.add_argument("--traffic-key", default=None)
    ap.add_argument("--traffic-out", default=None)
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print("| metric | " + " | ".join("`%s`" % r[ik].split("(")[0] for r in data) + " |")
    print("|---|" + "---|" * len(data))
    for name, label in METRICS:
        if name not in hdr:
            continue
        i = hdr.index(name)
        cells = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = ("%.4g" % f) if abs(f) < 1e6 else "%.4g" % f
            except ValueError:
                pass
            cells.append("%s %s" % (v, units[i]) if units[i] else v)
        print("| %s | %s |" % (label, " | ".join(cells)))
    if args.kernel and args.traffic_key and args.traffic_out:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in data:
            if args.kernel in r[ik]:
                total = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
                try:
                    with open(args.traffic_out) as f:
                        t = json.load(f)
                except Exception:
                    t = {}
                t.setdefault("k_accumulate<true>", {})[args.traffic_key] = int(total)
                with open(args.traffic_out, "w") as f:
                    json.dump(t, f, indent=1)
                break


if __name__ == "__main__":
    main()
