"""Turns an .ncu-rep capture (brought back in gpurun_out/) into the small text/JSON summaries kept in profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_r01_ns   # writes <prefix>_metrics.csv/.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_drain",
]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": rep, "kernels": {}}
    with open(prefix + "_metrics.csv", "w") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200r::", "")
            short = name.split("<")[0]
            entry = {}
            for k in KEYS:
                if k in idx and r[idx[k]] not in ("", "n/a"):
                    w.writerow([name, k, r[idx[k]], units[idx[k]]])
                    try:
                        entry[k] = float(r[idx[k]])
                    except ValueError:
                        entry[k] = r[idx[k]]
                    entry[k + "__unit"] = units[idx[k]]
            def to_bytes(key):
                v, u = entry.get(key, 0.0), entry.get(key + "__unit", "byte")
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            entry["dram_bytes_per_launch"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            out["kernels"].setdefault(short, entry)  # first captured launch of each kernel
    with open(prefix + "_metrics.json", "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", prefix + "_metrics.csv/.json", "kernels:", list(out["kernels"]))


if __name__ == "__main__":
    main()
