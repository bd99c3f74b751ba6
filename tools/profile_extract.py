"""Turn the .ncu-rep files a GPU run left in gpurun_out/ into the small, committed evidence
under profiles/:
    python tools/profile_extract.py gpurun_out/r02_ncu_prefill_step_raw.csv profiles/r02_ncu_prefill_step.csv
writes one row per captured launch with the columns the roofline discussion uses (duration,
DRAM bytes, tensor-pipe %, L2 hit rate, registers, grid, achieved occupancy, top stall reasons)."""
import csv
import subprocess
import sys

KEYS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct_active"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
    ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
]


def main(rep, out):
    if rep.endswith(".csv"):     # a raw page already exported on the GPU box (`ncu -i … --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k, _ in KEYS if k in hdr}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([name + (f" [{units[idx[k]]}]" if k in idx and units[idx[k]] else "")
                    for k, name in KEYS if k in idx])
        for r in rows[2:]:
            if len(r) != len(hdr):
                continue
            vals = []
            for k, _ in KEYS:
                if k not in idx:
                    continue
                v = r[idx[k]]
                if k == "Kernel Name":
                    v = v.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
                    v = v.replace("sb::", "").replace("void ", "")
                    v = v.split("(CUtensorMap")[0].split("(const ")[0].split("(sb::")[0][:80]
                vals.append(v)
            w.writerow(vals)
    print(out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
