"""Key counters per kernel out of an `ncu --set full` report (read on the CPU box):
    python scripts/ncu_summary.py gpurun_out/r2_kernels.ncu-rep profiles/r2_ncu_full_kernels.csv
Keeps the SECOND launch of every kernel/grid (the first one is the cold warm-up launch of scripts/profile_kernels.py)."""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launch"] + [f"{n} [{units[idx[m]]}]" if m in idx else n for m, n in COLS])
        for r in data:
            key = (r[idx["Kernel Name"]], r[idx["launch__grid_size"]])
            seen[key] = seen.get(key, 0) + 1
            if seen[key] != 2:
                continue
            w.writerow([r[idx["Kernel Name"]][:110], r[idx["ID"]]] + [r[idx[m]] if m in idx else "" for m, _ in COLS])
    print("wrote", out)


if __name__ == "__main__":
    main()
