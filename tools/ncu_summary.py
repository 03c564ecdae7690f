"""Condense an ncu report (--set full) into the handful of per-launch metrics quoted in DESIGN.md:  python tools/ncu_summary.py rep.ncu-rep out.csv"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["%s [%s]" % (n, units[i]) if units[i] else n for n, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
print("wrote", out, len(rows) - 2, "launches")
