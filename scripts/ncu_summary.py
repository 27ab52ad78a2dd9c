"""Summarise an .ncu-rep (raw page CSV) into a compact per-kernel table: python scripts/ncu_summary.py rep out.csv"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
idx = [(k, hdr.index(k)) for k in keep if k in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([k for k, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
for r in rows[2:]:
    d = {k: r[i] for k, i in idx}
    print(f"{d['Kernel Name'][:28]:28s} grid {d['Grid Size']:>14s} t {float(d['gpu__time_duration.sum']):7.1f}us "
          f"dram R {float(d['dram__bytes_read.sum']):6.1f} W {float(d['dram__bytes_write.sum']):6.1f} MB "
          f"dram% {float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):5.1f} "
          f"warps% {float(d['sm__warps_active.avg.pct_of_peak_sustained_active']):5.1f} regs {d['launch__registers_per_thread']:>3s} "
          f"inst {float(d['smsp__inst_executed.sum'])/1e6:6.1f}M issue% {float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']):5.1f} "
          f"L2hit {float(d['lts__t_sector_hit_rate.pct']):5.1f}")
