"""Per-kernel metric table of an ncu report (raw page).  usage: python profiles/ncu_kernels.py <report.ncu-rep> [name-filter]"""
import csv, subprocess, sys
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    if flt and flt not in r[ki]:
        continue
    print('----', r[ki])
    for w in want:
        if w in hdr:
            print(f"  {w}: {r[hdr.index(w)]} {units[hdr.index(w)]}")
    stalls = [(float(r[i]), hdr[i]) for i in range(len(hdr)) if 'smsp__average_warps_issue_stalled' in hdr[i]
              and '_per_issue_active' in hdr[i] and r[i] not in ('', 'n/a')]
    print("  stalls:", ", ".join(f"{n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                                for v, n in sorted(stalls, reverse=True)[:6]))
