"""profiles/summarize_ncu.py <report.ncu-rep> -- the metrics DESIGN.md quotes, one block per captured launch, from `ncu -i ... --page raw --csv`"""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "duration"), ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__cycles_active.avg", "active cycles / SMSP")]
STALL = "smsp__average_warps_issue_stalled_"
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("==", d.get("Kernel Name", "?")[:110])
    for k, name in WANT:
        if k in d:
            print("   %-32s %s %s" % (name, d[k], u.get(k, "")))
    st = sorted(((float(v), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v), reverse=True)
    print("   stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
