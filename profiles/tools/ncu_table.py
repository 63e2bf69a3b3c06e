"""Summarise an `ncu --csv --page raw` capture: one row per kernel with the metrics the roofline argument uses."""
import csv, sys
lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.reader(lines[start:]))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdB"), ("dram__bytes_write.sum", "wrB"),
        ("smsp__inst_executed.sum", "winst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wf%"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "l1wf"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit"), ("lts__t_sector_hit_rate.pct", "l2hit"),
        ("launch__registers_per_thread", "regs"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
        ("lts__t_sectors_srcunit_tex.sum", "l2sect")]
ik = hdr.index("Kernel Name")
print("| kernel | " + " | ".join(n for _, n in want) + " |")
print("|---|" + "---|" * len(want))
for r in rows[2:]:
    vals = []
    for m, _ in want:
        vals.append(r[hdr.index(m)] if m in hdr else "-")
    print("| `" + r[ik][:48] + "` | " + " | ".join(vals) + " |")
