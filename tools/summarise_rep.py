#!/usr/bin/env python
"""tools/summarise_rep.py -- key ncu metrics of every distinct kernel in one or more .ncu-rep files, as a markdown table.
    python tools/summarise_rep.py "title" out.md rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, subprocess, sys

title, out_path, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
        ("smsp__inst_executed_op_shfl.sum" if False else "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %")]
seen = {}
units = None
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units_row = rows[0], rows[1]
    ni = h.index("Kernel Name")
    for r in rows[2:]:
        k = r[ni].split("(")[0].replace("pm::", "")
        if k not in seen:
            seen[k] = {n: (r[h.index(n)], units_row[h.index(n)]) for n, _ in want if n in h}
with open(out_path, "w") as f:
    f.write("# %s\n\n" % title)
    f.write("| metric | " + " | ".join(seen) + " |\n|---|" + "---|" * len(seen) + "\n")
    for n, label in want:
        if not any(n in v for v in seen.values()):
            continue
        unit = next(v[n][1] for v in seen.values() if n in v)
        f.write("| %s [%s] | " % (label, unit) + " | ".join(seen[k].get(n, ("", ""))[0] for k in seen) + " |\n")
    f.write("\nSources (scratch, not tracked): %s\n" % ", ".join(reps))
print(open(out_path).read())
