"""Summarise an .ncu-rep: headline metrics + hottest source lines.  python tools/ncu_summary.py rep [n_lines]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_lsu.sum", "sm__cycles_elapsed.avg"]
for vals in rows[2:]:
    print("== kernel:", vals[hdr.index("Kernel Name")][:90])
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"  {h:70s} {v} {u}")
        elif "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(v) > 0.25:
                    print(f"  {h:70s} {v}")
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
ia, isrc, ismp, ith = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
data = [r for r in rows[hi + 1:] if len(r) > ia and r[ia].isdigit()]
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[ismp] or 0) for r in data)
print(f"total warp-inst {tot}  samples {ts}  sass lines {len(data)}")
# aggregate samples / instructions over windows of consecutive SASS lines to find hot regions
W = 24
best = []
for i in range(0, len(data), W):
    blk = data[i:i + W]
    best.append((sum(int(r[ismp] or 0) for r in blk), sum(int(r[ia]) for r in blk), i))
for smp, ins, i in sorted(best, reverse=True)[:topn]:
    blk = data[i:i + W]
    ops = " ".join(sorted(set(r[isrc].split()[1 if r[isrc].lstrip().startswith("@") else 0].split(".")[0] for r in blk if r[isrc].split())))
    print(f"sass[{i:5d}:{i+W:5d}] samples {smp/ts*100:5.1f}%  inst {ins/tot*100:5.1f}%  thr~{blk[0][ith]:>5}  {ops[:120]}")
