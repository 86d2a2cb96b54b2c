"""Opcode mix per CUDA source line of one kernel in an ncu report (SASS -> line via nvdisasm -g).
python tools/ncu_ops.py rep kernel_substr [top_lines] [lib.so]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(root, "syconn_b200", "libsyk.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines_by_fn = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    fn, cur, seq = None, None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn = m.group(1)
            seq = lines_by_fn.setdefault(fn, [])
            cur = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            seq.append(cur)
which = int(os.environ.get("NCU_RESULT", "0"))  # index of the result inside the report (the page lists them back to back)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
hdr = rows[hi]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:end] if len(r) > ia and r[ia].isdigit()]
cands = [k for k in lines_by_fn if ksub in k and len(lines_by_fn[k]) == len(data)]
if not cands:
    sys.exit("no function with matching SASS length (profile taken with another build?)")
seq = lines_by_fn[cands[0]]
tot = sum(int(r[ia]) for r in data)
per_line = collections.defaultdict(collections.Counter)
smp = collections.Counter()
for r, loc in zip(data, seq):
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    per_line[loc][m.group(1) if m else "?"] += int(r[ia])
    smp[loc] += int(r[ismp] or 0)
ts = sum(smp.values())
srcs = {}
print(f"{cands[0][:60]}: {tot} warp instructions, {ts} samples")
for loc, ops in sorted(per_line.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    n = sum(ops.values())
    text = ""
    if loc:
        pth = os.path.join(root, "syconn_b200", "csrc", loc[0])
        if os.path.exists(pth):
            srcs.setdefault(loc[0], open(pth).read().splitlines())
            text = srcs[loc[0]][loc[1] - 1].strip()[:70]
    mix = " ".join(f"{o}:{100 * c / tot:.1f}" for o, c in ops.most_common(5))
    print(f"{str(loc and loc[1]):>5} inst {100 * n / tot:4.1f}% smp {100 * smp[loc] / ts:4.1f}% | {mix} | {text}")
