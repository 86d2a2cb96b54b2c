"""Aggregate an ncu SASS profile by CUDA source line.
NCU_RESULT=i python tools/ncu_lines.py rep kernel_substr [min_pct] [lib.so]   (i = result index in a multi-kernel report; needs the same libsyk.so the profile was taken with)"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, ksub = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "syconn_b200", "libsyk.so")
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
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
which = int(os.environ.get("NCU_RESULT", "0"))  # index of the result inside a multi-kernel report (listed back to back)
heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
hdr = rows[hi]
ia, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:end] if len(r) > ia and r[ia].isdigit()]
cands = [k for k in lines_by_fn if ksub in k and len(lines_by_fn[k]) == len(data)] or [k for k in lines_by_fn if ksub in k]
fn = cands[0]
seq = lines_by_fn[fn]
print("function", fn, "sass", len(seq), "profile rows", len(data))
agg = {}
for r, loc in zip(data, seq):
    a = agg.setdefault(loc, [0, 0])
    a[0] += int(r[ia])
    a[1] += int(r[ismp] or 0)
tot = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
srcs = {}
for loc in sorted(k for k in agg if k):
    if loc[0] not in srcs:
        for root in ("syconn_b200/csrc", "include"):
            pth = os.path.join(os.path.dirname(so), "..", root, loc[0])
            if os.path.exists(pth):
                srcs[loc[0]] = open(pth).read().splitlines()
    a = agg[loc]
    if a[0] > tot * minpct / 100 or a[1] > ts * minpct / 100:
        text = srcs.get(loc[0], [""] * 10000)[loc[1] - 1].strip()[:100] if loc[0] in srcs else ""
        print(f"{loc[0]}:{loc[1]:4d} inst {a[0]/tot*100:5.1f}% smp {a[1]/ts*100:5.1f}%  {text}")
