"""Per-kernel SASS instruction-class counts of libsyk.so (evidence of the Blackwell-native paths: UTMALDG = TMA tensor loads,
SYNCS = mbarrier, ATOMG.E.CAS.128 = 128-bit pair-key claim, MATCH.ANY = warp grouping, REDG = fire-and-forget atomics).
  python tools/sass_grep.py [lib.so] > profiles/r2_sass_classes.txt"""
import os
import re
import subprocess
import sys
import tempfile

so = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "syconn_b200/libsyk.so")
pats = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "ATOMG.E.CAS.128", "ATOMG", "ATOMS", "REDG", "RED.E", "MATCH.ANY", "REDUX", "VOTE", "SHFL",
        "BAR.SYNC", "LDS", "STS", "LDG", "STG", "PRMT", "LOP3", "IADD3", "IMAD", "POPC", "FLO", "HMMA", "UTCHMMA"]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
print(f"# SASS instruction classes per kernel of {os.path.basename(so)} (cuobjdump -xelf all; nvdisasm -c); counts are static instructions")
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    fn, counts, total = None, {}, 0
    def flush():
        if fn and total:
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)[:110]
            hits = "  ".join(f"{k}={v}" for k, v in counts.items() if v)
            print(f"{name}\n    {total} instructions: {hits}")
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            flush()
            fn, counts, total = m.group(1), {p: 0 for p in pats}, 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if fn and m:
            total += 1
            op = m.group(1)
            for p in pats:
                if op.startswith(p) or (p in ("ATOMG.E.CAS.128", "MATCH.ANY") and p in op):
                    counts[p] += 1
    flush()
