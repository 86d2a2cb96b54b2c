"""One line: map_subcell_extract_props (cell + 3 organelle channels, 512^3 uint64, x fastest) for the library named by
SYK_LIB_NAME.  Development aid."""
import os
import sys
import torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit
S = 512
cell = dev.synth_labels((S, S, S), pitch=(32, 32, 16), seed=1, order="F")
subs = torch.empty((3, S, S, S), dtype=torch.int64, device="cuda").permute(0, 3, 2, 1)
for c in range(3):
    dev.synth_labels((S, S, S), pitch=(12, 12, 6), seed=1, kind=1 + c, density16=1, out=subs[c])
sts = [dev.IdTable(1 << 18) for _ in range(3)]
pts = [dev.PairTable(1 << 18) for _ in range(3)]


def run():
    for t in sts + pts:
        t.clear()
    dev.map_subcell_extract_props(None, sts, pts, cell, subs)


tmin, tmed = timeit(run, n=7, warm=3)
print(f"{os.environ.get('SYK_LIB_NAME', 'libsyk.so')}: 3 organelle scans min {tmin:.3f} ms med {tmed:.3f} ms  "
      f"pairs {[p.export().shape[0] for p in pts]} ids {[t.count()[0] for t in sts]}", flush=True)
