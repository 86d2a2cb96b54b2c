"""BASELINE config 5: high-cardinality stress (1e7 distinct ids) and stencil-size sweep.  Prints one line per case."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit

# ---- 1e7+ distinct ids: every voxel of a 224^3 volume its own id (11.2M ids), then 2x2x2-voxel objects ----
S = 224
ar = torch.arange(S ** 3, dtype=torch.int64, device="cuda")
for name, lab in (("unique id per voxel (11.2M ids)", (ar * 2654435761 + 1).reshape(S, S, S)),
                  ("2x2x2 voxel objects (1.4M ids)", dev.synth_labels((S, S, S), pitch=(2, 2, 2), warp_amp=0, seed=3))):
    tab = dev.IdTable(1 << 25)
    def run():
        tab.clear()
        dev.find_object_properties(tab, lab)
    tmin, _ = timeit(run, n=3, warm=1)
    n, ovf = tab.count()
    ids = torch.unique(lab)
    ok = n == int((ids != 0).sum()) and not ovf
    print(f"props {name}: {tmin:.2f} ms  {S**3/tmin/1e6:.2f} GVox/s  ids={n} parity_count={ok}", flush=True)
    tab.close()
# ---- stencil sweep on a 256^3 (+halo) uint32 volume, pitch 32x32x16 ----
for st in ((3, 3, 3), (5, 5, 3), (7, 7, 3), (9, 9, 5), (13, 13, 7), (15, 15, 9), (17, 17, 9)):
    shape = tuple(256 + s - 1 for s in st)
    seg = dev.synth_labels(shape, pitch=(32, 32, 16), seed=1, dtype=torch.int32, order="F")
    out = dev.detect_cs(seg, st)
    tmin, _ = timeit(lambda: dev.detect_cs(seg, st, out=out), n=3, warm=1)
    print(f"detect_cs stencil {st}: {tmin:.3f} ms  {256**3/tmin/1e6:.2f} GVox/s  contacts={float((out != 0).float().mean()):.3f}", flush=True)
# ---- near-random labels (worst case for the window histogram): 96^3, stencil 13x13x7 -> generic kernel fallback ----
rnd = torch.randint(1, 2 ** 31 - 1, (96 + 12, 96 + 12, 96 + 6), dtype=torch.int32, device="cuda")
out = dev.detect_cs(rnd, (13, 13, 7))
tmin, _ = timeit(lambda: dev.detect_cs(rnd, (13, 13, 7), out=out), n=2, warm=1)
print(f"detect_cs random labels 96^3: {tmin:.1f} ms  {96**3/tmin/1e6:.4f} GVox/s", flush=True)
