"""One line: find_object_properties on a 512^3 uint64 cell chunk and on a cropped, aligned 512^3 contact volume (x fastest) for the
library named by SYK_LIB_NAME.  Development aid."""
import os
import sys
import torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from syconn_b200.chunked import ExtractionPipeline
from tools.quick_bench import timeit
S = 512
cell = dev.synth_labels((S, S, S), pitch=(32, 32, 16), seed=1, order="F")
halo = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=(32, 32, 16), seed=0, dtype=torch.int32, order="F")
class P: pass
p = P(); p.stencil = (13, 13, 7); p.cs_out = None
cs = dev.detect_cs(halo, (13, 13, 7), out=ExtractionPipeline._cs_buffer(p, halo))
crop = cs[6:-6, 6:-6, 6:-6]
tab = dev.IdTable(1 << 19)
def run(v):
    tab.clear()
    dev.find_object_properties(tab, v)
t1, _ = timeit(lambda: run(cell), n=7, warm=3)
t2, _ = timeit(lambda: run(crop), n=7, warm=3)
print(f"{os.environ.get('SYK_LIB_NAME', 'libsyk.so')}: props cell {t1:.3f} ms  contacts {t2:.3f} ms", flush=True)
