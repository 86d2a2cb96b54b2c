"""One line: detect_cs time on the production chunk (536x536x530 uint32, x fastest, stencil 13x13x7) for the library named by
SYK_LIB_NAME (variants built with SYK_NVCC_EXTRA).  Development aid."""
import os
import sys
import torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit
S = 512
PITCH = tuple(int(x) for x in os.environ.get("SYK_PITCH", "32,32,16").split(","))
seg = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=PITCH, seed=0, dtype=torch.int32, order="F")
out = dev.detect_cs(seg)
tmin, tmed = timeit(lambda: dev.detect_cs(seg, out=out), n=7, warm=3)
print(f"{os.environ.get('SYK_LIB_NAME', 'libsyk.so')} pitch {PITCH}: detect_cs min {tmin:.3f} ms med {tmed:.3f} ms  checksum {int(out.sum().item()) & 0xFFFFFFFF:08x}", flush=True)
