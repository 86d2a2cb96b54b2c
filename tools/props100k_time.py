"""find_object_properties on a 512^3 uint64 cube with ~1e5 ids (BASELINE config 2).  python tools/props100k_time.py [pitch]"""
import sys
import torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit
p = int(sys.argv[1]) if len(sys.argv) > 1 else 11
lab = dev.synth_labels((512, 512, 512), pitch=(p, p, p), seed=5, order="F")
tab = dev.IdTable(1 << 19)


def run():
    tab.clear()
    dev.find_object_properties(tab, lab)


tmin, tmed = timeit(run, n=5, warm=2)
n, ovf = tab.count()
print(f"props 512^3 pitch {p}: min {tmin:.3f} ms med {tmed:.3f} ms  ids {n}  {512 ** 3 * 8 / tmin / 1e6:.0f} GB/s", flush=True)
