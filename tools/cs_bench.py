import sys, os, torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit
for S in (256, 512):
    for pitch in ((32, 32, 16), (64, 64, 32), (16, 16, 8)):
        for order in ("C", "F"):
            seg = dev.synth_labels((S + 12, S + 12, S + 6), pitch=pitch, seed=1, dtype=torch.int32, order=order)
            out = dev.detect_cs(seg)
            os.environ["SYK_CS_DEBUG"] = "1"
            dev.detect_cs(seg, out=out)
            del os.environ["SYK_CS_DEBUG"]
            tmin, tmed = timeit(lambda: dev.detect_cs(seg, out=out), n=5, warm=2)
            print(f"detect_cs {S}^3 pitch {pitch} {order}: min {tmin:.3f} ms med {tmed:.3f} ms  {S**3/tmin/1e6:.2f} GVox/s", flush=True)
