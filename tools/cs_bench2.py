import sys, os, torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit
for S in (512,):
    for pitch in ((32, 32, 16),):
        for order in ("F",):
            seg = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=pitch, seed=0, dtype=torch.int32, order=order)
            out = dev.detect_cs(seg)
            os.environ["SYK_CS_DEBUG"] = "1"
            dev.detect_cs(seg, out=out)
            del os.environ["SYK_CS_DEBUG"]
            ts = [timeit(lambda: dev.detect_cs(seg, out=out), n=5, warm=2) for _ in range(3)]
            print(os.environ.get("SYK_LIB_NAME"), f"detect_cs chunk-like {S} pitch {pitch} {order}: min {min(t[0] for t in ts):.3f} ms med {sorted(t[1] for t in ts)[1]:.3f} ms", flush=True)
