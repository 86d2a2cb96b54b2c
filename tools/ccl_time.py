"""Time label_components (row f4) on the bench's 512^3 uint8 stress volume; under ncu this gives the per-kernel launch list.
python tools/ccl_time.py [reps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from syconn_b200 import device as dev

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
prob = (dev.synth_labels((512, 512, 512), pitch=(12, 12, 6), seed=2, kind=7, density16=1, order="F") != 0).to(torch.uint8) * 200
lab, n = dev.label_components(prob, 128)
ts = []
for _ in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.label_components(prob, 128, out=lab)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"label_components 512^3: min {min(ts):.3f} ms med {sorted(ts)[len(ts) // 2]:.3f} ms  components {int(n)}  checksum {int(lab.long().sum())}")
