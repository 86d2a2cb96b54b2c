"""Run one kernel a few times for ncu captures: python tools/prof_target.py props|map|cs [size] [pitch] [order]."""
import sys

import torch

sys.path.insert(0, ".")
from syconn_b200 import device as dev  # noqa: E402

what = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
P = int(sys.argv[3]) if len(sys.argv) > 3 else 32
order = sys.argv[4] if len(sys.argv) > 4 else "C"
pitch = (P, P, max(P // 2, 1)) if P >= 16 else (P, P, P)
reps = 3
if what == "props":
    lab = dev.synth_labels((S, S, S), pitch=pitch, seed=1, order=order)
    tab = dev.IdTable(1 << 19)
    for _ in range(reps):
        tab.clear()
        dev.find_object_properties(tab, lab)
elif what == "map":
    cell = dev.synth_labels((S, S, S), pitch=pitch, seed=1, order=order)
    if order == "F":
        subs = torch.empty((3, S, S, S), dtype=torch.int64, device="cuda").permute(0, 3, 2, 1)
    else:
        subs = torch.empty((3, S, S, S), dtype=torch.int64, device="cuda")
    for c in range(3):
        dev.synth_labels((S, S, S), pitch=(12, 12, 6), seed=1, kind=1 + c, density16=1, out=subs[c])
    ct = dev.IdTable(1 << 18)
    sts = [dev.IdTable(1 << 18) for _ in range(3)]
    pts = [dev.PairTable(1 << 18) for _ in range(3)]
    for _ in range(reps):
        ct.clear()
        for t in sts + pts:
            t.clear()
        dev.map_subcell_extract_props(ct, sts, pts, cell, subs)
    print('digest', ct.count()[0], [t.count()[0] for t in sts], [int(p.export()[:, 2].sum()) for p in pts])
elif what == "cs":
    seg = dev.synth_labels((S + 12, S + 12, S + 6), pitch=pitch, seed=1, dtype=torch.int32, order=order)
    out = dev.detect_cs(seg)
    for _ in range(reps):
        dev.detect_cs(seg, out=out)
torch.cuda.synchronize()
print("done", what)
