"""f2 at production size: detect_cs on one 536x536x530 haloed chunk (x fastest), find_object_properties of the contacts,
then syk_close_contacts (n_closings 6, cs_dilation 2).  Prints timings; `--check` compares a 160^3 corner with the oracle."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit

edge = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 512
shape = (edge + 24, edge + 24, edge + 18)  # chunk + overlap 6 + stencil offset on every side (cs_extraction_steps.py:383-387)
seg = dev.synth_labels(shape, pitch=(32, 32, 16), seed=1, dtype=torch.int32, order="F")
cs0 = dev.detect_cs(seg, (13, 13, 7))
tab = dev.IdTable(1 << 21)
dev.find_object_properties(tab, cs0)
recs = dev.records_numpy(tab.export(dev.geoms([[0, 0, 0]], [list(cs0.shape)])))
recs = recs[np.argsort(recs["id"])]
ids = recs["id"].copy()
bbox = np.stack([recs["bb_min"], recs["bb_max"]], axis=1).astype(np.int32)
ext = bbox[:, 1] - bbox[:, 0] + 12
print(f"contacts {tuple(cs0.shape)}: {len(ids)} ids, contact voxels {float((cs0 != 0).float().mean()):.3f}, "
      f"padded box volume / volume = {ext.prod(axis=1).sum() / cs0.numel():.2f}, max box {ext.max(axis=0)}", flush=True)
cs = cs0.clone()


def run():
    cs.copy_(cs0)
    dev.close_contacts(cs, ids, bbox, 6, 2)


t_copy, _ = timeit(lambda: cs.copy_(cs0), n=3, warm=1)
tmin, _ = timeit(run, n=3, warm=1)
print(f"close_contacts: {tmin - t_copy:.2f} ms (+ {t_copy:.2f} ms copy)  {cs0.numel() / (tmin - t_copy) / 1e6:.2f} GVox/s; "
      f"filled voxels {int(((cs != 0) & (cs0 == 0)).sum())}", flush=True)
if "--check" in sys.argv:
    from oracle import oracle
    n = 120
    sub0 = cs0[:n, :n, :n].contiguous().cpu().numpy().view(np.uint64)
    bb = oracle.find_object_properties(sub0)[1]
    t0 = time.time()
    want = oracle.close_contact_sites(sub0.copy(), bb, 6, 2)
    t_cpu = time.time() - t0
    subg = torch.from_numpy(sub0.view(np.int64).copy()).cuda()
    k = np.fromiter(bb.keys(), np.uint64, len(bb))
    b = np.array(list(bb.values()), np.int32)
    dev.close_contacts(subg, k, b, 6, 2)
    print("parity on a %d^3 corner: %s (%d ids, CPU restatement %.1f s)" % (n, np.array_equal(subg.cpu().numpy().view(np.uint64), want),
                                                                            len(bb), t_cpu), flush=True)
