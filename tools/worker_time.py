"""Stage times of the contact-site worker body (rows f1 / f2) on one production chunk: wall clock (host round trips count)
and, under ncu, the launch list.  python tools/worker_time.py [reps]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from syconn_b200 import device as dev
from syconn_b200.chunked import ExtractionPipeline

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ST = (13, 13, 7)
S = 512
halo = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=(32, 32, 16), seed=0, dtype=torch.int32, order="F")
oshape = (S + 12, S + 12, S + 12)
masks = [(dev.synth_labels(oshape, (506, -6, 1018), (12, 12, 6), 4, 0, kind, 1 if kind == 3 else 4, order="F") != 0).to(torch.uint8)
         for kind in (3, 4, 5)]
pipe = ExtractionPipeline(0, ST, chunk_table_capacity=1 << 18, log_capacity=1 << 20, pair_log_capacity=1 << 10, with_syn=True, cs_dilation=2)
overlap = 6


def stages():
    t = {}

    def mark(name, t0):
        torch.cuda.synchronize()
        t[name] = t.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cs = dev.detect_cs(halo, ST, out=pipe._cs_buffer(halo))
    pipe.cs_out = cs
    t0 = mark("detect_cs", t0)
    pipe.t_cs.clear()
    dev.find_object_properties(pipe.t_cs, cs)
    t0 = mark("props_full", t0)
    rec = pipe.t_cs.export(dev.geoms([[0, 0, 0]], [list(cs.shape)]), sort=True)
    t0 = mark("export_sorted_device", t0)
    dev.close_contacts_records(cs, rec, overlap, 2)
    t0 = mark("close_contacts", t0)
    crop = tuple(slice(overlap, n - overlap) for n in cs.shape)
    cs_c = cs[crop]
    sj, asym, sym = (m[crop] for m in masks)
    pipe.t_cs.clear()
    pipe.t_syn.clear()
    vox = dev.extract_cs_syntype(pipe.t_cs, cs_c, sj, asym, sym, origin=(0, 0, 0), chunk_seq=0, syn_table=pipe.t_syn)
    t0 = mark("extract_cs_syntype+syn", t0)
    pipe._append(pipe.t_cs, 1, pipe.logs["cs"], (0, 0, 0), cs_c.shape)
    pipe._append(pipe.t_syn, pipe.kinds.index("syn"), pipe.logs["syn"], (0, 0, 0), cs_c.shape)
    t0 = mark("append_cs_syn", t0)
    t["n_cs"] = int(rec.shape[0])
    t["n_synvox"] = int(vox.shape[0])
    return t


pipe.reset()
stages()
acc = []
for _ in range(reps):
    pipe.reset()
    acc.append(stages())
keys = [k for k in acc[0] if not k.startswith("n_")]
tot = 0.0
for k in keys:
    v = min(a[k] for a in acc)
    tot += v
    print(f"{k:22s} {v:8.3f} ms")
print(f"{'total':22s} {tot:8.3f} ms   objects {acc[0]['n_cs']}  syn voxels {acc[0]['n_synvox']}")
