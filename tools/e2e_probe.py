"""Development aid: host-buffer calls of one chunk, serial, with blocking (default) or spinning (SYK_SPIN_WAIT=1) stream waits."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from syconn_b200.extraction import _host
from syconn_b200.extraction.find_object_properties import detect_cs
S = 512
cell = dev.synth_labels((S, S, S), pitch=(32, 32, 16), seed=1, order="F")
subs = torch.empty((3, S, S, S), dtype=torch.int64, device="cuda").permute(0, 3, 2, 1)
for c in range(3):
    dev.synth_labels((S, S, S), pitch=(12, 12, 6), seed=1, kind=1 + c, density16=1, out=subs[c])
halo = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=(32, 32, 16), seed=0, dtype=torch.int32, order="F")
def pinned(t):
    h = torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, pin_memory=True); h.copy_(t); return h
hc, hs, hh = pinned(cell), pinned(subs), pinned(halo)
ho = torch.empty((S + 12, S + 12, S + 12), dtype=torch.int64, pin_memory=True)
c_np, s_np, h_np, o_np = hc.numpy().view(np.uint64), hs.numpy().view(np.uint64), hh.numpy().view(np.uint32), ho.numpy().view(np.uint64)
for rep in range(3):
    t0 = time.perf_counter(); detect_cs(h_np, (13, 13, 7), out=o_np); t1 = time.perf_counter()
    _host.find_object_properties_records(o_np); t2 = time.perf_counter()
    _host.map_subcell_records(c_np, s_np); t3 = time.perf_counter()
    print(f"{'spin' if os.environ.get('SYK_SPIN_WAIT') else 'block'}: detect_cs {t1-t0:.3f} s  props {t2-t1:.3f} s  map {t3-t2:.3f} s", flush=True)
