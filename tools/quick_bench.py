"""Kernel-level timings on one GPU (CUDA events on torch's current stream).  Development aid, not the bench."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from syconn_b200 import device as dev  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    res = {}
    for name, pitch in (("pitch11", (11, 11, 11)), ("pitch32", (32, 32, 16))):
        for order in ("C", "F"):
            lab = dev.synth_labels((S, S, S), pitch=pitch, seed=1, order=order)
            tab = dev.IdTable(1 << 19)

            def run():
                tab.clear()
                dev.find_object_properties(tab, lab)
            tmin, tmed = timeit(run)
            n, ovf = tab.count()
            res[f"props_{name}_{order}"] = dict(ms=tmin, ms_med=tmed, gvox_s=S ** 3 / tmin / 1e6, gbs=S ** 3 * 8 / tmin / 1e6, ids=n, ovf=ovf)
            print(f"props {name} {order}: {tmin:.3f} ms  {S**3*8/tmin/1e6:.0f} GB/s  ids={n}", flush=True)
            del lab
    # mapping: cell + 3 organelles
    cell = dev.synth_labels((S, S, S), pitch=(32, 32, 16), seed=1)
    subs = torch.stack([dev.synth_labels((S, S, S), pitch=(12, 12, 6), seed=1, kind=1 + c, density16=1) for c in range(3)])
    ct = dev.IdTable(1 << 18)
    sts = [dev.IdTable(1 << 18) for _ in range(3)]
    pts = [dev.PairTable(1 << 18) for _ in range(3)]

    def run_map():
        ct.clear()
        for t in sts + pts:
            t.clear()
        dev.map_subcell_extract_props(ct, sts, pts, cell, subs)
    tmin, tmed = timeit(run_map)
    res["map3"] = dict(ms=tmin, gvox_s=S ** 3 / tmin / 1e6, gbs=S ** 3 * 32 / tmin / 1e6, fg=float((subs != 0).float().mean()))
    print(f"map C=3: {tmin:.3f} ms  {S**3*32/tmin/1e6:.0f} GB/s  fg={res['map3']['fg']:.3f} pairs={[p.export().shape[0] for p in pts]}", flush=True)
    del subs
    # contact sites
    Sc = min(S, 256)
    for name, pitch in (("pitch32", (32, 32, 16)), ("pitch64", (64, 64, 32))):
        for order in ("C", "F"):
            seg = dev.synth_labels((Sc + 12, Sc + 12, Sc + 6), pitch=pitch, seed=1, dtype=torch.int32, order=order)
            out = dev.detect_cs(seg)
            tmin, tmed = timeit(lambda: dev.detect_cs(seg, out=out), n=3, warm=1)
            frac = float((out != 0).float().mean())
            bfrac = float(dev.detect_seg_boundaries(seg).float().mean())
            res[f"cs_{name}_{order}"] = dict(ms=tmin, gvox_s=Sc ** 3 / tmin / 1e6, contact_frac=frac, boundary_frac=bfrac)
            print(f"detect_cs {name} {order} {Sc}^3: {tmin:.3f} ms  {Sc**3/tmin/1e6:.2f} GVox/s  boundary={bfrac:.3f} contacts={frac:.3f}", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
