"""Development aid: the marching tier-1 kernel (syk_cs_march.cuh) against the round-1 kernels (default; the marching kernel is opt-in with SYK_CS_MARCH=1) and the
C oracle, on small volumes of both memory orders and on the production chunk; prints the first mismatches and timings.
  python tools/cs_check.py [--big] [--time]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syconn_b200 import device as dev  # noqa: E402
from tools.quick_bench import timeit  # noqa: E402


def set_march(on):
    if on:
        os.environ["SYK_CS_MARCH"] = "1"
    else:
        os.environ.pop("SYK_CS_MARCH", None)


def run(seg, st, march):
    set_march(march)
    out = dev.detect_cs(seg, st)
    torch.cuda.synchronize()
    return out


def report(tag, a, b):
    same = bool(torch.equal(a, b))
    print(f"{tag}: {'OK' if same else 'MISMATCH'}  nonzero {int(torch.count_nonzero(b))}", flush=True)
    if not same:
        d = torch.nonzero(a != b)
        print(f"   {d.shape[0]} voxels differ; first: {d[:8].tolist()}")
        for idx in d[:8].tolist():
            print(f"   at {idx}: got {int(a[tuple(idx)]):#x} want {int(b[tuple(idx)]):#x}")
        for ax in range(3):
            vals, cnt = torch.unique(d[:, ax], return_counts=True)
            print(f"   axis {ax}: {len(vals)} distinct coordinates, min {int(vals.min())} max {int(vals.max())}")
    return same


def main():
    ok = True
    cases = [((80, 72, 90), (16, 16, 8), "F", (13, 13, 7)), ((80, 72, 90), (16, 16, 8), "C", (13, 13, 7)),
             ((150, 100, 140), (32, 32, 16), "F", (13, 13, 7)), ((150, 100, 140), (24, 20, 12), "C", (13, 13, 7)),
             ((100, 100, 100), (7, 7, 5), "F", (13, 13, 7)), ((64, 200, 96), (40, 40, 40), "F", (13, 13, 7))]
    try:
        from oracle import oracle
    except Exception:
        oracle = None
    for shape, pitch, order, st in cases:
        seg = dev.synth_labels(shape, origin=(5, -3, 11), pitch=pitch, seed=1, dtype=torch.int32, order=order)
        new, old = run(seg, st, True), run(seg, st, False)
        ok &= report(f"{shape} pitch {pitch} order {order}: march vs round-1 kernels", new, old)
        if oracle is not None:
            want = torch.from_numpy(oracle.detect_cs(seg.cpu().numpy().view(np.uint32), st).view(np.int64)).cuda()
            ok &= report(f"{shape} pitch {pitch} order {order}: march vs oracle", new, want)
    if "--big" in sys.argv:
        S = 512
        for pitch in ((32, 32, 16), (16, 16, 8)):
            seg = dev.synth_labels((S + 24, S + 24, S + 18), origin=(500, -12, 1015), pitch=pitch, seed=0, dtype=torch.int32, order="F")
            os.environ["SYK_CS_DEBUG"] = "1"
            new = run(seg, (13, 13, 7), True)
            os.environ.pop("SYK_CS_DEBUG")
            old = run(seg, (13, 13, 7), False)
            ok &= report(f"production chunk pitch {pitch}: march vs round-1 kernels", new, old)
            if "--time" in sys.argv:
                for march in (True, False):
                    set_march(march)
                    tmin, tmed = timeit(lambda: dev.detect_cs(seg, (13, 13, 7), out=new), n=7, warm=3)
                    print(f"   pitch {pitch} {'march' if march else 'round-1'}: min {tmin:.3f} ms med {tmed:.3f} ms", flush=True)
            del new, old, seg
    print("CS_CHECK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
