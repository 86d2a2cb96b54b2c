import sys, torch
sys.path.insert(0, ".")
from syconn_b200 import device as dev
from tools.quick_bench import timeit
S = 512
tab = dev.IdTable(1 << 18)
pt = dev.PairTable(1 << 18)
cell = dev.synth_labels((S, S, S), pitch=(32, 32, 16), seed=1)
for name, lab in (("zeros", torch.zeros((S, S, S), dtype=torch.int64, device="cuda")),
                  ("organelle 6%", dev.synth_labels((S, S, S), pitch=(12, 12, 6), seed=1, kind=1, density16=1)),
                  ("cell", cell)):
    def run():
        tab.clear()
        dev.find_object_properties(tab, lab)
    tmin, tmed = timeit(run, n=7, warm=3)
    print(f"props on {name}: {tmin:.3f} ms  {S**3*8/tmin/1e6:.0f} GB/s", flush=True)
    if name != "cell":
        sub = lab[None]
        def run2():
            tab.clear(); pt.clear()
            dev.map_subcell_extract_props(None, [tab], [pt], cell, sub)
        tmin, tmed = timeit(run2, n=7, warm=3)
        print(f"  org-mode map(1 channel, no cell props) on {name}: {tmin:.3f} ms  {S**3*8/tmin/1e6:.0f} GB/s", flush=True)
a = torch.empty((S, S, S), dtype=torch.int64, device="cuda"); b = torch.empty_like(a)
tmin, _ = timeit(lambda: b.copy_(a), n=7, warm=3)
print(f"torch copy: {tmin:.3f} ms {2*S**3*8/tmin/1e6:.0f} GB/s (r+w)")
tmin, _ = timeit(lambda: tab.clear(), n=7, warm=3)
print(f"table clear 2^18: {tmin:.3f} ms")
