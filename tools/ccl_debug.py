import sys, os
import numpy as np, scipy.ndimage, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syconn_b200 import device as dev
sys.path.insert(0, "tests")
from test_gpu_ccl import _blobs, _canon
for shape in [(40, 37, 45), (8, 8, 40), (4, 4, 32), (2, 3, 70)]:
    prob, thr = _blobs(shape, 3)
    for t in (thr, 0):
        want, n_want = scipy.ndimage.label(prob > t)
        got, n = dev.label_components(torch.from_numpy(prob).cuda(), t)
        g = got.cpu().numpy()
        bad = np.argwhere(g != want)
        cg, _ = _canon(g); cw, _ = _canon(want)
        print(shape, t, "n", n, n_want, "mismatch voxels", len(bad), "fg mask equal", np.array_equal(g != 0, want != 0),
              "partition equal", np.array_equal(cg, cw), "max label", g.max(), flush=True)
        if len(bad):
            for b in bad[:5]:
                print("   at", b.tolist(), "got", g[tuple(b)], "want", want[tuple(b)])
