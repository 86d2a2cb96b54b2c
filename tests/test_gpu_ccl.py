"""Row f4, first slice: GPU connected components against scipy.ndimage.label (the reference's call,
syconn/extraction/object_extraction_steps.py:350-352) -- bit-exact label volumes -- and chunked labelling + stitching
against the components of the whole volume."""
import numpy as np
import pytest
import scipy.ndimage
import torch

pytestmark = pytest.mark.gpu


def _blobs(shape, seed, density=0.35, smooth=2):
    rng = np.random.default_rng(seed)
    a = rng.random(shape).astype(np.float32)
    a = scipy.ndimage.uniform_filter(a, smooth)
    lo, hi = a.min(), a.max()
    return ((a - lo) / (hi - lo) * 255).astype(np.uint8), int(np.quantile((a - lo) / (hi - lo) * 255, 1 - density))


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("shape", [(40, 37, 45), (64, 64, 64), (5, 130, 33), (1, 1, 70)])
def test_label_components_equals_scipy(shape, order):
    from syconn_b200 import device as dev
    prob, thr = _blobs(shape, 3)
    for t in (thr, 0, 254):
        want, n_want = scipy.ndimage.label(prob > t)
        h = prob if order == "C" else np.asfortranarray(prob)
        d = torch.from_numpy(np.ascontiguousarray(h.transpose(2, 1, 0))).cuda().permute(2, 1, 0) if order == "F" else torch.from_numpy(h).cuda()
        assert tuple(d.shape) == shape
        got, n = dev.label_components(d, t)
        assert n == n_want
        assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), want), (shape, order, t)


def test_label_components_edge_cases():
    from syconn_b200 import device as dev
    z = torch.zeros((9, 8, 7), dtype=torch.uint8, device="cuda")
    lab, n = dev.label_components(z)
    assert n == 0 and int(lab.abs().sum()) == 0
    o = torch.ones((9, 8, 70), dtype=torch.uint8, device="cuda")
    lab, n = dev.label_components(o)
    assert n == 1 and bool((lab == 1).all())
    # 3-D checkerboard: every foreground voxel is its own component under 6-connectivity
    idx = np.indices((12, 11, 40)).sum(axis=0) % 2
    lab, n = dev.label_components(torch.from_numpy(idx.astype(np.uint8)).cuda())
    want, n_want = scipy.ndimage.label(idx)
    assert n == n_want == int(idx.sum()) and np.array_equal(lab.cpu().numpy(), want)
    # a snake that winds through the volume: one component, long union chains; plus uint32 / uint64 inputs and a threshold
    s = np.zeros((20, 20, 64), np.uint32)
    for x in range(0, 20, 2):
        s[x, :, 0 if (x // 2) % 2 else 63] = 7
        s[x, 0 if (x // 2) % 2 else 19, :] = 7
        s[x, :, :][:, ::1][(np.arange(20) % 2 == 0)] = 7
        if x + 1 < 20:
            s[x + 1, 0, 0] = 7
    for dt in (np.uint32, np.uint64):
        want, n_want = scipy.ndimage.label(s > 5)
        lab, n = dev.label_components(torch.from_numpy(s.astype(dt).view(np.int32 if dt == np.uint32 else np.int64)).cuda(), 5)
        assert n == n_want and np.array_equal(lab.cpu().numpy(), want)


def test_label_components_into_foreign_layout():
    """Labels written into a caller's array whose memory order differs from the input's (strided store path)."""
    from syconn_b200 import device as dev
    rng = np.random.default_rng(5)
    a = (rng.random((37, 29, 45)) < 0.45).astype(np.uint8)
    want, n_want = scipy.ndimage.label(a)
    vol = torch.from_numpy(np.asfortranarray(a).T.copy()).cuda().permute(2, 1, 0)  # x fastest in memory
    assert vol.stride(0) == 1
    out = torch.full(a.shape, -1, dtype=torch.int32, device="cuda")                # z fastest
    lab, n = dev.label_components(vol, out=out)
    assert lab is out and n == n_want and np.array_equal(out.cpu().numpy(), want)
    big = torch.full((37, 29, 90), -1, dtype=torch.int32, device="cuda")           # every second element of a larger buffer
    lab, n = dev.label_components(vol, out=big[:, :, ::2])
    assert n == n_want and np.array_equal(big[:, :, ::2].cpu().numpy(), want) and bool((big[:, :, 1::2] == -1).all())


def test_label_components_big_volume_properties():
    """256 x 256 x 320 (21 M voxels): equality with scipy on a volume that needs many CTAs and long-range unions."""
    from syconn_b200 import device as dev
    prob, thr = _blobs((256, 256, 320), 11, density=0.45, smooth=3)
    want, n_want = scipy.ndimage.label(prob > thr)
    got, n = dev.label_components(torch.from_numpy(prob).cuda(), thr)
    assert n == n_want and np.array_equal(got.cpu().numpy(), want)


def _canon(lab):
    """relabel by first occurrence (C order): two label volumes describe the same partition iff their canon forms match"""
    flat = lab.reshape(-1)
    u, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty(len(u), np.int64)
    rank[order] = np.arange(len(u))
    out = rank[inv]
    zero = np.flatnonzero(u == 0)
    if len(zero):   # keep background distinguishable: compare (is_bg, class)
        return out.reshape(lab.shape), (flat == 0).reshape(lab.shape)
    return out.reshape(lab.shape), np.zeros(lab.shape, bool)


@pytest.mark.parametrize("chunk", [(32, 32, 32), (40, 24, 64)])
def test_chunked_labelling_and_stitching_equals_whole_volume(chunk):
    from syconn_b200.chunked import ChunkPlan
    from syconn_b200.extraction.object_extraction_steps import extract_components_chunked
    shape = (96, 80, 128)
    prob, thr = _blobs(shape, 5, density=0.4, smooth=3)
    want, n_want = scipy.ndimage.label(prob > thr)
    vol = torch.from_numpy(prob).cuda()
    ov = (2, 2, 2)
    padded = torch.zeros([s + 2 * o for s, o in zip(shape, ov)], dtype=torch.uint8, device="cuda")
    padded[ov[0]:-ov[0], ov[1]:-ov[1], ov[2]:-ov[2]] = vol

    def load(off, size):
        return padded[off[0] + ov[0]:off[0] + ov[0] + size[0], off[1] + ov[1]:off[1] + ov[1] + size[1],
                      off[2] + ov[2]:off[2] + ov[2] + size[2]]
    plan = ChunkPlan(shape, chunk)
    out, n = extract_components_chunked(load, plan, thr, overlap=ov, stitch_overlap=(1, 1, 1))
    got = np.zeros(shape, np.int64)
    for seq, lab in out.items():
        o, s = plan.offsets[seq], plan.sizes[seq]
        got[o[0]:o[0] + s[0], o[1]:o[1] + s[1], o[2]:o[2] + s[2]] = lab.cpu().numpy()
    assert n == n_want
    cg, bg_g = _canon(got)
    cw, bg_w = _canon(want)
    assert np.array_equal(bg_g, bg_w) and np.array_equal(cg, cw)


def test_label_components_degenerate_shapes_and_views():
    from syconn_b200 import device as dev
    rng = np.random.default_rng(8)
    for shape in ((1, 1, 1), (1, 1, 70), (1, 33, 1), (40, 1, 1), (2, 2, 2), (3, 65, 31), (5, 4, 129)):
        a = (rng.random(shape) < 0.6).astype(np.uint8)
        want, n_want = scipy.ndimage.label(a)
        for fortran in (False, True):
            t = torch.from_numpy(a).cuda()
            if fortran:
                t = t.permute(2, 1, 0).contiguous().permute(2, 1, 0)
            lab, n = dev.label_components(t)
            assert n == n_want and np.array_equal(lab.cpu().numpy(), want), (shape, fortran)
    # a strided view of a larger buffer (every second row, a sub-range of columns), uint16 with a threshold
    big = torch.from_numpy(rng.integers(0, 1000, (12, 40, 60)).astype(np.int16)).cuda()
    view = big[1:11, 2:38:2, 7:55]
    want, n_want = scipy.ndimage.label(view.cpu().numpy() > 400)
    lab, n = dev.label_components(view, 400)
    assert n == n_want and np.array_equal(lab.cpu().numpy(), want)
