"""Row f4, morphology between threshold and labelling: syk_binary_morph_ops against the oracle restatement of the
reference's apply_morphological_operations (pinned to the reference's outputs in test_oracle_pins.py) and against the
golden vectors directly."""
import numpy as np
import pytest
import torch

from oracle import oracle
from test_oracle_pins import _morph_golden_cases

pytestmark = pytest.mark.gpu

CONFIG_OPS = {  # syconn/handler/config.yml:130-140
    "mi": ["binary_opening", "binary_closing"] + ["binary_erosion"] * 4,
    "sj": ["binary_opening", "binary_closing", "binary_erosion"],
    "er": ["binary_dilation"] * 3 + ["binary_erosion"] * 3,
}


def blobs(shape, density, grow, seed):
    import scipy.ndimage
    rng = np.random.default_rng(seed)
    v = rng.random(shape) < density
    if grow:
        v = scipy.ndimage.binary_dilation(v, iterations=grow)
    return v.astype(np.uint8)


def test_golden_vectors():
    from syconn_b200.proc import image
    for tag, ops, scaling, vin, vout, g in _morph_golden_cases():
        st = image.get_aniso_struct(np.array(scaling))
        assert st.dtype == np.float64 and np.array_equal(st.astype(np.uint8), g["struct_%d_%d_%d" % scaling])
        t = torch.from_numpy(vin.copy()).cuda()
        r = image.apply_morphological_operations(t, list(ops), mop_kwargs=dict(structure=st))
        assert r is t and np.array_equal(t.cpu().numpy(), vout), tag


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("shape", [(40, 37, 45), (70, 33, 31), (9, 8, 100)])
def test_random_volumes_vs_oracle(shape, order):
    from syconn_b200.proc import image
    k = 0
    for name, ops in list(CONFIG_OPS.items()) + [("single", ["binary_closing"]), ("runs", ["binary_dilation", "binary_dilation", "binary_opening"])]:
        for density, grow in ((0.004, 3), (0.02, 2), (0.5, 0)):
            for scaling in ((10, 10, 20), (10, 10, 10)):
                k += 1
                v = blobs(shape, density, grow, seed=k)
                if k % 3 == 0:
                    v[: shape[0] // 3] = 0          # foreground box away from the volume faces on one side
                if k % 4 == 0:
                    v[:, :, -5:] = 0
                st = oracle.get_aniso_struct(np.array(scaling))
                want = oracle.apply_morphological_operations(v.copy(), ops, st)
                t = torch.from_numpy(v).cuda()
                if order == "F":  # x fastest in memory
                    t = t.permute(2, 1, 0).contiguous().permute(2, 1, 0)
                image.apply_morphological_operations(t, ops, mop_kwargs=dict(structure=st))
                assert np.array_equal(t.cpu().numpy(), want), (name, density, grow, scaling, shape, order)


def test_edge_cases_and_errors():
    from syconn_b200 import _lib
    from syconn_b200.proc import image
    st = image.get_aniso_struct((10, 10, 20))
    z = torch.zeros((12, 11, 10), dtype=torch.uint8, device="cuda")
    image.apply_morphological_operations(z, ["binary_closing", "binary_erosion"], dict(structure=st))
    assert int(z.sum()) == 0
    one = torch.zeros((12, 11, 40), dtype=torch.uint8, device="cuda")
    one[5, 5, 33] = 1
    w = one.clone()
    image.apply_morphological_operations(w, ["binary_dilation"], dict(structure=st))
    assert torch.equal(w, one)  # a dilation never leaves the object's bounding box in the reference
    image.apply_morphological_operations(w, ["binary_erosion"], dict(structure=st))
    assert int(w.sum()) == 0
    # default structure (scipy's cross), iterations override, int32 volumes, host arrays
    v = blobs((33, 30, 35), 0.01, 3, seed=7)
    want = oracle.apply_morphological_operations(v.copy(), ["binary_closing"] * 3, None)
    t = torch.from_numpy(v.astype(np.int32)).cuda()
    image.apply_morphological_operations(t, ["binary_closing"], dict(iterations=3))
    assert np.array_equal(t.cpu().numpy(), want)
    h = v.copy()
    r = image.apply_morphological_operations(h, ["binary_closing"] * 3)
    assert r is h and np.array_equal(h, want)
    # the reference's error for anything that is neither erosion nor dilation; multi-label overlays are no device path
    with pytest.raises(NotImplementedError):
        image.apply_morphological_operations(t, ["binary_fill_holes"], dict(structure=st))
    multi = torch.from_numpy(v * 3).cuda()
    with pytest.raises(_lib.SykError):
        image.apply_morphological_operations(multi, ["binary_erosion"], dict(structure=st))
    asym = np.zeros((3, 3, 3))
    asym[1, 1, 1] = asym[0, 1, 1] = 1
    with pytest.raises(_lib.SykError):
        image.apply_morphological_operations(torch.from_numpy(v).cuda(), ["binary_erosion"], dict(structure=asym))


def test_production_sized_chunk_properties():
    """512^3: idempotence of opening / closing, extensivity order, and a 128^3 corner against the oracle."""
    from syconn_b200 import device as dev
    from syconn_b200.proc import image
    prob = (dev.synth_labels((512, 512, 512), pitch=(12, 12, 6), seed=2, kind=7, density16=1, order="F") != 0).to(torch.uint8)
    st = image.get_aniso_struct((10, 10, 20))
    opened = image.apply_morphological_operations(prob.clone(), ["binary_opening"], dict(structure=st))
    again = image.apply_morphological_operations(opened.clone(), ["binary_opening"], dict(structure=st))
    assert torch.equal(opened, again) and bool((opened <= prob).all())
    eroded = image.apply_morphological_operations(prob.clone(), ["binary_erosion"], dict(structure=st))
    assert bool((eroded <= opened).all()) and 0 < int(eroded.sum()) < int(opened.sum())
    corner = prob[:128, :128, :128].clone()
    want = oracle.apply_morphological_operations(corner.cpu().numpy().copy(), CONFIG_OPS["sj"], st)
    image.apply_morphological_operations(corner, CONFIG_OPS["sj"], dict(structure=st))
    assert np.array_equal(corner.cpu().numpy(), want)


@pytest.mark.parametrize("kind", ["mi", "sj", "er"])
def test_segmentation_chunk_with_morphology_and_seeds(kind):
    """object_segmentation_chunk with an erosion-free op list == scipy.ndimage.label of the oracle's mask; the watershed
    branch's mask / markers (min_seed_vx clean-up included) == the oracle restatement."""
    import scipy.ndimage
    from syconn_b200.extraction import object_extraction_steps as oes
    st = oracle.get_aniso_struct(np.array((10, 10, 20)))
    rng = np.random.default_rng(11)
    prob = (scipy.ndimage.binary_dilation(rng.random((72, 64, 50)) < 0.004, iterations=4) * 200).astype(np.uint8)
    prob[rng.random(prob.shape) < 0.02] = 255  # speckle that the opening removes
    ops = CONFIG_OPS[kind]
    first = ops.index("binary_erosion")
    tmp = (prob > 128).astype(np.uint8)
    t = torch.from_numpy(prob).cuda()
    # erosion-free prefix: morphology + labelling
    want, n_want = scipy.ndimage.label(oracle.apply_morphological_operations(tmp.copy(), ops[:first], st))
    lab, n = oes.object_segmentation_chunk(t, 128, ops[:first], st)
    assert n == n_want and np.array_equal(lab.cpu().numpy(), want)
    with pytest.raises(NotImplementedError):
        oes.object_segmentation_chunk(t, 128, ops, st)
    # watershed seeds
    for min_size in (1, 10, 50):
        mask_w, markers_w = oracle.watershed_seeds(tmp, ops, st, min_size)
        mask, markers, n = oes.watershed_seeds_chunk(t, 128, ops, st, min_seed_vx=min_size)
        assert np.array_equal(mask.cpu().numpy(), mask_w)
        assert np.array_equal(markers.cpu().numpy().astype(np.uint32), markers_w), (kind, min_size)
        assert n == int(markers_w.max())
    assert torch.equal(t.cpu(), torch.from_numpy(prob))  # the probability map itself is untouched


def test_degenerate_shapes_and_strided_views():
    """Extents of 1, rows shorter / longer than a 32-voxel word, padding larger than the volume, non-dense views."""
    from syconn_b200.proc import image
    st = oracle.get_aniso_struct(np.array((10, 10, 20)))
    rng = np.random.default_rng(21)
    k = 0
    for shape in ((1, 1, 1), (1, 5, 40), (7, 1, 33), (3, 3, 3), (2, 70, 5), (65, 4, 6)):
        for ops in (["binary_closing"] * 6, ["binary_dilation", "binary_erosion"], ["binary_opening", "binary_closing"], ["binary_erosion"]):
            k += 1
            v = (rng.random(shape) < 0.6).astype(np.uint8)
            want = oracle.apply_morphological_operations(v.copy(), ops, st)
            for fortran in (False, True):
                t = torch.from_numpy(v).cuda()
                if fortran:
                    t = t.permute(2, 1, 0).contiguous().permute(2, 1, 0)
                image.apply_morphological_operations(t, ops, dict(structure=st))
                assert np.array_equal(t.cpu().numpy(), want), (shape, ops, fortran)
    # a view into a larger buffer: every second plane and a sub-range of rows; the rest of the buffer must stay untouched
    big = torch.full((20, 30, 50), 7, dtype=torch.uint8, device="cuda")
    view = big[2:18:2, 5:25, 3:47]
    v = (rng.random(tuple(view.shape)) < 0.55).astype(np.uint8)
    view.copy_(torch.from_numpy(v).cuda())
    ops = ["binary_opening", "binary_closing", "binary_erosion"]
    want = oracle.apply_morphological_operations(v.copy(), ops, st)
    image.apply_morphological_operations(view, ops, dict(structure=st))
    assert np.array_equal(view.cpu().numpy(), want)
    mask = torch.ones_like(big, dtype=torch.bool)
    mask[2:18:2, 5:25, 3:47] = False
    assert bool((big[mask] == 7).all())
