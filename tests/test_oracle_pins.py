"""Pins the CPU oracle (oracle/syk_oracle.c) against (a) the reference's known-answer tests, re-stated from
/root/reference/tests/test_segmentation_analysis.py, (b) golden vectors produced by the reference's own compiled
code (tests/golden/make_golden.py) and (c), when oracle/_ref was built, the reference modules live."""
import numpy as np
import pytest

from oracle import oracle, ref
from syconn_b200.synth import synth_labels
from helpers import assert_props_equal, assert_props_equal_arrays, check_dict_types, map_to_rows

STENCIL = np.array([13, 13, 7], dtype=np.int32)  # tests/test_segmentation_analysis.py:16 (config default)


def gen_sample_seg(distance_between_cube, stencil, cube_size):
    """Analytic expectation of reference tests/test_segmentation_analysis.py:100-123 (_gen_sample_seg)."""
    offset = stencil // 2
    a = np.amax(offset + 1)
    edge_s = np.amax(stencil + distance_between_cube + cube_size)
    sample = np.zeros((edge_s, edge_s, edge_s), dtype=np.uint32)
    c, d = cube_size, distance_between_cube
    sample[a:a + c, a:a + c, a:a + c] = 4
    sample[a + d[0]:a + d[0] + c, a + d[1]:a + d[1] + c, a + d[2]:a + d[2] + c] = 5
    o_o = np.maximum(0, d - offset)
    oshape = np.array(sample.shape) + 1 - stencil
    o = offset
    m = np.zeros(tuple(oshape), dtype=np.uint32)
    m[a - o[0] + o_o[0]:a + c - o[0], a - o[1] + o_o[1]:a + c - o[1], a - o[2] + o_o[2]:a + c - o[2]] = 1
    m[a + d[0] - o[0]:a + d[0] + c - o[0] - o_o[0], a + d[1] - o[1]:a + d[1] + c - o[1] - o_o[1],
      a + d[2] - o[2]:a + d[2] + c - o[2] - o_o[2]] = 1
    m[a - o[0] + 1:a + c - o[0] - 1, a - o[1] + 1:a + c - o[1] - 1, a - o[2] + 1:a + c - o[2] - 1] = 0
    m[a + d[0] - o[0] + 1:a + d[0] + c - o[0] - 1, a + d[1] - o[1] + 1:a + d[1] + c - o[1] - 1,
      a + d[2] - o[2] + 1:a + d[2] + c - o[2] - 1] = 0
    return sample, 4 * m, 5 * m


def check_known_find_object_properties(fn):
    """reference tests/test_segmentation_analysis.py:19-52."""
    sample = np.array([[[0, 1], [1, 1]], [[5, 2], [2, 1]]], np.uint64)
    rc, bb, cnt = fn(sample)
    el, count = np.unique(sample, return_counts=True)
    assert 0 not in rc and 0 not in bb and 0 not in cnt
    count, el = count[el != 0], el[el != 0]
    for e, n in zip(el, count):
        assert cnt[int(e)] == n
        ll = rc[int(e)]
        assert sample[ll[0], ll[1], ll[2]] == e
        mask = np.transpose(np.where(sample == e))
        assert np.all(mask.min(axis=0) == bb[int(e)][0])
        assert np.all(mask.max(axis=0) + 1 == bb[int(e)][1])
    # exact first-voxel semantics
    assert rc[1] == [0, 0, 1] and rc[5] == [1, 0, 0] and rc[2] == [1, 0, 1]


def check_known_detect_cs(fn):
    """reference tests/test_segmentation_analysis.py:55-76,126-129."""
    for d in ([0, 6, 0], [6, 0, 0], [0, 0, 6]):
        sample, lo, hi = gen_sample_seg(np.array(d), STENCIL, 5)
        out = np.asarray(fn(sample))
        assert out.dtype == np.uint64
        assert np.array_equal(hi.astype(np.uint32), out.astype(np.uint32))
        assert np.array_equal(lo.astype(np.uint32), (out // 2 ** 32).astype(np.uint32))


def check_known_boundary(fn):
    """reference tests/test_segmentation_analysis.py:162-169."""
    b = np.asarray(fn(np.arange(1000).reshape((10, 10, 10)))).flatten()
    assert b[0] == 0 and np.all(b[1:])
    assert not np.any(fn(np.zeros((10, 10, 10))))


def test_known_answers_oracle():
    check_known_find_object_properties(oracle.find_object_properties)
    check_known_detect_cs(oracle.detect_cs)
    check_known_boundary(oracle.detect_seg_boundaries)


def test_oracle_vs_golden_props(golden):
    g = golden
    got = oracle.find_object_properties(g["fop_in"])
    check_dict_types(got)
    assert_props_equal_arrays(got, g["fop_ids"], g["fop_sizes"], g["fop_bbox"], g["fop_rep"], "fop")
    got = oracle.find_object_properties(g["fop_in"].transpose(2, 1, 0))
    assert_props_equal_arrays(got, g["fopT_ids"], g["fopT_sizes"], g["fopT_bbox"], g["fopT_rep"], "fop strided")
    got = oracle.find_object_properties(g["fop32_in"])
    assert_props_equal_arrays(got, g["fop32_ids"], g["fop32_sizes"], g["fop32_bbox"], g["fop32_rep"], "fop u32")


def test_oracle_vs_golden_map(golden):
    g = golden
    cp, sp, md = oracle.map_subcell_extract_props(g["map_cell"], g["map_subs"])
    assert_props_equal_arrays(cp, g["map_cell_ids"], g["map_cell_sizes"], g["map_cell_bbox"], g["map_cell_rep"], "cell")
    mc = oracle.map_subcell_C(g["map_cell"], g["map_subs"])
    for c in range(3):
        assert_props_equal_arrays((sp[0][c], sp[1][c], sp[2][c]), g[f"map_sub{c}_ids"], g[f"map_sub{c}_sizes"],
                                  g[f"map_sub{c}_bbox"], g[f"map_sub{c}_rep"], f"sub{c}")
        assert np.array_equal(map_to_rows(md[c]), g[f"map_pairs{c}"])
        assert np.array_equal(map_to_rows(mc[c]), g[f"mapC_pairs{c}"])


def test_oracle_vs_golden_cs(golden):
    g = golden
    assert np.array_equal(oracle.detect_seg_boundaries(g["cs_in"]), g["cs_bdry"])
    for st in ((13, 13, 7), (7, 7, 3), (5, 5, 3), (3, 3, 3), (1, 1, 1), (3, 5, 7)):
        assert np.array_equal(oracle.detect_cs(g["cs_in"], st), g["cs_out_%d_%d_%d" % st]), st
    for st in ((5, 5, 3), (3, 3, 3)):
        assert np.array_equal(oracle.detect_cs(g["tie_in"], st), g["tie_out_%d_%d_%d" % st]), st
    assert np.array_equal(oracle.detect_cs(g["many_in"], (5, 5, 3)), g["many_out_5_5_3"])
    edges = np.ones(g["tie_in"].shape, np.uint32)
    assert np.array_equal(oracle.process_block_nonzero(edges, g["tie_in"], (7, 7, 3)), g["pbn_forced_7_7_3"])


def test_oracle_edge_cases():
    assert oracle.find_object_properties(np.zeros((3, 4, 5), np.uint64)) == ({}, {}, {})
    r = oracle.map_subcell_extract_props(np.zeros((3, 3, 3), np.uint64), np.zeros((2, 3, 3, 3), np.uint64))
    assert r == ([{}, {}, {}], [[{}, {}], [{}, {}], [{}, {}]], [{}, {}])
    # organelle over background: present in props, absent from mapping
    cell = np.zeros((2, 2, 2), np.uint64)
    sub = np.zeros((1, 2, 2, 2), np.uint64)
    sub[0, 1, 1, 0] = 9
    cp, sp, md = oracle.map_subcell_extract_props(cell, sub)
    assert sp[2][0] == {9: 1} and md == [{}]
    with pytest.raises(AssertionError):
        oracle.process_block_nonzero(np.ones((5, 5, 5), np.uint32), np.ones((5, 5, 5), np.uint32), (4, 3, 3))
    with pytest.raises(ValueError):
        oracle.process_block_nonzero(np.ones((5, 5, 5), np.uint64), np.ones((5, 5, 5), np.uint64), (3, 3, 3))
    with pytest.raises(AssertionError):
        oracle.map_subcell_extract_props(np.zeros((3, 3, 3), np.uint64), np.zeros((1, 3, 3, 4), np.uint64))
    # tie-break: smallest id wins (block_processing_C.pyx:34-39)
    w = np.ones((3, 3, 3), np.uint32)
    w[0, 0, 0], w[2, 2, 2] = 3, 2
    assert int(oracle.process_block_nonzero(np.ones_like(w), w, (3, 3, 3))[0, 0, 0]) == (1 << 32) + 2
    w[0, 0, 1] = 3
    assert int(oracle.process_block_nonzero(np.ones_like(w), w, (3, 3, 3))[0, 0, 0]) == (1 << 32) + 3


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_vs_reference_live(seed):
    """Differential check of the restatement against the reference's compiled Cython modules."""
    rng = np.random.default_rng(seed)
    shape = tuple(int(s) for s in rng.integers(9, 33, size=3))
    cell = synth_labels(shape, pitch=(7, 6, 5), warp_amp=int(rng.integers(0, 5)), seed=seed)
    cell[rng.random(shape) < 0.03] = rng.integers(0, 50)
    layout = [cell, cell.transpose(1, 0, 2).copy().transpose(1, 0, 2), np.asfortranarray(cell)][seed % 3]
    assert_props_equal(oracle.find_object_properties(layout), ref.find_object_properties(layout), "fop")
    subs = np.stack([synth_labels(shape, pitch=(4, 5, 3), seed=seed, kind=1 + c, density16=4) for c in range(2)])
    o, r = oracle.map_subcell_extract_props(cell, subs), ref.map_subcell_extract_props(cell, subs)
    assert_props_equal(o[0], r[0], "cell")
    for c in range(2):
        assert_props_equal([o[1][k][c] for k in range(3)], [r[1][k][c] for k in range(3)], f"sub{c}")
        assert np.array_equal(map_to_rows(o[2][c]), map_to_rows(r[2][c]))
        assert np.array_equal(map_to_rows(oracle.map_subcell_C(cell, subs)[c]), map_to_rows(ref.map_subcell_C(cell, subs)[c]))
    seg = (cell & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    edges = oracle.detect_seg_boundaries(seg).astype(np.uint32)
    for st in ((3, 3, 3), (5, 7, 3), (7, 7, 3)):
        if all(shape[i] >= st[i] for i in range(3)):
            assert np.array_equal(oracle.process_block_nonzero(edges, seg, st),
                                  np.asarray(ref.process_block_nonzero(edges, seg, st)))


def test_oracle_vs_golden_syntype(golden):
    from helpers import check_syntype_against_golden
    g = golden
    check_syntype_against_golden(oracle.extract_cs_syntype(g["syn_cs"], g["syn_mask"], g["syn_asym"], g["syn_sym"], g["syn_off"]), g)


def test_oracle_vs_golden_cs64(golden64):
    """the numba 64-bit variants (first-seen tie-break, XYZC output, per-pair properties)"""
    from helpers import check_cs64_against_golden
    check_cs64_against_golden(oracle.detect_cs_64bit, oracle.find_object_properties_cs_64bit, golden64)
    # reference tests/test_segmentation_analysis.py:132-135 (test_detect_cs_64bit): same analytic volumes as test_detect_cs
    for d in ([0, 6, 0], [6, 0, 0], [0, 0, 6]):
        sample, lo, hi = gen_sample_seg(np.array(d), STENCIL, 5)
        out = oracle.detect_cs_64bit(sample.astype(np.uint64), tuple(STENCIL))
        assert np.array_equal(out[..., 0], lo) and np.array_equal(out[..., 1], hi)


def test_oracle_closing_vs_scipy():
    """f2: the restated binary closing / dilation equals scipy.ndimage (the reference's call, cs_extraction_steps.py:451-458)
    on random masks, and the restated worker loop equals the literal one."""
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    for _ in range(40):
        shp = tuple(int(v) for v in rng.integers(1, 14, size=3))
        m = rng.random(shp) < rng.choice([0.05, 0.3, 0.7])
        n, d = int(rng.integers(0, 7)), int(rng.integers(0, 3))
        want = ndi.binary_closing(m, iterations=n) if n > 0 else m
        if d > 0:
            want = ndi.binary_dilation(want, iterations=d)
        assert np.array_equal(oracle.binary_closing_dilation(m, n, d), want), (shp, n, d)
    seg = synth_labels((40, 36, 30), pitch=(12, 10, 7), warp_amp=3, seed=2, dtype=np.uint32)
    cs = oracle.detect_cs(seg, (5, 5, 3))
    bb = oracle.find_object_properties(cs)[1]
    a = oracle.close_contact_sites(cs.copy(), bb, 3, 2, use_scipy=True)
    assert np.array_equal(a, oracle.close_contact_sites(cs.copy(), bb, 3, 2)) and (a != cs).any()


def _morph_golden_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "morph_golden.npz"))
    shape = (30, 27, 22)
    n = int(np.prod(shape))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_cases", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    src = open(spec.origin).read()
    ns = {}
    exec(src[src.index("MORPH_CASES = {"):src.index("def morph_input")], ns)  # the op lists, without importing the script
    for key in g.files:
        if key.startswith("in_"):
            name = key[3:key.rindex("_")]
            vin = np.unpackbits(g[key])[:n].reshape(shape).astype(np.uint8)
            vout = np.unpackbits(g["out_" + key[3:]])[:n].reshape(shape).astype(np.uint8)
            yield key[3:], ns["MORPH_CASES"][name], tuple(int(x) for x in g["sc_" + key[3:]]), vin, vout, g


def test_oracle_morphology_vs_golden():
    """Row f4: get_aniso_struct / apply_morphological_operations restatements against the reference's outputs."""
    seen = 0
    for tag, ops, scaling, vin, vout, g in _morph_golden_cases():
        st = oracle.get_aniso_struct(np.array(scaling))
        assert np.array_equal(st.astype(np.uint8), g["struct_%d_%d_%d" % scaling]), scaling
        got = oracle.apply_morphological_operations(vin.copy(), ops, st)
        assert np.array_equal(got, vout), tag
        seen += 1
    assert seen == 18
