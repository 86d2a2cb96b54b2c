"""Shared comparison helpers for the parity tests (order-insensitive, bit-exact)."""
import numpy as np


def props_dicts_to_arrays(dicts):
    rc, bb, sz = dicts
    assert set(rc) == set(bb) == set(sz)
    ids = np.array(sorted(sz.keys()), np.uint64)
    return (ids, np.array([sz[int(k)] for k in ids], np.int64),
            np.array([bb[int(k)] for k in ids], np.int64).reshape(-1, 2, 3),
            np.array([rc[int(k)] for k in ids], np.int64).reshape(-1, 3))


def assert_props_equal(got, want, what=""):
    g, w = props_dicts_to_arrays(got), props_dicts_to_arrays(want)
    for name, a, b in zip(("ids", "sizes", "bbox", "rep"), g, w):
        assert np.array_equal(a, b), f"{what}: {name} differ"


def assert_props_equal_arrays(got_dicts, ids, sizes, bbox, rep, what=""):
    g = props_dicts_to_arrays(got_dicts)
    for name, a, b in zip(("ids", "sizes", "bbox", "rep"), g, (ids, sizes, bbox, rep)):
        assert np.array_equal(a, b.astype(a.dtype)), f"{what}: {name} differ"


def map_to_rows(md):
    rows = sorted((s, c, n) for s, d in md.items() for c, n in d.items())
    return np.array(rows, np.uint64).reshape(-1, 3)


def check_dict_types(dicts):
    """keys int, rep list[int] len 3, bbox [[int]*3,[int]*3], size int."""
    rc, bb, sz = dicts
    for k in list(sz)[:50]:
        assert type(k) is int and type(sz[k]) is int
        assert type(rc[k]) is list and len(rc[k]) == 3 and all(type(v) is int for v in rc[k])
        assert type(bb[k]) is list and len(bb[k]) == 2 and all(type(v) is int for r in bb[k] for v in r)


def check_syntype_against_golden(res, g):
    """extract_cs_syntype result vs the golden arrays produced by the reference (tests/golden/make_golden.py)."""
    cs_p, syn_p, asym, sym, vox = res
    assert_props_equal_arrays(tuple(cs_p), g["syn_csprops_ids"], g["syn_csprops_sizes"], g["syn_csprops_bbox"], g["syn_csprops_rep"], "cs")
    assert_props_equal_arrays(tuple(syn_p), g["syn_synprops_ids"], g["syn_synprops_sizes"], g["syn_synprops_bbox"], g["syn_synprops_rep"], "syn")
    assert np.array_equal(np.array(sorted(dict(asym).items()), np.uint64).reshape(-1, 2), g["syn_asymcnt"])
    assert np.array_equal(np.array(sorted(dict(sym).items()), np.uint64).reshape(-1, 2), g["syn_symcnt"])
    vox = dict(vox)
    ids = np.array(sorted(vox), np.uint64)
    assert np.array_equal(ids, g["syn_vox_ids"])
    assert np.array_equal(np.array([len(vox[int(k)]) for k in ids], np.int64), g["syn_vox_len"])
    assert np.array_equal(np.array([c for k in ids for c in vox[int(k)]], np.int64).reshape(-1, 3), g["syn_vox_xyz"])  # scan order
