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


def pair_props_to_arrays(dicts):
    """nested dicts d[id0][id1] of find_object_properties_cs_64bit -> (keys [N, 2], sizes, bbox, rep), sorted by key."""
    rc, bb, sz = dicts
    keys = sorted((int(a), int(b)) for a, d in sz.items() for b in d)
    assert keys == sorted((int(a), int(b)) for a, d in rc.items() for b in d) == sorted((int(a), int(b)) for a, d in bb.items() for b in d)
    return (np.array(keys, np.uint64).reshape(-1, 2), np.array([int(sz[a][b]) for a, b in keys], np.int64),
            np.array([np.asarray(bb[a][b]) for a, b in keys], np.int64).reshape(-1, 2, 3),
            np.array([np.asarray(rc[a][b]) for a, b in keys], np.int64).reshape(-1, 3))


def check_cs64_against_golden(detect_cs_64bit, find_props_64bit, g):
    """detect_cs_64bit / find_object_properties_cs_64bit (numba, find_object_properties.py:197-269,347-421) against the
    vectors the reference produced (tests/golden/make_golden.py main64).  ``detect_cs_64bit(arr, stencil)``."""
    for st in ((13, 13, 7), (7, 7, 3), (5, 5, 3)):
        out = detect_cs_64bit(g["seg_in"], st)
        assert out.dtype == np.uint64 and np.array_equal(out, g["seg_out_%d_%d_%d" % st]), st
    for st in ((5, 5, 3), (3, 3, 3)):
        assert np.array_equal(detect_cs_64bit(g["tie_in"], st), g["tie_out_%d_%d_%d" % st]), st
    for tag, vol in (("seg", g["seg_out_7_7_3"]), ("tie", g["tie_out_5_5_3"])):
        got = pair_props_to_arrays(find_props_64bit(vol))
        for name, a in zip(("keys", "sizes", "bbox", "rep"), got):
            assert np.array_equal(a, g[f"props_{tag}_{name}"]), f"{tag}: {name} differ"


def reference_contact_site_chunk(oracle, data, sj_d, asym_d, sym_d, offset, stencil, cs_dilation):
    """The chunk-loop body of _contact_site_extraction_thread (cs_extraction_steps.py:381-486) composed from the oracle's
    restatements; ids are closed in ascending order (the reference's order is its hash-map order)."""
    overlap = int(max(np.array(stencil) // 2))
    contacts = np.asarray(oracle.detect_cs(data, stencil))
    bb = oracle.find_object_properties(contacts)[1]
    oracle.close_contact_sites(contacts, {k: bb[k] for k in sorted(bb)}, overlap, cs_dilation)
    c = (slice(overlap, -overlap),) * 3
    res = oracle.extract_cs_syntype(contacts[c], sj_d[c], asym_d[c], sym_d[c], np.asarray(offset) + overlap)
    cs_seg = contacts[c].copy()
    syn = contacts.copy()
    syn[sj_d == 0] = 0
    return res[0], res[1], res[2], res[3], res[4], cs_seg, syn[c]
