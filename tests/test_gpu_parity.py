"""GPU parity tests: the CUDA path, called through the C ABI (host-buffer entry points behind the reference-named
Python shims, and the device-buffer entry points), against the CPU oracle and the committed golden vectors.
Bit-exact: everything here is integer work."""
import numpy as np
import pytest

from helpers import assert_props_equal, assert_props_equal_arrays, check_dict_types, map_to_rows
from test_oracle_pins import (check_known_boundary, check_known_detect_cs, check_known_find_object_properties)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from oracle import oracle
    from syconn_b200.extraction import block_processing_C as bpc
    from syconn_b200.extraction import find_object_properties as fop
    from syconn_b200.extraction import find_object_properties_C as fopc
    from syconn_b200.synth import synth_labels
    return dict(oracle=oracle, bpc=bpc, fop=fop, fopc=fopc, synth=synth_labels)


def test_reference_known_answers(mods):
    check_known_find_object_properties(mods["fop"].find_object_properties)
    check_known_detect_cs(mods["fop"].detect_cs)
    check_known_boundary(mods["fop"].detect_seg_boundaries)


def test_golden_props(mods, golden):
    g, f = golden, mods["fop"].find_object_properties
    got = f(g["fop_in"])
    check_dict_types(got)
    assert_props_equal_arrays(got, g["fop_ids"], g["fop_sizes"], g["fop_bbox"], g["fop_rep"], "fop")
    assert_props_equal_arrays(f(g["fop_in"].transpose(2, 1, 0)), g["fopT_ids"], g["fopT_sizes"], g["fopT_bbox"],
                              g["fopT_rep"], "fop strided")
    assert_props_equal_arrays(f(g["fop32_in"]), g["fop32_ids"], g["fop32_sizes"], g["fop32_bbox"], g["fop32_rep"], "u32")


def test_golden_map(mods, golden):
    g = golden
    cp, sp, md = mods["fop"].map_subcell_extract_props(g["map_cell"], g["map_subs"])
    assert_props_equal_arrays(cp, g["map_cell_ids"], g["map_cell_sizes"], g["map_cell_bbox"], g["map_cell_rep"], "cell")
    mc = mods["fopc"].map_subcell_C(g["map_cell"], g["map_subs"])
    for c in range(3):
        assert_props_equal_arrays((sp[0][c], sp[1][c], sp[2][c]), g[f"map_sub{c}_ids"], g[f"map_sub{c}_sizes"],
                                  g[f"map_sub{c}_bbox"], g[f"map_sub{c}_rep"], f"sub{c}")
        assert np.array_equal(map_to_rows(md[c]), g[f"map_pairs{c}"])
        assert np.array_equal(map_to_rows(mc[c]), g[f"mapC_pairs{c}"])


def test_golden_cs(mods, golden):
    g, fop = golden, mods["fop"]
    assert np.array_equal(fop.detect_seg_boundaries(g["cs_in"]), g["cs_bdry"])
    for st in ((13, 13, 7), (7, 7, 3), (5, 5, 3), (3, 3, 3), (1, 1, 1), (3, 5, 7)):
        assert np.array_equal(fop.detect_cs(g["cs_in"], st), g["cs_out_%d_%d_%d" % st]), st
    for st in ((5, 5, 3), (3, 3, 3)):
        assert np.array_equal(fop.detect_cs(g["tie_in"], st), g["tie_out_%d_%d_%d" % st]), st
    assert np.array_equal(fop.detect_cs(g["many_in"], (5, 5, 3)), g["many_out_5_5_3"])
    edges = np.ones(g["tie_in"].shape, np.uint32)
    assert np.array_equal(mods["bpc"].process_block_nonzero(edges, g["tie_in"], (7, 7, 3)), g["pbn_forced_7_7_3"])
    # x-fastest memory layout (production: ZYX memory seen as XYZ)
    segF = np.ascontiguousarray(g["cs_in"].transpose(2, 1, 0)).transpose(2, 1, 0)
    assert np.array_equal(fop.detect_cs(segF), g["cs_out_13_13_7"])


def test_edge_cases(mods):
    fop, fopc, bpc = mods["fop"], mods["fopc"], mods["bpc"]
    assert fop.find_object_properties(np.zeros((3, 4, 5), np.uint64)) == ({}, {}, {})
    assert fop.find_object_properties(np.zeros((0, 4, 5), np.uint64)) == ({}, {}, {})
    r = fop.map_subcell_extract_props(np.zeros((3, 3, 3), np.uint64), np.zeros((2, 3, 3, 3), np.uint64))
    assert r == ([{}, {}, {}], [[{}, {}], [{}, {}], [{}, {}]], [{}, {}])
    cell = np.zeros((2, 2, 2), np.uint64)
    sub = np.zeros((1, 2, 2, 2), np.uint64)
    sub[0, 1, 1, 0] = 9
    cp, sp, md = fop.map_subcell_extract_props(cell, sub)
    assert sp[2][0] == {9: 1} and sp[0][0] == {9: [1, 1, 0]} and md == [{}]
    w = np.ones((3, 3, 3), np.uint32)
    w[0, 0, 0], w[2, 2, 2] = 3, 2
    assert int(bpc.process_block_nonzero(np.ones_like(w), w, (3, 3, 3))[0, 0, 0]) == (1 << 32) + 2
    assert bpc.kernel(w, 1) == (1 << 32) + 2
    w[0, 0, 1] = 3
    assert int(bpc.process_block_nonzero(np.ones_like(w), w, (3, 3, 3))[0, 0, 0]) == (1 << 32) + 3
    # window with only centre/background -> 0 ; background centre with forced edge -> partner id, high word 0
    z = np.zeros((3, 3, 3), np.uint32)
    z[1, 1, 1] = 5
    assert int(bpc.process_block_nonzero(np.ones_like(z), z, (3, 3, 3))[0, 0, 0]) == 0
    z[1, 1, 1], z[0, 0, 0] = 0, 9
    assert int(bpc.process_block_nonzero(np.ones_like(z), z, (3, 3, 3))[0, 0, 0]) == 9
    # two half spaces, all-ones edge mask
    h = np.full((9, 9, 5), 5, np.uint32)
    h[5:] = 9
    out = bpc.process_block_nonzero(np.ones_like(h), h, (7, 7, 3))
    assert out.shape == (3, 3, 3) and np.all(out == 0x500000009)
    # ids above 2^32 in 64-bit volumes and > 2^31 voxel... sizes are plain ints
    big = np.full((4, 4, 4), 2 ** 63 + 5, np.uint64)
    rc, bb, sz = fop.find_object_properties(big)
    assert sz == {2 ** 63 + 5: 64} and bb[2 ** 63 + 5] == [[0, 0, 0], [4, 4, 4]]


@pytest.mark.parametrize("seed,shape", [(0, (33, 70, 41)), (1, (64, 64, 64)), (2, (5, 130, 97)), (3, (96, 17, 1))])
def test_props_vs_oracle_random(mods, seed, shape):
    rng = np.random.default_rng(seed)
    oracle, synth = mods["oracle"], mods["synth"]
    v = synth(shape, pitch=(9, 7, 6), warp_amp=seed, seed=seed)
    v[rng.random(shape) < 0.02] = rng.integers(0, 2 ** 40)
    for lay in (v, np.asfortranarray(v), v.transpose(1, 0, 2).copy().transpose(1, 0, 2), v[::-1, :, ::2]):
        assert_props_equal(mods["fop"].find_object_properties(lay), oracle.find_object_properties(lay), f"{seed}")
    v32 = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert_props_equal(mods["fop"].find_object_properties(v32), oracle.find_object_properties(v32), "u32")


def test_props_high_cardinality(mods):
    """every voxel its own id (hash-table stress, private-table early flush path)."""
    v = (np.arange(48 * 40 * 36, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1)).reshape(48, 40, 36)
    assert_props_equal(mods["fop"].find_object_properties(v), mods["oracle"].find_object_properties(v), "unique ids")


@pytest.mark.parametrize("seed,shape,nsub", [(0, (40, 50, 37), 3), (1, (64, 64, 64), 2), (2, (20, 33, 70), 5), (3, (16, 16, 16), 1)])
def test_map_vs_oracle_random(mods, seed, shape, nsub):
    rng = np.random.default_rng(seed)
    oracle, synth = mods["oracle"], mods["synth"]
    cell = synth(shape, pitch=(10, 9, 7), warp_amp=2, seed=seed)
    subs = np.stack([synth(shape, pitch=(5, 6, 4), warp_amp=3, seed=seed, kind=1 + c, density16=3 + c) for c in range(nsub)])
    subs[:, rng.random(shape) < 0.01] = 77
    for c_arr, s_arr in ((cell, subs), (np.asfortranarray(cell), np.ascontiguousarray(subs.transpose(0, 3, 2, 1)).transpose(0, 3, 2, 1))):
        o = oracle.map_subcell_extract_props(c_arr, s_arr)
        g = mods["fop"].map_subcell_extract_props(c_arr, s_arr)
        assert_props_equal(g[0], o[0], "cell")
        for c in range(nsub):
            assert_props_equal([g[1][k][c] for k in range(3)], [o[1][k][c] for k in range(3)], f"sub{c}")
            assert np.array_equal(map_to_rows(g[2][c]), map_to_rows(o[2][c])), f"pairs {c}"
        mc, oc = mods["fopc"].map_subcell_C(c_arr, s_arr), oracle.map_subcell_C(c_arr, s_arr)
        for c in range(nsub):
            assert np.array_equal(map_to_rows(mc[c]), map_to_rows(oc[c]))


@pytest.mark.parametrize("seed,shape,st", [(0, (40, 44, 39), (13, 13, 7)), (1, (30, 30, 30), (7, 7, 3)),
                                           (2, (21, 37, 50), (5, 5, 3)), (3, (70, 20, 19), (3, 3, 3)),
                                           (4, (24, 24, 24), (17, 17, 9)), (5, (16, 20, 33), (1, 3, 1)),
                                           (6, (48, 40, 36), (15, 15, 9)), (7, (40, 36, 44), (9, 15, 15)),
                                           (8, (44, 40, 40), (13, 13, 7))])  # 16-byte rows in both orders: specialised kernels
def test_detect_cs_vs_oracle_random(mods, seed, shape, st):
    rng = np.random.default_rng(seed)
    oracle, synth = mods["oracle"], mods["synth"]
    seg = synth(shape, pitch=(8, 9, 5), warp_amp=3, seed=seed, dtype=np.uint32)
    seg[rng.random(shape) < 0.03] = rng.integers(0, 2 ** 32 - 1)
    want = oracle.detect_cs(seg, st)
    assert np.array_equal(mods["fop"].detect_cs(seg, st), want)
    assert np.array_equal(mods["fop"].detect_cs(np.asfortranarray(seg), st), want)
    # uint64 input: boundary mask on the 64-bit values, window histogram on the ids narrowed like .astype(np.uint32)
    # (find_object_properties.py:466-468); here even the background carries high bits, i.e. it is a 64-bit id of its own
    seg64 = seg.astype(np.uint64) | (np.uint64(3) << np.uint64(40))
    assert np.array_equal(mods["fop"].detect_cs(seg64, st), oracle.detect_cs(seg64, st))
    seg64 = np.where(seg != 0, seg64, np.uint64(0))
    assert np.array_equal(mods["fop"].detect_cs(seg64, st), want)
    edges = oracle.detect_seg_boundaries(seg).astype(np.uint32)
    assert np.array_equal(mods["bpc"].process_block_nonzero(edges, seg, st), want)
    assert np.array_equal(mods["fop"].detect_seg_boundaries(seg), edges.astype(bool))


def test_detect_cs_with_props_extension(mods):
    """detect_cs(..., return_props=True) == (detect_cs, find_object_properties(contacts)) of the reference worker."""
    seg = mods["synth"]((60, 52, 47), pitch=(14, 12, 8), warp_amp=3, seed=11, dtype=np.uint32)
    cs, props = mods["fop"].detect_cs(seg, (7, 7, 3), return_props=True)
    want = mods["oracle"].detect_cs(seg, (7, 7, 3))
    assert np.array_equal(cs, want)
    assert_props_equal(props, mods["oracle"].find_object_properties(want), "cs props")


def test_host_calls_from_worker_threads(mods):
    """The host entry points are issued from several threads at once (each thread owns a stream inside libsyk):
    results must equal the serial ones."""
    from concurrent.futures import ThreadPoolExecutor
    fop, synth = mods["fop"], mods["synth"]
    segs = [synth((72, 64, 56), origin=(100 * i, 7, 3), pitch=(14, 12, 8), warp_amp=3, seed=20 + i, dtype=np.uint32) for i in range(6)]
    labs = [synth((64, 48, 40), origin=(5, 90 * i, 1), pitch=(9, 9, 6), seed=30 + i) for i in range(6)]
    want_cs = [fop.detect_cs(s, (7, 7, 3)) for s in segs]
    want_pr = [fop.find_object_properties(l) for l in labs]

    def job(i):
        return fop.detect_cs(segs[i], (7, 7, 3)), fop.find_object_properties(labs[i])
    with ThreadPoolExecutor(max_workers=4) as pool:
        for _ in range(3):
            got = list(pool.map(job, range(6)))
            for i, (cs, pr) in enumerate(got):
                assert np.array_equal(cs, want_cs[i])
                assert_props_equal(pr, want_pr[i], f"props thread job {i}")
    assert np.array_equal(want_cs[0], mods["oracle"].detect_cs(segs[0], (7, 7, 3)))


def test_detect_cs_random_labels_overflow_path(mods):
    """near-random labels: > 32 distinct ids per window (block-cooperative hash fallback)."""
    rng = np.random.default_rng(7)
    seg = rng.integers(1, 2 ** 32 - 1, size=(20, 20, 18), dtype=np.uint64).astype(np.uint32)
    seg[rng.random(seg.shape) < 0.3] = 11
    for st in ((5, 5, 3), (7, 7, 3)):
        assert np.array_equal(mods["fop"].detect_cs(seg, st), mods["oracle"].detect_cs(seg, st))


def test_device_api_and_synth_twin(mods):
    """device-buffer entry points on torch tensors; the CUDA generator is bit-identical to the NumPy twin."""
    import torch
    from syconn_b200 import device as dev
    oracle, synth = mods["oracle"], mods["synth"]
    shape, origin = (50, 45, 70), (-6, 506, 1000)
    for order in ("C", "F"):
        for kind, dt, nd in ((0, torch.int64, np.uint64), (2, torch.int64, np.uint64), (0, torch.int32, np.uint32)):
            t = dev.synth_labels(shape, origin, pitch=(11, 9, 7), warp_amp=4, seed=5, kind=kind, density16=5, dtype=dt, order=order)
            n = synth(shape, origin, pitch=(11, 9, 7), warp_amp=4, seed=5, kind=kind, density16=5, dtype=nd, order=order)
            assert np.array_equal(t.cpu().numpy().view(nd), n), (order, kind, dt)
    lab = dev.synth_labels(shape, origin, pitch=(11, 9, 7), seed=5)
    tab = dev.IdTable(1 << 12)
    dev.find_object_properties(tab, lab, origin=origin, chunk_seq=3)
    recs = dev.records_numpy(tab.export(dev.geoms([[0, 0, 0]] * 3 + [origin], [[1, 1, 1]] * 3 + [shape])))
    want = oracle.find_object_properties_arrays(lab.cpu().numpy().view(np.uint64))
    o = np.argsort(recs["id"])
    w = np.argsort(want[0])
    assert np.array_equal(recs["id"][o], want[0][w]) and np.array_equal(recs["count"][o].astype(np.int64), want[1][w])
    assert np.array_equal(recs["bb_min"][o], want[2][w][:, 0] + np.array(origin)) and np.array_equal(recs["bb_max"][o], want[2][w][:, 1] + np.array(origin))
    assert np.array_equal(recs["rep"][o], want[3][w] + np.array(origin)) and np.all(recs["chunk_seq"] == 3)
    # fused detect_cs on device, both layouts
    seg = dev.synth_labels((60, 50, 40), pitch=(12, 10, 6), seed=9, dtype=torch.int32)
    want = oracle.detect_cs(seg.cpu().numpy().view(np.uint32))
    assert np.array_equal(dev.detect_cs(seg).cpu().numpy().view(np.uint64), want)
    segF = dev.synth_labels((60, 50, 40), pitch=(12, 10, 6), seed=9, dtype=torch.int32, order="F")
    outF = dev.detect_cs(segF)
    assert outF.stride(0) == 1 and np.array_equal(outF.cpu().numpy().view(np.uint64), want)


def test_randomized_sweep_vs_oracle(mods):
    """Seeded sweep over shapes, memory layouts, dtypes, label alphabets and stencils (all three stages)."""
    rng = np.random.default_rng(2024)
    oracle, fop, fopc = mods["oracle"], mods["fop"], mods["fopc"]
    perms = [(0, 1, 2), (2, 1, 0), (1, 0, 2), (0, 2, 1)]
    for case in range(24):
        shape = tuple(int(s) for s in rng.integers(1, 40, size=3))
        nlab = int(rng.choice([2, 5, 40, 5000]))
        dt = np.uint64 if case % 3 else np.uint32
        hi = 2 ** 40 if dt == np.uint64 else 2 ** 32 - 1
        alphabet = np.concatenate([[0], rng.integers(1, hi, size=nlab)]).astype(dt)
        blocky = mods["synth"](shape, pitch=tuple(int(p) for p in rng.integers(2, 9, size=3)), warp_amp=int(rng.integers(0, 6)),
                                       seed=case)
        vol = alphabet[(blocky % np.uint64(len(alphabet))).astype(np.int64)]
        noise = rng.random(shape) < 0.05
        vol[noise] = alphabet[rng.integers(0, len(alphabet), size=int(noise.sum()))]
        perm = perms[case % 4]
        lay = np.ascontiguousarray(vol.transpose(perm)).transpose(np.argsort(perm))  # same logical array, other memory order
        assert np.array_equal(lay, vol)
        assert_props_equal(fop.find_object_properties(lay), oracle.find_object_properties(vol), f"case {case} props")
        nsub = int(rng.integers(1, 4))
        subs = np.stack([np.where(rng.random(shape) < 0.3, alphabet[rng.integers(0, len(alphabet), size=shape)], 0).astype(dt)
                         for _ in range(nsub)])
        g, o = fop.map_subcell_extract_props(lay, subs), oracle.map_subcell_extract_props(vol, subs)
        assert_props_equal(g[0], o[0], f"case {case} cell")
        for c in range(nsub):
            assert_props_equal([g[1][k][c] for k in range(3)], [o[1][k][c] for k in range(3)], f"case {case} sub{c}")
            assert np.array_equal(map_to_rows(g[2][c]), map_to_rows(o[2][c])), f"case {case} pairs{c}"
        st = tuple(int(s) for s in rng.choice([1, 3, 5, 7, 9, 13], size=3))
        if all(shape[i] >= st[i] for i in range(3)):
            seg = lay.astype(np.uint32) if dt == np.uint32 else lay
            want = oracle.detect_cs(np.ascontiguousarray(vol), st)
            assert np.array_equal(fop.detect_cs(seg, st), want), f"case {case} detect_cs {st} {shape}"


def test_extract_cs_syntype(mods, golden):
    """"next" row f1: block_processing_C.extract_cs_syntype against the reference's golden vector and the oracle."""
    from helpers import check_syntype_against_golden
    g, bpc, oracle = golden, mods["bpc"], mods["oracle"]
    check_syntype_against_golden(bpc.extract_cs_syntype(g["syn_cs"], g["syn_mask"], g["syn_asym"], g["syn_sym"], g["syn_off"]), g)
    rng = np.random.default_rng(5)
    for shape, dt in (((33, 40, 29), np.uint64), ((20, 64, 17), np.uint32), ((5, 5, 5), np.uint64)):
        cs = (mods["synth"](shape, pitch=(6, 5, 4), seed=3) % np.uint64(7)).astype(dt) * dt(3)
        syn = ((rng.random(shape) < 0.3) * rng.integers(1, 4, size=shape)).astype(np.uint8)
        asym, sym = rng.integers(0, 3, size=shape).astype(np.uint8), rng.integers(0, 3, size=shape).astype(np.uint8)
        for lay in (lambda a: a, np.asfortranarray):
            got = bpc.extract_cs_syntype(lay(cs), lay(syn), asym, lay(sym), [1, 2, 3])
            want = oracle.extract_cs_syntype(cs, syn, asym, sym, [1, 2, 3])
            assert_props_equal(tuple(got[0]), tuple(want[0]), "cs")
            assert_props_equal(tuple(got[1]), tuple(want[1]), "syn")
            assert got[2] == want[2] and got[3] == want[3] and got[4] == want[4]
    empty = bpc.extract_cs_syntype(np.zeros((4, 4, 4), np.uint64), np.ones((4, 4, 4), np.uint8), np.ones((4, 4, 4), np.uint8),
                                   np.ones((4, 4, 4), np.uint8), [0, 0, 0])
    assert empty == ([{}, {}, {}], [{}, {}, {}], {}, {}, {})


def test_cs64_variants(mods, golden64):
    """detect_cs_64bit / detect_contact_partners / find_object_properties_cs_64bit (the numba twins,
    find_object_properties.py:197-269,347-421): golden vectors of the reference, then seeded volumes against the oracle."""
    from helpers import check_cs64_against_golden, pair_props_to_arrays
    from syconn_b200 import global_params
    fop, oracle = mods["fop"], mods["oracle"]
    keep = list(global_params.config['cell_objects']['cs_filtersize'])

    def dcs(arr, st):
        global_params.config['cell_objects']['cs_filtersize'] = list(st)
        try:
            return fop.detect_cs_64bit(arr)
        finally:
            global_params.config['cell_objects']['cs_filtersize'] = keep
    check_cs64_against_golden(dcs, fop.find_object_properties_cs_64bit, golden64)
    rng = np.random.default_rng(3)
    for seed, shape, st in ((0, (30, 28, 26), (13, 13, 7)), (1, (21, 33, 18), (5, 7, 3)), (2, (16, 16, 40), (3, 3, 3))):
        seg = mods["synth"](shape, pitch=(7, 8, 5), warp_amp=3, seed=seed)
        seg[rng.random(shape) < 0.04] = rng.integers(1, 2 ** 50)
        want = oracle.detect_cs_64bit(seg, st)
        for lay in (seg, np.asfortranarray(seg)):
            assert np.array_equal(dcs(lay, st), want), (seed, st)
        # explicit edge mask (bool / uint32) and uint32 labels
        o = np.array(st) // 2
        off = np.array([(-o[0], o[0]), (-o[1], o[1]), (-o[2], o[2])])
        edges = rng.random(shape) < 0.2
        want_e = oracle.detect_contact_partners(seg, edges, off)
        assert np.array_equal(fop.detect_contact_partners(seg, edges, off), want_e)
        assert np.array_equal(fop.detect_contact_partners(seg, edges.astype(np.uint32), off), want_e)
        seg32 = (seg & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        assert np.array_equal(fop.detect_contact_partners(seg32, edges, off), oracle.detect_contact_partners(seg32, edges, off))
        got, wantp = fop.find_object_properties_cs_64bit(want), oracle.find_object_properties_cs_64bit(want)
        for a, b in zip(pair_props_to_arrays(got), pair_props_to_arrays(wantp)):
            assert np.array_equal(a, b)
    # near-random labels: more than 32 distinct ids per window (block-cooperative fallback keeps the first-seen rank)
    noise = rng.integers(1, 2 ** 40, size=(14, 13, 12)).astype(np.uint64)
    noise[rng.random(noise.shape) < 0.4] = 5
    assert np.array_equal(dcs(noise, (5, 5, 3)), oracle.detect_cs_64bit(noise, (5, 5, 3)))
    assert fop.find_object_properties_cs_64bit(np.zeros((3, 3, 3, 2), np.uint64)) == ({}, {}, {})
    with pytest.raises(NotImplementedError):
        fop.detect_contact_partners(noise, None, np.array([(-1, 2), (-1, 1), (-1, 1)]))


def test_close_contact_sites(mods, monkeypatch):
    """"next" row f2: the per-contact closing / dilation loop (cs_extraction_steps.py:439-461) against the oracle's
    literal restatement, in the same id order; shared-memory path, HBM path and HBM batches."""
    from syconn_b200.extraction.cs_extraction_steps import close_contact_sites
    oracle, fop = mods["oracle"], mods["fop"]
    seg = mods["synth"]((64, 60, 52), pitch=(14, 12, 8), warp_amp=3, seed=1, dtype=np.uint32)
    cs = oracle.detect_cs(seg, (7, 7, 3))
    bb = oracle.find_object_properties(cs)[1]
    rng = np.random.default_rng(0)
    keys = list(bb)
    rng.shuffle(keys)
    bb_shuffled = {k: bb[k] for k in keys}
    for n_close, n_dil in ((3, 2), (6, 2), (0, 2), (2, 0), (1, 1)):
        for order in (bb, bb_shuffled):
            want = oracle.close_contact_sites(cs.copy(), order, n_close, n_dil)
            got = close_contact_sites(cs.copy(), order, n_close, n_dil)
            assert np.array_equal(got, want), (n_close, n_dil)
    want = oracle.close_contact_sites(cs.copy(), bb, 3, 2)
    assert (want != cs).sum() > 1000                                   # the closing does fill gaps in this volume
    assert np.array_equal(oracle.close_contact_sites(cs.copy(), bb, 3, 2, use_scipy=True), want)
    # memory layouts: x-fastest (production), a non-dense view
    csF = np.asfortranarray(cs)
    assert np.array_equal(close_contact_sites(csF, bb, 3, 2), want) and csF.flags.f_contiguous
    wide = np.zeros((cs.shape[0], cs.shape[1], cs.shape[2] + 3), np.uint64)
    wide[:, :, :-3] = cs
    assert np.array_equal(close_contact_sites(wide[:, :, :-3], bb, 3, 2), want) and np.array_equal(wide[:, :, :-3], want)
    # defaults from the config (n_closings = max(cs_filtersize // 2) = 6, cs_dilation = 2), boxes from the library itself
    got = close_contact_sites(cs.copy())
    assert np.array_equal(got, oracle.close_contact_sites(cs.copy(), {k: bb[k] for k in sorted(bb)}, 6, 2))
    # HBM path (boxes above the shared-memory threshold) in several batches
    monkeypatch.setenv("SYK_MORPH_SMALL", "64")
    monkeypatch.setenv("SYK_MORPH_BATCH", "5000")
    assert np.array_equal(close_contact_sites(cs.copy(), bb_shuffled, 3, 2), oracle.close_contact_sites(cs.copy(), bb_shuffled, 3, 2))
    monkeypatch.setenv("SYK_MORPH_SMALL", "0")
    assert np.array_equal(close_contact_sites(np.asfortranarray(cs), bb, 6, 2), oracle.close_contact_sites(cs.copy(), bb, 6, 2))
    monkeypatch.delenv("SYK_MORPH_SMALL")
    monkeypatch.delenv("SYK_MORPH_BATCH")
    # one object spanning the whole volume with holes (box wider than one 32-voxel word, clipped on every side), uint32
    big = (rng.random((40, 37, 100)) < 0.15).astype(np.uint32) * np.uint32(7)
    big[5:9, 5:9, 40:60] = 9
    bbb = oracle.find_object_properties(big)[1]
    assert np.array_equal(close_contact_sites(big.copy(), bbb, 2, 1), oracle.close_contact_sites(big.copy(), bbb, 2, 1))
    # ids >= 2^63 and the no-op cases
    hi = cs.copy()
    hi[hi != 0] |= np.uint64(1) << np.uint64(63)
    bbh = oracle.find_object_properties(hi)[1]
    assert np.array_equal(close_contact_sites(hi.copy(), bbh, 2, 1), oracle.close_contact_sites(hi.copy(), bbh, 2, 1))
    assert np.array_equal(close_contact_sites(cs.copy(), bb, 0, 0), cs)
    assert np.array_equal(close_contact_sites(cs.copy(), {}, 3, 2), cs)
    z = np.zeros((4, 4, 4), np.uint64)
    assert not close_contact_sites(z).any()


def test_contact_site_extraction_chunk(mods):
    """the numeric body of the contact-site worker for one chunk (detect_cs -> props -> closing -> extract_cs_syntype),
    device resident, against the same composition of the oracle's restatements"""
    from helpers import reference_contact_site_chunk
    from syconn_b200.extraction.cs_extraction_steps import contact_site_extraction_chunk
    oracle = mods["oracle"]
    rng = np.random.default_rng(4)
    for st, dil, size, lay in (((7, 7, 3), 2, (40, 36, 30), "zyx"), ((5, 5, 5), 1, (33, 31, 35), "xyz")):
        so, ov = np.array(st) // 2, max(np.array(st) // 2)
        full = tuple(int(size[i] + 2 * ov + 2 * so[i]) for i in range(3))
        data = mods["synth"](full, origin=(100, 50, 7), pitch=(13, 11, 8), warp_amp=3, seed=6, dtype=np.uint32)
        oshape = tuple(full[i] - st[i] + 1 for i in range(3))
        sj = ((rng.random(oshape) < 0.4) * rng.integers(1, 3, size=oshape)).astype(np.uint8)
        asym, sym = rng.integers(0, 3, size=oshape).astype(np.uint8), rng.integers(0, 3, size=oshape).astype(np.uint8)
        if lay == "zyx":  # production: ZYX memory seen as XYZ (kd.load_seg(...).swapaxes(0, 2))
            data, sj, asym, sym = (np.ascontiguousarray(a.transpose(2, 1, 0)).transpose(2, 1, 0) for a in (data, sj, asym, sym))
        offset = np.array([100, 50, 7]) + so                     # position of the contact volume in the dataset
        got = contact_site_extraction_chunk(data, sj, asym, sym, offset, st, dil)
        want = reference_contact_site_chunk(oracle, data, sj, asym, sym, offset, st, dil)
        assert_props_equal(tuple(got[0]), tuple(want[0]), "cs")
        assert_props_equal(tuple(got[1]), tuple(want[1]), "syn")
        assert got[2] == want[2] and got[3] == want[3] and got[4] == want[4]
        assert got[5].dtype == np.uint64 and np.array_equal(got[5], want[5]) and np.array_equal(got[6], want[6])
        assert len(want[4]) > 5 and (want[5] != 0).sum() > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["F", "C"])
def test_detect_cs_marching_variant_vs_oracle(order, monkeypatch):
    """The opt-in u -> v -> w marching tier 1 (csrc/syk_cs_march.cuh, SYK_CS_MARCH=1; TMA plane loads, two barriers per
    plane) is bit-exact too: production stencil, both memory orders, a pitch that forces slot recycling and tier-2 hand-over."""
    import torch
    from oracle import oracle
    from syconn_b200 import device as dev
    monkeypatch.setenv("SYK_CS_MARCH", "1")
    for shape, pitch in (((90, 80, 100), (16, 16, 8)), ((150, 100, 140), (32, 32, 16)), ((100, 100, 100), (7, 7, 5))):
        seg = dev.synth_labels(shape, origin=(5, -3, 11), pitch=pitch, seed=2, dtype=torch.int32, order=order)
        got = dev.detect_cs(seg, (13, 13, 7)).cpu().numpy().view(np.uint64)
        want = oracle.detect_cs(seg.cpu().numpy().view(np.uint32), (13, 13, 7))
        assert np.array_equal(got, want), (shape, pitch, order)


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["C", "F"])
def test_process_block_nonzero_explicit_edges_fast_path(order):
    """process_block_nonzero(edges, arr, stencil) with a caller-supplied mask (INTEGRATION.md mode 1: the reference's own
    numba detect_seg_boundaries + the shadowed Cython module) runs the box-sum kernels too: the mask only replaces the
    fused boundary test.  Random masks flag background centres and interior voxels as well (block_processing_C.pyx:66-73)."""
    import torch
    from oracle import oracle
    from syconn_b200 import device as dev
    rng = np.random.default_rng(5)
    for shape, pitch, st in (((70, 64, 90), (16, 16, 8), (13, 13, 7)), ((90, 60, 70), (24, 20, 12), (7, 7, 3)),
                             ((64, 64, 64), (9, 9, 9), (5, 5, 3))):
        seg = dev.synth_labels(shape, origin=(3, 1, -2), pitch=pitch, seed=4, dtype=torch.int32, order=order)
        seg_np = seg.cpu().numpy().view(np.uint32)
        for kind in ("boundary", "random"):
            if kind == "boundary":
                edges_np = oracle.detect_seg_boundaries(seg_np).astype(np.uint32)
            else:
                edges_np = (rng.random(shape) < 0.15).astype(np.uint32) * np.uint32(7)
            want = oracle.process_block_nonzero(edges_np, np.ascontiguousarray(seg_np), st)
            for edt in (torch.int32, torch.uint8):
                e = torch.from_numpy((edges_np != 0).astype(np.uint8) if edt == torch.uint8 else edges_np.view(np.int32)).cuda()
                if order == "F":
                    e = e.permute(2, 1, 0).contiguous().permute(2, 1, 0)
                got = dev.process_block_nonzero(e, seg, st).cpu().numpy().view(np.uint64)
                assert np.array_equal(got, want), (shape, st, kind, edt, order)


@pytest.mark.gpu
def test_detect_cs_uint64_boundaries_use_all_64_bits():
    """find_object_properties.py:466-468: detect_seg_boundaries sees the 64-bit ids, the window histogram the ids narrowed to
    uint32.  Two neighbours that are equal modulo 2^32 still form a boundary, whose voxels then report the most frequent
    OTHER narrowed id of the window."""
    import torch
    from oracle import oracle
    from syconn_b200 import device as dev
    from syconn_b200.extraction.find_object_properties import detect_cs
    a = np.zeros((30, 24, 26), np.uint64)
    a[:, :12, :] = np.uint64(5) | (np.uint64(1) << np.uint64(32))
    a[:, 12:, :] = np.uint64(5) | (np.uint64(2) << np.uint64(32))     # same low word, different id
    a[:, 13:, 13:] = np.uint64(9)                                     # a third id inside the window of the 5|5 interface
    for st in ((13, 13, 7), (5, 5, 3)):
        want = oracle.detect_cs(a, st)
        assert want[:, 11 - st[1] // 2, :].any()                      # the equal-mod-2^32 interface does produce contacts
        got_host = detect_cs(a, st)
        got_dev = dev.detect_cs(torch.from_numpy(a.view(np.int64)).cuda(), st).cpu().numpy().view(np.uint64)
        assert np.array_equal(got_host, want) and np.array_equal(got_dev, want), st


@pytest.mark.gpu
def test_detect_cs_wide_stencils_march_along_the_other_axis():
    """Stencils wider than 15 along the natural v axis (17 x 17 x 9 on x-fastest data) stay on the box-sum path by marching
    along v instead of u (csrc/syk_cs.cu::cs_plan); C-order data with 17 along both slow axes uses the generic kernel.
    Both must equal the oracle."""
    import torch
    from oracle import oracle
    from syconn_b200 import device as dev
    for order in ("F", "C"):
        for st in ((17, 17, 9), (9, 17, 17), (17, 9, 17), (15, 17, 5)):
            seg = dev.synth_labels((60, 58, 56), origin=(1, 2, 3), pitch=(16, 16, 8), seed=6, dtype=torch.int32, order=order)
            got = dev.detect_cs(seg, st).cpu().numpy().view(np.uint64)
            want = oracle.detect_cs(seg.cpu().numpy().view(np.uint32), st)
            assert np.array_equal(got, want), (order, st)


def test_close_contacts_from_device_records(mods, monkeypatch):
    """syk_close_contacts_records (boxes planned on the device from a sorted table export) == the oracle's loop in ascending
    id order; shared-memory classes, HBM path, both memory orders; sorted export == host-sorted export."""
    import torch
    from syconn_b200 import device as dev
    oracle = mods["oracle"]
    for seed, shape, pitch, st in ((1, (64, 60, 52), (14, 12, 8), (7, 7, 3)), (2, (50, 70, 90), (20, 9, 30), (5, 5, 3))):
        seg = mods["synth"](shape, pitch=pitch, warp_amp=3, seed=seed, dtype=np.uint32)
        cs = oracle.detect_cs(seg, st)
        bb = oracle.find_object_properties(cs)[1]
        order = {k: bb[k] for k in sorted(bb)}
        for n_close, n_dil in ((3, 2), (6, 2), (0, 2), (2, 0)):
            want = oracle.close_contact_sites(cs.copy(), order, n_close, n_dil)
            for small in (None, "64", "0"):
                if small is None:
                    monkeypatch.delenv("SYK_MORPH_SMALL", raising=False)
                else:
                    monkeypatch.setenv("SYK_MORPH_SMALL", small)
                    monkeypatch.setenv("SYK_MORPH_BATCH", "5000")
                for fortran in (False, True):
                    t = torch.from_numpy(cs.view(np.int64).copy()).cuda()
                    if fortran:
                        t = t.permute(2, 1, 0).contiguous().permute(2, 1, 0)
                    table = dev.IdTable(1 << 14)
                    dev.find_object_properties(table, t)
                    g = dev.geoms([[0, 0, 0]], [list(t.shape)])
                    rec = table.export(g, sort=True)
                    host = dev.records_numpy(table.export(g))
                    assert np.array_equal(dev.records_numpy(rec), host[np.argsort(host["id"])])
                    dev.close_contacts_records(t, rec, n_close, n_dil)
                    assert np.array_equal(t.cpu().numpy().view(np.uint64), want), (seed, n_close, n_dil, small, fortran)
                    table.close()
    monkeypatch.delenv("SYK_MORPH_SMALL", raising=False)
    monkeypatch.delenv("SYK_MORPH_BATCH", raising=False)
    # a bounding box outside the volume is refused
    bad = rec.clone()
    bad[0, 5] = 1 << 20  # bb_max words
    with pytest.raises(Exception):
        dev.close_contacts_records(t, bad, 1, 1)


def test_extract_cs_syntype_fused_syn_props(mods):
    """syk_extract_cs_syntype_props: the synaptic props accumulated by the voxel-compaction pass itself == the second return
    value of the oracle's extract_cs_syntype (block_processing_C.pyx:117-137); layouts, cropped views, offsets, chunk_seq."""
    import torch
    from syconn_b200 import device as dev
    from syconn_b200.extraction._host import records_to_dicts
    oracle = mods["oracle"]
    rng = np.random.default_rng(9)
    for shape, off in (((33, 40, 29), (1, 2, 3)), ((20, 64, 70), (100, -5, 7)), ((6, 5, 4), (0, 0, 0))):
        cs = (mods["synth"](shape, pitch=(6, 5, 4), seed=4) % np.uint64(9)) * np.uint64(5)
        syn = ((rng.random(shape) < 0.35) * rng.integers(1, 4, size=shape)).astype(np.uint8)
        asym, sym = rng.integers(0, 3, size=shape).astype(np.uint8), rng.integers(0, 3, size=shape).astype(np.uint8)
        want = oracle.extract_cs_syntype(cs, syn, asym, sym, list(off))

        def shifted(props, o):  # the Cython props are block-local (the worker adds the offset when merging, :482-483)
            rc, bb, sz = props
            o = np.asarray(o)
            return ({k: (np.asarray(v) + o).tolist() for k, v in rc.items()}, {k: (np.asarray(v) + o).tolist() for k, v in bb.items()}, sz)
        want = (shifted(want[0], off), shifted(want[1], off)) + tuple(want[2:])
        for fortran in (False, True):
            def put(a):
                t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).cuda()
                return t.permute(2, 1, 0).contiguous().permute(2, 1, 0) if fortran else t
            t_cs, t_syn = dev.IdTable(1 << 10), dev.IdTable(1 << 10)
            vox = dev.extract_cs_syntype(t_cs, put(cs), put(syn), put(asym), put(sym), origin=off, chunk_seq=3, syn_table=t_syn)
            g = dev.geoms([list(off)] * 4, [list(shape)] * 4)
            got_cs = records_to_dicts(dev.records_numpy(t_cs.export(g)))
            got_syn = records_to_dicts(dev.records_numpy(t_syn.export(g)))
            assert_props_equal(tuple(got_cs), tuple(want[0]), "cs")
            assert_props_equal(tuple(got_syn), tuple(want[1]), "syn")
            assert int(vox.shape[0]) == sum(len(v) for v in want[4].values())
            assert set(dev.records_numpy(t_syn.export(g))["chunk_seq"].tolist()) <= {3}
            t_cs.close()
            t_syn.close()
    # a cropped view (strided) of larger volumes, as the worker passes them
    big = (mods["synth"]((40, 44, 48), pitch=(7, 6, 5), seed=6) % np.uint64(11)) * np.uint64(3)
    mask = (rng.random(big.shape) < 0.3).astype(np.uint8)
    crop = (slice(6, -6),) * 3
    want = oracle.extract_cs_syntype(np.ascontiguousarray(big[crop]), np.ascontiguousarray(mask[crop]), np.ascontiguousarray(mask[crop]),
                                     np.ascontiguousarray(mask[crop]), [10, 20, 30])
    tb, tm = torch.from_numpy(big.view(np.int64)).cuda(), torch.from_numpy(mask).cuda()
    t_cs, t_syn = dev.IdTable(1 << 10), dev.IdTable(1 << 10)
    dev.extract_cs_syntype(t_cs, tb[crop], tm[crop], tm[crop], tm[crop], origin=(10, 20, 30), syn_table=t_syn)
    g = dev.geoms([[10, 20, 30]], [list(big[crop].shape)])
    assert_props_equal(tuple(records_to_dicts(dev.records_numpy(t_syn.export(g)))), tuple(shifted(want[1], (10, 20, 30))), "syn crop")
