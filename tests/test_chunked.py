"""Chunk plan, record/pair folding (reference merge semantics) and the hash-owner exchange.
CPU: plan + folds against the oracle's merge_prop_dicts / merge_map_dicts; world_size-2 gloo all-to-all.
GPU: the whole chunked pipeline on a small volume against the oracle run chunk by chunk."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from syconn_b200._lib import PAIR_DTYPE, RECORD_DTYPE
from syconn_b200.chunked import ChunkPlan, cs_halo_geometry, exchange_logs, owner_of, reduce_pairs, reduce_records
from syconn_b200.synth import synth_labels


def test_chunk_plan_partition():
    plan = ChunkPlan((100, 64, 70), (32, 32, 32))
    assert len(plan) == 4 * 2 * 3
    assert plan.offsets[0] == (0, 0, 0) and plan.offsets[1] == (0, 0, 32) and plan.sizes[-1] == (4, 32, 6)
    for world in (1, 2, 3, 5, 8):
        parts = [plan.chunks_of_rank(r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(len(plan)))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        assert all(p == list(range(p[0], p[0] + len(p))) for p in parts if p)
    lo, ls, oo, os_ = cs_halo_geometry((512, 0, 1024), (512, 512, 512))
    assert lo == [500, -12, 1015] and ls == [536, 536, 530] and oo == [506, -6, 1018] and os_ == [524, 524, 524]


def _oracle_chunk_log(vol, plan, seqs):
    """per-(id, chunk) records from the oracle, chunk by chunk, with global coordinates."""
    rows = []
    acc = oracle.new_prop_acc()
    for s in seqs:
        off, size = plan.offsets[s], plan.sizes[s]
        sl = tuple(slice(off[i], off[i] + size[i]) for i in range(3))
        ids, sizes, bbox, rep = oracle.find_object_properties_arrays(vol[sl])
        r = np.zeros(len(ids), RECORD_DTYPE)
        r["id"], r["count"], r["chunk_seq"] = ids, sizes, s
        r["bb_min"], r["bb_max"], r["rep"] = bbox[:, 0] + off, bbox[:, 1] + off, rep + off
        rows.append(r)
        oracle.merge_prop_dicts([acc, list(oracle.find_object_properties(vol[sl]))], offset=np.array(off))
    return np.concatenate(rows), acc


def test_reduce_records_matches_reference_merge():
    """reduce_records == merge_prop_dicts (sd_proc.py:1248-1273) + the writers' final reduction (:939-945)."""
    vol = synth_labels((48, 40, 36), pitch=(14, 12, 9), seed=4)
    plan = ChunkPlan(vol.shape, (16, 16, 16))
    log, acc = _oracle_chunk_log(vol, plan, range(len(plan)))
    rng = np.random.default_rng(0)
    red = reduce_records(log[rng.permutation(len(log))])
    rc, bb, sz = acc
    assert red["id"].tolist() == sorted(sz)
    for i, k in enumerate(red["id"].tolist()):
        assert red["size"][i] == sz[k]
        assert red["rep_coord"][i].tolist() == list(rc[k])          # last chunk's first voxel
        bbs = np.array(bb[k])                                        # per-chunk boxes, chunk order
        assert np.array_equal(red["bbs"][i], bbs)
        assert np.array_equal(red["bounding_box"][i], [bbs[:, 0].min(axis=0), bbs[:, 1].max(axis=0)])


def test_reduce_pairs_matches_reference_merge():
    cell = synth_labels((40, 32, 24), pitch=(12, 10, 8), seed=5)
    sub = synth_labels((40, 32, 24), pitch=(5, 5, 4), seed=5, kind=1, density16=5)
    plan = ChunkPlan(cell.shape, (16, 16, 16))
    tot, rows = {}, []
    for s in range(len(plan)):
        off, size = plan.offsets[s], plan.sizes[s]
        sl = tuple(slice(off[i], off[i] + size[i]) for i in range(3))
        md = oracle.map_subcell_C(cell[sl], sub[sl][None])[0]
        oracle.merge_map_dicts([tot, {k: dict(v) for k, v in md.items()}])
        for a, d in md.items():
            for b, n in d.items():
                rows.append((a, b, n, 0))
    red = reduce_pairs(np.array(rows, dtype=PAIR_DTYPE))
    want = sorted((a, b, n) for a, d in tot.items() for b, n in d.items())
    assert list(zip(red["sub_id"].tolist(), red["cell_id"].tolist(), red["count"].tolist())) == want


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_logs(rank, world):
    """Per-(id, chunk) record log and pair log of this rank's chunks (oracle, chunk by chunk) on a 3x2x2 chunk grid."""
    cell = synth_labels((48, 32, 32), pitch=(14, 12, 9), seed=4)
    sub = synth_labels((48, 32, 32), pitch=(5, 5, 4), seed=4, kind=1, density16=5)
    plan = ChunkPlan(cell.shape, (16, 16, 16))
    seqs = plan.chunks_of_rank(rank, world)
    log, _ = _oracle_chunk_log(cell, plan, seqs)
    rows = []
    for s in seqs:
        off, size = plan.offsets[s], plan.sizes[s]
        sl = tuple(slice(off[i], off[i] + size[i]) for i in range(3))
        for a, d in oracle.map_subcell_C(cell[sl], sub[sl][None])[0].items():
            rows += [(a, b, n, 0) for b, n in d.items()]
    return log, np.array(rows, dtype=PAIR_DTYPE)


def _bucket(arr, key, world):
    """host stand-in of syk_records_bucket / syk_pairs_bucket: rows grouped by owner rank + per-owner counts"""
    o = owner_of(arr[key], world)
    order = np.argsort(o, kind="stable")
    mat = arr[order].view(np.int64).reshape(len(arr), -1)
    return torch.from_numpy(mat.copy()), torch.from_numpy(np.bincount(o, minlength=world).astype(np.int64))


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        log, pairs = _sharded_logs(rank, world)
        b0, c0 = _bucket(log, "id", world)
        b1, c1 = _bucket(pairs, "sub_id", world)
        empty = torch.zeros((0, 8), dtype=torch.int64)                     # an empty log must pass through too
        got = exchange_logs([b0, b1, empty], [c0, c1, torch.zeros(world, dtype=torch.int64)], world)
        recs = got[0].numpy().view(RECORD_DTYPE).reshape(-1)
        prs = got[1].numpy().view(PAIR_DTYPE).reshape(-1)
        ok = got[2].shape == (0, 8) and bool(np.all(owner_of(recs["id"], world) == rank)) and \
            bool(np.all(owner_of(prs["sub_id"], world) == rank))
        q.put((rank, ok, recs, prs))
    finally:
        dist.destroy_process_group()


def test_exchange_logs_gloo_world2():
    """The product's exchange (chunked.exchange_logs, what ExtractionPipeline.finish calls) under gloo with 2 ranks:
    owner fold of the exchanged logs == single-rank fold of all chunks (sd_proc.py:511-556, 1248-1322)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=180) for _ in ps], key=lambda r: r[0])
    [p.join(timeout=60) for p in ps]
    assert [r[:2] for r in res] == [(0, True), (1, True)]
    log_all, pairs_all = _sharded_logs(0, 1)
    want, want_p = reduce_records(log_all), reduce_pairs(pairs_all)
    parts = [reduce_records(r[2]) for r in res]
    pparts = [reduce_pairs(r[3]) for r in res]
    ids = np.concatenate([p["id"] for p in parts])
    o = np.argsort(ids)
    assert np.array_equal(ids[o], want["id"]) and len(set(ids.tolist())) == len(ids)
    for f in ("size", "bounding_box", "rep_coord"):
        assert np.array_equal(np.concatenate([p[f] for p in parts])[o], want[f]), f
    bbs = [b for p in parts for b in p["bbs"]]
    assert all(np.array_equal(bbs[i], w) for i, w in zip(o, want["bbs"]))
    key = np.concatenate([np.stack([p["sub_id"], p["cell_id"], p["count"].astype(np.uint64)], 1) for p in pparts])
    key = key[np.lexsort((key[:, 1], key[:, 0]))]
    assert np.array_equal(key, np.stack([want_p["sub_id"], want_p["cell_id"], want_p["count"].astype(np.uint64)], 1))


def test_owner_of_matches_device_hash_constants():
    """owner_of is a pure function of the id with every rank in range; the GPU test checks it against syk_records_bucket."""
    ids = np.arange(1, 5000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for w in (1, 2, 3, 8):
        o = owner_of(ids, w)
        assert o.min() >= 0 and o.max() < w and (w == 1 or len(np.unique(o)) == w)


@pytest.mark.gpu
def test_pipeline_small_volume_vs_oracle():
    """All three stages over a 2x2x2 chunk grid on one GPU: device logs -> owner fold == oracle chunk by chunk."""
    from syconn_b200 import device as dev
    from syconn_b200._lib import GEOM_DTYPE
    from syconn_b200.chunked import ExtractionPipeline
    E, st, nsub = 48, (5, 5, 3), 2
    plan = ChunkPlan((2 * E, 2 * E, 2 * E), (E, E, E))
    pipe = ExtractionPipeline(nsub, st, chunk_table_capacity=1 << 12, log_capacity=1 << 14, pair_log_capacity=1 << 14)
    geoms = {"cell": np.zeros(len(plan), GEOM_DTYPE), "cs": np.zeros(len(plan), GEOM_DTYPE)}
    want_cs, want_cell, want_sub, want_pairs = [], [], [[] for _ in range(nsub)], [{} for _ in range(nsub)]
    for rep in range(2):  # second pass checks reset()
        pipe.reset()
        for s in range(len(plan)):
            off, size = plan.offsets[s], plan.sizes[s]
            lo, ls, oo, os_ = cs_halo_geometry(off, size, st)
            geoms["cell"][s], geoms["cs"][s] = (off, size), (off, size)
            cell = dev.synth_labels(size, off, (13, 11, 7), 3, 2, 0, order="F")
            subs = torch.stack([dev.synth_labels(size, off, (6, 5, 4), 3, 2, 1 + c, 3) for c in range(nsub)])
            halo = dev.synth_labels(ls, lo, (13, 11, 7), 3, 2, 0, dtype=torch.int32, order="F")
            pipe.process_chunk(s, off, cell, subs, halo)
            if rep == 0:
                hn = synth_labels(ls, lo, (13, 11, 7), 3, 2, 0, dtype=np.uint32)
                ov = max(x // 2 for x in st)       # the worker merges the props of the CROPPED contacts (:465-486)
                cs = np.ascontiguousarray(oracle.detect_cs(hn, st)[ov:-ov, ov:-ov, ov:-ov])
                assert cs.shape == tuple(size)
                for arrs, origin, dst in ((oracle.find_object_properties_arrays(cs), off, want_cs),
                                          (oracle.find_object_properties_arrays(cell.cpu().numpy().view(np.uint64)), off, want_cell)):
                    dst.append((s, origin, arrs))
                c_, s_, p_ = oracle.map_subcell_extract_props_arrays(cell.cpu().numpy().view(np.uint64),
                                                                     subs.cpu().numpy().view(np.uint64))
                for c in range(nsub):
                    want_sub[c].append((s, off, s_[c]))
                    for a, b, n in zip(*[x.tolist() for x in p_[c]]):
                        want_pairs[c][(a, b)] = want_pairs[c].get((a, b), 0) + n
        owned, owned_pairs = pipe.finish()
        final, final_pairs = pipe.reduce_on_device(owned, owned_pairs, geoms)

    def want_log(lst):
        rows = []
        for s, origin, (ids, sizes, bbox, rep) in lst:
            r = np.zeros(len(ids), RECORD_DTYPE)
            r["id"], r["count"], r["chunk_seq"] = ids, sizes, s
            r["bb_min"], r["bb_max"], r["rep"] = bbox[:, 0] + origin, bbox[:, 1] + origin, rep + origin
            rows.append(r)
        return np.concatenate(rows)

    for kind, lst in (("cs", want_cs), ("cell", want_cell), ("sub0", want_sub[0]), ("sub1", want_sub[1])):
        w = reduce_records(want_log(lst))
        glog = dev.records_numpy(owned[kind])
        g = reduce_records(glog)
        assert len(glog) == sum(len(x[2][0]) for x in lst), kind
        for f in ("id", "size", "bounding_box", "rep_coord"):
            assert np.array_equal(g[f], w[f]), (kind, f)
        assert all(np.array_equal(a, b) for a, b in zip(g["bbs"], w["bbs"])), kind
        fin = dev.records_numpy(final[kind])
        o = np.argsort(fin["id"])
        assert np.array_equal(fin["id"][o], w["id"]) and np.array_equal(fin["count"][o].astype(np.int64), w["size"])
        assert np.array_equal(fin["bb_min"][o], w["bounding_box"][:, 0]) and np.array_equal(fin["bb_max"][o], w["bounding_box"][:, 1])
        assert np.array_equal(fin["rep"][o], w["rep_coord"]), kind
    for c in range(nsub):
        g = reduce_pairs(dev.pairs_numpy(owned_pairs[c]))
        got = dict(zip(zip(g["sub_id"].tolist(), g["cell_id"].tolist()), g["count"].tolist()))
        assert got == want_pairs[c]
        f = dev.pairs_numpy(final_pairs[c])
        assert dict(zip(zip(f["sub_id"].tolist(), f["cell_id"].tolist()), f["count"].tolist())) == want_pairs[c]
    # bucketing: every record lands in exactly one owner bucket, same owner for equal ids
    b, counts = dev.bucket_records(owned["cell"], 4)
    bn, cn = dev.records_numpy(b), counts.tolist()
    assert sum(cn) == owned["cell"].shape[0]
    pos = 0
    for o, n in enumerate(cn):
        assert np.all(owner_of(bn["id"][pos:pos + n], 4) == o)      # host mirror == device hash
        pos += n
    assert sorted(bn["id"].tolist()) == sorted(dev.records_numpy(owned["cell"])["id"].tolist())


@pytest.mark.gpu
def test_pipeline_reports_table_overflow():
    """A chunk with more ids than the per-chunk table holds must fail loudly, never drop records silently."""
    from syconn_b200 import _lib, device as dev
    from syconn_b200.chunked import ExtractionPipeline
    pipe = ExtractionPipeline(0, (3, 3, 3), chunk_table_capacity=1024, log_capacity=1 << 14, pair_log_capacity=16)
    cell = (torch.arange(24 * 24 * 24, dtype=torch.int64, device="cuda") + 1).reshape(24, 24, 24)   # 13824 ids
    subs = torch.empty((0, 24, 24, 24), dtype=torch.int64, device="cuda")
    pipe.reset()
    pipe.process_chunk(0, (0, 0, 0), cell, subs, None)
    with pytest.raises(_lib.SykError):
        pipe.finish()


def test_sd_proc_merge_mirrors():
    """syconn_b200.proc.sd_proc: the dict-level merge helpers behave like the reference's (restated in the oracle), and the
    array-level reductions of the device pipeline give the same merged result."""
    import copy
    from collections import defaultdict
    from oracle import oracle
    from syconn_b200.chunked import reduce_pairs, reduce_records
    from syconn_b200.proc import sd_proc
    from syconn_b200.synth import synth_labels
    vol = synth_labels((48, 40, 36), pitch=(9, 8, 7), warp_amp=2, seed=4)
    subs = np.stack([synth_labels((48, 40, 36), pitch=(5, 4, 4), seed=4, kind=1 + c, density16=5) for c in range(2)])
    chunks = [((0, 0, 0), (24, 40, 36)), ((24, 0, 0), (24, 40, 36))]
    per_chunk, per_chunk_maps = [], []
    for off, size in chunks:
        sl = tuple(slice(o, o + s) for o, s in zip(off, size))
        cp, sp, md = oracle.map_subcell_extract_props(vol[sl], subs[(slice(None),) + sl])
        per_chunk.append((off, cp))
        per_chunk_maps.append(md)
    # reference-style worker loop: merge_prop_dicts([acc, chunk_props], offset) per chunk
    acc_ours, acc_ref = sd_proc.new_prop_dicts(), [{}, defaultdict(list), {}]
    for off, cp in per_chunk:
        sd_proc.merge_prop_dicts([acc_ours, copy.deepcopy(cp)], np.array(off))
        oracle.merge_prop_dicts([acc_ref, copy.deepcopy(cp)], np.array(off))
    assert acc_ours[0] == acc_ref[0] and dict(acc_ours[1]) == dict(acc_ref[1]) and acc_ours[2] == acc_ref[2]
    maps_ours, maps_ref = copy.deepcopy(per_chunk_maps), copy.deepcopy(per_chunk_maps)
    for c in range(2):
        a, b = [m[c] for m in maps_ours], [m[c] for m in maps_ref]
        sd_proc.merge_map_dicts(a)
        oracle.merge_map_dicts(b)
        assert a[0] == b[0]
        inv = sd_proc.invert_mdc(a[0])
        assert all(inv[cid][sid] == n for sid, d in a[0].items() for cid, n in d.items())
        ratios = copy.deepcopy(a[0])
        sd_proc.convert_nvox2ratio_mapdict(ratios)
        assert all(abs(sum(d.values()) - 1.0) < 1e-12 for d in ratios.values())
        # array-level reduction of the device pipeline == dict-level merge
        log = np.concatenate([sd_proc.map_dict_to_pairs(m[c]) for m in per_chunk_maps])
        assert sd_proc.reduced_to_map_dict(reduce_pairs(log)) == a[0]
    recs = []
    for seq, (off, cp) in enumerate(per_chunk):
        r = sd_proc.prop_dicts_to_records(cp, chunk_seq=seq)
        for f in ("bb_min", "bb_max", "rep"):
            r[f] += np.array(off, np.int32)
        recs.append(r)
    red = sd_proc.reduced_to_prop_dicts(reduce_records(np.concatenate(recs)))
    assert red[0] == acc_ref[0] and dict(red[1]) == dict(acc_ref[1]) and red[2] == acc_ref[2]


def test_dataset_analysis_cache(tmp_path):
    """f3 (numpy column caches only): arrays written by write_dataset_analysis_cache read back with the dtypes of
    dataset_analysis (sd_proc.py:244-251); mapping ids / ratios follow the writers' dict logic (:1064-1084)."""
    from oracle import oracle
    from syconn_b200.chunked import reduce_pairs, reduce_records
    from syconn_b200.proc import sd_proc
    from syconn_b200.synth import synth_labels
    vol = synth_labels((40, 36, 32), pitch=(9, 8, 7), warp_amp=2, seed=7)
    subs = np.stack([synth_labels((40, 36, 32), pitch=(5, 4, 4), seed=7, kind=1, density16=6)])
    cp, sp, md = oracle.map_subcell_extract_props(vol, subs)
    red_cell = reduce_records(sd_proc.prop_dicts_to_records(cp))
    red_org = reduce_records(sd_proc.prop_dicts_to_records([sp[0][0], sp[1][0], sp[2][0]]))
    pairs = reduce_pairs(sd_proc.map_dict_to_pairs(md[0]))
    keep = red_org["size"] >= 4                                   # organelle dataset after a size threshold
    m_ids, m_rat = sd_proc.cell_mapping_attributes(red_cell["id"], pairs, red_org["id"][keep], red_org["size"][keep])
    # the writers' dict logic: invert -> drop unknown organelles -> normalise by the organelle size -> invert back
    size_dc = dict(zip(red_org["id"][keep].tolist(), red_org["size"][keep].tolist()))
    want = {}
    for sid, d in md[0].items():
        if sid in size_dc:
            for cid, n in d.items():
                want.setdefault(cid, {})[sid] = n / size_dc[sid]
    for cid, ids, rat in zip(red_cell["id"].tolist(), m_ids, m_rat):
        assert dict(zip(ids, rat)) == want.get(cid, {})
    paths = sd_proc.write_dataset_analysis_cache(str(tmp_path / "sv_0"), red_cell,
                                                 extra={"mapping_mi_id": m_ids, "mapping_mi_ratio": m_rat})
    assert sorted(os.path.basename(p) for p in paths) == ["bounding_boxs.npy", "ids.npy", "mapping_mi_ids.npy",
                                                          "mapping_mi_ratios.npy", "rep_coords.npy", "sizes.npy"]
    ids = np.load(tmp_path / "sv_0" / "ids.npy")
    assert ids.dtype == np.uint64 and np.array_equal(ids, red_cell["id"])
    assert np.load(tmp_path / "sv_0" / "sizes.npy").dtype == np.int64
    bb = np.load(tmp_path / "sv_0" / "bounding_boxs.npy")
    assert bb.dtype == np.int32 and bb.shape == (len(ids), 2, 3)
    rc = np.load(tmp_path / "sv_0" / "rep_coords.npy")
    assert rc.dtype == np.int32 and rc.shape == (len(ids), 3)
    back = np.load(tmp_path / "sv_0" / "mapping_mi_ids.npy", allow_pickle=True)
    assert back.dtype == object and list(back[0]) == list(m_ids[0])
    for k, i in zip(ids.tolist(), range(len(ids))):
        assert cp[2][k] == int(np.load(tmp_path / "sv_0" / "sizes.npy")[i]) if i < 3 else True


@pytest.mark.gpu
def test_pipeline_contact_site_worker_body_vs_oracle():
    """ExtractionPipeline(with_syn=True): detect_cs -> closing / dilation -> crop -> extract_cs_syntype per chunk on the device,
    merged across a 2x2x2 chunk grid == the same composition of the oracle's restatements per chunk
    (cs_extraction_steps.py:381-486) merged with merge_prop_dicts (:481-486).  The merged sizes do not double-count the
    overlap between neighbouring chunks."""
    from helpers import reference_contact_site_chunk
    from syconn_b200 import device as dev
    from syconn_b200.chunked import ExtractionPipeline
    E, st, dil = 40, (7, 7, 3), 2
    so, ov = [s // 2 for s in st], max(s // 2 for s in st)
    plan = ChunkPlan((2 * E, 2 * E, 2 * E), (E, E, E))
    rng = np.random.default_rng(9)
    G = 2 * E + 2 * ov                                                     # global masks incl. the outer overlap
    sj_g = ((rng.random((G, G, G)) < 0.4) * rng.integers(1, 3, size=(G, G, G))).astype(np.uint8)
    asym_g, sym_g = rng.integers(0, 3, size=(G, G, G)).astype(np.uint8), rng.integers(0, 3, size=(G, G, G)).astype(np.uint8)
    pipe = ExtractionPipeline(0, st, chunk_table_capacity=1 << 14, log_capacity=1 << 15, pair_log_capacity=16, with_syn=True,
                              cs_dilation=dil)
    pipe.reset()
    acc_cs, acc_syn, asym_tot, sym_tot, vox_tot = oracle.new_prop_acc(), oracle.new_prop_acc(), {}, {}, {}
    empty = torch.empty((0, E, E, E), dtype=torch.int64, device="cuda")
    for s in range(len(plan)):
        off, size = plan.offsets[s], plan.sizes[s]
        lo, ls, oo, os_ = cs_halo_geometry(off, size, st)
        halo = dev.synth_labels(ls, lo, (13, 11, 7), 3, 2, 0, dtype=torch.int32, order="F")
        cell = dev.synth_labels(size, off, (13, 11, 7), 3, 2, 0, order="F")
        msl = tuple(slice(off[i], off[i] + size[i] + 2 * ov) for i in range(3))        # mask block at offset - overlap
        masks = [torch.from_numpy(np.ascontiguousarray(m[msl])).cuda() for m in (sj_g, asym_g, sym_g)]
        pipe.process_chunk(s, off, cell, empty, halo, syn_masks=masks)
        data = halo.cpu().numpy().view(np.uint32)
        want = reference_contact_site_chunk(oracle, data, sj_g[msl], asym_g[msl], sym_g[msl], np.array(off) - ov, st, dil)
        # the props are chunk-local; the worker adds offset + overlap when it merges them (cs_extraction_steps.py:481-482)
        oracle.merge_prop_dicts([acc_cs, [dict(want[0][0]), dict(want[0][1]), dict(want[0][2])]], offset=np.array(off))
        oracle.merge_prop_dicts([acc_syn, [dict(want[1][0]), dict(want[1][1]), dict(want[1][2])]], offset=np.array(off))
        for k, n in want[2].items():
            asym_tot[k] = asym_tot.get(k, 0) + n
        for k, n in want[3].items():
            sym_tot[k] = sym_tot.get(k, 0) + n
        for k, v in want[4].items():
            vox_tot.setdefault(k, []).extend([tuple(int(c) for c in x) for x in v])
    owned, _ = pipe.finish()
    for kind, acc in (("cs", acc_cs), ("syn", acc_syn)):
        red = reduce_records(dev.records_numpy(owned[kind]))
        rc, bb, sz = acc
        assert red["id"].tolist() == sorted(sz) and len(sz) > 20, kind
        for i, k in enumerate(red["id"].tolist()):
            assert red["size"][i] == sz[k] and red["rep_coord"][i].tolist() == list(rc[k]), (kind, k)
            assert np.array_equal(red["bbs"][i], np.array(bb[k])), (kind, k)
    # synaptic voxel tuples of all chunks: coordinates, sym / asym counts
    got_vox, got_asym, got_sym = {}, {}, {}
    for seq, off, shape, vox in pipe.syn_voxels:
        v = vox.cpu().numpy().view(_lib_synvox()).reshape(-1)
        lin = v["lin"].astype(np.int64)
        xyz = np.stack([lin // (shape[1] * shape[2]), lin // shape[2] % shape[1], lin % shape[2]], axis=1) + np.array(off)
        for k, c, f in zip(v["id"].tolist(), xyz.tolist(), v["flags"].tolist()):
            got_vox.setdefault(k, []).append(tuple(c))
            got_asym[k] = got_asym.get(k, 0) + (f & 1)
            got_sym[k] = got_sym.get(k, 0) + ((f >> 1) & 1)
    assert {k: sorted(v) for k, v in got_vox.items()} == {k: sorted(v) for k, v in vox_tot.items()}
    assert {k: n for k, n in got_asym.items() if n} == {k: n for k, n in asym_tot.items() if n}
    assert {k: n for k, n in got_sym.items() if n} == {k: n for k, n in sym_tot.items() if n}


def _lib_synvox():
    from syconn_b200._lib import SYNVOX_DTYPE
    return SYNVOX_DTYPE


@pytest.mark.gpu
def test_pipeline_small_object_drop_mapping_inversion_and_voxel_index():
    """min_obj_vx drop inside the worker (sd_proc.py:650-661, :667-680), the device-side mapping inversion / normalisation
    (:1054-1084) and the per-object voxel index (:940-946) against the oracle driven like the reference worker."""
    from syconn_b200 import device as dev
    from syconn_b200._lib import GEOM_DTYPE
    from syconn_b200.chunked import ExtractionPipeline
    E, nsub = 40, 2
    min_vx = {"cell": 208, "sub0": 36, "sub1": 37}       # cuts into the size distributions (cells ~ 210, organelles ~ 36 voxels)
    plan = ChunkPlan((2 * E, 2 * E, E), (E, E, E))
    pipe = ExtractionPipeline(nsub, (3, 3, 3), chunk_table_capacity=1 << 14, log_capacity=1 << 15, pair_log_capacity=1 << 15,
                              min_obj_vx=min_vx)
    geoms = {"cell": np.zeros(len(plan), GEOM_DTYPE), "cs": np.zeros(len(plan), GEOM_DTYPE)}
    pipe.reset()
    acc = {k: oracle.new_prop_acc() for k in ("cell", "sub0", "sub1")}
    maps = [{} for _ in range(nsub)]
    dropped = {k: 0 for k in acc}
    for s in range(len(plan)):
        off, size = plan.offsets[s], plan.sizes[s]
        geoms["cell"][s] = geoms["cs"][s] = (off, size)
        cell = dev.synth_labels(size, off, (7, 6, 5), 2, 2, 0, order="F")
        subs = torch.stack([dev.synth_labels(size, off, (4, 3, 3), 2, 2, 1 + c, 4) for c in range(nsub)])
        pipe.process_chunk(s, off, cell, subs, None)
        cn, sn = cell.cpu().numpy().view(np.uint64), subs.cpu().numpy().view(np.uint64)
        cp, sp, md = oracle.map_subcell_extract_props(cn, sn)

        def faces(a):
            return set(np.unique(np.concatenate([a[0].ravel(), a[-1].ravel(), a[:, 0].ravel(), a[:, -1].ravel(),
                                                 a[:, :, 0].ravel(), a[:, :, -1].ravel()])).tolist())
        props = {"cell": [dict(d) for d in cp], "sub0": [dict(sp[k][0]) for k in range(3)], "sub1": [dict(sp[k][1]) for k in range(3)]}
        for kind, arr in (("cell", cn), ("sub0", sn[0]), ("sub1", sn[1])):
            rc, bb, sz = props[kind]
            for ix in set(sz) - faces(arr):
                if sz[ix] < min_vx[kind]:
                    del rc[ix], bb[ix], sz[ix]
                    dropped[kind] += 1
                    if kind != "cell":
                        md[int(kind[3])].pop(ix, None)
            oracle.merge_prop_dicts([acc[kind], props[kind]], offset=np.array(off))
        for c in range(nsub):
            oracle.merge_map_dicts([maps[c], {k: dict(v) for k, v in md[c].items()}])
    assert all(n > 0 for n in dropped.values()), dropped
    owned, owned_pairs = pipe.finish()
    final, final_pairs = pipe.reduce_on_device(owned, owned_pairs, geoms)
    for kind in acc:
        red = reduce_records(dev.records_numpy(owned[kind]))
        rc, bb, sz = acc[kind]
        assert red["id"].tolist() == sorted(sz), kind
        assert all(red["size"][i] == sz[k] for i, k in enumerate(red["id"].tolist()))
        # per-object voxel index on the device == the per-chunk boxes of merge_prop_dicts, chunk order
        ids, start, boxes = (t.cpu().numpy() for t in ExtractionPipeline.voxel_index(owned[kind]))
        for i, k in enumerate(ids.view(np.uint64).tolist()):
            assert np.array_equal(boxes[start[i]:start[i + 1]], np.array(bb[k])), (kind, k)
    inv = pipe.invert_mapping(final_pairs)
    for c in range(nsub):
        sizes = acc[f"sub{c}"][2]
        want = sorted((cid, sid, n / sizes[sid]) for sid, d in maps[c].items() for cid, n in d.items())
        cid, sid, ratio = (t.cpu().numpy() for t in inv[c])
        got = sorted(zip(cid.view(np.uint64).tolist(), sid.view(np.uint64).tolist(), ratio.tolist()))
        assert got == want and len(want) > 50


@pytest.mark.gpu
def test_device_pipeline_to_storage_files(tmp_path):
    """device pipeline -> reduce_on_device + voxel_index -> write_segmentation_objects -> files read back like the reference
    reads them (oracle/storage_ref.py) == the oracle worker's merged dicts (sd_proc.py:788-1000)."""
    from oracle import storage_ref
    from syconn_b200 import device as dev
    from syconn_b200._lib import GEOM_DTYPE
    from syconn_b200.chunked import ExtractionPipeline
    from syconn_b200.proc import sd_proc
    E = 32
    plan = ChunkPlan((2 * E, 2 * E, E), (E, E, E))
    pipe = ExtractionPipeline(1, (3, 3, 3), chunk_table_capacity=1 << 14, log_capacity=1 << 15, pair_log_capacity=1 << 15)
    geoms = {"cell": np.zeros(len(plan), GEOM_DTYPE), "cs": np.zeros(len(plan), GEOM_DTYPE)}
    pipe.reset()
    acc, maps = oracle.new_prop_acc(), {}
    for s in range(len(plan)):
        off, size = plan.offsets[s], plan.sizes[s]
        geoms["cell"][s] = geoms["cs"][s] = (off, size)
        cell = dev.synth_labels(size, off, (9, 8, 7), 2, 5, 0, order="F")
        subs = torch.stack([dev.synth_labels(size, off, (5, 4, 4), 2, 5, 1, 5)]) * 1009      # spread over storage folders
        pipe.process_chunk(s, off, cell, subs, None)
        cp, sp, md = oracle.map_subcell_extract_props(cell.cpu().numpy().view(np.uint64), subs.cpu().numpy().view(np.uint64))
        oracle.merge_prop_dicts([acc, [sp[0][0], sp[1][0], sp[2][0]]], offset=np.array(off))
        oracle.merge_map_dicts([maps, {k: dict(v) for k, v in md[0].items()}])
    owned, owned_pairs = pipe.finish()
    final, final_pairs = pipe.reduce_on_device(owned, owned_pairs, geoms)
    red = sd_proc.reduced_from_device(dev.records_numpy(final["sub0"]), ExtractionPipeline.voxel_index(owned["sub0"]))
    mapping = sd_proc.reduced_to_map_dict(reduce_pairs(dev.pairs_numpy(final_pairs[0])))
    folders = sd_proc.write_segmentation_objects(str(tmp_path / "mi_0"), red, mapping=mapping, min_obj_vx=1, n_folders_fs=100)
    rc, bb, sz = acc
    seen = set()
    for folder in folders:
        attr = storage_ref.read_attr_dict(os.path.join(folder, "attr_dict.pkl"))
        bbs, sizes, reps, _ = storage_ref.read_voxel_dyn(os.path.join(folder, "voxel.pkl"))
        for k, a in attr.items():
            assert a["size"] == sz[k] == sizes[k] and a["rep_coord"].tolist() == list(rc[k]) == reps[k].tolist()
            assert np.array_equal(bbs[k], np.array(bb[k]))
            assert dict(zip(a["mapping_ids"], a["mapping_ratios"])) == {c: n / sz[k] for c, n in maps.get(k, {}).items()}
            seen.add(k)
    assert seen == set(sz) and len(seen) > 50


def test_seed_cleanup_table_matches_oracle_loop():
    """Host logic of the marker clean-up against the loop of object_extraction_steps.py:325-343 restated on np.unique output."""
    from syconn_b200.extraction.object_extraction_steps import seed_cleanup_table
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(1, 30))
        sizes = np.concatenate([[0], rng.integers(1, 12, n)])
        min_size = int(rng.integers(2, 12))
        vol = np.repeat(np.arange(n + 1), np.concatenate([[5 if trial % 2 else 0], sizes[1:]])).astype(np.uint8)
        vol = vol.reshape(1, 1, -1)
        ixs, cnt = np.unique(vol, return_counts=True)
        m = (ixs != 0) & (cnt < min_size)
        dels, keep = np.sort(ixs[m]), np.sort(ixs[~m])
        want = {int(d): 0 for d in dels}
        ii = len(keep) - 1
        for d in dels:
            if len(keep) == 0 or (d > keep[ii]) or (keep[ii] == 0) or (ii < 0):
                break
            want[int(keep[ii])] = int(d)
            ii -= 1
        table = seed_cleanup_table(sizes, min_size)
        for lab in range(1, n + 1):
            assert table[lab] == want.get(lab, lab), (trial, lab)
