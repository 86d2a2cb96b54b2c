"""Bit-exact parity at PRODUCTION size: the CUDA path against the C oracle, voxel for voxel and record for record.

* detect_cs on one real 536 x 536 x 530 -> 524^3 chunk (BASELINE config 4 geometry, cs_extraction_steps.py:376-391),
  both memory orders, supervoxel pitches 32x32x16 (tier 1 of the fast path) and 16x16x8 (tier 2 + slot recycling).
  The oracle runs in a process pool over x-slabs (its ~2 MVoxels/s per core would need minutes on one core); every
  worker regenerates its haloed slab from the coordinate-hashed generator and compares it with the GPU result in
  shared memory.
* map_subcell_extract_props on one 512^3 cell + 3 organelle chunk (config 3 geometry), both memory orders and pitches:
  cell / organelle records and overlap pairs against the oracle (x-slabs folded with the reference's own semantics:
  first voxel in scan order, min/max boxes, summed sizes).

Follows the expected-value style of the reference's tests/test_segmentation_analysis.py:55-76,126-129 (exact array
equality)."""
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ST = (13, 13, 7)
HALO_SHAPE, HALO_ORIGIN = (536, 536, 530), (500, -12, 1015)
SLAB = 16


def _cs_slab(args):
    """oracle.detect_cs on output planes [x0, x1) of the chunk, compared with the GPU result at `path`."""
    x0, x1, pitch, order, path, oshape = args
    from oracle import oracle
    from syconn_b200.synth import synth_labels
    shape = (x1 - x0 + ST[0] - 1, HALO_SHAPE[1], HALO_SHAPE[2])
    origin = (HALO_ORIGIN[0] + x0, HALO_ORIGIN[1], HALO_ORIGIN[2])
    seg = synth_labels(shape, origin, pitch, 4, 0, 0, dtype=np.uint32, order=order)
    want = np.asarray(oracle.detect_cs(seg, ST))
    got = np.lib.format.open_memmap(path, mode="r")[x0:x1]
    bad = np.argwhere(got != want)
    first = None
    if len(bad):
        i = tuple(bad[0])
        first = (int(bad[0][0]) + x0, int(bad[0][1]), int(bad[0][2]), int(got[i]), int(want[i]))
    return len(bad), first, int(np.count_nonzero(want))


@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("pitch", [(32, 32, 16), (16, 16, 8)])
def test_detect_cs_production_chunk_bit_exact(pitch, order, tmp_path_factory):
    from oracle import oracle
    from syconn_b200 import device as dev
    oracle.build()
    seg = dev.synth_labels(HALO_SHAPE, HALO_ORIGIN, pitch, 4, 0, 0, dtype=torch.int32, order=order)
    out = dev.detect_cs(seg, ST)
    oshape = tuple(HALO_SHAPE[i] - ST[i] + 1 for i in range(3))
    assert tuple(out.shape) == oshape
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path_factory.mktemp("cs"))
    path = os.path.join(shm, f"syk_cs_exact_{os.getpid()}.npy")
    try:
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint64, shape=oshape)
        mm[...] = out.cpu().numpy().view(np.uint64)
        mm.flush()
        del mm
        tasks = [(x0, min(x0 + SLAB, oshape[0]), pitch, order, path, oshape) for x0 in range(0, oshape[0], SLAB)]
        with ProcessPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            res = list(ex.map(_cs_slab, tasks, chunksize=1))
    finally:
        if os.path.exists(path):
            os.remove(path)
    n_bad = sum(r[0] for r in res)
    firsts = [r[1] for r in res if r[1] is not None]
    assert n_bad == 0, f"{n_bad} voxels differ from the oracle; first (x, y, z, got, want): {firsts[:3]}"
    nz = sum(r[2] for r in res)
    assert nz == int(torch.count_nonzero(out)) and nz > 0.1 * np.prod(oshape)   # a real workload, not an empty volume


# ---------------------------------------------------------------------------------------------- map_subcell at 512^3
def _map_slab(args):
    x0, x1, pitch, sub_pitch, order, n_sub = args
    from oracle import oracle
    from syconn_b200.synth import synth_labels
    shape, origin = (x1 - x0, 512, 512), (x0, 0, 0)
    cell = synth_labels(shape, origin, pitch, 4, 0, 0, order=order)
    subs = np.stack([synth_labels(shape, origin, sub_pitch, 4, 0, 1 + c, 1, order=order) for c in range(n_sub)])
    cell_o, subs_o, pairs_o = oracle.map_subcell_extract_props_arrays(cell, subs)
    return x0, cell_o, subs_o, pairs_o


def _fold_objs(parts):
    """[(x0, (ids, sizes, bbox, rep))] of x-slabs -> whole-volume arrays sorted by id: sizes summed, boxes min/max, rep of the
    FIRST slab holding the id (x is the slowest axis of the reference's scan, find_object_properties_C.pyx:30-48)."""
    ids = np.concatenate([p[1][0] for p in parts])
    sizes = np.concatenate([p[1][1] for p in parts]).astype(np.int64)
    off = np.concatenate([np.full(len(p[1][0]), p[0], np.int64) for p in parts])
    bbox = np.concatenate([p[1][2] for p in parts]).astype(np.int64)
    rep = np.concatenate([p[1][3] for p in parts]).astype(np.int64)
    bbox[:, :, 0] += off[:, None]
    rep[:, 0] += off
    o = np.lexsort((off, ids))
    ids, sizes, bbox, rep = ids[o], sizes[o], bbox[o], rep[o]
    uid, start = np.unique(ids, return_index=True)
    return (uid, np.add.reduceat(sizes, start), np.minimum.reduceat(bbox[:, 0], start, axis=0),
            np.maximum.reduceat(bbox[:, 1], start, axis=0), rep[start])


def _fold_pairs(parts):
    sub = np.concatenate([p[0] for p in parts])
    cell = np.concatenate([p[1] for p in parts])
    cnt = np.concatenate([p[2] for p in parts]).astype(np.int64)
    o = np.lexsort((cell, sub))
    sub, cell, cnt = sub[o], cell[o], cnt[o]
    new = np.ones(len(sub), bool)
    new[1:] = (sub[1:] != sub[:-1]) | (cell[1:] != cell[:-1])
    start = np.flatnonzero(new)
    return sub[start], cell[start], np.add.reduceat(cnt, start) if len(start) else cnt[:0]


def _check_records(rec, want, what):
    uid, sizes, bmin, bmax, rep = want
    o = np.argsort(rec["id"])
    assert np.array_equal(rec["id"][o], uid), f"{what}: id sets differ"
    assert np.array_equal(rec["count"][o].astype(np.int64), sizes), f"{what}: sizes differ"
    assert np.array_equal(rec["bb_min"][o], bmin) and np.array_equal(rec["bb_max"][o], bmax), f"{what}: boxes differ"
    assert np.array_equal(rec["rep"][o], rep), f"{what}: rep_coords differ"


@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("pitch", [(32, 32, 16), (16, 16, 8)])
def test_map_subcell_512_chunk_bit_exact(pitch, order):
    from oracle import oracle
    from syconn_b200 import device as dev
    from syconn_b200.extraction import _host
    oracle.build()
    n_sub, S, sub_pitch = 3, 512, (12, 12, 6)
    cell = dev.synth_labels((S, S, S), (0, 0, 0), pitch, 4, 0, 0, order=order)
    if order == "F":
        subs = torch.empty((n_sub, S, S, S), dtype=torch.int64, device="cuda").permute(0, 3, 2, 1)
    else:
        subs = torch.empty((n_sub, S, S, S), dtype=torch.int64, device="cuda")
    for c in range(n_sub):
        dev.synth_labels((S, S, S), (0, 0, 0), sub_pitch, 4, 0, 1 + c, 1, out=subs[c])
    # through the host-buffer C ABI (what the reference-named shims call)
    cell_rec, sub_recs, pair_recs = _host.map_subcell_records(cell.cpu().numpy().view(np.uint64), subs.cpu().numpy().view(np.uint64))
    ncpu = os.cpu_count() or 1
    slab = max(8, S // max(1, min(ncpu, 32)))
    tasks = [(x0, min(x0 + slab, S), pitch, sub_pitch, order, n_sub) for x0 in range(0, S, slab)]
    with ProcessPoolExecutor(max_workers=ncpu) as ex:
        res = sorted(ex.map(_map_slab, tasks, chunksize=1), key=lambda r: r[0])
    _check_records(cell_rec, _fold_objs([(r[0], r[1]) for r in res]), "cell")
    assert len(cell_rec) > 1000
    for c in range(n_sub):
        _check_records(sub_recs[c], _fold_objs([(r[0], r[2][c]) for r in res]), f"organelle {c}")
        ws, wc, wn = _fold_pairs([r[3][c] for r in res])
        p = pair_recs[c]
        o = np.lexsort((p["cell_id"], p["sub_id"]))
        assert np.array_equal(p["sub_id"][o], ws) and np.array_equal(p["cell_id"][o], wc), f"pairs {c}: keys differ"
        assert np.array_equal(p["count"][o].astype(np.int64), wn), f"pairs {c}: counts differ"
        assert len(ws) > 100
