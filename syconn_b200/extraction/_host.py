"""Host-side glue shared by the reference-facing shims: NumPy views -> (pointer, shape, strides) for the C ABI,
C records -> the reference's dict structures."""
import ctypes as C

import numpy as np

from .. import _lib


def dense_view(a):
    """Return an array sharing the logical content of ``a`` that is a dense block (any axis permutation,
    positive strides, no padding).  Views such as ``zyx.swapaxes(0, 2)`` are passed through untouched."""
    if a.size == 0:
        return a
    isz = a.itemsize
    order = sorted(range(a.ndim), key=lambda i: a.strides[i])
    expect = isz
    ok = True
    for i in order:
        if a.shape[i] == 1:
            continue
        if a.strides[i] != expect:
            ok = False
            break
        expect *= a.shape[i]
    return a if ok else np.ascontiguousarray(a)


def estrides(a):
    return [s // a.itemsize for s in a.strides]


def check_label_array(a, name, ndim):
    if not isinstance(a, np.ndarray):
        a = np.asarray(a)
    if a.dtype not in (np.uint32, np.uint64):
        # same exception type and wording family as the Cython fused-type buffer check
        raise ValueError(f"Buffer dtype mismatch, expected 'uint64_t' or 'uint32_t' but got '{a.dtype}' ({name})")
    if a.ndim != ndim:
        raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    return a


def records_to_dicts(rec):
    """records -> (rep_coords, bounding_box, sizes): keys int, coords lists of int, bbox [[min],[max+1]]
    (find_object_properties_C.pyx:42-49)."""
    ids = rec["id"].tolist()
    rc = dict(zip(ids, rec["rep"].tolist()))
    bb = dict(zip(ids, np.stack([rec["bb_min"], rec["bb_max"]], axis=1).tolist())) if len(ids) else {}
    sz = dict(zip(ids, rec["count"].astype(np.int64).tolist()))
    return rc, bb, sz


def pairs_to_dict(pairs):
    """pairs -> {sub_id: {cell_id: count}} (find_object_properties_C.pyx:163-174)."""
    out = {}
    for s, c, n in zip(pairs["sub_id"].tolist(), pairs["cell_id"].tolist(), pairs["count"].tolist()):
        d = out.get(s)
        if d is None:
            out[s] = {c: n}
        else:
            d[c] = n
    return out


def syntype_to_dicts(cs_records, v, shape, offset):
    """contact-site records + synaptic voxel tuples (``syk_synvox_t``) of one ``extract_cs_syntype`` call -> the reference's
    return value ``[rc, bb, size], [rc_syn, bb_syn, size_syn], cs_asym, cs_sym, voxels_syn``
    (block_processing_C.pyx:119-158): per id the synaptic voxels in scan order with ``offset`` added, their bounding box,
    first voxel and count, and the number of them with ``asym_mask == 1`` / ``sym_mask == 1``."""
    cs_props = records_to_dicts(cs_records)
    rc_syn, bb_syn, size_syn, cs_asym, cs_sym, voxels = {}, {}, {}, {}, {}, {}
    if len(v):
        v = v[np.lexsort((v["lin"], v["id"]))]                   # per id, reference scan order
        sy, sz = int(shape[1]), int(shape[2])
        lin = v["lin"].astype(np.int64)
        xyz = np.stack([lin // (sy * sz), (lin // sz) % sy, lin % sz], axis=1)
        ids, start = np.unique(v["id"], return_index=True)
        end = np.append(start[1:], len(v))
        mn = np.minimum.reduceat(xyz, start, axis=0)
        mx = np.maximum.reduceat(xyz, start, axis=0) + 1
        n_asym = np.add.reduceat((v["flags"] & 1).astype(np.int64), start)
        n_sym = np.add.reduceat(((v["flags"] >> 1) & 1).astype(np.int64), start)
        off = np.array([int(offset[0]), int(offset[1]), int(offset[2])], np.int64)
        shifted = (xyz + off).tolist()
        for i, k in enumerate(ids.tolist()):
            s, e = int(start[i]), int(end[i])
            rc_syn[k] = xyz[s].tolist()
            bb_syn[k] = [mn[i].tolist(), mx[i].tolist()]
            size_syn[k] = e - s
            voxels[k] = shifted[s:e]
            if n_asym[i]:
                cs_asym[k] = int(n_asym[i])
            if n_sym[i]:
                cs_sym[k] = int(n_sym[i])
    return [cs_props[0], cs_props[1], cs_props[2]], [rc_syn, bb_syn, size_syn], cs_asym, cs_sym, voxels


def find_object_properties_records(chunk, capacity_hint=0):
    chunk = dense_view(check_label_array(chunk, "chunk", 3))
    L = _lib.load()
    out = C.c_void_p()
    n = C.c_uint64()
    _lib.check(L.syk_find_object_properties_host(chunk.ctypes.data, chunk.itemsize, _lib.i64(chunk.shape),
                                                 _lib.i64(estrides(chunk)), capacity_hint, C.byref(out), C.byref(n)))
    return _lib.take_array(out.value, n.value, _lib.RECORD_DTYPE)


def map_subcell_records(ch, subcell_chs, props_too=True, capacity_hint=0):
    ch = check_label_array(ch, "ch", 3)
    subcell_chs = check_label_array(subcell_chs, "subcell_chs", 4)
    if ch.dtype != subcell_chs.dtype:
        raise ValueError("Buffer dtype mismatch: ch and subcell_chs must share the dtype")
    sh = ch.shape
    for ii in range(subcell_chs.shape[0]):
        s = subcell_chs[ii].shape
        assert (s[0] == sh[0]) & (s[1] == sh[1]) & (s[2] == sh[2]), \
            "Segmentation of cells and subcellular structures must have same shape. {} {} {} {} {} {}".format(
                s[0], s[1], s[2], sh[0], sh[1], sh[2])
    n_sub = subcell_chs.shape[0]
    L = _lib.load()
    cell_rec = np.empty(0, _lib.RECORD_DTYPE)
    sub_recs, pair_recs = [], []
    ch = dense_view(ch)
    # the kernel takes up to 4 organelle channels per launch
    for c0 in range(0, max(n_sub, 1), 4):
        grp = dense_view(subcell_chs[c0:c0 + 4])
        k = grp.shape[0]
        cell_out, n_cell = C.c_void_p(), C.c_uint64()
        sub_out = (C.c_void_p * max(k, 1))()
        n_sub_out = (C.c_uint64 * max(k, 1))()
        pairs_out = (C.c_void_p * max(k, 1))()
        n_pairs_out = (C.c_uint64 * max(k, 1))()
        do_cell = props_too and c0 == 0
        # props of the cell channel are only needed once; later groups run with the cell table switched off by
        # asking for the organelle props via a second call in "props" mode and dropping the cell records
        _lib.check(L.syk_map_subcell_extract_props_host(
            ch.ctypes.data, _lib.i64(estrides(ch)), grp.ctypes.data if k else None,
            _lib.i64(estrides(grp) if k else [0, 0, 0, 0]), k, ch.itemsize, _lib.i64(sh), 1 if props_too else 0,
            capacity_hint, C.byref(cell_out), C.byref(n_cell), sub_out, n_sub_out, pairs_out, n_pairs_out))
        rec = _lib.take_array(cell_out.value, n_cell.value, _lib.RECORD_DTYPE)
        if do_cell:
            cell_rec = rec
        for c in range(k):
            sub_recs.append(_lib.take_array(sub_out[c], n_sub_out[c], _lib.RECORD_DTYPE))
            pair_recs.append(_lib.take_array(pairs_out[c], n_pairs_out[c], _lib.PAIR_DTYPE))
    return cell_rec, sub_recs, pair_recs
