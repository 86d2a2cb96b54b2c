"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``oracle/syk_oracle.c`` (plain-C CPU restatement of the reference's
hot path).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module; ``syconn_b200`` never does.

Parity status: PINNED.  The restatement is checked against (a) the reference's own
known-answer tests (``/root/reference/tests/test_segmentation_analysis.py:19-52,100-135,162-169``,
re-stated in ``tests/test_oracle_pins.py``) and (b) golden vectors produced by the reference's
compiled Cython/numba code (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

The functions mirror the reference API (same return structures):
  * find_object_properties      -- syconn/extraction/find_object_properties_C.pyx:24-49
  * map_subcell_extract_props   -- syconn/extraction/find_object_properties_C.pyx:112-192
  * map_subcell_C               -- syconn/extraction/find_object_properties_C.pyx:72-109
  * detect_seg_boundaries       -- syconn/extraction/find_object_properties.py:424-455
  * process_block_nonzero       -- syconn/extraction/block_processing_C.pyx:53-75 (+ kernel :21-49)
  * detect_cs                   -- syconn/extraction/find_object_properties.py:458-472
  * merge_prop_dicts / merge_map_dicts -- syconn/proc/sd_proc.py:1248-1273, :1300-1322
  * close_contact_sites         -- syconn/extraction/cs_extraction_steps.py:439-461 (+ scipy.ndimage restated)
"""
import ctypes as C
import os
import subprocess
from collections import defaultdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "syk_oracle.c")
_lib = None

CS_FILTERSIZE = (13, 13, 7)  # syconn/handler/config.yml:148


def build(force=False):
    """Compile the C restatement (gcc only; a few seconds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-o", _SO, _SRC])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        p, i64p = C.c_void_p, C.POINTER(C.c_int64)
        L.orc_detect_seg_boundaries.argtypes = [p, C.c_int, i64p, i64p, p]
        L.orc_process_block_nonzero.argtypes = [p, C.c_int, i64p, p, C.c_int, i64p, i64p,
                                                C.POINTER(C.c_int32), p]
        L.orc_find_object_properties.argtypes = [p, C.c_int, i64p, i64p]
        L.orc_find_object_properties.restype = p
        L.orc_map_subcell_extract_props.argtypes = [p, i64p, p, i64p, C.c_int, C.c_int, i64p, C.c_int]
        L.orc_map_subcell_extract_props.restype = p
        L.orc_result_free.argtypes = [p]
        L.orc_result_nobj.argtypes = [p, C.c_int]
        L.orc_result_nobj.restype = C.c_uint64
        L.orc_result_objs.argtypes = [p, C.c_int, p, p, p, p]
        L.orc_result_npairs.argtypes = [p, C.c_int]
        L.orc_result_npairs.restype = C.c_uint64
        L.orc_result_pairs.argtypes = [p, C.c_int, p, p, p]
        L.orc_extract_cs_syntype.argtypes = [p, C.c_int, i64p, p, i64p, p, i64p, p, i64p, i64p]
        L.orc_extract_cs_syntype.restype = p
        L.orc_syntype_result.argtypes = [p]
        L.orc_syntype_result.restype = p
        L.orc_syntype_nvox.argtypes = [p]
        L.orc_syntype_nvox.restype = C.c_uint64
        L.orc_syntype_voxels.argtypes = [p, p, p, p, p]
        L.orc_syntype_free.argtypes = [p]
        _lib = L
    return _lib


def _i64(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def _estrides(a):
    assert all(s % a.itemsize == 0 for s in a.strides)
    return [s // a.itemsize for s in a.strides]


def _check_label_dtype(a, name="chunk"):
    if a.dtype not in (np.uint32, np.uint64):
        raise ValueError(f"Buffer dtype mismatch, expected 'uint64_t' or 'uint32_t' but got {a.dtype} ({name})")


# ------------------------------------------------------------------------------------------ arrays
def _objs(L, h, tab):
    n = int(L.orc_result_nobj(h, tab))
    ids = np.empty(n, np.uint64)
    sizes = np.empty(n, np.int64)
    bbox = np.empty((n, 2, 3), np.int32)
    rep = np.empty((n, 3), np.int32)
    if n:
        L.orc_result_objs(h, tab, ids.ctypes.data, sizes.ctypes.data, bbox.ctypes.data, rep.ctypes.data)
    return ids, sizes, bbox, rep


def _pairs(L, h, ch):
    n = int(L.orc_result_npairs(h, ch))
    sub = np.empty(n, np.uint64)
    cell = np.empty(n, np.uint64)
    cnt = np.empty(n, np.int64)
    if n:
        L.orc_result_pairs(h, ch, sub.ctypes.data, cell.ctypes.data, cnt.ctypes.data)
    return sub, cell, cnt


def find_object_properties_arrays(chunk):
    """-> (ids u64[N], sizes i64[N], bbox i32[N,2,3], rep i32[N,3]) in first-occurrence order."""
    chunk = np.asarray(chunk)
    _check_label_dtype(chunk)
    assert chunk.ndim == 3
    L = lib()
    h = L.orc_find_object_properties(chunk.ctypes.data, chunk.itemsize, _i64(chunk.shape), _i64(_estrides(chunk)))
    try:
        return _objs(L, h, 0)
    finally:
        L.orc_result_free(h)


def map_subcell_extract_props_arrays(ch, subcell_chs, props_too=True):
    """-> (cell_objs, [objs_c], [pairs_c]) with objs = (ids,sizes,bbox,rep), pairs = (sub,cell,cnt)."""
    ch = np.asarray(ch)
    subcell_chs = np.asarray(subcell_chs)
    _check_label_dtype(ch, "ch")
    if subcell_chs.dtype != ch.dtype:
        raise ValueError("Buffer dtype mismatch (ch and subcell_chs must share the dtype)")
    assert ch.ndim == 3 and subcell_chs.ndim == 4
    sh = ch.shape
    for ii in range(subcell_chs.shape[0]):
        s = subcell_chs[ii].shape
        assert s == sh, ("Segmentation of cells and subcellular structures must have same shape. "
                         "{} {} {} {} {} {}".format(s[0], s[1], s[2], sh[0], sh[1], sh[2]))
    L = lib()
    nsub = subcell_chs.shape[0]
    h = L.orc_map_subcell_extract_props(ch.ctypes.data, _i64(_estrides(ch)), subcell_chs.ctypes.data,
                                        _i64(_estrides(subcell_chs)), ch.itemsize, nsub, _i64(sh),
                                        1 if props_too else 0)
    try:
        cell = _objs(L, h, 0)
        subs = [_objs(L, h, 1 + i) for i in range(nsub)]
        pairs = [_pairs(L, h, i) for i in range(nsub)]
    finally:
        L.orc_result_free(h)
    return cell, subs, pairs


# ------------------------------------------------------------------------------------------ dict API
def _objs_to_dicts(objs):
    ids, sizes, bbox, rep = objs
    ids_l = ids.tolist()
    rc = dict(zip(ids_l, rep.tolist()))
    bb = dict(zip(ids_l, bbox.tolist()))
    sz = dict(zip(ids_l, sizes.tolist()))
    return rc, bb, sz


def _pairs_to_dict(pairs):
    sub, cell, cnt = pairs
    out = {}
    for s, c, n in zip(sub.tolist(), cell.tolist(), cnt.tolist()):
        out.setdefault(s, {})[c] = n
    return out


def find_object_properties(chunk):
    return _objs_to_dicts(find_object_properties_arrays(chunk))


def map_subcell_extract_props(ch, subcell_chs):
    cell, subs, pairs = map_subcell_extract_props_arrays(ch, subcell_chs, True)
    rc, bb, sz = _objs_to_dicts(cell)
    sd = [_objs_to_dicts(s) for s in subs]
    return [rc, bb, sz], [[d[0] for d in sd], [d[1] for d in sd], [d[2] for d in sd]], \
        [_pairs_to_dict(p) for p in pairs]


def map_subcell_C(ch, subcell_chs):
    _, _, pairs = map_subcell_extract_props_arrays(ch, subcell_chs, False)
    return [_pairs_to_dict(p) for p in pairs]


def detect_seg_boundaries(arr):
    arr = np.asarray(arr)
    a = arr
    if a.dtype not in (np.uint32, np.uint64):
        # the numba reference accepts any integer/float array; compare as 64-bit patterns
        a = np.ascontiguousarray(arr).astype(np.int64).view(np.uint64) if arr.dtype.kind in "iu" \
            else np.ascontiguousarray(arr, dtype=np.float64).view(np.uint64)
    out = np.empty(a.shape, np.uint8)
    rc = lib().orc_detect_seg_boundaries(a.ctypes.data, a.itemsize, _i64(a.shape), _i64(_estrides(a)),
                                         out.ctypes.data)
    assert rc == 0
    return out.view(np.bool_)


def process_block_nonzero(edges, arr, stencil1=(7, 7, 3)):
    edges = np.asarray(edges)
    arr = np.asarray(arr)
    if edges.dtype != np.uint32 or arr.dtype != np.uint32:
        raise ValueError("Buffer dtype mismatch, expected 'uint32_t'")
    st = [int(s) for s in stencil1]
    assert (st[0] % 2 + st[1] % 2 + st[2] % 2) == 3
    oshape = tuple(max(0, arr.shape[i] - st[i] + 1) for i in range(3))
    out = np.zeros(oshape, np.uint64)
    rc = lib().orc_process_block_nonzero(edges.ctypes.data, 4, _i64(_estrides(edges)), arr.ctypes.data, 4,
                                         _i64(_estrides(arr)), _i64(arr.shape), (C.c_int32 * 3)(*st),
                                         out.ctypes.data)
    assert rc == 0, rc
    return out


def detect_cs(arr, stencil=CS_FILTERSIZE):
    edges = detect_seg_boundaries(arr).astype(np.uint32, copy=False)
    arr = np.asarray(arr).astype(np.uint32, copy=False)
    return process_block_nonzero(edges, arr, stencil)


def detect_contact_partners(seg_arr, edge_arr, offset):
    """syconn/extraction/find_object_properties.py:371-421 (numba), restated with NumPy per flagged voxel: histogram of
    the window without 0 and the centre id; the winner is the id with the largest count, ties going to the id met first
    in the x, y, z scan of the window (the reference iterates a numba typed.Dict, which keeps insertion order, and
    only replaces the candidate on a strictly larger count, :408-413).  Small volumes only (pure Python loop)."""
    seg = np.asarray(seg_arr)
    edge = np.asarray(edge_arr)
    off = np.asarray(offset)
    nx, ny, nz = seg.shape[:3]
    out = np.zeros((nx + off[0, 0] - off[0, 1], ny + off[1, 0] - off[1, 1], nz + off[2, 0] - off[2, 1], 2), np.uint64)
    for xx in range(-off[0, 0], nx - off[0, 1]):
        for yy in range(-off[1, 0], ny - off[1, 1]):
            for zz in range(-off[2, 0], nz - off[2, 1]):
                if edge[xx, yy, zz] == 0:
                    continue
                c = seg[xx, yy, zz]
                w = seg[xx + off[0, 0]:xx + off[0, 1] + 1, yy + off[1, 0]:yy + off[1, 1] + 1,
                        zz + off[2, 0]:zz + off[2, 1] + 1].reshape(-1)
                w = w[(w != 0) & (w != c)]
                if w.size == 0:
                    continue
                ids, first, cnt = np.unique(w, return_index=True, return_counts=True)
                cand = np.flatnonzero(cnt == cnt.max())
                m = ids[cand[np.argmin(first[cand])]]
                out[xx + off[0, 0], yy + off[1, 0], zz + off[2, 0]] = (m, c) if c > m else (c, m)
    return out


def detect_cs_64bit(arr, stencil=CS_FILTERSIZE):
    """syconn/extraction/find_object_properties.py:347-368."""
    o = np.asarray(stencil) // 2
    return detect_contact_partners(arr, detect_seg_boundaries(arr), np.array([(-o[0], o[0]), (-o[1], o[1]), (-o[2], o[2])]))


def find_object_properties_cs_64bit(cs_seg):
    """syconn/extraction/find_object_properties.py:197-269: nested dicts d[id0][id1] of rep coord (first voxel in x, y, z
    order), bounding box [[min], [max + 1]] and size for every pair with id0 != 0."""
    cs = np.asarray(cs_seg)
    rep, bb, sz = {}, {}, {}
    xs, ys, zs = np.nonzero(cs[..., 0])
    for x, y, z in zip(xs.tolist(), ys.tolist(), zs.tolist()):  # np.nonzero is in C (x, y, z) order
        k0, k1 = int(cs[x, y, z, 0]), int(cs[x, y, z, 1])
        if k1 not in sz.setdefault(k0, {}):
            sz[k0][k1] = 1
            rep.setdefault(k0, {})[k1] = np.array([x, y, z], np.int64)
            bb.setdefault(k0, {})[k1] = np.array([(x, y, z), (x + 1, y + 1, z + 1)], np.int64)
        else:
            sz[k0][k1] += 1
            b = bb[k0][k1]
            b[0] = np.minimum(b[0], (x, y, z))
            b[1] = np.maximum(b[1], (x + 1, y + 1, z + 1))
    return rep, bb, sz


def extract_cs_syntype(cs_seg, syn_mask, asym_mask, sym_mask, offset):
    """syconn/extraction/block_processing_C.pyx:78-158 ("next" row f1) ->
    ([rc, bb, size], [rc_syn, bb_syn, size_syn], cs_asym, cs_sym, voxels_syn)."""
    cs_seg = np.asarray(cs_seg)
    _check_label_dtype(cs_seg, "cs_seg")
    masks = [np.asarray(m) for m in (syn_mask, asym_mask, sym_mask)]
    for m in masks:
        if m.dtype != np.uint8:
            raise ValueError("Buffer dtype mismatch, expected 'uint8_t'")
    L = lib()
    h = L.orc_extract_cs_syntype(cs_seg.ctypes.data, cs_seg.itemsize, _i64(_estrides(cs_seg)), masks[0].ctypes.data,
                                 _i64(_estrides(masks[0])), masks[1].ctypes.data, _i64(_estrides(masks[1])),
                                 masks[2].ctypes.data, _i64(_estrides(masks[2])), _i64(cs_seg.shape))
    try:
        res = L.orc_syntype_result(h)
        cs_p, syn_p = _objs_to_dicts(_objs(L, res, 0)), _objs_to_dicts(_objs(L, res, 1))
        n = int(L.orc_syntype_nvox(h))
        key, xyz = np.empty(n, np.uint64), np.empty((n, 3), np.int32)
        asym, sym = np.empty(n, np.uint8), np.empty(n, np.uint8)
        if n:
            L.orc_syntype_voxels(h, key.ctypes.data, xyz.ctypes.data, asym.ctypes.data, sym.ctypes.data)
    finally:
        L.orc_syntype_free(h)
    off = [int(o) for o in offset]
    vox, cs_asym, cs_sym = {}, {}, {}
    for k, c, a, s_ in zip(key.tolist(), xyz.tolist(), asym.tolist(), sym.tolist()):
        vox.setdefault(k, []).append([c[0] + off[0], c[1] + off[1], c[2] + off[2]])
        if a:
            cs_asym[k] = cs_asym.get(k, 0) + 1
        if s_:
            cs_sym[k] = cs_sym.get(k, 0) + 1
    return [cs_p[0], cs_p[1], cs_p[2]], [syn_p[0], syn_p[1], syn_p[2]], cs_asym, cs_sym, vox


# ------------------------------------------------------------------------------------------ merges (a6)
def _binary_step(mask, erode):
    """One iteration of scipy.ndimage.binary_dilation / binary_erosion with the default structure
    (generate_binary_structure(3, 1): the 6-neighbourhood cross) and border_value = 0: a voxel of the result is the
    OR (dilation) / AND (erosion) of itself and its six face neighbours, neighbours outside the array counting as 0."""
    out = mask.copy()
    for ax in range(3):
        for sh in (1, -1):
            nb = np.zeros_like(mask)
            src = [slice(None)] * 3
            dst = [slice(None)] * 3
            if sh == 1:
                src[ax], dst[ax] = slice(0, -1), slice(1, None)
            else:
                src[ax], dst[ax] = slice(1, None), slice(0, -1)
            nb[tuple(dst)] = mask[tuple(src)]
            out = (out & nb) if erode else (out | nb)
    return out


def binary_closing_dilation(mask, n_closings, n_dilations):
    """scipy.ndimage.binary_closing(mask, iterations=n_closings) followed by binary_dilation(iterations=n_dilations),
    restated (scipy is the third-party home of this arithmetic; the reference pins scipy < 1.9, environment.yml:41):
    closing = n dilations then n erosions, each with border_value 0."""
    m = np.asarray(mask) != 0
    for _ in range(n_closings):
        m = _binary_step(m, False)
    for _ in range(n_closings):
        m = _binary_step(m, True)
    for _ in range(n_dilations):
        m = _binary_step(m, False)
    return m


def close_contact_sites(contacts, bb_dc, n_closings, cs_dilation, use_scipy=False):
    """syconn/extraction/cs_extraction_steps.py:439-461 ("next" row f2): per contact id, in the iteration order of
    ``bb_dc``, the id's mask inside its bounding box padded by ``n_closings`` (clipped to the volume) is closed and
    dilated; the result is written where the volume still holds background (or the id itself).  In place, like the
    reference.  ``use_scipy``: call scipy.ndimage literally as the reference does (pins the restatement above)."""
    for ix in bb_dc.keys():
        obj_start, obj_end = np.array(bb_dc[ix])
        obj_start -= n_closings
        obj_start[obj_start < 0] = 0
        obj_end += n_closings
        sl = tuple(slice(obj_start[ii], obj_end[ii], None) for ii in range(3))
        sub_vol = contacts[sl]
        binary_mask = (sub_vol == ix).astype(np.int8, copy=False)
        if use_scipy:
            import scipy.ndimage
            res = scipy.ndimage.binary_closing(binary_mask, iterations=n_closings) if n_closings > 0 else binary_mask
            if cs_dilation > 0:
                res = scipy.ndimage.binary_dilation(res, iterations=cs_dilation)
        else:
            res = binary_closing_dilation(binary_mask, n_closings, cs_dilation)
        proc_mask = ((binary_mask == 1) | (sub_vol == 0)) & (res == 1)
        contacts[sl][proc_mask] = np.uint64(ix)  # == res[proc_mask] * ix (:461); np.uint64 keeps ids >= 2^63 exact
    return contacts


# ------------------------------------------------------------------------------------------ organelle morphology (f4)
def get_aniso_struct(scaling):
    """syconn/proc/image.py:522-539: 5 x 5 x 3 element -- single centre voxels above and below, and in the middle plane
    the result of ``scaling[2] // scaling[0]`` cross dilations of the centre pixel of a 5 x 5 array."""
    import scipy.ndimage
    aniso = scaling[2] // scaling[0]
    assert scaling[1] // scaling[0] == 1 and aniso >= 1
    centre = np.zeros((5, 5))
    centre[2, 2] = 1
    disc = centre != 0
    for _ in range(int(aniso)):  # == binary_dilation(centre, iterations=aniso), one step at a time (see _scipy_binary_op)
        disc = scipy.ndimage.binary_dilation(disc)
    return np.stack([centre, disc.astype(np.float64), centre], axis=2)


def _scipy_binary_op(name, mask, n, kw):
    """``getattr(scipy.ndimage, name)(mask, iterations=n, **kw)`` evaluated as single-iteration calls: n erosions, n
    dilations, n erosions + n dilations (opening) or n dilations + n erosions (closing) -- scipy's own definition of
    ``iterations``.  The literal multi-iteration call is avoided on purpose: scipy 1.18.1's iterated code path corrupts the heap
    ("double free or corruption") when the array is smaller than the structuring element, which the per-object boxes of
    the reference's loop often are; the golden vectors (made by the reference's literal calls) pin the equivalence."""
    import scipy.ndimage
    seq = {"binary_erosion": "e" * n, "binary_dilation": "d" * n, "binary_opening": "e" * n + "d" * n,
           "binary_closing": "d" * n + "e" * n}[name]
    m = np.asarray(mask) != 0
    for step in seq:
        m = (scipy.ndimage.binary_erosion if step == "e" else scipy.ndimage.binary_dilation)(m, **kw)
    return m


def apply_morphological_operations(vol, morph_ops, structure=None):
    """syconn/proc/image.py:485-507 with _count_subsequent_mops (:510-519) and _multi_mop_findobjects (:358-437), restated
    for the ops the organelle worker uses (binary erosion / dilation / opening / closing; scipy.ndimage is the third-party
    home of the arithmetic): runs of equal ops become one call with that many iterations; every call goes object by
    object through ``find_objects`` boxes -- dilation / closing on the box zero-padded by the iteration count and cropped
    back, writing the object's own voxels and background only; erosion / opening on the box itself, writing the object's
    own voxels only.  In place, like the reference."""
    import scipy.ndimage
    runs = []
    for name in morph_ops:
        if runs and runs[-1][0] == name:
            runs[-1][1] += 1
        else:
            runs.append([name, 1])
    kw = {} if structure is None else {"structure": structure}
    for name, n in runs:
        grows = ("closing" in name) or ("dilation" in name)
        shrinks = ("erosion" in name) or ("opening" in name)
        assert grows != shrinks, name
        boxes = scipy.ndimage.find_objects(vol)
        for ix in np.unique(vol[vol != 0]):
            box = boxes[int(ix) - 1]
            sub = vol[box]
            if grows:
                sub = np.pad(sub, n)
            own = (sub == ix).astype(np.int32)
            res = _scipy_binary_op(name, own, n, kw)
            if grows:
                inner = (slice(n, -n),) * 3
                res, own, sub = res[inner], own[inner], sub[inner]
                target = (own == 1) | (sub == 0)
            else:
                target = own == 1
            vol[box][target] = res[target] * ix
    return vol


def watershed_seeds(tmp_data, morph_ops, structure, min_size):
    """syconn/extraction/object_extraction_steps.py:313-343, restated: the mask after the ops before the first erosion, and the
    markers = scipy.ndimage.label of the mask after the remaining ops, cleaned of markers below ``min_size`` voxels with
    the freed ids refilled from the largest kept ids.  (This stretch of the worker cannot be run on its own -- it sits
    between KnossosDataset I/O and the vigra / skimage watershed -- so it is pinned by restatement only.)"""
    import scipy.ndimage
    first = morph_ops.index("binary_erosion")
    mask = apply_morphological_operations(tmp_data.copy(), morph_ops[:first], structure)
    markers = apply_morphological_operations(mask.copy(), morph_ops[first:], structure)
    markers = scipy.ndimage.label(markers)[0].astype(np.uint32)
    if min_size > 1:
        ixs, cnt = np.unique(markers, return_counts=True)
        m = (ixs != 0) & (cnt < min_size)
        ixs_del = np.sort(ixs[m])
        ixs_keep = np.sort(ixs[~m])
        label_m = {ix_del: 0 for ix_del in ixs_del}
        ii = len(ixs_keep) - 1
        for ix_del in ixs_del:
            if (ix_del > ixs_keep[ii]) or (ixs_keep[ii] == 0) or (ii < 0):
                break
            label_m[ixs_keep[ii]] = ix_del
            ii -= 1
        out = markers.copy()
        for k, v in label_m.items():  # relabel_vol (block_processing_C.pyx:161-170): one look-up per voxel, no chaining
            out[markers == k] = v
        markers = out
    return mask, markers


def merge_prop_dicts(prop_dicts, offset=None):
    """syconn/proc/sd_proc.py:1248-1273: rc overwritten by later chunks, bboxes appended, sizes summed."""
    tot_rc, tot_bb, tot_size = prop_dicts[0]
    for el in prop_dicts[1:]:
        if len(el[0]) == 0:
            continue
        if offset is not None:
            for k in el[0]:
                el[0][k] = [el[0][k][ii] + offset[ii] for ii in range(3)]
        tot_rc.update(el[0])
        for k, v in el[1].items():
            bb = v if offset is None else [[v[0][ii] + offset[ii] for ii in range(3)],
                                           [v[1][ii] + offset[ii] for ii in range(3)]]
            tot_bb[k].append(bb)
        for k, v in el[2].items():
            tot_size[k] = tot_size.get(k, 0) + v


def merge_map_dicts(map_dicts):
    """syconn/proc/sd_proc.py:1300-1322."""
    tot_map = map_dicts[0]
    for el in map_dicts[1:]:
        for sc_id, sc_dc in el.items():
            if sc_id in tot_map:
                for cellsv_id, n in sc_dc.items():
                    tot_map[sc_id][cellsv_id] = tot_map[sc_id].get(cellsv_id, 0) + n
            else:
                tot_map[sc_id] = sc_dc


def new_prop_acc():
    return [{}, defaultdict(list), {}]
