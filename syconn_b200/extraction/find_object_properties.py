"""Drop-in for ``syconn.extraction.find_object_properties`` (the import boundary the reference's callers use:
extraction/cs_extraction_steps.py:40-41, proc/sd_proc.py:28)."""
import numpy as np

from .. import _lib, global_params
from ._host import dense_view, estrides
from .block_processing_C import process_block_nonzero  # noqa: F401
from .find_object_properties_C import find_object_properties, map_subcell_extract_props  # noqa: F401


def detect_seg_boundaries(arr):
    """syconn/extraction/find_object_properties.py:424-455: 6-neighbourhood boundary mask (bool)."""
    arr = np.asarray(arr)
    a = arr
    if a.dtype not in (np.uint32, np.uint64):
        # numba accepts any numeric array; equality of values == equality of their 64-bit patterns
        a = arr.astype(np.int64).view(np.uint64) if arr.dtype.kind in "iub" else \
            np.ascontiguousarray(arr, dtype=np.float64).view(np.uint64)
    assert a.ndim == 3
    out = np.zeros(a.shape, np.uint8)
    if a.size:
        a = dense_view(a)
        _lib.check(_lib.load().syk_detect_seg_boundaries_host(a.ctypes.data, a.itemsize, _lib.i64(a.shape),
                                                              _lib.i64(estrides(a)), out.ctypes.data))
    return out.view(np.bool_)


def detect_cs(arr, stencil=None, out=None, return_props=False):
    """syconn/extraction/find_object_properties.py:458-472.  Boundary mask and partner stencil are fused in one
    kernel (``syk_detect_cs_host``); uint64 input is narrowed to uint32 exactly like the caller-side
    ``.astype(np.uint32)`` (cs_extraction_steps.py:385-387).  ``stencil`` overrides the config default; ``out`` may be a
    preallocated C-contiguous uint64 array of the output shape (e.g. pinned memory).  ``return_props=True`` also returns
    ``find_object_properties(contacts)`` (the next call of the reference worker, cs_extraction_steps.py:439), computed
    while the contact volume is still on the GPU: ``(contacts, (rep_coords, bounding_box, sizes))``; ``return_props="records"``
    returns the raw ``syk_record_t`` array instead of the three dicts."""
    arr = np.asarray(arr)
    if arr.dtype not in (np.uint32, np.uint64):
        raise ValueError(f"Buffer dtype mismatch, expected 'uint32_t' but got '{arr.dtype}'")
    if stencil is None:
        stencil = global_params.config['cell_objects']['cs_filtersize']
    st = [int(s) for s in stencil]
    assert (st[0] % 2 + st[1] % 2 + st[2] % 2) == 3
    oshape = tuple(max(0, arr.shape[i] - st[i] + 1) for i in range(3))
    if out is None:
        out = np.empty(oshape, np.uint64)
    assert out.shape == oshape and out.dtype == np.uint64 and out.flags.c_contiguous
    if out.size == 0:
        return (out, ({}, {}, {})) if return_props else out
    arr = dense_view(arr)
    if return_props:
        import ctypes as C
        from ._host import records_to_dicts
        rec, n = C.c_void_p(), C.c_uint64()
        _lib.check(_lib.load().syk_detect_cs_props_host(arr.ctypes.data, arr.itemsize, _lib.i64(arr.shape), _lib.i64(estrides(arr)),
                                                        _lib.i32(st), out.ctypes.data, C.byref(rec), C.byref(n)))
        rec = _lib.take_array(rec.value, n.value, _lib.RECORD_DTYPE)
        return out, (rec if return_props == "records" else records_to_dicts(rec))
    _lib.check(_lib.load().syk_detect_cs_host(arr.ctypes.data, arr.itemsize, _lib.i64(arr.shape), _lib.i64(estrides(arr)),
                                              _lib.i32(st), out.ctypes.data))
    return out


def detect_contact_partners(seg_arr, edge_arr, offset):
    """syconn/extraction/find_object_properties.py:371-421 (numba).  Most frequent foreign id inside the window
    ``offset`` around every voxel flagged in ``edge_arr``; ties go to the id met first in the x, y, z scan of the window.
    Returns the "valid" XYZC (C = 2) uint64 volume of sorted partner ids.  The GPU path supports the symmetric windows
    the reference itself uses (``offset = [(-o, o)] * 3``, :363-366); anything else raises ``NotImplementedError``."""
    seg_arr = np.asarray(seg_arr)
    offset = np.asarray(offset)
    if offset.shape != (3, 2) or np.any(offset[:, 0] != -offset[:, 1]) or np.any(offset[:, 1] < 0):
        raise NotImplementedError("libsyk implements the symmetric windows of detect_cs_64bit (offset = [(-o, o)] * 3)")
    if seg_arr.dtype not in (np.uint32, np.uint64):
        seg_arr = seg_arr.astype(np.uint64)
    st = [int(2 * o + 1) for o in offset[:, 1]]
    oshape = tuple(max(0, seg_arr.shape[i] - st[i] + 1) for i in range(3))
    out = np.zeros(oshape + (2,), np.uint64)
    if out.size == 0:
        return out
    seg = dense_view(seg_arr)
    if edge_arr is None:
        ep, eb, es = None, 0, None
    else:
        edges = np.asarray(edge_arr)
        assert edges.shape == seg_arr.shape
        edges = dense_view(edges.view(np.uint8) if edges.dtype == np.bool_ else
                           edges if edges.dtype in (np.uint8, np.uint32) else (edges != 0).view(np.uint8))
        ep, eb, es = edges.ctypes.data, edges.itemsize, _lib.i64(estrides(edges))
    _lib.check(_lib.load().syk_detect_contact_partners_host(ep, eb, es, seg.ctypes.data, seg.itemsize, _lib.i64(estrides(seg)),
                                                            _lib.i64(seg.shape), _lib.i32(st), out.ctypes.data))
    return out


def detect_cs_64bit(arr):
    """syconn/extraction/find_object_properties.py:347-368: ``detect_seg_boundaries`` + ``detect_contact_partners``
    with the stencil of ``global_params.config['cell_objects']['cs_filtersize']``; 4-D XYZC result (C = 2)."""
    stencil = np.array(global_params.config['cell_objects']['cs_filtersize'])
    assert np.sum(stencil % 2) == 3
    o = stencil // 2
    return detect_contact_partners(arr, None, np.array([(-o[0], o[0]), (-o[1], o[1]), (-o[2], o[2])]))


def find_object_properties_cs_64bit(cs_seg):
    """syconn/extraction/find_object_properties.py:197-269: representative coordinate, bounding box and size of every
    partner pair of an XYZC (C = 2) contact volume, as nested dicts ``d[id0][id1]`` (int64 arrays / int)."""
    import ctypes as C
    cs_seg = np.asarray(cs_seg)
    assert cs_seg.ndim == 4 and cs_seg.shape[3] == 2
    if cs_seg.dtype != np.uint64:
        cs_seg = cs_seg.astype(np.uint64)
    rep_coords, bounding_box, sizes = {}, {}, {}
    if cs_seg.size == 0:
        return rep_coords, bounding_box, sizes
    cs = dense_view(cs_seg)
    rec, par, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    _lib.check(_lib.load().syk_find_object_properties_cs_64bit_host(cs.ctypes.data, _lib.i64(cs.shape[:3]), _lib.i64(estrides(cs)),
                                                                    C.byref(rec), C.byref(par), C.byref(n)))
    recs = _lib.take_array(rec.value, n.value, _lib.RECORD_DTYPE)
    pairs = _lib.take_array(par.value, 2 * n.value, np.dtype("<u8")).reshape(-1, 2)
    bbs = np.stack([recs["bb_min"], recs["bb_max"]], axis=1).astype(np.int64) if len(recs) else np.zeros((0, 2, 3), np.int64)
    reps = recs["rep"].astype(np.int64)
    for i in range(len(recs)):
        k0, k1 = int(pairs[i, 0]), int(pairs[i, 1])
        rep_coords.setdefault(k0, {})[k1] = reps[i]
        bounding_box.setdefault(k0, {})[k1] = bbs[i]
        sizes.setdefault(k0, {})[k1] = int(recs["count"][i])
    return rep_coords, bounding_box, sizes


def extract_cs_syntype_64bit(cs_seg, syn_mask, asym_mask, sym_mask):
    """syconn/extraction/find_object_properties.py:23-195 -- dead code in the reference (no caller; its numba typing
    fails at the first call).  Not provided; use ``block_processing_C.extract_cs_syntype``."""
    raise NotImplementedError("extract_cs_syntype_64bit has no caller in the reference and does not compile there; "
                              "use block_processing_C.extract_cs_syntype")


def convert_nvox2ratio_syntype(syn_cnts, sym_cnts, asym_cnts):
    """syconn/extraction/find_object_properties.py:272-299: per contact id, sym / asym voxel counts divided by the
    synaptic voxel count (0 when the id has no such voxels).  Returns ``(asym_ratio, sym_ratio)``."""
    sym_ratio = {k: (sym_cnts[k] / n if k in sym_cnts else 0) for k, n in syn_cnts.items()}
    asym_ratio = {k: (asym_cnts[k] / n if k in asym_cnts else 0) for k, n in syn_cnts.items()}
    return asym_ratio, sym_ratio


def merge_type_dicts(type_dicts):
    """syconn/extraction/find_object_properties.py:302-320: sum the per-id counts into ``type_dicts[0]`` (in place)."""
    tot = type_dicts[0]
    for d in type_dicts[1:]:
        for k, cnt in d.items():
            tot[k] = tot[k] + cnt if k in tot else cnt


def merge_voxel_dicts(voxel_dicts, key_to_str=False):
    """syconn/extraction/find_object_properties.py:323-344: concatenate the per-id voxel lists into
    ``voxel_dicts[0]`` (in place); arrays become lists, keys optionally strings."""
    tot = voxel_dicts[0]
    for d in voxel_dicts[1:]:
        for k, vxs in d.items():
            if key_to_str:
                k = str(k)
            if k in tot:
                tot[k].extend(vxs)
            else:
                tot[k] = vxs.tolist() if isinstance(vxs, np.ndarray) else vxs
