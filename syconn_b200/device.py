"""Device-resident API: libsyk kernels on torch CUDA tensors (PyTorch only owns the buffers and the stream).

Array-level counterpart of the reference's dict-returning functions; outputs use the dtypes of
``dataset_analysis`` (syconn/proc/sd_proc.py:244-251) via the 64-byte ``syk_record_t`` records.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GEOM_DTYPE, PAIR_DTYPE, RECORD_DTYPE, check, i32, i64


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _elem_bytes(t):
    if t.dtype in (torch.int64, torch.uint64):
        return 8
    if t.dtype in (torch.int32, torch.uint32):
        return 4
    raise ValueError(f"label tensors must be 32/64-bit integers, got {t.dtype}")


def geoms(origins, shapes):
    g = np.zeros(len(origins), GEOM_DTYPE)
    g["origin"] = np.asarray(origins, np.int64).reshape(-1, 3)
    g["shape"] = np.asarray(shapes, np.int64).reshape(-1, 3)
    return g


class IdTable:
    """Device hash table id -> (count, bbox, rep) (``syk_table_t``)."""

    def __init__(self, capacity):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.syk_table_create(C.byref(h), int(capacity)))
        self.h = h
        self.capacity = int(self._L.syk_table_capacity(h))

    def clear(self, stream=None):
        check(self._L.syk_table_clear(self.h, _stream_ptr(stream)))

    def count(self, stream=None):
        n, ovf = C.c_uint64(), C.c_int()
        check(self._L.syk_table_count(self.h, _stream_ptr(stream), C.byref(n), C.byref(ovf)))
        return n.value, bool(ovf.value)

    def export(self, geom_array, max_records=None, stream=None, sort=False):
        """-> int64 CUDA tensor [n, 8] (one ``syk_record_t`` per row); use ``records_numpy`` to view it.
        ``sort``: ascending ids (radix sort on the device)."""
        if max_records is None:
            max_records, ovf = self.count(stream)
            if ovf:
                raise _lib.SykError(_lib.SYK_EOVERFLOW, "id table overflow")
        out = torch.empty((max(int(max_records), 1), 8), dtype=torch.int64, device="cuda")
        n = C.c_uint64()
        g = np.ascontiguousarray(geom_array)
        check(self._L.syk_table_export(self.h, g.ctypes.data, len(g), out.data_ptr(), int(max_records), C.byref(n),
                                       _stream_ptr(stream)))
        if sort and n.value > 1:
            check(self._L.syk_records_sort_by_id(out.data_ptr(), n.value, _stream_ptr(stream)))
        return out[:n.value]

    def merge_records(self, recs, stream=None):
        if recs.shape[0]:
            assert recs.is_cuda and recs.dtype == torch.int64 and recs.is_contiguous()
            check(self._L.syk_table_merge_records(self.h, recs.data_ptr(), recs.shape[0], _stream_ptr(stream)))

    def close(self):
        if self.h:
            self._L.syk_table_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PairTable:
    """Device hash table (sub_id, cell_id) -> overlap count (``syk_pairs_t``)."""

    def __init__(self, capacity):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.syk_pairs_create(C.byref(h), int(capacity)))
        self.h = h
        self.capacity = int(self._L.syk_pairs_capacity(h))

    def clear(self, stream=None):
        check(self._L.syk_pairs_clear(self.h, _stream_ptr(stream)))

    def export(self, max_pairs=None, stream=None):
        """-> int64 CUDA tensor [n, 4] (``syk_pair_t`` rows)."""
        if max_pairs is None:
            max_pairs = self.capacity
        out = torch.empty((max(int(max_pairs), 1), 4), dtype=torch.int64, device="cuda")
        n = C.c_uint64()
        check(self._L.syk_pairs_export(self.h, out.data_ptr(), int(max_pairs), C.byref(n), _stream_ptr(stream)))
        return out[:n.value]

    def merge(self, pairs, stream=None):
        if pairs.shape[0]:
            assert pairs.is_cuda and pairs.dtype == torch.int64 and pairs.is_contiguous()
            check(self._L.syk_pairs_merge(self.h, pairs.data_ptr(), pairs.shape[0], _stream_ptr(stream)))

    def close(self):
        if self.h:
            self._L.syk_pairs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def records_numpy(recs):
    """int64 tensor [n, 8] -> structured numpy array (RECORD_DTYPE)."""
    return recs.cpu().numpy().view(RECORD_DTYPE).reshape(-1)


def pairs_numpy(pairs):
    return pairs.cpu().numpy().view(PAIR_DTYPE).reshape(-1)


def _strides(t):
    return list(t.stride())


def find_object_properties(table, labels, origin=(0, 0, 0), chunk_seq=0, stream=None):
    """syk_find_object_properties on a CUDA tensor [X,Y,Z] (any strides)."""
    check(_lib.load().syk_find_object_properties(table.h, labels.data_ptr(), _elem_bytes(labels), i64(labels.shape),
                                                 i64(_strides(labels)), i64(origin), int(chunk_seq), _stream_ptr(stream)))


def map_subcell_extract_props(cell_table, sub_tables, pair_tables, cell, subcell, origin=(0, 0, 0), chunk_seq=0,
                              stream=None):
    """syk_map_subcell_extract_props: cell [X,Y,Z], subcell [C,X,Y,Z] CUDA tensors.  ``cell_table``/``sub_tables``
    may be None (map_subcell_C mode)."""
    n_sub = subcell.shape[0]
    assert tuple(subcell.shape[1:]) == tuple(cell.shape), \
        "Segmentation of cells and subcellular structures must have same shape."
    assert subcell.dtype == cell.dtype
    L = _lib.load()
    vp = C.c_void_p
    for c0 in range(0, n_sub, 4):
        k = min(4, n_sub - c0)
        subs = (vp * k)(*[subcell[c0 + c].data_ptr() for c in range(k)])
        pts = (vp * k)(*[pair_tables[c0 + c].h for c in range(k)])
        sts = (vp * k)(*[sub_tables[c0 + c].h for c in range(k)]) if sub_tables is not None else None
        ct = cell_table.h if (cell_table is not None and c0 == 0) else None
        check(L.syk_map_subcell_extract_props(ct, sts, pts, cell.data_ptr(), i64(_strides(cell)), subs,
                                              i64(_strides(subcell)[1:]), k, _elem_bytes(cell), i64(cell.shape),
                                              i64(origin), int(chunk_seq), _stream_ptr(stream)))
    if n_sub == 0 and cell_table is not None:
        find_object_properties(cell_table, cell, origin, chunk_seq, stream)


def detect_cs(arr, stencil=(13, 13, 7), out=None, stream=None):
    """Fused detect_seg_boundaries + process_block_nonzero -> uint64 contact ids (int64 tensor), valid-cropped.
    The output is laid out like the input (same fastest axis) so that stores stay coalesced."""
    st = [int(s) for s in stencil]
    oshape = [arr.shape[i] - st[i] + 1 for i in range(3)]
    if out is None:
        order = sorted(range(3), key=lambda a: -abs(arr.stride(a)))  # slowest .. fastest
        phys = torch.empty([max(oshape[a], 0) for a in order], dtype=torch.int64, device=arr.device)
        inv = [order.index(a) for a in range(3)]
        out = phys.permute(inv)
    if min(oshape) <= 0:
        return out
    check(_lib.load().syk_detect_cs(arr.data_ptr(), _elem_bytes(arr), i64(arr.shape), i64(_strides(arr)), i32(st),
                                    out.data_ptr(), i64(_strides(out)), _stream_ptr(stream)))
    return out


def process_block_nonzero(edges, arr, stencil=(7, 7, 3), stream=None):
    st = [int(s) for s in stencil]
    oshape = [max(arr.shape[i] - st[i] + 1, 0) for i in range(3)]
    out = torch.empty(oshape, dtype=torch.int64, device=arr.device)
    if min(oshape) <= 0:
        return out
    eb = 1 if edges.dtype in (torch.uint8, torch.bool, torch.int8) else 4
    check(_lib.load().syk_process_block_nonzero(edges.data_ptr(), eb, i64(_strides(edges)), arr.data_ptr(), _elem_bytes(arr),
                                                i64(_strides(arr)), i64(arr.shape), i32(st), out.data_ptr(),
                                                i64(_strides(out)), _stream_ptr(stream)))
    return out


def detect_seg_boundaries(arr, stream=None):
    out = torch.empty(tuple(arr.shape), dtype=torch.uint8, device=arr.device)
    check(_lib.load().syk_detect_seg_boundaries(arr.data_ptr(), _elem_bytes(arr), i64(arr.shape), i64(_strides(arr)),
                                                out.data_ptr(), _stream_ptr(stream)))
    return out


def extract_cs_syntype(table, cs, syn, asym, sym, max_vox=None, origin=(0, 0, 0), chunk_seq=0, stream=None, syn_table=None):
    """syk_extract_cs_syntype(_props) on CUDA tensors (any strides, e.g. cropped views): contact-site props go to ``table``,
    the props of the synaptic parts to ``syn_table`` (optional), the synaptic voxel tuples are returned as an int64 tensor
    [n, 4] (``syk_synvox_t`` rows: id, lin, flags, pad).
    Synchronises (the tuple count is read back; the call is repeated once when ``max_vox`` was too small -- the tables
    are cleared for the repeat, so pass empty tables)."""
    for m in (syn, asym, sym):
        assert m.dtype == torch.uint8 and tuple(m.shape) == tuple(cs.shape)
    L = _lib.load()
    n_max = int(max_vox) if max_vox else cs.numel() // 16 + 4096
    while True:
        vox = torch.empty((n_max, 4), dtype=torch.int64, device=cs.device)
        counter = torch.zeros(2, dtype=torch.int64, device=cs.device)
        check(L.syk_extract_cs_syntype_props(table.h, syn_table.h if syn_table is not None else None, cs.data_ptr(), _elem_bytes(cs),
                                             i64(cs.shape), i64(_strides(cs)), syn.data_ptr(), i64(_strides(syn)), asym.data_ptr(),
                                             i64(_strides(asym)), sym.data_ptr(), i64(_strides(sym)), i64(origin), int(chunk_seq),
                                             vox.data_ptr(), n_max, counter.data_ptr(), _stream_ptr(stream)))
        n = int(counter[0].item())
        if n <= n_max:
            return vox[:n]
        table.clear(stream)
        if syn_table is not None:
            syn_table.clear(stream)
        n_max = n


def close_contacts(cs, ids, bbox, n_closings=6, cs_dilation=2, stream=None):
    """syk_close_contacts on a CUDA contact volume (in place): ``ids`` uint64 [n] and ``bbox`` int32 [n, 2, 3]
    (min, exclusive max) are HOST arrays in processing order (cs_extraction_steps.py:439-461)."""
    import numpy as np
    ids = np.ascontiguousarray(ids, np.uint64)
    bbox = np.ascontiguousarray(bbox, np.int32).reshape(-1, 2, 3)
    assert len(ids) == len(bbox)
    check(_lib.load().syk_close_contacts(cs.data_ptr(), _elem_bytes(cs), i64(cs.shape), i64(_strides(cs)), ids.ctypes.data,
                                         bbox.ctypes.data, len(ids), int(n_closings), int(cs_dilation), _stream_ptr(stream)))
    return cs


def close_contacts_records(cs, records, n_closings=6, cs_dilation=2, stream=None):
    """syk_close_contacts_records: like ``close_contacts`` with the boxes taken from DEVICE records (int64 tensor [n, 8] of a
    table export in volume-local coordinates, sorted by id) -- no host round trip of the box list."""
    assert records.is_cuda and records.dtype == torch.int64 and records.is_contiguous()
    check(_lib.load().syk_close_contacts_records(cs.data_ptr(), _elem_bytes(cs), i64(cs.shape), i64(_strides(cs)), records.data_ptr(),
                                                 records.shape[0], int(n_closings), int(cs_dilation), _stream_ptr(stream)))
    return cs


def synth_labels(shape, origin=(0, 0, 0), pitch=(32, 32, 16), warp_amp=4, seed=0, kind=0, density16=1,
                 dtype=torch.int64, order="C", out=None, stream=None):
    """Device twin of ``syconn_b200.synth.synth_labels`` (bit-identical)."""
    if out is None:
        if order == "C":
            out = torch.empty(tuple(shape), dtype=dtype, device="cuda")
        else:
            out = torch.empty(tuple(shape)[::-1], dtype=dtype, device="cuda").permute(2, 1, 0)
    check(_lib.load().syk_synth_labels(out.data_ptr(), _elem_bytes(out), i64(out.shape), i64(_strides(out)), i64(origin),
                                       i32(pitch), int(warp_amp), int(seed), int(kind), int(density16),
                                       _stream_ptr(stream)))
    return out


def bucket_records(recs, n_owners, stream=None):
    """Reorder records into per-owner buckets -> (bucketed tensor, counts list)."""
    out = torch.empty_like(recs)
    counts = torch.zeros(n_owners, dtype=torch.int64, device=recs.device)
    check(_lib.load().syk_records_bucket(recs.data_ptr(), recs.shape[0], n_owners, out.data_ptr(), counts.data_ptr(),
                                         _stream_ptr(stream)))
    return out, counts


def bucket_pairs(pairs, n_owners, stream=None):
    out = torch.empty_like(pairs)
    counts = torch.zeros(n_owners, dtype=torch.int64, device=pairs.device)
    check(_lib.load().syk_pairs_bucket(pairs.data_ptr(), pairs.shape[0], n_owners, out.data_ptr(), counts.data_ptr(),
                                       _stream_ptr(stream)))
    return out, counts


def label_components(vol, threshold=0, out=None, stream=None):
    """syk_label_components: ``scipy.ndimage.label(vol > threshold)`` (6-connectivity) on a CUDA tensor [X,Y,Z] of any
    integer dtype and strides -> (int32 label tensor laid out like ``vol``, number of components)."""
    eb = vol.element_size()
    if out is None:
        order = sorted(range(3), key=lambda a: -abs(vol.stride(a)))
        phys = torch.empty([vol.shape[a] for a in order], dtype=torch.int32, device=vol.device)
        out = phys.permute([order.index(a) for a in range(3)])
    n = C.c_uint64()
    check(_lib.load().syk_label_components(vol.data_ptr(), eb, i64(vol.shape), i64(_strides(vol)), int(threshold), out.data_ptr(),
                                           i64(_strides(out)), C.byref(n), _stream_ptr(stream)))
    return out, int(n.value)


def label_overlap_pairs(pairs, a, b, a_offset=0, b_offset=0, stream=None):
    """syk_label_overlap_pairs: count the (a + a_offset, b + b_offset) pairs of two int32 label blocks of equal shape."""
    assert tuple(a.shape) == tuple(b.shape) and a.dtype == torch.int32 and b.dtype == torch.int32
    check(_lib.load().syk_label_overlap_pairs(pairs.h, a.data_ptr(), i64(_strides(a)), b.data_ptr(), i64(_strides(b)), i64(a.shape),
                                              int(a_offset), int(b_offset), _stream_ptr(stream)))


MORPH_OPS = {"binary_erosion": 0, "binary_dilation": 1, "binary_opening": 2, "binary_closing": 3}


def binary_morph_ops(vol, ops, iterations, structure, stream=None):
    """syk_binary_morph_ops: the reference's per-object morphology (``proc/image.py:358-437``) on a 0/1 CUDA volume
    [X,Y,Z], in place.  ``ops``: names from MORPH_OPS (or their codes), ``iterations`` one count per op, ``structure`` a
    point-symmetric 3-D array-like with odd extents."""
    st = np.ascontiguousarray(np.asarray(structure) != 0, dtype=np.uint8)
    assert st.ndim == 3, "3-D structuring element expected"
    codes = (C.c_int32 * len(ops))(*[MORPH_OPS[o] if isinstance(o, str) else int(o) for o in ops])
    its = (C.c_int32 * len(ops))(*[int(i) for i in iterations])
    check(_lib.load().syk_binary_morph_ops(vol.data_ptr(), vol.element_size(), i64(vol.shape), i64(_strides(vol)), st.tobytes(),
                                           i64(st.shape), codes, its, len(ops), _stream_ptr(stream)))
    return vol


def label_map(labels, lut, stream=None):
    """syk_label_map: int32 label tensor [X,Y,Z] mapped in place through the dense int32 table ``lut`` (labels >= len(lut)
    and 0 stay)."""
    assert labels.dtype == torch.int32 and lut.dtype == torch.int32 and lut.is_cuda and lut.is_contiguous()
    check(_lib.load().syk_label_map(labels.data_ptr(), i64(labels.shape), i64(_strides(labels)), lut.data_ptr(), lut.numel(),
                                    _stream_ptr(stream)))
    return labels
