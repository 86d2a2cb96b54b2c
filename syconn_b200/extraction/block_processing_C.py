"""Drop-in for ``syconn.extraction.block_processing_C`` (Cython module): ``process_block_nonzero`` and ``kernel``."""
import numpy as np

from .. import _lib
from ._host import dense_view, estrides


def _u32_3d(a, name):
    a = np.asarray(a)
    if a.dtype != np.uint32:
        raise ValueError(f"Buffer dtype mismatch, expected 'uint32_t' but got '{a.dtype}' ({name})")
    if a.ndim != 3:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 3, got {a.ndim})")
    return a


def process_block_nonzero(edges, arr, stencil1=(7, 7, 3)):
    """syconn/extraction/block_processing_C.pyx:53-75: valid-mode partner stencil on boundary voxels.
    Returns a uint64 array of shape ``arr.shape - stencil + 1`` (the reference returns a cython.view.array that
    callers wrap with ``np.asarray``, cs_extraction_steps.py:391)."""
    edges = _u32_3d(edges, "edges")
    arr = _u32_3d(arr, "arr")
    st = [int(stencil1[0]), int(stencil1[1]), int(stencil1[2])]
    assert (st[0] % 2 + st[1] % 2 + st[2] % 2) == 3
    oshape = tuple(max(0, arr.shape[i] - st[i] + 1) for i in range(3))
    out = np.zeros(oshape, np.uint64)
    if out.size == 0:
        return out
    edges = dense_view(edges)
    arr = dense_view(arr)
    L = _lib.load()
    _lib.check(L.syk_process_block_nonzero_host(edges.ctypes.data, 4, _lib.i64(estrides(edges)), arr.ctypes.data, 4,
                                                _lib.i64(estrides(arr)), _lib.i64(arr.shape), _lib.i32(st),
                                                out.ctypes.data))
    return out


def extract_cs_syntype(cs_seg, syn_mask, asym_mask, sym_mask, offset):
    """syconn/extraction/block_processing_C.pyx:78-158: per contact-site id the cs props, the props of its synaptic part
    (``syn_mask != 0``), the number of ``asym_mask == 1`` / ``sym_mask == 1`` synaptic voxels and the synaptic voxel
    list (scan order, ``offset`` added).  Returns
    ``[rc, bb, size], [rc_syn, bb_syn, size_syn], cs_asym, cs_sym, voxels_syn`` like the reference."""
    import ctypes as C
    from ._host import check_label_array, syntype_to_dicts
    cs_seg = check_label_array(cs_seg, "cs_seg", 3)
    masks = []
    for name, m in (("syn_mask", syn_mask), ("asym_mask", asym_mask), ("sym_mask", sym_mask)):
        m = np.asarray(m)
        if m.dtype != np.uint8:
            raise ValueError(f"Buffer dtype mismatch, expected 'uint8_t' but got '{m.dtype}' ({name})")
        assert m.shape == cs_seg.shape, "cs_seg, syn_mask, sym_mask and asym_mask must all have the same shape"
        masks.append(dense_view(m))
    cs_seg = dense_view(cs_seg)
    L = _lib.load()
    rec, n_rec, vox, n_vox = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
    _lib.check(L.syk_extract_cs_syntype_host(
        cs_seg.ctypes.data, cs_seg.itemsize, _lib.i64(cs_seg.shape), _lib.i64(estrides(cs_seg)),
        masks[0].ctypes.data, _lib.i64(estrides(masks[0])), masks[1].ctypes.data, _lib.i64(estrides(masks[1])),
        masks[2].ctypes.data, _lib.i64(estrides(masks[2])), C.byref(rec), C.byref(n_rec), C.byref(vox), C.byref(n_vox)))
    return syntype_to_dicts(_lib.take_array(rec.value, n_rec.value, _lib.RECORD_DTYPE),
                            _lib.take_array(vox.value, n_vox.value, _lib.SYNVOX_DTYPE), cs_seg.shape, offset)


def kernel(chunk, center_id):
    """syconn/extraction/block_processing_C.pyx:21-49: one window -> packed partner id (Python int)."""
    chunk = _u32_3d(chunk, "chunk")
    sh = chunk.shape
    # odd-pad so that the window is the full stencil around a synthetic centre is not needed: run the window as a
    # 1-output valid stencil with the edge flag forced on and the centre id substituted
    st = [sh[0] | 1, sh[1] | 1, sh[2] | 1]
    pad = np.zeros(st, np.uint32)
    pad[:sh[0], :sh[1], :sh[2]] = chunk
    c = (st[0] // 2, st[1] // 2, st[2] // 2)
    orig = int(pad[c])
    cid = int(np.uint32(center_id))
    # the centre voxel itself counts for its own id in the reference; keep the histogram intact by only
    # redirecting which id is treated as centre
    if orig != cid:
        return _kernel_general(chunk, cid)
    edges = np.zeros(st, np.uint32)
    edges[c] = 1
    return int(process_block_nonzero(edges, pad, st)[0, 0, 0])


def _kernel_general(chunk, cid):
    # rarely used form (centre id differs from the voxel at the geometric centre): relabel-free evaluation by
    # counting through the stencil kernel on a volume whose geometric centre is moved outside the window
    sh = chunk.shape
    st = [2 * sh[0] + 1, 2 * sh[1] + 1, 2 * sh[2] + 1]
    pad = np.zeros(st, np.uint32)
    pad[:sh[0], :sh[1], :sh[2]] = chunk
    c = (st[0] // 2, st[1] // 2, st[2] // 2)  # outside the copied block
    pad[c] = cid
    edges = np.zeros(st, np.uint32)
    edges[c] = 1
    return int(process_block_nonzero(edges, pad, st)[0, 0, 0])
