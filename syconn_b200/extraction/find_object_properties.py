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
