"""ctypes binding of ``libsyk.so`` (the C ABI declared in ``include/syk.h``).

The library is built in-tree by ``syconn_b200/csrc/build.py`` (nvcc, sm_100a).  There is no CPU fallback:
a missing library raises ``ImportError`` and every compute entry point raises ``SykError`` without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("SYK_LIB_NAME", "libsyk.so"))

SYK_OK, SYK_EINVAL, SYK_ECUDA, SYK_ENODEV, SYK_EOVERFLOW, SYK_ENOMEM = 0, -1, -2, -3, -4, -5

# numpy views of the C structs (include/syk.h)
RECORD_DTYPE = np.dtype([("id", "<u8"), ("count", "<u8"), ("rep_key", "<u8"), ("bb_min", "<i4", (3,)),
                         ("bb_max", "<i4", (3,)), ("rep", "<i4", (3,)), ("chunk_seq", "<u4")])
SYNVOX_DTYPE = np.dtype([("id", "<u8"), ("lin", "<u8"), ("flags", "<u8"), ("_pad", "<u8")])
PAIR_DTYPE = np.dtype([("sub_id", "<u8"), ("cell_id", "<u8"), ("count", "<u8"), ("_pad", "<u8")])
GEOM_DTYPE = np.dtype([("origin", "<i8", (3,)), ("shape", "<i8", (3,))])
assert RECORD_DTYPE.itemsize == 64 and PAIR_DTYPE.itemsize == 32 and GEOM_DTYPE.itemsize == 48

EXPORTS = [
    "syk_version", "syk_last_error", "syk_device_count", "syk_set_device",
    "syk_table_create", "syk_table_destroy", "syk_table_clear", "syk_table_capacity", "syk_table_count",
    "syk_table_export", "syk_table_append_records", "syk_table_merge_records", "syk_records_bucket", "syk_records_decode_rep",
    "syk_pairs_create", "syk_pairs_capacity", "syk_pairs_destroy", "syk_pairs_clear", "syk_pairs_export", "syk_pairs_append", "syk_pairs_merge",
    "syk_pairs_bucket",
    "syk_find_object_properties", "syk_map_subcell_extract_props", "syk_detect_seg_boundaries",
    "syk_process_block_nonzero", "syk_detect_cs", "syk_extract_cs_syntype", "syk_synth_labels",
    "syk_find_object_properties_host", "syk_map_subcell_extract_props_host", "syk_detect_cs_host", "syk_detect_cs_props_host",
    "syk_process_block_nonzero_host", "syk_extract_cs_syntype_host", "syk_detect_seg_boundaries_host", "syk_free",
    "syk_detect_contact_partners", "syk_cs64_unpack", "syk_dense_relabel",
    "syk_detect_contact_partners_host", "syk_find_object_properties_cs_64bit_host",
    "syk_close_contacts", "syk_close_contacts_host", "syk_close_contacts_records",
    "syk_lz4_compress_bound", "syk_lz4_compress_block", "syk_lz4_decompress_block",
    "syk_label_components", "syk_label_overlap_pairs", "syk_binary_morph_ops", "syk_label_map", "syk_extract_cs_syntype_props", "syk_records_sort_by_id",
    "syk_table_append_records_min_vx", "syk_pairs_append_min_vx", "syk_pairs_attach_size",
]


class SykError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsyk error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libsyk.so (fails loudly when the CUDA extension was not built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m syconn_b200.csrc.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64p, i32p, u64p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
    u64, u32, ci = C.c_uint64, C.c_uint32, C.c_int
    L.syk_version.restype = ci
    L.syk_last_error.restype = C.c_char_p
    L.syk_device_count.restype = ci
    L.syk_set_device.argtypes = [ci]
    L.syk_table_create.argtypes = [C.POINTER(vp), u64]
    L.syk_table_destroy.argtypes = [vp]
    L.syk_table_clear.argtypes = [vp, vp]
    L.syk_table_capacity.argtypes = [vp]
    L.syk_table_capacity.restype = u64
    L.syk_table_count.argtypes = [vp, vp, u64p, C.POINTER(ci)]
    L.syk_table_export.argtypes = [vp, vp, u32, vp, u64, u64p, vp]
    L.syk_table_append_records.argtypes = [vp, vp, vp, u64, vp, vp]
    L.syk_pairs_append.argtypes = [vp, vp, u64, vp, vp]
    L.syk_table_append_records_min_vx.argtypes = [vp, vp, vp, u64, vp, u64, vp]
    L.syk_pairs_append_min_vx.argtypes = [vp, vp, vp, u64, vp, u64, vp, vp]
    L.syk_pairs_attach_size.argtypes = [vp, u64, vp, vp]
    L.syk_table_merge_records.argtypes = [vp, vp, u64, vp]
    L.syk_records_bucket.argtypes = [vp, u64, u32, vp, vp, vp]
    L.syk_records_decode_rep.argtypes = [vp, u64, vp, u32, vp]
    L.syk_pairs_create.argtypes = [C.POINTER(vp), u64]
    L.syk_pairs_capacity.argtypes = [vp]
    L.syk_pairs_capacity.restype = u64
    L.syk_pairs_destroy.argtypes = [vp]
    L.syk_pairs_clear.argtypes = [vp, vp]
    L.syk_pairs_export.argtypes = [vp, vp, u64, u64p, vp]
    L.syk_pairs_merge.argtypes = [vp, vp, u64, vp]
    L.syk_pairs_bucket.argtypes = [vp, u64, u32, vp, vp, vp]
    L.syk_find_object_properties.argtypes = [vp, vp, ci, i64p, i64p, i64p, u32, vp]
    L.syk_map_subcell_extract_props.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, i64p, C.POINTER(vp), i64p, ci, ci,
                                                i64p, i64p, u32, vp]
    L.syk_detect_seg_boundaries.argtypes = [vp, ci, i64p, i64p, vp, vp]
    L.syk_process_block_nonzero.argtypes = [vp, ci, i64p, vp, ci, i64p, i64p, i32p, vp, i64p, vp]
    L.syk_detect_cs.argtypes = [vp, ci, i64p, i64p, i32p, vp, i64p, vp]
    L.syk_extract_cs_syntype.argtypes = [vp, vp, ci, i64p, i64p, vp, i64p, vp, i64p, vp, i64p, i64p, u32, vp, u64, vp, vp]
    L.syk_extract_cs_syntype_props.argtypes = [vp, vp, vp, ci, i64p, i64p, vp, i64p, vp, i64p, vp, i64p, i64p, u32, vp, u64, vp, vp]
    L.syk_records_sort_by_id.argtypes = [vp, u64, vp]
    L.syk_extract_cs_syntype_host.argtypes = [vp, ci, i64p, i64p, vp, i64p, vp, i64p, vp, i64p, C.POINTER(vp), u64p,
                                              C.POINTER(vp), u64p]
    L.syk_synth_labels.argtypes = [vp, ci, i64p, i64p, i64p, i32p, C.c_int32, u64, ci, ci, vp]
    L.syk_find_object_properties_host.argtypes = [vp, ci, i64p, i64p, u64, C.POINTER(vp), u64p]
    L.syk_map_subcell_extract_props_host.argtypes = [vp, i64p, vp, i64p, ci, ci, i64p, ci, u64, C.POINTER(vp), u64p,
                                                     C.POINTER(vp), u64p, C.POINTER(vp), u64p]
    L.syk_detect_cs_host.argtypes = [vp, ci, i64p, i64p, i32p, vp]
    L.syk_detect_cs_props_host.argtypes = [vp, ci, i64p, i64p, i32p, vp, C.POINTER(vp), u64p]
    L.syk_process_block_nonzero_host.argtypes = [vp, ci, i64p, vp, ci, i64p, i64p, i32p, vp]
    L.syk_detect_seg_boundaries_host.argtypes = [vp, ci, i64p, i64p, vp]
    L.syk_detect_contact_partners.argtypes = [vp, ci, i64p, vp, i64p, i64p, i32p, vp, i64p, vp]
    L.syk_cs64_unpack.argtypes = [vp, u64, vp, vp, vp]
    L.syk_dense_relabel.argtypes = [vp, vp, ci, u64, vp, vp, u64, u64p, vp]
    L.syk_detect_contact_partners_host.argtypes = [vp, ci, i64p, vp, ci, i64p, i64p, i32p, vp]
    L.syk_find_object_properties_cs_64bit_host.argtypes = [vp, i64p, i64p, C.POINTER(vp), C.POINTER(vp), u64p]
    L.syk_close_contacts.argtypes = [vp, ci, i64p, i64p, vp, vp, u64, ci, ci, vp]
    L.syk_close_contacts_host.argtypes = [vp, ci, i64p, i64p, vp, vp, u64, ci, ci]
    L.syk_lz4_compress_bound.argtypes = [u64]
    L.syk_lz4_compress_bound.restype = u64
    L.syk_lz4_compress_block.argtypes = [vp, u64, vp, u64, u64p]
    L.syk_lz4_decompress_block.argtypes = [C.c_char_p, u64, vp, u64, u64p]
    L.syk_label_components.argtypes = [vp, ci, i64p, i64p, u64, vp, i64p, u64p, vp]
    L.syk_label_overlap_pairs.argtypes = [vp, vp, i64p, vp, i64p, i64p, u64, u64, vp]
    L.syk_binary_morph_ops.argtypes = [vp, ci, i64p, i64p, C.c_char_p, i64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), ci, vp]
    L.syk_close_contacts_records.argtypes = [vp, ci, i64p, i64p, vp, u64, ci, ci, vp]
    L.syk_label_map.argtypes = [vp, i64p, i64p, vp, u64, vp]
    L.syk_free.argtypes = [vp]
    L.syk_free.restype = None
    for name in EXPORTS:
        f = getattr(L, name)
        if name not in ("syk_version", "syk_last_error", "syk_device_count", "syk_table_capacity", "syk_pairs_capacity", "syk_free", "syk_lz4_compress_bound"):
            f.restype = ci
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SykError(rc, load().syk_last_error().decode())
    return rc


def i64(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def i32(vals):
    return (C.c_int32 * len(vals))(*[int(v) for v in vals])


def take_array(ptr, n, dtype):
    """Copy a malloc'ed C array of n structs into a numpy array and free it."""
    L = load()
    if not ptr or n == 0:
        if ptr:
            L.syk_free(ptr)
        return np.empty(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n).copy()
    L.syk_free(ptr)
    return arr
