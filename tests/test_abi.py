"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/syk.h declares,
refuses to compute without a GPU (no CPU fallback), and the Python shims keep the reference's error behaviour."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "syk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(syk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from syconn_b200 import _lib
    from syconn_b200.csrc import build
    build.build()
    L = _lib.load()
    decl = _declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/syk.h but not exported by libsyk.so"
    assert set(decl) == set(_lib.EXPORTS)
    assert L.syk_version() == 100


def test_struct_layouts_match_header():
    from syconn_b200 import _lib
    assert _lib.RECORD_DTYPE.itemsize == 64
    assert _lib.RECORD_DTYPE.fields["bb_min"][1] == 24 and _lib.RECORD_DTYPE.fields["rep"][1] == 48
    assert _lib.RECORD_DTYPE.fields["chunk_seq"][1] == 60
    assert _lib.PAIR_DTYPE.itemsize == 32 and _lib.GEOM_DTYPE.itemsize == 48


def test_no_cpu_fallback_without_gpu():
    from syconn_b200 import _lib
    L = _lib.load()
    if L.syk_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert L.syk_table_create(C.byref(h), 1024) == _lib.SYK_ENODEV
    from syconn_b200.extraction.find_object_properties import find_object_properties
    with pytest.raises(_lib.SykError):
        find_object_properties(np.ones((4, 4, 4), np.uint64))


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "syconn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "syk_oracle" not in txt, f


def test_shim_error_behaviour():
    from syconn_b200.extraction.block_processing_C import process_block_nonzero
    from syconn_b200.extraction.find_object_properties import detect_cs, find_object_properties, map_subcell_extract_props
    with pytest.raises(ValueError):  # reference: ValueError("Buffer dtype mismatch ...") for uint64 input
        process_block_nonzero(np.ones((5, 5, 5), np.uint64), np.ones((5, 5, 5), np.uint64), (3, 3, 3))
    with pytest.raises(AssertionError):  # block_processing_C.pyx:57
        process_block_nonzero(np.ones((5, 5, 5), np.uint32), np.ones((5, 5, 5), np.uint32), (4, 3, 3))
    with pytest.raises(ValueError):
        find_object_properties(np.ones((3, 3, 3), np.float32))
    with pytest.raises(AssertionError):  # find_object_properties_C.pyx:134-137
        map_subcell_extract_props(np.zeros((3, 3, 3), np.uint64), np.zeros((1, 3, 3, 4), np.uint64))
    with pytest.raises(ValueError):  # fused dtype must agree
        map_subcell_extract_props(np.zeros((3, 3, 3), np.uint64), np.zeros((1, 3, 3, 3), np.uint32))
    with pytest.raises(AssertionError):
        detect_cs(np.zeros((9, 9, 9), np.uint32), stencil=(3, 2, 3))


def test_new_shims_host_logic():
    """argument checks and the trivial cases of the f2 / 64-bit shims run before any device call"""
    from syconn_b200.extraction import cs_extraction_steps as ces
    from syconn_b200.extraction.find_object_properties import (convert_nvox2ratio_syntype, detect_contact_partners,
                                                               extract_cs_syntype_64bit, find_object_properties_cs_64bit,
                                                               merge_type_dicts, merge_voxel_dicts)
    vol = np.zeros((4, 5, 6), np.uint64)
    assert ces.close_contact_sites(vol, {}, 3, 2) is vol                       # no ids: untouched, no device needed
    assert ces.close_contact_sites(vol, {7: [[0, 0, 0], [1, 1, 1]]}, 0, 0) is vol   # no iterations: untouched
    with pytest.raises(ValueError):
        ces.close_contact_sites(np.zeros((4, 4, 4), np.float32), {}, 3, 2)
    ids, bb = ces._boxes({9: [[1, 2, 3], [4, 5, 6]], 2 ** 63 + 1: [[0, 0, 0], [1, 1, 1]]})
    assert ids.dtype == np.uint64 and ids.tolist() == [9, 2 ** 63 + 1]         # dict order kept, 64-bit ids exact
    assert bb.dtype == np.int32 and bb.shape == (2, 2, 3) and bb[0].tolist() == [[1, 2, 3], [4, 5, 6]]
    with pytest.raises(ValueError):
        ces.contact_site_extraction_chunk(np.zeros((30, 30, 30), np.uint64), None, None, None, (0, 0, 0), (7, 7, 3), 2)
    with pytest.raises(NotImplementedError):                                    # asymmetric windows are not provided
        detect_contact_partners(np.zeros((5, 5, 5), np.uint64), None, np.array([(-1, 2), (-1, 1), (-1, 1)]))
    with pytest.raises(NotImplementedError):                                    # dead code in the reference
        extract_cs_syntype_64bit(None, None, None, None)
    assert find_object_properties_cs_64bit(np.zeros((0, 3, 3, 2), np.uint64)) == ({}, {}, {})
    # pure-Python helpers of the facade (find_object_properties.py:272-344)
    asym, sym = convert_nvox2ratio_syntype({1: 4, 2: 10}, {1: 1}, {2: 5})
    assert sym == {1: 0.25, 2: 0} and asym == {1: 0, 2: 0.5}
    tot = [{1: 2}, {1: 3, 5: 1}]
    merge_type_dicts(tot)
    assert tot[0] == {1: 5, 5: 1}
    vx = [{1: [[0, 0, 0]]}, {1: [[1, 1, 1]], 2: np.array([[2, 2, 2]])}]
    merge_voxel_dicts(vx)
    assert vx[0] == {1: [[0, 0, 0], [1, 1, 1]], 2: [[2, 2, 2]]}
