"""the module-level drop-in.  The reference's callers import
``syconn.extraction.{block_processing_C, find_object_properties_C, find_object_properties}``; shadowing those names in
``sys.modules`` with the shims of ``syconn_b200.extraction`` must be all it takes.

* CPU (this container, where /root/reference exists): the reference's OWN ``find_object_properties.py`` is executed with the two
  Cython module names shadowed by the shims (mode 1 of INTEGRATION.md) -- its ``from .block_processing_C import ...`` lines
  bind the shims, and calling through it without a GPU fails loudly (no CPU fallback).
* GPU box (no /root/reference there): the same shadowing, driven the way the reference's callers and its facade do --
  ``detect_seg_boundaries`` -> ``.astype(uint32)`` -> ``process_block_nonzero(edges, arr, stencil)``
  (find_object_properties.py:466-472) -- against the golden vectors that the reference itself produced.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

REF_FACADE = "/root/reference/syconn/extraction/find_object_properties.py"
NAMES = ("syconn", "syconn.global_params", "syconn.extraction", "syconn.extraction.block_processing_C",
         "syconn.extraction.find_object_properties_C", "syconn.extraction.find_object_properties")


@pytest.fixture
def shadowed(monkeypatch):
    """sys.modules as INTEGRATION.md section 1 sets it up (plus the stub package modules that stand in for the parts of
    SyConn that are not installed here)."""
    import syconn_b200.extraction.block_processing_C as bpc
    import syconn_b200.extraction.find_object_properties_C as fopc
    for n in NAMES:
        monkeypatch.delitem(sys.modules, n, raising=False)
    syconn = types.ModuleType("syconn")
    syconn.__path__ = []
    gp = types.ModuleType("syconn.global_params")
    gp.config = {"cell_objects": {"cs_filtersize": [13, 13, 7]}}
    ext = types.ModuleType("syconn.extraction")
    ext.__path__ = []
    syconn.global_params, syconn.extraction = gp, ext
    for n, m in (("syconn", syconn), ("syconn.global_params", gp), ("syconn.extraction", ext),
                 ("syconn.extraction.block_processing_C", bpc), ("syconn.extraction.find_object_properties_C", fopc)):
        monkeypatch.setitem(sys.modules, n, m)
    return bpc, fopc, gp


@pytest.mark.skipif(not os.path.exists(REF_FACADE), reason="needs the reference checkout (build container only)")
def test_reference_facade_binds_the_shims(shadowed):
    pytest.importorskip("numba")
    bpc, fopc, gp = shadowed
    spec = importlib.util.spec_from_file_location("syconn.extraction.find_object_properties", REF_FACADE)
    facade = importlib.util.module_from_spec(spec)
    sys.modules["syconn.extraction.find_object_properties"] = facade
    spec.loader.exec_module(facade)                      # the reference's own source, untouched
    assert facade.process_block_nonzero is bpc.process_block_nonzero
    assert facade.find_object_properties is fopc.find_object_properties
    assert facade.map_subcell_extract_props is fopc.map_subcell_extract_props
    seg = np.zeros((20, 20, 12), np.uint32)
    seg[:, :10], seg[:, 10:] = 5, 9
    import torch
    if not torch.cuda.is_available():                    # no GPU here: the reference's detect_cs reaches libsyk and fails loudly
        from syconn_b200._lib import SykError
        with pytest.raises(SykError):
            facade.detect_cs(seg)
    else:
        from oracle import oracle
        assert np.array_equal(np.asarray(facade.detect_cs(seg)), oracle.detect_cs(seg, (13, 13, 7)))


@pytest.mark.gpu
def test_shadowed_modules_serve_the_reference_call_sequence(shadowed, golden):
    """what find_object_properties.py:458-472 and cs_extraction_steps.py:385-391,439 do, through the shadowed module names"""
    import syconn_b200.extraction.find_object_properties as fop
    sys.modules["syconn.extraction.find_object_properties"] = fop      # mode 2: the facade too
    saved = list(fop.global_params.config["cell_objects"]["cs_filtersize"])
    try:
        bpc = importlib.import_module("syconn.extraction.block_processing_C")
        facade = importlib.import_module("syconn.extraction.find_object_properties")
        gp = importlib.import_module("syconn.global_params")
        seg = golden["cs_in"]
        assert np.array_equal(facade.detect_seg_boundaries(seg), golden["cs_bdry"].astype(bool))   # the reference's numba mask
        for st in ((13, 13, 7), (7, 7, 3), (5, 5, 3)):
            gp.config["cell_objects"]["cs_filtersize"] = list(st)
            fop.global_params.config["cell_objects"]["cs_filtersize"] = list(st)
            want = golden["cs_out_%d_%d_%d" % st]                                        # produced by the reference itself
            edges = facade.detect_seg_boundaries(seg).astype(np.uint32)                 # mode 1: the caller's own mask ...
            via_edges = np.asarray(bpc.process_block_nonzero(edges, seg.astype(np.uint32), st))   # ... and the shadowed Cython name
            fused = facade.detect_cs(seg.astype(np.uint32))
            assert np.array_equal(via_edges, fused)
            assert np.array_equal(fused, want)
            rc, bb, sz = facade.find_object_properties(fused)
            assert set(sz) == set(np.unique(fused).tolist()) - {0}
    finally:
        fop.global_params.config["cell_objects"]["cs_filtersize"] = saved
        sys.modules.pop("syconn.extraction.find_object_properties", None)
