"""Generate tests/golden/*.npz by running the REFERENCE's own code (this container only).

  * the two Cython modules compiled from /root/reference by oracle/build_ref.py (oracle/_ref/*.so)
  * /root/reference/syconn/extraction/find_object_properties.py (numba detect_seg_boundaries, detect_cs)
    imported under a stub ``syconn`` package (global_params.config with the config.yml default stencil)

Run:  python tests/golden/make_golden.py      (needs /root/reference; the fixtures are committed)
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref  # noqa: E402
from syconn_b200.synth import synth_labels  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_python():
    assert build_ref.build()
    syconn = types.ModuleType("syconn")
    syconn.__path__ = []
    gp = types.ModuleType("syconn.global_params")
    gp.config = {"cell_objects": {"cs_filtersize": [13, 13, 7]}}
    ext = types.ModuleType("syconn.extraction")
    ext.__path__ = []
    syconn.global_params = gp
    sys.modules.update({"syconn": syconn, "syconn.global_params": gp, "syconn.extraction": ext,
                        "syconn.extraction.block_processing_C": ref._load("block_processing_C"),
                        "syconn.extraction.find_object_properties_C": ref._load("find_object_properties_C")})
    spec = importlib.util.spec_from_file_location("syconn.extraction.find_object_properties",
                                                  os.path.join(REF, "syconn/extraction/find_object_properties.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["syconn.extraction.find_object_properties"] = m
    spec.loader.exec_module(m)
    return m, gp


def props_arrays(dicts):
    rc, bb, sz = dicts
    ids = np.array(sorted(sz.keys()), np.uint64)
    return (ids, np.array([sz[int(k)] for k in ids], np.int64),
            np.array([bb[int(k)] for k in ids], np.int32).reshape(-1, 2, 3),
            np.array([rc[int(k)] for k in ids], np.int32).reshape(-1, 3))


def map_arrays(md):
    rows = sorted((s, c, n) for s, d in md.items() for c, n in d.items())
    a = np.array(rows, np.uint64).reshape(-1, 3)
    return a


def noisy(vol, rng, frac, hi):
    m = rng.random(vol.shape) < frac
    v = vol.copy()
    v[m] = rng.integers(0, hi, size=int(m.sum())).astype(vol.dtype)
    return v


def main():
    fop, gp = load_reference_python()
    rng = np.random.default_rng(1234)
    out = {}

    # ---- find_object_properties -------------------------------------------------------------
    v = noisy(synth_labels((24, 20, 28), pitch=(7, 6, 5), warp_amp=3, seed=1), rng, 0.02, 9)
    out["fop_in"] = v
    for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays(ref.find_object_properties(v))):
        out["fop_" + k] = a
    vt = v.transpose(2, 1, 0)  # strided view: results follow the logical index order
    for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays(ref.find_object_properties(vt))):
        out["fopT_" + k] = a
    v32 = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    out["fop32_in"] = v32
    for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays(ref.find_object_properties(v32))):
        out["fop32_" + k] = a

    # ---- map_subcell_extract_props ----------------------------------------------------------
    cell = synth_labels((20, 24, 18), pitch=(8, 7, 6), warp_amp=2, seed=2)
    subs = np.stack([synth_labels((20, 24, 18), pitch=(5, 4, 3), warp_amp=2, seed=2, kind=1 + c, density16=3)
                     for c in range(3)])
    out["map_cell"] = cell
    out["map_subs"] = subs
    cp, sp, md = ref.map_subcell_extract_props(cell, subs)
    for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays(cp)):
        out["map_cell_" + k] = a
    for c in range(3):
        for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays((sp[0][c], sp[1][c], sp[2][c]))):
            out[f"map_sub{c}_" + k] = a
        out[f"map_pairs{c}"] = map_arrays(md[c])
        out[f"mapC_pairs{c}"] = map_arrays(ref.map_subcell_C(cell, subs)[c])

    # ---- detect_seg_boundaries / detect_cs / process_block_nonzero --------------------------------
    seg = synth_labels((30, 28, 26), pitch=(9, 8, 5), warp_amp=3, seed=3, dtype=np.uint32)
    out["cs_in"] = seg
    out["cs_bdry"] = np.asarray(fop.detect_seg_boundaries(seg))
    for st in ((13, 13, 7), (7, 7, 3), (5, 5, 3), (3, 3, 3), (1, 1, 1), (3, 5, 7)):
        gp.config["cell_objects"]["cs_filtersize"] = list(st)
        out["cs_out_%d_%d_%d" % st] = np.asarray(fop.detect_cs(seg))
    gp.config["cell_objects"]["cs_filtersize"] = [13, 13, 7]
    # tie-breaking stress: 6-label noise, every voxel a boundary
    tie = rng.integers(0, 6, size=(14, 15, 13)).astype(np.uint32)
    out["tie_in"] = tie
    for st in ((5, 5, 3), (3, 3, 3)):
        gp.config["cell_objects"]["cs_filtersize"] = list(st)
        out["tie_out_%d_%d_%d" % st] = np.asarray(fop.detect_cs(tie))
    gp.config["cell_objects"]["cs_filtersize"] = [13, 13, 7]
    # many distinct ids per window (> 32) and large ids
    many = rng.integers(0, 2 ** 32 - 1, size=(12, 12, 10), dtype=np.uint64).astype(np.uint32)
    many[rng.random(many.shape) < 0.5] = 7
    out["many_in"] = many
    gp.config["cell_objects"]["cs_filtersize"] = [5, 5, 3]
    out["many_out_5_5_3"] = np.asarray(fop.detect_cs(many))
    gp.config["cell_objects"]["cs_filtersize"] = [13, 13, 7]
    # forced edge mask (incl. background centres): process_block_nonzero alone
    edges = np.ones(tie.shape, np.uint32)
    out["pbn_forced_7_7_3"] = np.asarray(ref.process_block_nonzero(edges, tie, (7, 7, 3)))
    # strided input (x fastest in memory)
    segF = np.ascontiguousarray(seg.transpose(2, 1, 0)).transpose(2, 1, 0)
    assert np.array_equal(np.asarray(fop.detect_cs(segF)), out["cs_out_13_13_7"])

    # ---- extract_cs_syntype ("next" row f1) ----------------------------------------------------------------------
    cs = out["cs_out_7_7_3"].copy()                               # a real contact volume (uint64 packed ids)
    syn = ((rng.random(cs.shape) < 0.35) * rng.integers(1, 3, size=cs.shape)).astype(np.uint8)
    asym = rng.integers(0, 3, size=cs.shape).astype(np.uint8)
    sym = rng.integers(0, 3, size=cs.shape).astype(np.uint8)
    off = np.array([7, -4, 120])
    r = ref.extract_cs_syntype(cs, syn, asym, sym, off)
    out["syn_cs"], out["syn_mask"], out["syn_asym"], out["syn_sym"], out["syn_off"] = cs, syn, asym, sym, off
    for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays(r[0])):
        out["syn_csprops_" + k] = a
    for k, a in zip(("ids", "sizes", "bbox", "rep"), props_arrays(r[1])):
        out["syn_synprops_" + k] = a
    out["syn_asymcnt"] = np.array(sorted(dict(r[2]).items()), np.uint64).reshape(-1, 2)
    out["syn_symcnt"] = np.array(sorted(dict(r[3]).items()), np.uint64).reshape(-1, 2)
    vox = dict(r[4])
    out["syn_vox_ids"] = np.array(sorted(vox), np.uint64)
    out["syn_vox_len"] = np.array([len(vox[int(k)]) for k in out["syn_vox_ids"]], np.int64)
    out["syn_vox_xyz"] = np.array([c for k in out["syn_vox_ids"] for c in vox[int(k)]], np.int64).reshape(-1, 3)

    np.savez_compressed(os.path.join(OUT, "hotpath_golden.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(OUT, "hotpath_golden.npz")), "bytes")
    main64(fop, gp, rng)
    main_morph()


MORPH_CASES = {  # op lists of syconn/handler/config.yml:130-140 plus single ops / repeated runs
    "mi": ["binary_opening", "binary_closing"] + ["binary_erosion"] * 4,
    "sj": ["binary_opening", "binary_closing", "binary_erosion"],
    "er": ["binary_dilation"] * 3 + ["binary_erosion"] * 3,
    "closing2": ["binary_closing", "binary_closing"],
    "dilation1": ["binary_dilation"],
    "opening3": ["binary_opening"] * 3,
}


def morph_input(case_index, density, scaling):
    """0/1 volume of a golden morphology case: random voxels thickened into blobs, empty margin on one side."""
    import scipy.ndimage
    rng = np.random.default_rng(100 + case_index)
    v = rng.random((30, 27, 22)) < density * 0.004
    v = scipy.ndimage.binary_dilation(v, iterations=3)
    v[:2] = False
    return v.astype(np.uint8)


def main_morph():
    """apply_morphological_operations / get_aniso_struct of the reference's syconn/proc/image.py (imported under a stub
    ``syconn.proc`` package) -> morph_golden.npz.  Note: with this image's scipy (1.18.1) the reference's literal
    ``iterations=n`` calls corrupt the heap when an object's box is smaller than the 5 x 5 x 3 element (reproducible with
    scipy alone); the golden volumes keep their foreground boxes larger than that, and the oracle evaluates the
    iterations one at a time."""
    import logging
    proc = types.ModuleType("syconn.proc")
    proc.__path__ = []
    proc.log_proc = logging.getLogger("reference")
    sys.modules["syconn.proc"] = proc
    spec = importlib.util.spec_from_file_location("syconn.proc.image", os.path.join(REF, "syconn/proc/image.py"))
    img = importlib.util.module_from_spec(spec)
    sys.modules["syconn.proc.image"] = img
    spec.loader.exec_module(img)
    out = {}
    scalings = [(10, 10, 20), (9, 9, 20), (10, 10, 10), (10, 10, 45)]
    for sc in scalings:
        out["struct_%d_%d_%d" % sc] = img.get_aniso_struct(np.array(sc)).astype(np.uint8)
    i = 0
    for name, ops in MORPH_CASES.items():
        for density, sc in ((1.0, scalings[0]), (2.5, scalings[1]), (5.0, scalings[2])):
            v = morph_input(i, density, sc)
            st = img.get_aniso_struct(np.array(sc))
            res = img.apply_morphological_operations(v.copy(), list(ops), mop_kwargs=dict(structure=st))
            out["in_%s_%d" % (name, i)] = np.packbits(v)
            out["out_%s_%d" % (name, i)] = np.packbits(res)
            out["sc_%s_%d" % (name, i)] = np.array(sc)
            i += 1
    np.savez_compressed(os.path.join(OUT, "morph_golden.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(OUT, "morph_golden.npz")), "bytes")


def pair_props_arrays(dicts):
    rc, bb, sz = dicts
    keys = sorted((int(k0), int(k1)) for k0, d in sz.items() for k1 in d.keys())
    return (np.array(keys, np.uint64).reshape(-1, 2), np.array([int(sz[a][b]) for a, b in keys], np.int64),
            np.array([np.asarray(bb[a][b]) for a, b in keys], np.int64).reshape(-1, 2, 3),
            np.array([np.asarray(rc[a][b]) for a, b in keys], np.int64).reshape(-1, 3))


def main64(fop, gp, rng):
    """64-bit variants (numba): detect_cs_64bit / find_object_properties_cs_64bit -> cs64_golden.npz"""
    out = {}
    seg = synth_labels((26, 24, 22), pitch=(9, 8, 5), warp_amp=3, seed=5)
    seg[seg != 0] += np.uint64(3) << np.uint64(40)            # ids beyond 32 bits
    out["seg_in"] = seg
    for st in ((13, 13, 7), (7, 7, 3), (5, 5, 3)):
        gp.config["cell_objects"]["cs_filtersize"] = list(st)
        out["seg_out_%d_%d_%d" % st] = np.asarray(fop.detect_cs_64bit(seg))
    tie = rng.integers(0, 6, size=(13, 14, 12)).astype(np.uint64)  # every window full of ties
    tie[tie != 0] += np.uint64(1) << np.uint64(33)
    out["tie_in"] = tie
    for st in ((5, 5, 3), (3, 3, 3)):
        gp.config["cell_objects"]["cs_filtersize"] = list(st)
        out["tie_out_%d_%d_%d" % st] = np.asarray(fop.detect_cs_64bit(tie))
    gp.config["cell_objects"]["cs_filtersize"] = [13, 13, 7]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):               # the numba function prints debug lines
        pr = fop.find_object_properties_cs_64bit(out["seg_out_7_7_3"])
        pt = fop.find_object_properties_cs_64bit(out["tie_out_5_5_3"])
    for tag, r in (("seg", pr), ("tie", pt)):
        for k, a in zip(("keys", "sizes", "bbox", "rep"), pair_props_arrays(r)):
            out[f"props_{tag}_" + k] = a
    np.savez_compressed(os.path.join(OUT, "cs64_golden.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(OUT, "cs64_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
