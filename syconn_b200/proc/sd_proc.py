"""COMPATIBILITY GLUE, not the compute path: host-side mirrors of the dict-level merge helpers of ``syconn.proc.sd_proc`` that
the extraction workers call around the hot path (same names, argument meaning and in-place behaviour), plus converters between the reference's dict structures and
the array/record form of the device pipeline (``syconn_b200.chunked``).

These functions only re-arrange per-chunk RESULTS in Python, exactly where the reference does it in Python; all voxel work
stays in libsyk.  On the device-resident path the same reductions run as ``syk_table_merge_records`` / ``syk_pairs_merge``
(``chunked.ExtractionPipeline``); the functions here exist so that code written against the reference's worker API keeps
working when it merges the dicts returned by the shims.
"""
from collections import defaultdict
from typing import Dict, List, Optional

import numpy as np

from .._lib import PAIR_DTYPE, RECORD_DTYPE


def _shift(coord, offset):
    return [int(c) + int(o) for c, o in zip(coord, offset)]


def merge_prop_dicts(prop_dicts: List[List[dict]], offset: Optional[np.ndarray] = None):
    """syconn/proc/sd_proc.py:1248-1273.  Fold the property triples ``[rep_coords, bounding_boxes, sizes]`` of
    ``prop_dicts[1:]`` into ``prop_dicts[0]`` (in place, nothing is returned).  Per id: the representative coordinate of
    the LATEST triple wins (:1261), every bounding box is appended to the id's list -- the accumulator's second dict must
    map to lists, e.g. ``defaultdict(list)`` (:1268) -- and sizes add up.  With ``offset`` the coordinates of the incoming
    triples are translated first; like the reference, the incoming representative coordinates are rewritten in place."""
    acc_rep, acc_boxes, acc_size = prop_dicts[0]
    for rep, boxes, sizes in prop_dicts[1:]:
        if not rep:
            continue
        if offset is not None:
            for key in rep:
                rep[key] = _shift(rep[key], offset)
        acc_rep.update(rep)
        for key, box in boxes.items():
            acc_boxes[key].append(box if offset is None else [_shift(box[0], offset), _shift(box[1], offset)])
        for key, n_vox in sizes.items():
            acc_size[key] = acc_size[key] + n_vox if key in acc_size else n_vox


def merge_map_dicts(map_dicts: List[Dict[int, Dict[int, int]]]):
    """syconn/proc/sd_proc.py:1300-1322.  Fold ``{sub_id: {cell_id: n_overlap_voxels}}`` dicts into ``map_dicts[0]`` in
    place: counts of the same (organelle, cell) pair add up; an organelle seen for the first time adopts the incoming
    inner dict as is (no copy, like the reference)."""
    acc = map_dicts[0]
    for incoming in map_dicts[1:]:
        for sub_id, per_cell in incoming.items():
            mine = acc.get(sub_id)
            if mine is None:
                acc[sub_id] = per_cell
                continue
            for cell_id, n_vox in per_cell.items():
                mine[cell_id] = mine.get(cell_id, 0) + n_vox


def convert_nvox2ratio_mapdict(map_dc):
    """syconn/proc/sd_proc.py:1275-1285: overlap voxel counts -> fractions of the organelle's overlapping voxels (in place)."""
    for per_cell in map_dc.values():
        total = np.sum(list(per_cell.values()))
        for cell_id in per_cell:
            per_cell[cell_id] = per_cell[cell_id] / total


def invert_mdc(mapping_dict):
    """syconn/proc/sd_proc.py:1288-1297: ``{sub_id: {cell_id: v}}`` -> ``{cell_id: {sub_id: v}}``."""
    inverted = {}
    for outer, inner in mapping_dict.items():
        for key, value in inner.items():
            inverted.setdefault(key, {})[outer] = value
    return inverted


def new_prop_dicts():
    """the accumulator the workers start from (``[{}, defaultdict(list), {}]``, sd_proc.py:605, cs_extraction_steps.py:370)"""
    return [{}, defaultdict(list), {}]


# ---------------------------------------------------------------------------------------------- dicts <-> arrays
def reduced_to_prop_dicts(red):
    """Output of ``chunked.reduce_records`` -> the merged dicts a reference worker ends up with after
    ``merge_prop_dicts`` over its chunks: ``rep_coords {id: [x, y, z]}``, ``bounding_boxes {id: [bb_chunk0, bb_chunk1, ...]}``
    (chunk order), ``sizes {id: n}``."""
    ids = red["id"].tolist()
    rc = dict(zip(ids, red["rep_coord"].tolist()))
    bb = defaultdict(list)
    for k, b in zip(ids, red["bbs"]):
        bb[k] = b.tolist()
    sz = dict(zip(ids, red["size"].tolist()))
    return [rc, bb, sz]


def reduced_to_map_dict(red):
    """Output of ``chunked.reduce_pairs`` -> ``{sub_id: {cell_id: count}}`` (the result of ``merge_map_dicts``)."""
    out = {}
    for s, c, n in zip(red["sub_id"].tolist(), red["cell_id"].tolist(), red["count"].tolist()):
        out.setdefault(s, {})[c] = n
    return out


def prop_dicts_to_records(prop_dicts, chunk_seq=0):
    """``[rep_coords, bounding_box, sizes]`` of ONE chunk call (bounding boxes not yet appended) -> ``syk_record_t`` array
    that the host-side ``reduce_records`` accepts.  ``rep_key`` is left 0 (only the decoded ``rep`` is carried), so these
    records are NOT valid input for the device fold ``IdTable.merge_records``, which recomputes ``rep`` from ``rep_key``."""
    rc, bb, sz = prop_dicts
    rec = np.zeros(len(sz), RECORD_DTYPE)
    for i, k in enumerate(sz):
        rec["id"][i] = k
        rec["count"][i] = sz[k]
        rec["bb_min"][i] = bb[k][0]
        rec["bb_max"][i] = bb[k][1]
        rec["rep"][i] = rc[k]
    rec["chunk_seq"] = chunk_seq
    return rec


def map_dict_to_pairs(map_dc):
    """``{sub_id: {cell_id: count}}`` -> ``syk_pair_t`` array."""
    rows = [(s, c, n) for s, d in map_dc.items() for c, n in d.items()]
    out = np.zeros(len(rows), PAIR_DTYPE)
    if rows:
        a = np.array(rows, np.uint64)
        out["sub_id"], out["cell_id"], out["count"] = a[:, 0], a[:, 1], a[:, 2]
    return out


# ---------------------------------------------------------------------------------------------- dataset_analysis caches
def cell_mapping_attributes(cell_ids, pairs_red, organelle_ids, organelle_sizes):
    """Per cell supervoxel the organelle ids it overlaps and the overlap ratios, as the writers store them
    (syconn/proc/sd_proc.py:1064-1084 -> ``mapping_{organelle}_ids`` / ``mapping_{organelle}_ratios``, :1172-1177): the
    ratio is ``overlap voxels / total voxels of the organelle object``; organelles that are not part of the organelle
    dataset (``organelle_ids``, e.g. removed by the size threshold) are dropped (:1072-1074).

    ``pairs_red`` is the output of ``chunked.reduce_pairs`` (arrays ``sub_id, cell_id, count``).  Returns two object
    arrays aligned with ``cell_ids``: lists of organelle ids (ascending) and lists of float ratios."""
    cell_ids = np.asarray(cell_ids, np.uint64)
    organelle_ids = np.asarray(organelle_ids, np.uint64)
    organelle_sizes = np.asarray(organelle_sizes, np.int64)
    order = np.argsort(organelle_ids)
    sub, cell, cnt = pairs_red["sub_id"], pairs_red["cell_id"], pairs_red["count"]
    pos = np.searchsorted(organelle_ids[order], sub)
    pos[pos >= len(order)] = 0
    known = (organelle_ids[order][pos] == sub) if len(order) else np.zeros(len(sub), bool)
    sub, cell, cnt, pos = sub[known], cell[known], cnt[known], pos[known]
    ratio = cnt / organelle_sizes[order][pos] if len(sub) else np.zeros(0)
    o = np.lexsort((sub, cell))
    sub, cell, ratio = sub[o], cell[o], ratio[o]
    ids_out = np.empty(len(cell_ids), dtype=object)
    rat_out = np.empty(len(cell_ids), dtype=object)
    lo = np.searchsorted(cell, cell_ids, side="left")
    hi = np.searchsorted(cell, cell_ids, side="right")
    sub_l, ratio_l = sub.tolist(), ratio.tolist()
    for i, (a, b) in enumerate(zip(lo.tolist(), hi.tolist())):
        ids_out[i] = sub_l[a:b]
        rat_out[i] = ratio_l[a:b]
    return ids_out, rat_out


def write_dataset_analysis_cache(sd_path, red, extra=None):
    """Write the ``.npy`` column caches that ``dataset_analysis`` produces and ``SegmentationDataset`` reads back
    (syconn/proc/sd_proc.py:138-155, :244-251; reps/segmentation.py:1601-1621, :1765-1784): ``ids.npy`` uint64,
    ``sizes.npy`` int64, ``bounding_boxs.npy`` int32 [N, 2, 3], ``rep_coords.npy`` int32 [N, 3], plus one
    ``{attribute}s.npy`` per entry of ``extra`` (e.g. ``mapping_mi_ids`` / ``mapping_mi_ratios`` object arrays from
    ``cell_mapping_attributes``).  ``red`` is the output of ``chunked.reduce_records``.  Returns the written paths."""
    import os
    os.makedirs(sd_path, exist_ok=True)
    n = len(red["id"])
    cols = {"id": np.asarray(red["id"], np.uint64), "size": np.asarray(red["size"], np.int64),
            "bounding_box": np.asarray(red["bounding_box"], np.int32).reshape(n, 2, 3),
            "rep_coord": np.asarray(red["rep_coord"], np.int32).reshape(n, 3)}
    for k, v in (extra or {}).items():
        assert len(v) == n, f"attribute {k} has {len(v)} entries for {n} objects"
        cols[k] = v
    paths = []
    for k, v in cols.items():
        p = os.path.join(sd_path, f"{k}s.npy")
        np.save(p, v)
        paths.append(p)
    return paths


# ---------------------------------------------------------------------------------------------- storage writers (f3)
def subfold_from_ix(ix, n_folders):
    """reps/rep_helper.py:143-163 (``use_new_subfold``): the storage sub-folder of object ``ix``."""
    assert n_folders % 10 == 0
    order = int(np.log10(n_folders))
    ix = int(int(ix) // 1e3 % n_folders)  # float floor division like the reference (div_base = 1e3): ids >= 2^53 round
    id_str = "{num:0{w}d}".format(num=ix, w=order)
    subfold = "/"
    for idx in range(0, order, 2):
        subfold += "%s/" % id_str[idx: idx + 2]
    return subfold


def write_segmentation_objects(sd_path, red, mapping=None, min_obj_vx=1, n_folders_fs=10000, voxeldata_path=None,
                               extra_attrs=None):
    """The writer loop of ``_write_props_to_sc_thread`` / ``_write_props_to_sv_thread`` (syconn/proc/sd_proc.py:788-1000,
    :1003-1215) on the reduced records of ``chunked.reduce_records``: one ``attr_dict.pkl`` (``AttributeDict``) and one
    ``voxel.pkl`` (``VoxelStorageDyn``, voxel_mode=False) per storage folder ``<sd_path>/so_storage_<n>/<subfold>/``
    (reps/segmentation.py:336-416).  Per object with ``size >= min_obj_vx`` (:929-931): ``rep_coord`` int32 [3] (:939),
    ``bounding_box`` [2, 3] = min / max over the object's per-chunk boxes (:940-944), ``size`` (:945), the per-chunk boxes
    themselves as the object's voxel index ``voxel_dc[id] = bbs`` with its size and rep coord (:946-950), and

      * organelle objects: ``mapping_ids`` / ``mapping_ratios`` from ``mapping = {sub_id: {cell_id: count}}``, ratios
        normalised by the object's size (:929-937);
      * cell supervoxels: ``extra_attrs = {name: object array aligned with red['id']}``, e.g. ``mapping_mi_ids`` /
        ``mapping_mi_ratios`` from ``cell_mapping_attributes`` (:1166-1171).

    Returns the list of written folder paths."""
    import os
    from ..backend.storage import AttributeDict, VoxelStorageDyn
    ids = np.asarray(red["id"], np.uint64)
    by_folder = defaultdict(list)
    for i, k in enumerate(ids.tolist()):
        by_folder[subfold_from_ix(k, n_folders_fs)].append(i)
    base = os.path.join(sd_path, "so_storage_%d" % n_folders_fs)
    written = []
    for sub, rows in by_folder.items():
        folder = base + sub
        os.makedirs(folder, exist_ok=True)
        attr_dc = AttributeDict(folder + "attr_dict.pkl", read_only=False, disable_locking=True)
        voxel_dc = VoxelStorageDyn(folder + "voxel", voxel_mode=False, read_only=False, disable_locking=True,
                                   voxeldata_path=voxeldata_path)
        for i in rows:
            k = int(ids[i])
            size = int(red["size"][i])
            if size < min_obj_vx:
                continue
            if mapping is not None:
                m = mapping.get(k)
                attr_dc[k]["mapping_ids"] = list(m.keys()) if m else []
                attr_dc[k]["mapping_ratios"] = [v / size for v in m.values()] if m else []
            for name, col in (extra_attrs or {}).items():
                attr_dc[k][name] = col[i]
            rp = np.array(red["rep_coord"][i], dtype=np.int32)
            bbs = np.asarray(red["bbs"][i]).reshape(-1, 2, 3).astype(np.int64)   # np.concatenate of lists of lists (:940)
            attr_dc[k]["rep_coord"] = rp
            attr_dc[k]["bounding_box"] = np.array([bbs[:, 0].min(axis=0), bbs[:, 1].max(axis=0)])
            attr_dc[k]["size"] = size
            voxel_dc[k] = bbs
            voxel_dc.increase_object_size(k, size)
            voxel_dc.set_object_repcoord(k, rp)
        attr_dc.push()
        voxel_dc.push()
        written.append(folder)
    return written


def reduced_from_device(final_records, voxel_index):
    """Bridge from the device pipeline to the writers: ``final_records`` is the structured array of one kind returned by
    ``ExtractionPipeline.reduce_on_device`` (``device.records_numpy``), ``voxel_index`` the ``(ids, start, boxes)`` triple of
    ``ExtractionPipeline.voxel_index`` (device tensors or arrays).  Returns the dict ``write_segmentation_objects`` /
    ``write_dataset_analysis_cache`` take (``id, size, bounding_box, rep_coord, bbs``), sorted by id."""
    ids, start, boxes = (np.asarray(t.cpu().numpy() if hasattr(t, "cpu") else t) for t in voxel_index)
    ids = ids.view(np.uint64)
    o = np.argsort(final_records["id"])
    rec = final_records[o]
    v = np.argsort(ids)
    assert np.array_equal(ids[v], rec["id"]), "voxel index and final records describe different id sets"
    bbs = [boxes[start[i]:start[i + 1]].astype(np.int64) for i in v.tolist()]
    return dict(id=rec["id"].copy(), size=rec["count"].astype(np.int64),
                bounding_box=np.stack([rec["bb_min"], rec["bb_max"]], axis=1).astype(np.int32),
                rep_coord=rec["rep"].astype(np.int32), bbs=bbs)
