"""Organelle instance segmentation, first slice (row f4): threshold -> connected components per chunk on the GPU ->
unique labels -> stitching of the components that cross chunk borders -> merged labels.

Mirrors the array-level core of ``syconn/extraction/object_extraction_steps.py``:

  * ``object_segmentation_chunk``   <- ``_object_segmentation_thread`` :204-366: threshold (:302-303), the morphology hook
                                      (:312, :354-356, ``proc/image.py``) and ``scipy.ndimage.label`` (:350-358)
  * ``watershed_seeds_chunk``       <- the first half of the watershed branch (:313-343): mask from the ops before the
                                      first erosion, markers = components of the eroded mask with the ``min_seed_vx``
                                      clean-up.  The distance transform (vigra) and ``skimage.segmentation.watershed``
                                      (:345-348) are not built -- neither library exists here to pin them -- and neither
                                      is the Gaussian smoothing (vigra, :297-298)
  * ``make_unique_labels``          <- :369-443 (per-chunk label offsets = running sum of the component counts)
  * ``make_stitch_list``            <- :446-617 (co-located label pairs in the 2 * stitch_overlap slab around the border to
                                      the +x / +y / +z neighbour; ``overlap_thresh`` = 0)
  * ``make_merge_list``             <- :620-656 (connected components of the pair graph; the reference keeps an arbitrary
                                      member of every class, here the smallest id)
  * ``apply_merge_list``            <- :659-737 (crop the overlap, map the ids)

The voxel work (labelling, pair counting, id mapping) runs in libsyk on device tensors; the pair graph is a host-side
union-find on a few thousand pairs, where the reference uses networkx.
"""
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import device as dev
from ..chunked import ChunkPlan


def _threshold_mask(prob: torch.Tensor, threshold: int) -> torch.Tensor:
    """``np.array(prob > threshold, dtype=np.uint8)`` (:302-303; threshold 0 leaves the data as it is in the reference --
    it must then already be 0/1), laid out like ``prob``."""
    return (prob > threshold).to(torch.uint8) if threshold != 0 else prob


def object_segmentation_chunk(prob: torch.Tensor, threshold: int = 0, morph_ops: Sequence[str] = (), structure=None
                              ) -> Tuple[torch.Tensor, int]:
    """One (overlap-extended) chunk of one organelle type: threshold, the type's ``extract_morph_op`` list (without
    erosion -- with ``binary_erosion`` the reference switches to the watershed branch, see ``watershed_seeds_chunk``),
    connected components -> (int32 labels numbered like ``scipy.ndimage.label``, number of components)."""
    if "binary_erosion" in morph_ops:
        raise NotImplementedError("op lists with binary_erosion select the reference's watershed branch; "
                                  "watershed_seeds_chunk builds its mask and markers")
    if len(morph_ops) == 0:
        return dev.label_components(prob, threshold)
    from ..proc.image import apply_morphological_operations
    mask = _threshold_mask(prob, threshold)
    if mask is prob:
        mask = prob.clone()  # the reference hands a copy to the morphology (:354)
    apply_morphological_operations(mask, list(morph_ops), dict(structure=structure))
    return dev.label_components(mask, 0)


def seed_cleanup_table(sizes: np.ndarray, min_size: int) -> np.ndarray:
    """Label table of the marker clean-up (:325-343): markers smaller than ``min_size`` go to 0 and the holes they leave in
    the id space are refilled from the top -- the largest kept id takes the smallest freed id, and so on, while the kept id
    is still larger than the freed one.  ``sizes[l]`` = voxel count of marker l (index 0 unused)."""
    n = len(sizes) - 1
    ids = np.arange(1, n + 1)
    present = sizes[1:] > 0
    small = ids[present & (sizes[1:] < min_size)]
    kept = ids[present & (sizes[1:] >= min_size)]
    table = np.arange(n + 1, dtype=np.int32)
    table[small] = 0
    k = len(kept) - 1
    for freed in small:
        # (the reference's kept list also holds the background id 0 at its front; reaching it ends the loop there as well)
        if k < 0 or freed > kept[k]:
            break
        table[kept[k]] = freed
        k -= 1
    return table


def watershed_seeds_chunk(prob: torch.Tensor, threshold: int, morph_ops: Sequence[str], structure=None, min_seed_vx: int = 1
                          ) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Mask and markers of the watershed branch for an op list that contains ``binary_erosion`` ->
    (uint8 mask after the ops before the first erosion, int32 markers after the clean-up, largest marker id)."""
    from ..proc.image import apply_morphological_operations
    ops = list(morph_ops)
    first = ops.index("binary_erosion")
    mask = _threshold_mask(prob, threshold)
    if mask is prob:
        mask = prob.clone()
    apply_morphological_operations(mask, ops[:first], dict(structure=structure))
    eroded = apply_morphological_operations(mask.clone(), ops[first:], dict(structure=structure))
    markers, n = dev.label_components(eroded, 0)
    if min_seed_vx > 1 and n > 0:
        table = dev.IdTable(max(2 * n, 1 << 12))
        try:
            dev.find_object_properties(table, markers)
            rec = dev.records_numpy(table.export(dev.geoms([(0, 0, 0)], [tuple(markers.shape)])))
        finally:
            table.close()
        sizes = np.zeros(n + 1, np.int64)
        sizes[rec["id"].astype(np.int64)] = rec["count"]
        lut = seed_cleanup_table(sizes, int(min_seed_vx))
        dev.label_map(markers, torch.from_numpy(lut).to(markers.device))
        n = int(lut.max())
    return mask, markers, n


def make_unique_labels(counts: Sequence[int]) -> np.ndarray:
    """Label offset of every chunk: chunk k's label l becomes ``l + offsets[k]`` (labels stay 0 for background)."""
    return np.concatenate([[0], np.cumsum(np.asarray(counts, np.int64))[:-1]]).astype(np.int64)


def _neighbour(plan: ChunkPlan, seq: int, axis: int):
    off = list(plan.offsets[seq])
    off[axis] += plan.sizes[seq][axis]
    try:
        return plan.offsets.index(tuple(off))
    except ValueError:
        return -1


def make_stitch_list(plan: ChunkPlan, labels: Dict[int, torch.Tensor], offsets: np.ndarray, overlap: Sequence[int],
                     stitch_overlap: Sequence[int]) -> np.ndarray:
    """Unique (id_a, id_b) pairs of components that share a voxel in the stitch slab between a chunk and its neighbour in
    +x, +y or +z.  ``labels[seq]`` is the label block of chunk ``seq`` extended by ``overlap`` on every side."""
    table = dev.PairTable(1 << 16)
    rows = []
    for seq in labels:
        a = labels[seq]
        for axis in range(3):
            nb = _neighbour(plan, seq, axis)
            if nb < 0 or nb not in labels:
                continue
            b = labels[nb]
            ov, so = int(overlap[axis]), int(stitch_overlap[axis])
            sl_a = [slice(None)] * 3
            sl_b = [slice(None)] * 3
            sl_a[axis] = slice(a.shape[axis] - ov - so, a.shape[axis] - ov + so)
            sl_b[axis] = slice(ov - so, ov + so)
            ra, rb = a[tuple(sl_a)], b[tuple(sl_b)]
            if tuple(ra.shape) != tuple(rb.shape):   # ragged chunk grid: compare the common part
                m = [min(x, y) for x, y in zip(ra.shape, rb.shape)]
                ra, rb = ra[:m[0], :m[1], :m[2]], rb[:m[0], :m[1], :m[2]]
            table.clear()
            dev.label_overlap_pairs(table, ra, rb, int(offsets[seq]), int(offsets[nb]))
            p = dev.pairs_numpy(table.export())
            if len(p):
                rows.append(np.stack([p["sub_id"], p["cell_id"]], axis=1))
    table.close()
    if not rows:
        return np.empty((0, 2), np.int64)
    return np.unique(np.concatenate(rows).astype(np.int64), axis=0)


def make_merge_list(stitch_list: np.ndarray, max_label: int) -> np.ndarray:
    """id_changer[id] = representative (smallest id) of the id's class in the stitch graph; identity elsewhere."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    changer = np.arange(max_label + 1, dtype=np.int64)
    if len(stitch_list) == 0:
        return changer
    nodes, inv = np.unique(stitch_list.reshape(-1), return_inverse=True)
    e = inv.reshape(-1, 2)
    g = coo_matrix((np.ones(len(e), np.int8), (e[:, 0], e[:, 1])), shape=(len(nodes), len(nodes)))
    _, comp = connected_components(g, directed=False)
    rep = np.full(comp.max() + 1, np.iinfo(np.int64).max, np.int64)
    np.minimum.at(rep, comp, nodes)
    changer[nodes] = rep[comp]
    return changer


def apply_merge_list(label_block: torch.Tensor, offset: int, changer: torch.Tensor, overlap: Sequence[int]) -> torch.Tensor:
    """Crop ``overlap`` from every side of one chunk's label block and map its (unique) ids through ``changer``."""
    sl = tuple(slice(int(o), label_block.shape[i] - int(o)) for i, o in enumerate(overlap))
    lab = label_block[sl].to(torch.int64)
    lab = torch.where(lab != 0, lab + int(offset), lab)
    return changer[lab]


def extract_components_chunked(load_block: Callable[[Sequence[int], Sequence[int]], torch.Tensor], plan: ChunkPlan,
                               threshold: int = 0, overlap=(1, 1, 1), stitch_overlap=(1, 1, 1),
                               morph_ops: Sequence[str] = (), structure=None):
    """The whole first slice over a chunk grid.  ``load_block(offset, size)`` returns the probability block at ``offset``
    (may reach outside the volume: the caller pads with zeros, as the KnossosDataset does).  Returns
    ``({seq: int64 label tensor of the chunk (overlap cropped, stitched ids)}, n_objects)``."""
    labels, counts = {}, []
    for seq in range(len(plan)):
        off = [plan.offsets[seq][i] - overlap[i] for i in range(3)]
        size = [plan.sizes[seq][i] + 2 * overlap[i] for i in range(3)]
        lab, n = object_segmentation_chunk(load_block(off, size), threshold, morph_ops, structure)
        labels[seq] = lab
        counts.append(n)
    offsets = make_unique_labels(counts)
    stitch = make_stitch_list(plan, labels, offsets, overlap, stitch_overlap)
    changer = make_merge_list(stitch, int(np.sum(counts)))
    changer_dev = torch.from_numpy(changer).to(labels[0].device) if labels else None
    out = {seq: apply_merge_list(labels[seq], int(offsets[seq]), changer_dev, overlap) for seq in labels}
    # objects that are only seen inside overlaps can vanish with the crop; count what is left
    ids = torch.unique(torch.cat([torch.unique(v) for v in out.values()])) if out else torch.zeros(0)
    return out, int((ids != 0).sum().item())
