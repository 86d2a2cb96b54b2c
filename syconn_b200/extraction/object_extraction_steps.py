"""Organelle instance segmentation, first slice (row f4): threshold -> connected components per chunk on the GPU ->
unique labels -> stitching of the components that cross chunk borders -> merged labels.

Mirrors the array-level core of ``syconn/extraction/object_extraction_steps.py``:

  * ``object_segmentation_chunk``   <- ``_object_segmentation_thread`` :204-366, the ``scipy.ndimage.label`` branch (:350-352)
                                      after the threshold (:302-303); Gaussian smoothing (vigra), morphology and the
                                      watershed branch are not part of this slice
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


def object_segmentation_chunk(prob: torch.Tensor, threshold: int = 0) -> Tuple[torch.Tensor, int]:
    """``scipy.ndimage.label(prob > threshold)`` of one (overlap-extended) chunk -> (int32 labels, number of components)."""
    return dev.label_components(prob, threshold)


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
                               threshold: int = 0, overlap=(1, 1, 1), stitch_overlap=(1, 1, 1)):
    """The whole first slice over a chunk grid.  ``load_block(offset, size)`` returns the probability block at ``offset``
    (may reach outside the volume: the caller pads with zeros, as the KnossosDataset does).  Returns
    ``({seq: int64 label tensor of the chunk (overlap cropped, stitched ids)}, n_objects)``."""
    labels, counts = {}, []
    for seq in range(len(plan)):
        off = [plan.offsets[seq][i] - overlap[i] for i in range(3)]
        size = [plan.sizes[seq][i] + 2 * overlap[i] for i in range(3)]
        lab, n = object_segmentation_chunk(load_block(off, size), threshold)
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
