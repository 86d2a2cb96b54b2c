"""Morphology helpers of the organelle extraction (row f4) on the GPU.

Drop-in names of ``syconn/proc/image.py`` for the calls ``_object_segmentation_thread`` makes on the thresholded volume
(``object_extraction_steps.py:312-358``):

  * ``get_aniso_struct``                <- image.py:522-539
  * ``apply_morphological_operations``  <- image.py:485-507 (+ ``_count_subsequent_mops`` :510-519 and the single-object
                                           case of ``_multi_mop_findobjects`` :358-437)

The voxel work runs in ``syk_binary_morph_ops`` (csrc/syk_morph_vol.cu).  Only 0/1 volumes are a device path -- what the
organelle worker passes; the reference's per-object loop over multi-label overlays (``multi_dilation_backgroundonly``,
``binary_fill_holes``) is used elsewhere (glia / skeleton code) and raises here.
"""
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import device as dev

SUPPORTED_MOPS = tuple(dev.MORPH_OPS)


def get_aniso_struct(scaling: Union[tuple, np.ndarray]) -> np.ndarray:
    """5 x 5 x 3 kernel: the centre voxel in the planes z -+ 1 and, in the middle plane, the city-block disc of radius
    ``scaling[2] // scaling[0]`` clipped to the 5 x 5 window (what ``aniso`` cross dilations of the centre pixel inside a
    5 x 5 array give)."""
    aniso = int(scaling[2] // scaling[0])
    assert scaling[1] // scaling[0] == 1
    assert aniso >= 1
    d = np.abs(np.arange(5) - 2)
    struct = np.zeros((5, 5, 3), np.float64)
    struct[:, :, 1] = (d[:, None] + d[None, :]) <= aniso
    struct[2, 2, 0] = struct[2, 2, 2] = 1
    return struct


def _count_subsequent_mops(mops: Sequence[str]) -> Tuple[List[str], List[int]]:
    """Run-length encode the op list: a run of n equal ops is one op with n iterations."""
    names, counts = [], []
    for m in mops:
        if names and names[-1] == m:
            counts[-1] += 1
        else:
            names.append(m)
            counts.append(1)
    return names, counts


def apply_morphological_operations(vol, morph_ops: List[str], mop_kwargs: Optional[dict] = None):
    """Apply the ``scipy.ndimage`` binary ops named in ``morph_ops`` to the 0/1 volume ``vol`` the way the reference does
    (inside the foreground's bounding box; dilation / closing padded by the iteration count and cropped back).

    ``vol``: CUDA tensor [X,Y,Z] (updated in place and returned) or a NumPy array (uploaded, processed, downloaded into
    the same array).  ``mop_kwargs``: ``structure`` (default: the 6-neighbourhood cross, scipy's default) and optionally
    ``iterations`` (overrides every run length, as in the reference)."""
    if len(morph_ops) == 0:
        return vol
    kw = dict(mop_kwargs or {})
    structure = kw.pop("structure", None)
    forced_iters = kw.pop("iterations", None)
    if kw:
        raise TypeError(f"unsupported morphology keyword(s): {sorted(kw)}")
    if structure is None:
        structure = np.zeros((3, 3, 3), np.uint8)
        structure[1, 1, :] = structure[1, :, 1] = structure[:, 1, 1] = 1
    names, counts = _count_subsequent_mops(morph_ops)
    for n in names:
        if n not in dev.MORPH_OPS:
            msg = f"Only erosion or dilation allowed. Attempted to use morphological operation '{n}'."
            raise NotImplementedError(msg)
    if forced_iters is not None:
        counts = [int(forced_iters)] * len(names)
    if isinstance(vol, torch.Tensor):
        return dev.binary_morph_ops(vol, names, counts, structure)
    host = np.asarray(vol)
    t = torch.from_numpy(np.ascontiguousarray(host)).cuda()
    dev.binary_morph_ops(t, names, counts, structure)
    host[...] = t.cpu().numpy()
    return host
