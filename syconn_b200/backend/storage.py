"""Storage classes of row f3 with the reference's names and file formats (syconn/backend/storage.py:26-93, :208-421 and
syconn/backend/base.py FSBase): pickled dictionaries (``pickle.HIGHEST_PROTOCOL``, written to ``<path>.tmp`` and moved,
handler/basics.py:485-508) whose array values are lists of lz4 block strings.

Only what the extraction writers (syconn/proc/sd_proc.py:788-1215) and an unmodified ``SegmentationDataset`` /
``SegmentationObject`` need: ``AttributeDict``, ``CompressedStorage`` and ``VoxelStorageDyn`` with ``voxel_mode=False``
(bounding boxes, sizes, representative coordinates).  ``voxel_mode=True`` reads voxels back through a KnossosDataset
(knossos_utils, third party, not in this image) and raises.  File locking (fasteners) is not implemented:
``disable_locking`` must stay True, which is what every writer of the hot path passes.
"""
import os
import pickle
import shutil
from collections import defaultdict
from typing import Any, Optional, Union

import numpy as np

from ..handler.compression import arrtolz4string_list, lz4string_listtoarr


def write_obj2pkl(path: str, obj):
    """handler/basics.py:485-508"""
    with open(path + ".tmp", "wb") as f:
        pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)
    shutil.move(path + ".tmp", path)


def load_pkl2obj(path: str):
    with open(path, "rb") as f:
        return pickle.load(f)


class StorageClass:
    """File-backed dictionary (backend/base.py FSBase): ``pull`` on construction when the file exists, ``push`` to write."""

    def __init__(self, inp_p: Optional[str], cache_decomp: bool = False, read_only: bool = True,
                 disable_locking: bool = True, **kwargs):
        if not disable_locking:
            raise NotImplementedError("file locking (fasteners) is not available: pass disable_locking=True")
        self.read_only = read_only
        self._cache_decomp = cache_decomp
        self._cache_dc = {}
        self._dc_intern = {}
        self._path = inp_p
        if inp_p is not None:
            if type(inp_p) is not str:
                raise NotImplementedError(f"Unsupported initialization type {type(inp_p)} for 'FSBase'.")
            self.pull(inp_p)

    def __len__(self):
        return len(self._dc_intern)

    def __contains__(self, item):
        return item in self._dc_intern

    def __iter__(self):
        return iter(self._dc_intern)

    def __eq__(self, other):
        return isinstance(other, StorageClass) and self._dc_intern == other._dc_intern

    def __repr__(self):
        return repr(self._dc_intern)

    def keys(self):
        return self._dc_intern.keys()

    def items(self):
        for k in self._dc_intern.keys():
            yield k, self[k]

    def values(self):
        for k in self._dc_intern.keys():
            yield self[k]

    def push(self, dest: Optional[str] = None):
        dest = self._path if dest is None else dest
        if dest is None:
            return
        write_obj2pkl(dest, self._dc_intern)

    def pull(self, source: Optional[str] = None):
        source = self._path if source is None else source
        fold = os.path.split(source)[0]
        if fold and not os.path.isdir(fold) and not self.read_only:
            os.makedirs(fold, exist_ok=True)
        if os.path.isfile(source):
            self._dc_intern = load_pkl2obj(source)
        else:
            self._dc_intern = {}


class AttributeDict(StorageClass):
    """storage.py:26-50: object id -> attribute dictionary."""

    def __getitem__(self, item):
        try:
            return self._dc_intern[item]
        except KeyError:
            self._dc_intern[item] = {}
            return self._dc_intern[item]

    def __setitem__(self, key, value):
        self._dc_intern[key] = value

    def update(self, other, **kwargs):
        self._dc_intern.update(other, **kwargs)

    def copy_intern(self):
        return dict(self._dc_intern)


class CompressedStorage(StorageClass):
    """storage.py:52-93: key -> ``{"arr": [lz4 strings], "sh": shape with -1 first, "dt": dtype.str}``."""

    def __getitem__(self, item: Union[int, str]):
        try:
            return self._cache_dc[item]
        except KeyError:
            pass
        value_intern = self._dc_intern[item]
        decomp_arr = lz4string_listtoarr(value_intern["arr"], dtype=np.dtype(value_intern["dt"]), shape=value_intern["sh"])
        if self._cache_decomp:
            self._cache_dc[item] = decomp_arr
        return decomp_arr

    def __setitem__(self, key: Union[int, str], value: np.ndarray):
        if type(value) is not np.ndarray:
            raise ValueError("CompressedStorage supports np.array values only.")
        if self._cache_decomp:
            self._cache_dc[key] = value
        sh = list(value.shape)
        sh[0] = -1
        self._dc_intern[key] = {"arr": arrtolz4string_list(value), "sh": tuple(sh), "dt": value.dtype.str}

    def __delitem__(self, key):
        del self._dc_intern[key]
        if key in self._cache_dc:
            del self._cache_dc[key]


class VoxelStorageDyn(CompressedStorage):
    """storage.py:208-421 with ``voxel_mode=False``: object id -> all bounding boxes ``[N, 2, 3]`` that define the object
    (one per chunk it touches), plus the helper entries ``'meta'`` (path of the voxel data), ``'size'``, ``'rep_coord'`` and
    ``'voxel_cache'`` inside the same pickled dictionary."""

    def __init__(self, inp: str, voxel_mode: bool = True, voxeldata_path: Optional[str] = None, **kwargs):
        if not inp.endswith(".pkl"):
            inp = inp + ".pkl"
        super().__init__(inp, **kwargs)
        self.voxel_mode = voxel_mode
        if "meta" not in self._dc_intern:
            self._dc_intern["meta"] = dict(voxeldata_path=voxeldata_path)
        if "size" not in self._dc_intern:
            self._dc_intern["size"] = defaultdict(int)
        if "rep_coord" not in self._dc_intern:
            self._dc_intern["rep_coord"] = dict()
        if "voxel_cache" not in self._dc_intern:
            self._dc_intern["voxel_cache"] = dict()
        if voxeldata_path is not None and self._dc_intern["meta"]["voxeldata_path"] != voxeldata_path:
            self._dc_intern["meta"]["voxeldata_path"] = voxeldata_path
        if voxel_mode:
            raise NotImplementedError("voxel_mode=True reads voxels through knossos_utils.KnossosDataset, which is not "
                                      "available here; use voxel_mode=False (bounding boxes, sizes, rep coords)")

    def __setitem__(self, key: int, value: Any):
        if self.voxel_mode:
            raise RuntimeError("`VoxelStorageDyn.__setitem__` may only be used when `voxel_mode=False`.")
        return super().__setitem__(key, value)

    def object_size(self, item):
        if item not in self._dc_intern:
            raise KeyError('KeyError: Could not find key "{}" in `self._dc_intern`.`'.format(item))
        return self._dc_intern["size"][item]

    def increase_object_size(self, item, value):
        self._dc_intern["size"][item] += value

    def object_repcoord(self, item):
        if item not in self._dc_intern:
            raise KeyError('KeyError: Could not find key "{}" in `self._dc_intern`.`'.format(item))
        return self._dc_intern["rep_coord"][item]

    def set_object_repcoord(self, item, value):
        self._dc_intern["rep_coord"][item] = value

    def get_boundingdata(self, item: int):
        return super().__getitem__(item)

    def keys(self):
        return [k for k in self._dc_intern.keys() if (type(k) is str and k.isdigit()) or (type(k) is not str)]
