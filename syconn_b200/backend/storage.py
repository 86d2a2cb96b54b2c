"""Storage classes of row f3: the reference's names, constructor arguments and ON-DISK FORMAT
(syconn/backend/storage.py:26-93, :208-421 on top of syconn/backend/base.py FSBase), written in this repo's own way.

On disk every storage is ONE pickled ``dict`` (``pickle.HIGHEST_PROTOCOL``, written to ``<path>.tmp`` and renamed,
handler/basics.py:485-508):

  * ``AttributeDict``       ``{object id: {attribute name: value}}``
  * ``CompressedStorage``   ``{key: {"arr": [lz4 strings], "sh": shape with -1 as first entry, "dt": dtype.str}}``
  * ``VoxelStorageDyn``     a CompressedStorage of per-object box lists ``[N, 2, 3]`` plus the helper entries ``'meta'``
                            (``{'voxeldata_path': ...}``), ``'size'`` (defaultdict(int)), ``'rep_coord'``, ``'voxel_cache'``

Only what the extraction writers (syconn/proc/sd_proc.py:788-1215) and a reader of their files need is provided:
``voxel_mode=True`` of ``VoxelStorageDyn`` fetches voxels through a KnossosDataset (knossos_utils, not in this image) and
raises; file locking (fasteners) is not implemented, so ``disable_locking`` must stay True -- what every writer of the hot
path passes anyway.
"""
import os
import pickle
from collections import defaultdict
from typing import Optional

import numpy as np

from ..handler import compression as _codec


class StorageClass:
    """One pickled dictionary per file: read on construction when the file exists, written by ``push``."""

    def __init__(self, inp_p: Optional[str], cache_decomp: bool = False, read_only: bool = True,
                 disable_locking: bool = True, **_ignored):
        if not disable_locking:
            raise NotImplementedError("file locking (fasteners) is not available: pass disable_locking=True")
        if inp_p is not None and not isinstance(inp_p, str):
            raise NotImplementedError(f"Unsupported initialization type {type(inp_p)} for 'FSBase'.")
        self.read_only, self._cache_decomp = read_only, cache_decomp
        self._path, self._cache_dc, self._dc_intern = inp_p, {}, {}
        if inp_p is not None:
            self.pull()

    # -- file I/O
    def pull(self, source: Optional[str] = None):
        src = source or self._path
        folder = os.path.dirname(src)
        if folder and not self.read_only:
            os.makedirs(folder, exist_ok=True)
        self._dc_intern = {}
        if os.path.isfile(src):
            with open(src, "rb") as fh:
                self._dc_intern = pickle.load(fh)

    def push(self, dest: Optional[str] = None):
        dst = dest or self._path
        if dst is None:          # virtual storage: nothing to write
            return
        tmp = dst + ".tmp"
        with open(tmp, "wb") as fh:
            pickle.dump(self._dc_intern, fh, protocol=pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, dst)

    # -- read-only mapping protocol over the raw dictionary; subclasses define item access
    def keys(self):
        return self._dc_intern.keys()

    def values(self):
        return (self[k] for k in list(self._dc_intern))

    def items(self):
        return ((k, self[k]) for k in list(self._dc_intern))

    def __iter__(self):
        return iter(self._dc_intern)

    def __len__(self):
        return len(self._dc_intern)

    def __contains__(self, key):
        return key in self._dc_intern

    def __eq__(self, other):
        return isinstance(other, StorageClass) and other._dc_intern == self._dc_intern

    def __repr__(self):
        return f"{type(self).__name__}({self._path!r}, {len(self._dc_intern)} entries)"


class AttributeDict(StorageClass):
    """object id -> attribute dictionary; reading an unknown id creates its (empty) entry, like the reference."""

    def __getitem__(self, obj_id):
        return self._dc_intern.setdefault(obj_id, {})

    def __setitem__(self, obj_id, attrs):
        self._dc_intern[obj_id] = attrs

    def update(self, other, **kw):
        self._dc_intern.update(other, **kw)

    def copy_intern(self):
        return dict(self._dc_intern)


class CompressedStorage(StorageClass):
    """key -> NumPy array, kept as lz4 strings."""

    @staticmethod
    def _pack(arr: np.ndarray) -> dict:
        return {"arr": _codec.arrtolz4string_list(arr), "sh": (-1,) + tuple(arr.shape[1:]), "dt": arr.dtype.str}

    @staticmethod
    def _unpack(entry: dict) -> np.ndarray:
        return _codec.lz4string_listtoarr(entry["arr"], dtype=np.dtype(entry["dt"]), shape=entry["sh"])

    def __getitem__(self, key):
        if key in self._cache_dc:
            return self._cache_dc[key]
        arr = self._unpack(self._dc_intern[key])
        if self._cache_decomp:
            self._cache_dc[key] = arr
        return arr

    def __setitem__(self, key, value):
        if type(value) is not np.ndarray:
            raise ValueError("CompressedStorage supports np.array values only.")
        self._dc_intern[key] = self._pack(value)
        if self._cache_decomp:
            self._cache_dc[key] = value

    def __delitem__(self, key):
        self._dc_intern.pop(key)
        self._cache_dc.pop(key, None)


class VoxelStorageDyn(CompressedStorage):
    """Per object: all bounding boxes ``[N, 2, 3]`` that hold its voxels (one per chunk it touches) + size + representative
    coordinate.  Opened with ``voxel_mode=False`` by the writers."""

    _HELPERS = (("size", lambda: defaultdict(int)), ("rep_coord", dict), ("voxel_cache", dict))

    def __init__(self, inp: str, voxel_mode: bool = True, voxeldata_path: Optional[str] = None, **kwargs):
        super().__init__(inp if inp.endswith(".pkl") else inp + ".pkl", **kwargs)
        self.voxel_mode = voxel_mode
        meta = self._dc_intern.setdefault("meta", {"voxeldata_path": voxeldata_path})
        if voxeldata_path is not None:
            meta["voxeldata_path"] = voxeldata_path
        for name, make in self._HELPERS:
            if name not in self._dc_intern:
                self._dc_intern[name] = make()
        if voxel_mode:
            raise NotImplementedError("voxel_mode=True reads voxels through knossos_utils.KnossosDataset, which is not "
                                      "available here; use voxel_mode=False (bounding boxes, sizes, rep coords)")

    def _known(self, obj_id):
        if obj_id not in self._dc_intern:
            raise KeyError('KeyError: Could not find key "{}" in `self._dc_intern`.`'.format(obj_id))

    def __setitem__(self, obj_id, boxes):
        if self.voxel_mode:
            raise RuntimeError("`VoxelStorageDyn.__setitem__` may only be used when `voxel_mode=False`.")
        super().__setitem__(obj_id, boxes)

    def get_boundingdata(self, obj_id):
        return CompressedStorage.__getitem__(self, obj_id)

    def object_size(self, obj_id):
        self._known(obj_id)
        return self._dc_intern["size"][obj_id]

    def increase_object_size(self, obj_id, n_voxels):
        self._dc_intern["size"][obj_id] += n_voxels

    def object_repcoord(self, obj_id):
        self._known(obj_id)
        return self._dc_intern["rep_coord"][obj_id]

    def set_object_repcoord(self, obj_id, coord):
        self._dc_intern["rep_coord"][obj_id] = coord

    def keys(self):
        """object ids only (the helper entries have non-numeric string keys)"""
        return [k for k in self._dc_intern if not isinstance(k, str) or k.isdigit()]
