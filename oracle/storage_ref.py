"""oracle/storage_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Minimal re-statement of how the reference READS its storage files, used to check what ``syconn_b200``'s writers (row f3)
put on disk:

  * ``lz4_block_decode``      -- the published LZ4 block format (lz4_Block_format.md); independent, pure-Python decoder
                                (python-lz4, the reference's dependency, is not in this image)
  * ``lz4_decompress``        -- python-lz4 ``lz4.block.decompress`` with the default 4-byte size header
  * ``lz4string_listtoarr``   -- syconn/handler/compression.py:60-81,106-127
  * ``read_attr_dict``        -- AttributeDict.__getitem__ over the pickled dict, syconn/backend/storage.py:26-50
  * ``read_voxel_dyn``        -- VoxelStorageDyn(voxel_mode=False).__getitem__/object_size/object_repcoord/keys,
                                syconn/backend/storage.py:52-75, :208-421
  * ``subfold_from_ix``       -- syconn/reps/rep_helper.py:143-163

Parity status: UNPINNED for the compressed byte stream (no reference-written file and no liblz4 here); the decoded
content is checked exactly, and the decoder follows the format specification, not our encoder.
"""
import pickle
import struct

import numpy as np


def lz4_block_decode(src: bytes, n_out: int) -> bytes:
    out = bytearray()
    i, n = 0, len(src)
    while True:
        token = src[i]
        i += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]
        assert i + lit <= n, "literals run past the block"
        i += lit
        if i == n:
            break
        off = src[i] | (src[i + 1] << 8)
        i += 2
        assert 0 < off <= len(out), "offset outside the decoded data"
        ml = token & 15
        if ml == 15:
            while True:
                b = src[i]
                i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        for k in range(ml):          # byte-wise: overlapping matches repeat the pattern
            out.append(out[start + k])
    assert len(out) == n_out, f"decoded {len(out)} bytes, header says {n_out}"
    return bytes(out)


def lz4_decompress(string: bytes) -> bytes:
    try:
        from lz4.block import decompress          # the real thing, when present
        return decompress(string)
    except ImportError:
        (n,) = struct.unpack("<I", string[:4])
        return lz4_block_decode(string[4:], n)


def lz4string_listtoarr(str_lst, dtype, shape=None):
    if len(str_lst) == 0:
        return np.zeros((0,), dtype=dtype)
    parts = []
    for s in str_lst:
        if len(s) == 0:
            parts.append(np.zeros((0,), dtype=dtype))
            continue
        a = np.frombuffer(lz4_decompress(s), dtype=dtype)
        parts.append(a.reshape(shape) if shape is not None else a)
    return np.concatenate(parts)


def read_attr_dict(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def read_voxel_dyn(path):
    """-> ({id: bbs [N, 2, 3]}, sizes {id: int}, rep_coords {id: array}, meta)"""
    with open(path, "rb") as f:
        dc = pickle.load(f)
    keys = [k for k in dc.keys() if (type(k) is str and k.isdigit()) or (type(k) is not str)]
    bbs = {k: lz4string_listtoarr(dc[k]["arr"], np.dtype(dc[k]["dt"]), dc[k]["sh"]) for k in keys}
    return bbs, dict(dc["size"]), dict(dc["rep_coord"]), dc["meta"]


def subfold_from_ix(ix, n_folders):
    order = int(np.log10(n_folders))
    ix = int(ix // 1e3 % n_folders)
    id_str = '{num:0{w}d}'.format(num=ix, w=order)
    subfold = "/"
    for idx in range(0, order, 2):
        subfold += "%s/" % id_str[idx: idx + 2]
    return subfold
