"""lz4 array <-> byte-string helpers with the reference's names (syconn/handler/compression.py:39-127).

The reference calls python-lz4's ``lz4.block.compress`` / ``decompress`` (third party, absent from this image).  Here the
LZ4 block codec is ``libsyk.so``'s host-side ``syk_lz4_*`` (csrc/syk_lz4.cu, restated from the published block format);
``compress`` / ``decompress`` keep python-lz4's default framing (``store_size=True``: a 4-byte little-endian
uncompressed length in front of the block), so strings written here are read by ``lz4.block.decompress`` and vice versa.
The compressed bytes themselves are not claimed to equal liblz4's (parity of the byte stream: unpinned)."""
import ctypes as C
import struct
from typing import List, Optional, Tuple, Union

import numpy as np

from .. import _lib

LZ4_MAX_INPUT_SIZE = 0x7E000000


class LZ4BlockError(Exception):
    pass


def compress(data) -> bytes:
    """``lz4.block.compress(data)`` (mode 'default', store_size=True)."""
    data = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data
    n = len(data)
    if n > LZ4_MAX_INPUT_SIZE:
        raise OverflowError("Input too large for LZ4 API")
    L = _lib.load()
    cap = int(L.syk_lz4_compress_bound(n))
    dst = C.create_string_buffer(cap)
    out_n = C.c_uint64()
    src = (C.c_char * n).from_buffer_copy(data) if n else None
    _lib.check(L.syk_lz4_compress_block(src, n, dst, cap, C.byref(out_n)))
    return struct.pack("<I", n) + dst.raw[:out_n.value]


def decompress(data) -> bytes:
    """``lz4.block.decompress(data)`` for strings written with store_size=True."""
    data = bytes(data)
    if len(data) < 4:
        raise LZ4BlockError("Input source data size too small")
    (n,) = struct.unpack("<I", data[:4])
    L = _lib.load()
    dst = C.create_string_buffer(max(n, 1))
    out_n = C.c_uint64()
    body = data[4:]
    rc = L.syk_lz4_decompress_block(body, len(body), dst, n, C.byref(out_n))
    if rc != 0 or out_n.value != n:
        raise LZ4BlockError("Decompression failed: corrupt input or insufficient space in destination buffer")
    return dst.raw[:n]


def _as_array(arr):
    return np.array(arr) if isinstance(arr, list) else arr


def arrtolz4string(arr: np.ndarray) -> bytes:
    """array -> one lz4 string (``b""`` for an empty array); compression.py:39-57"""
    arr = _as_array(arr)
    return compress(arr.tobytes()) if len(arr) else b""


def lz4stringtoarr(string: bytes, dtype=np.float32, shape: Optional[Tuple[int]] = None) -> np.ndarray:
    """one lz4 string -> array of ``dtype`` (reshaped when ``shape`` is given); compression.py:60-81"""
    if not string:
        return np.zeros((0,), dtype=dtype)
    flat = np.frombuffer(decompress(string), dtype=dtype)
    return flat if shape is None else flat.reshape(shape)


def arrtolz4string_list(arr: np.ndarray) -> List[bytes]:
    """array -> list of lz4 strings: normally one; an array that is too large for a single LZ4 block is cut in halves along
    its first axis until every piece fits (compression.py:83-103 does the same by recursion)."""
    arr = _as_array(arr)
    if len(arr) == 0:
        return [b""]
    out, todo = [], [arr]
    while todo:
        piece = todo.pop()
        if piece.nbytes <= LZ4_MAX_INPUT_SIZE or len(piece) < 2:
            out.append(compress(piece.tobytes()) if len(piece) else b"")
        else:
            mid = len(piece) // 2
            todo += [piece[mid:], piece[:mid]]      # stack: the first half is compressed first
    return out


def lz4string_listtoarr(str_lst: Union[List[bytes], np.ndarray], dtype=np.float32,
                        shape: Optional[Tuple[int]] = None) -> np.ndarray:
    """list of lz4 strings -> one array (an array passes through unchanged); compression.py:106-127"""
    if isinstance(str_lst, np.ndarray):
        return str_lst
    pieces = [lz4stringtoarr(s, dtype=dtype, shape=shape) for s in str_lst]
    return np.concatenate(pieces) if pieces else np.zeros((0,), dtype=dtype)
