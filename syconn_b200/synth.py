"""Synthetic label volumes (bench/test inputs), NumPy twin of ``syk_synth_labels`` (csrc/syk_host.cu).

Integer-only and coordinate-hashed: a voxel's label depends only on its GLOBAL coordinate, the pitch, the warp
amplitude, the seed and the channel kind, so any chunk (with any halo) can be regenerated independently on the
CPU and on the GPU with identical bits.  Objects are cells of a warped grid ("wavy" supervoxels):

  kind 0   cell supervoxels, ids in [1, 2^32), ~3 % background cells (id 0)
  kind >=1 organelle channel: a fraction density16/16 of the grid cells is foreground with a 64-bit id
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(k):
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xff51afd7ed558ccd)
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xc4ceb9fe1a85ec53)
    k = k ^ (k >> np.uint64(33))
    return k


def _tri(t, P):
    return np.abs(t % (2 * P) - P)


def synth_labels(shape, origin=(0, 0, 0), pitch=(32, 32, 16), warp_amp=4, seed=0, kind=0, density16=1,
                 dtype=np.uint64, order="C"):
    """Return a label volume of logical shape ``shape`` whose voxel (x,y,z) sits at global ``origin + (x,y,z)``.
    ``order='C'``: z fastest in memory; ``order='F'``: x fastest in memory (ZYX memory seen as XYZ, the
    production layout of syconn/proc/sd_proc.py:629,641)."""
    dtype = np.dtype(dtype)
    assert dtype in (np.uint32, np.uint64)
    B = 1 << 20
    nx, ny, nz = (int(s) for s in shape)
    X = (np.arange(nx, dtype=np.int64) + int(origin[0]) + B)[:, None, None]
    Y = (np.arange(ny, dtype=np.int64) + int(origin[1]) + B)[None, :, None]
    Z = (np.arange(nz, dtype=np.int64) + int(origin[2]) + B)[None, None, :]
    amp = int(warp_amp)
    with np.errstate(over="ignore"):
        xw = X + (((_tri(Y, 41) + _tri(Z, 29)) * amp) >> 4)
        yw = Y + (((_tri(Z, 37) + _tri(X, 43)) * amp) >> 4)
        zw = Z + (((_tri(X, 31) + _tri(Y, 47)) * amp) >> 5)
        cx = (xw // int(pitch[0])).astype(np.uint64)
        cy = (yw // int(pitch[1])).astype(np.uint64)
        cz = (zw // int(pitch[2])).astype(np.uint64)
        salt = (np.uint64(seed) * np.uint64(0xD6E8FEB86659FD93) + np.uint64(kind) * np.uint64(0xA0761D6478BD642F))
        h = _mix64((cx * np.uint64(0x9E3779B97F4A7C15)) ^ (cy * np.uint64(0xC2B2AE3D27D4EB4F)) ^
                   (cz * np.uint64(0x165667B19E3779F9)) ^ salt)
        if kind == 0:
            ids = h >> np.uint64(32)
            ids = np.where(ids == 0, np.uint64(1), ids)
            out = np.where((h & np.uint64(31)) == 0, np.uint64(0), ids)
        else:
            ids = (h | np.uint64(1)) if dtype == np.uint64 else ((h >> np.uint64(32)) | np.uint64(1))
            out = np.where(((h >> np.uint64(8)) & np.uint64(15)) >= np.uint64(density16), np.uint64(0), ids)
    out = np.broadcast_to(out, (nx, ny, nz)).astype(dtype)
    if order == "F":
        out = np.ascontiguousarray(out.transpose(2, 1, 0)).transpose(2, 1, 0)
    else:
        out = np.ascontiguousarray(out)
    return out
