"""Chunked dataset extraction on one or several GPUs (one process per GPU).

Replaces, for the hot path only, the reference's map/reduce plumbing:

  * chunk list split across workers           syconn/proc/sd_proc.py:360-366, extraction/cs_extraction_steps.py:199-203
  * per-chunk kernels + per-worker dict merge  syconn/proc/sd_proc.py:579-684, cs_extraction_steps.py:376-486
  * reducers partitioned by an id hash         syconn/reps/rep_helper.py:143-163, sd_proc.py:511-556

Every rank runs its chunks through libsyk, keeps the per-(id, chunk) records in a device log, buckets them by
``owner = hash(id) mod world`` (organelle id for overlap pairs) and exchanges the buckets with one all-to-all
(NCCL over NVLink; gloo on CPU for tests).  The owner folds the received records with the reference's merge
semantics (size summed, bbox min/max, per-chunk bbox list kept in chunk order, rep_coord of the last chunk).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._lib import GEOM_DTYPE, PAIR_DTYPE, RECORD_DTYPE


# ------------------------------------------------------------------------------------------------ chunk plan
@dataclass
class ChunkPlan:
    """Chunk grid of a volume; chunk ``seq`` numbers follow the reference's chunk list order (x slowest)."""
    volume_shape: Tuple[int, int, int]
    chunk_size: Tuple[int, int, int] = (512, 512, 512)
    offsets: List[Tuple[int, int, int]] = field(default_factory=list)
    sizes: List[Tuple[int, int, int]] = field(default_factory=list)

    def __post_init__(self):
        if not self.offsets:
            vs, cs = self.volume_shape, self.chunk_size
            for x in range(0, vs[0], cs[0]):
                for y in range(0, vs[1], cs[1]):
                    for z in range(0, vs[2], cs[2]):
                        self.offsets.append((x, y, z))
                        self.sizes.append((min(cs[0], vs[0] - x), min(cs[1], vs[1] - y), min(cs[2], vs[2] - z)))

    def __len__(self):
        return len(self.offsets)

    def chunks_of_rank(self, rank: int, world: int) -> List[int]:
        """Contiguous block partition (``chunkify_successive``, sd_proc.py:366): neighbouring chunks stay on one GPU."""
        n = len(self)
        base, rem = divmod(n, world)
        start = rank * base + min(rank, rem)
        return list(range(start, start + base + (1 if rank < rem else 0)))


def cs_halo_geometry(offset, size, stencil=(13, 13, 7)):
    """Geometry of the contact-site call of one chunk (cs_extraction_steps.py:376-391): the loaded block is
    ``size + 2*overlap + 2*stencil_offset`` at ``offset - overlap - stencil_offset``; detect_cs returns the
    ``size + 2*overlap`` block at ``offset - overlap`` (overlap = max(stencil // 2))."""
    so = [s // 2 for s in stencil]
    overlap = max(so)
    load_off = [offset[i] - overlap - so[i] for i in range(3)]
    load_size = [size[i] + 2 * overlap + 2 * so[i] for i in range(3)]
    out_off = [offset[i] - overlap for i in range(3)]
    out_size = [size[i] + 2 * overlap for i in range(3)]
    return load_off, load_size, out_off, out_size


# ------------------------------------------------------------------------------------------------ exchange
_M64 = (1 << 64) - 1


def owner_of(ids, n_owners: int) -> np.ndarray:
    """Rank that owns each id in the final reduce: NumPy mirror of ``syk_owner_of`` (csrc/syk_table.cu), the
    counterpart of the reference's id -> reducer hash (syconn/reps/rep_helper.py:143-163).  Host-side bookkeeping
    only (e.g. "which rank's table holds object X"); the records themselves are bucketed on the device."""
    k = np.asarray(ids, np.uint64) ^ np.uint64(0x5bd1e9955bd1e995)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xff51afd7ed558ccd)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xc4ceb9fe1a85ec53)
        k ^= k >> np.uint64(33)
    return ((k >> np.uint64(20)) % np.uint64(n_owners)).astype(np.int64)


def exchange_logs(bucketed: Sequence[torch.Tensor], counts: Sequence[torch.Tensor], world: int, group=None):
    """The hash-owner exchange of ``ExtractionPipeline.finish`` (the reference's reduce keyed by object id,
    syconn/proc/sd_proc.py:511-556): ``bucketed[j]`` is log j of this rank with its rows grouped per destination rank
    (``counts[j][d]`` rows for rank d, rank order).  ONE count all-to-all for all logs, then one all-to-all-v per row
    width.  Returns the list of owned logs (rows received from all ranks, source-rank order).  Device-agnostic: NCCL
    on CUDA tensors, gloo on CPU tensors (tests/test_chunked.py drives this very function)."""
    n_logs = len(bucketed)
    if world == 1 or n_logs == 0:
        return list(bucketed)
    dev = bucketed[0].device
    send_mat = torch.stack([c.to(torch.int64) for c in counts], dim=1).contiguous()   # [dest, log] rows this rank sends
    recv_mat = torch.empty_like(send_mat)                                            # [source, log] rows it receives
    dist.all_to_all_single(recv_mat, send_mat, group=group)
    send_h, recv_h = send_mat.tolist(), recv_mat.tolist()                            # one host sync for both matrices
    out = [None] * n_logs
    widths = sorted(set(int(b.shape[1]) for b in bucketed))
    for width in widths:
        log_ids = [j for j in range(n_logs) if bucketed[j].shape[1] == width]
        dtype = bucketed[log_ids[0]].dtype
        parts, in_split = [], []
        offs = {j: 0 for j in log_ids}
        for d in range(world):
            tot = 0
            for j in log_ids:
                m = send_h[d][j]
                if m:
                    parts.append(bucketed[j][offs[j]:offs[j] + m])
                offs[j] += m
                tot += m
            in_split.append(tot)
        for j in log_ids:
            assert offs[j] == bucketed[j].shape[0], "bucket counts do not add up to the log length"
        send = torch.cat(parts) if parts else torch.empty((0, width), dtype=dtype, device=dev)
        out_split = [sum(recv_h[src][j] for j in log_ids) for src in range(world)]
        recv = torch.empty((sum(out_split), width), dtype=dtype, device=dev)
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=out_split, input_split_sizes=in_split, group=group)
        per_log = {j: [] for j in log_ids}
        pos = 0
        for src in range(world):
            for j in log_ids:
                m = recv_h[src][j]
                if m:
                    per_log[j].append(recv[pos:pos + m])
                pos += m
        for j, v in per_log.items():
            out[j] = torch.cat(v) if v else torch.empty((0, width), dtype=dtype, device=dev)
    return out


# ------------------------------------------------------------------------------------------------ host-side reductions
def reduce_records(log: np.ndarray):
    """Fold a per-(id, chunk) record log with the reference's merge semantics (sd_proc.py:1248-1273 and the final
    reduction :939-945): returns a dict of arrays ``id, size, bounding_box [N,2,3] int32, rep_coord [N,3] int32``
    plus the per-object voxel index ``bbs`` (list of [n_chunks_i, 2, 3] arrays, chunk order)."""
    if len(log) == 0:
        return dict(id=np.empty(0, np.uint64), size=np.empty(0, np.int64), bounding_box=np.empty((0, 2, 3), np.int32),
                    rep_coord=np.empty((0, 3), np.int32), bbs=[])
    o = np.lexsort((log["chunk_seq"], log["id"]))
    log = log[o]
    ids, start = np.unique(log["id"], return_index=True)
    end = np.append(start[1:], len(log))
    size = np.add.reduceat(log["count"].astype(np.int64), start)
    bmin = np.minimum.reduceat(log["bb_min"], start, axis=0)
    bmax = np.maximum.reduceat(log["bb_max"], start, axis=0)
    rep = log["rep"][end - 1]  # last chunk wins (dict.update, sd_proc.py:1261)
    bb_all = np.stack([log["bb_min"], log["bb_max"]], axis=1)
    bbs = [bb_all[s:e] for s, e in zip(start, end)]
    return dict(id=ids, size=size, bounding_box=np.stack([bmin, bmax], axis=1).astype(np.int32),
                rep_coord=rep.astype(np.int32), bbs=bbs)


def reduce_pairs(log: np.ndarray):
    """merge_map_dicts (sd_proc.py:1300-1322) on a pair log -> sorted (sub_id, cell_id, count) arrays."""
    if len(log) == 0:
        return dict(sub_id=np.empty(0, np.uint64), cell_id=np.empty(0, np.uint64), count=np.empty(0, np.int64))
    o = np.lexsort((log["cell_id"], log["sub_id"]))
    log = log[o]
    key_change = np.ones(len(log), bool)
    key_change[1:] = (log["sub_id"][1:] != log["sub_id"][:-1]) | (log["cell_id"][1:] != log["cell_id"][:-1])
    start = np.flatnonzero(key_change)
    return dict(sub_id=log["sub_id"][start], cell_id=log["cell_id"][start],
                count=np.add.reduceat(log["count"].astype(np.int64), start))


# ------------------------------------------------------------------------------------------------ device pipeline
class ExtractionPipeline:
    """Per-rank driver of the three hot-path stages over this rank's chunks (device-resident inputs)."""

    def __init__(self, n_sub: int, stencil=(13, 13, 7), chunk_table_capacity=1 << 18, log_capacity=1 << 21,
                 pair_log_capacity=1 << 21, rank=0, world=1, group=None, min_obj_vx=None, with_syn=False, cs_dilation=2,
                 sub_table_capacity=None):
        """``min_obj_vx``: {"cell": n, "sub0": n, ...} -- the worker's small-object drop (sd_proc.py:650-661, :667-680): objects
        that lie purely inside a chunk with fewer voxels are not reported (nor are their overlap pairs)."""
        from . import device as dev
        self.dev = dev
        self.n_sub, self.stencil, self.rank, self.world, self.group = n_sub, tuple(stencil), rank, world, group
        self.t_cell = dev.IdTable(chunk_table_capacity)
        self.t_cs = dev.IdTable(chunk_table_capacity)
        # per-chunk tables are cleared and scanned once per chunk: size the organelle / pair tables for the (far fewer) organelle
        # objects of a chunk when the caller knows them (an overflow is detected and reported, never silent)
        sub_cap = chunk_table_capacity if sub_table_capacity is None else sub_table_capacity
        self.t_sub = [dev.IdTable(sub_cap) for _ in range(n_sub)]
        self.t_pair = [dev.PairTable(sub_cap) for _ in range(n_sub)]
        self.kinds = ["cell", "cs"] + [f"sub{c}" for c in range(n_sub)] + (["syn"] if with_syn else [])
        self.min_obj_vx = dict(min_obj_vx or {})
        # with_syn: process_chunk(..., syn_masks=...) runs the whole numeric body of the contact-site worker
        # (cs_extraction_steps.py:381-486): closing / dilation of every contact, then extract_cs_syntype on the cropped volumes
        self.with_syn, self.cs_dilation = with_syn, int(cs_dilation)
        self.t_syn = dev.IdTable(chunk_table_capacity) if with_syn else None
        self.syn_voxels = []   # (chunk seq, offset, cropped shape, int64 tensor [n, 4] of syk_synvox_t rows) per chunk
        self.log_capacity, self.pair_log_capacity = log_capacity, pair_log_capacity
        self.logs = {k: torch.empty((log_capacity, 8), dtype=torch.int64, device="cuda") for k in self.kinds}
        self.pair_logs = [torch.empty((pair_log_capacity, 4), dtype=torch.int64, device="cuda") for _ in range(n_sub)]
        # one device counter per log (+ pair logs)
        self.counters = torch.zeros(len(self.kinds) + n_sub, dtype=torch.int64, device="cuda")
        self.cs_out = None
        self.launches = 0
        self._final = {}
        self.cs_events = None  # set to a list to collect (start, end) CUDA events around every detect_cs launch

    def reset(self):
        self.counters.zero_()
        self.launches = 0
        self.syn_voxels = []

    def _append(self, table, kind_idx, log, origin, shape):
        L = self.dev._lib.load()
        g = np.zeros(1, GEOM_DTYPE)
        g["origin"][0] = origin
        g["shape"][0] = shape
        min_vx = int(self.min_obj_vx.get(self.kinds[kind_idx], 0))
        self.dev.check(L.syk_table_append_records_min_vx(table.h, g.ctypes.data, log.data_ptr(), log.shape[0],
                                                         self.counters[kind_idx:].data_ptr(), min_vx, self.dev._stream_ptr()))
        self.launches += 2  # k_table_export + k_poison_on_overflow

    def process_chunk(self, seq, offset, cell, subcell, cell_halo, syn_masks=None):
        """One chunk: ``cell`` [X,Y,Z] and ``subcell`` [C,X,Y,Z] 64-bit labels at ``offset``; ``cell_halo`` the
        uint32 block of cs_halo_geometry (or None to skip contact sites).  ``syn_masks`` = (sj, asym, sym) uint8 tensors
        of the un-cropped contact volume (``size + 2 * overlap`` at ``offset - overlap``, cs_extraction_steps.py:394-434)
        switches the contact-site stage to the full worker body (needs ``with_syn=True``)."""
        dev, L = self.dev, self.dev._lib.load()
        if syn_masks is not None:
            assert self.with_syn and cell_halo is not None
            self._contact_site_worker(seq, offset, cell_halo, syn_masks)
            cell_halo = None
        # stage 1: contact sites (cs_extraction_steps.py:391) and the properties the worker merges across chunks:
        # those of the contact volume cropped by `overlap`, at the chunk's own offset (:465-486) -- neighbouring chunks
        # overlap by 2*overlap voxels in the un-cropped volume, which must not be counted twice
        if cell_halo is not None:
            so = [s // 2 for s in self.stencil]
            overlap = max(so)
            buf = self._cs_buffer(cell_halo)
            if self.cs_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self.cs_out = dev.detect_cs(cell_halo, self.stencil, out=buf)
            if self.cs_events is not None:
                e1.record()
                self.cs_events.append((e0, e1))
            crop = self.cs_out[overlap:self.cs_out.shape[0] - overlap, overlap:self.cs_out.shape[1] - overlap,
                               overlap:self.cs_out.shape[2] - overlap]
            self.t_cs.clear()
            dev.find_object_properties(self.t_cs, crop, origin=offset, chunk_seq=seq)
            self._append(self.t_cs, 1, self.logs["cs"], offset, crop.shape)
            self.launches += 3 + 1  # k_cs_fast tier 1, tier 2, k_detect_cs (list mode) + k_scan on the contact volume
        # stages 2+3: cell / organelle properties and overlap mapping (sd_proc.py:646-684)
        self.t_cell.clear()
        for t in self.t_sub + self.t_pair:
            t.clear()
        dev.map_subcell_extract_props(self.t_cell, self.t_sub, self.t_pair, cell, subcell, origin=offset, chunk_seq=seq)
        self.launches += 1 + self.n_sub  # k_scan on the cell volume + one organelle-first k_scan per channel
        self._append(self.t_cell, 0, self.logs["cell"], offset, cell.shape)
        for c in range(self.n_sub):
            self._append(self.t_sub[c], 2 + c, self.logs[f"sub{c}"], offset, cell.shape)
            g = np.zeros(1, GEOM_DTYPE)
            g["origin"][0], g["shape"][0] = offset, cell.shape
            dev.check(L.syk_pairs_append_min_vx(self.t_pair[c].h, self.t_sub[c].h, g.ctypes.data,
                                                int(self.min_obj_vx.get(f"sub{c}", 0)), self.pair_logs[c].data_ptr(),
                                                self.pair_logs[c].shape[0], self.counters[len(self.kinds) + c:].data_ptr(),
                                                dev._stream_ptr()))
            self.launches += 2  # k_pairs_export + k_poison_on_overflow

    def _contact_site_worker(self, seq, offset, cell_halo, syn_masks):
        """detect_cs (:391) -> find_object_properties of the un-cropped contacts for the closing boxes (:439) -> per-id closing
        + dilation in ascending id order (:440-461) -> extract_cs_syntype on the volumes cropped by ``overlap`` at the
        chunk's own offset (:465-470).  Logs the cs records and the records of the synaptic parts (``syn``); the synaptic
        voxel tuples (id, linear index, sym / asym flags) stay on the device in ``self.syn_voxels``."""
        dev = self.dev
        overlap = max(s // 2 for s in self.stencil)
        self.cs_out = dev.detect_cs(cell_halo, self.stencil, out=self._cs_buffer(cell_halo))
        full_shape = tuple(self.cs_out.shape)
        self.t_cs.clear()
        dev.find_object_properties(self.t_cs, self.cs_out)
        rec = self.t_cs.export(dev.geoms([[0, 0, 0]], [list(full_shape)]), sort=True)   # ascending ids; stays on the device
        if rec.shape[0]:
            dev.close_contacts_records(self.cs_out, rec, overlap, self.cs_dilation)
        crop = tuple(slice(overlap, n - overlap) for n in full_shape)
        cs_c = self.cs_out[crop]
        sj, asym, sym = (m[crop] for m in syn_masks)
        self.t_cs.clear()
        self.t_syn.clear()
        vox = dev.extract_cs_syntype(self.t_cs, cs_c, sj, asym, sym, origin=offset, chunk_seq=seq, syn_table=self.t_syn)
        self._append(self.t_cs, 1, self.logs["cs"], offset, cs_c.shape)
        self._append(self.t_syn, self.kinds.index("syn"), self.logs["syn"], offset, cs_c.shape)
        self.syn_voxels.append((seq, tuple(offset), tuple(cs_c.shape), vox))
        self.launches += 3 + 2 + 4 + 4 + 3 + 1   # detect_cs tiers, props + export, sort, morphology, extract_cs_syntype (+ syn props), export

    def _cs_buffer(self, cell_halo):
        """Contact volume of one chunk, laid out like the input (same fastest axis) with the row pitch padded to a multiple
        of 128 bytes and the base shifted so that the rows of the CROPPED view (`overlap` voxels in) start on 128-byte
        boundaries: the TMA row boxes of the property scan then fetch whole sectors only (the natural 524-element rows
        cost 1.19x the algorithmic DRAM reads, profiles/r1_traffic.json)."""
        oshape = [cell_halo.shape[i] - self.stencil[i] + 1 for i in range(3)]
        if self.cs_out is not None and list(self.cs_out.shape) == oshape and \
                (self.cs_out.stride(0) == 1) == (cell_halo.stride(0) == 1):
            return self.cs_out
        order = sorted(range(3), key=lambda a: -abs(cell_halo.stride(a)))  # slowest .. fastest axis of the input
        n_slow, n_mid, n_fast = (oshape[a] for a in order)
        overlap = max(s // 2 for s in self.stencil)
        pitch = (n_fast + 15) // 16 * 16                                   # int64 elements: 16 x 8 B = 128 B
        lead = (-overlap) % 16
        flat = torch.empty(lead + n_slow * n_mid * pitch + 16, dtype=torch.int64, device=cell_halo.device)
        lead += (-(flat.data_ptr() // 8 + lead + overlap)) % 16            # whatever the allocator's own alignment
        phys = flat[lead:lead + n_slow * n_mid * pitch].view(n_slow, n_mid, pitch)[:, :, :n_fast]
        inv = [order.index(a) for a in range(3)]
        return phys.permute(inv)

    def finish(self):
        """Bucket the logs by owner, exchange them (all-to-all) and fold them on the owner.
        Returns {kind: int64 tensor [n, 8]} of this rank's owned per-(id, chunk) records and the owned pair logs."""
        dev = self.dev
        n = self.counters.tolist()  # the only host sync of the step
        if any(v >> 62 for v in n):
            raise dev._lib.SykError(dev._lib.SYK_EOVERFLOW, "a per-chunk table overflowed: raise chunk_table_capacity")
        for i, k in enumerate(self.kinds):
            if n[i] > self.log_capacity:
                raise dev._lib.SykError(dev._lib.SYK_EOVERFLOW, f"record log '{k}' too small ({n[i]} > {self.log_capacity})")
        for c in range(self.n_sub):
            if n[len(self.kinds) + c] > self.pair_log_capacity:
                raise dev._lib.SykError(dev._lib.SYK_EOVERFLOW, "pair log too small")
        owned, owned_pairs = {}, []
        if self.world == 1:
            for i, k in enumerate(self.kinds):
                owned[k] = self.logs[k][:n[i]]
            for c in range(self.n_sub):
                owned_pairs.append(self.pair_logs[c][:n[len(self.kinds) + c]])
            return owned, owned_pairs
        # ---- hash-owner exchange (exchange_logs): bucket every log by owner on the device, then exchange ----
        W, nk = self.world, len(self.kinds)
        bucketed, cnts = [], []
        for i, k in enumerate(self.kinds):
            b, c = dev.bucket_records(self.logs[k][:n[i]], W)
            bucketed.append(b)
            cnts.append(c)
        for c in range(self.n_sub):
            b, cc = dev.bucket_pairs(self.pair_logs[c][:n[nk + c]], W)
            bucketed.append(b)
            cnts.append(cc)
        self.launches += 3 * len(bucketed)
        got = exchange_logs(bucketed, cnts, W, self.group)
        for i, k in enumerate(self.kinds):
            owned[k] = got[i]
        owned_pairs = [got[nk + c] for c in range(self.n_sub)]
        self.launches += 3
        return owned, owned_pairs

    def reduce_on_device(self, owned, owned_pairs, geoms_by_kind=None, capacity=None):
        """Owner-side fold of the owned records into final tables (syk_table_merge_records / syk_pairs_merge).
        ``geoms_by_kind[kind]`` is the GEOM array of ALL chunks (indexed by chunk seq) used to decode rep_coord.
        Returns ({kind: records tensor}, [pairs tensor]) with one row per object / per (sub, cell) pair."""
        dev = self.dev
        out, outp = {}, []
        for k, recs in owned.items():
            need = max(1 << 16, 4 * recs.shape[0]) if capacity is None else capacity
            t = self._final.get(k)
            if t is None or t.capacity < need:   # persistent owner tables: no cudaMalloc/cudaFree inside a step
                t = self._final[k] = dev.IdTable(need)
            else:
                t.clear()
            t.merge_records(recs)
            g = np.zeros(0, GEOM_DTYPE) if geoms_by_kind is None else geoms_by_kind["cs" if k == "cs" else "cell"]
            out[k] = t.export(g, max_records=max(recs.shape[0], 1))
            self.launches += 2  # k_merge_records + k_table_export
        for c, p in enumerate(owned_pairs):
            need = max(1 << 16, 4 * p.shape[0]) if capacity is None else capacity
            t = self._final.get(("pairs", c))
            if t is None or t.capacity < need:
                t = self._final[("pairs", c)] = dev.PairTable(need)
            else:
                t.clear()
            t.merge(p)
            outp.append(t.export(max_pairs=max(p.shape[0], 1)))
            self.launches += 2  # k_pairs_merge + k_pairs_export
        return out, outp

    # ------------------------------------------------------------------------------------------ writers' inputs
    @staticmethod
    def voxel_index(owned_records: torch.Tensor):
        """Per-object voxel index of the writers (sd_proc.py:940-946, :1173-1178: ``bbs = np.concatenate(prop_dict[1][id])``,
        the object's per-chunk bounding boxes in chunk order) from the owned per-(id, chunk) record log, on the device:
        returns ``(ids int64 [N], start int64 [N + 1], boxes int32 [M, 2, 3])`` with the boxes of object ``i`` at
        ``boxes[start[i]:start[i + 1]]``, sorted by chunk sequence number."""
        if owned_records.shape[0] == 0:
            z = torch.zeros(0, dtype=torch.int64, device=owned_records.device)
            return z, torch.zeros(1, dtype=torch.int64, device=owned_records.device), \
                torch.zeros((0, 2, 3), dtype=torch.int32, device=owned_records.device)
        rec32 = owned_records.view(torch.int32).view(-1, 16)            # syk_record_t as 16 x int32
        seq = rec32[:, 15].to(torch.int64) & 0xFFFFFFFF
        o1 = torch.sort(seq, stable=True).indices                       # chunk order ...
        o2 = torch.sort(owned_records[o1, 0], stable=True).indices      # ... within every id (ids grouped as int64 bit patterns)
        order = o1[o2]
        ids_sorted = owned_records[order, 0]
        first = torch.ones(ids_sorted.shape[0], dtype=torch.bool, device=ids_sorted.device)
        first[1:] = ids_sorted[1:] != ids_sorted[:-1]
        start = torch.nonzero(first).flatten()
        boxes = rec32[order][:, 6:12].reshape(-1, 2, 3).contiguous()
        start = torch.cat([start, torch.tensor([ids_sorted.shape[0]], dtype=torch.int64, device=start.device)])
        return ids_sorted[start[:-1]], start, boxes

    def invert_mapping(self, final_pairs):
        """The second, smaller exchange of the reduce (sd_proc.py:1054-1084): every final (organelle, cell, count) pair gets
        the organelle's total size from its owner's final table (``syk_pairs_attach_size``), the pairs are re-bucketed by
        the owner of the CELL id and exchanged, and organelles that are not part of the organelle dataset (size 0: removed
        by the size threshold, :1072-1074) are dropped.  Must follow ``reduce_on_device``.  Returns, per organelle channel,
        ``(cell_id int64 [n], sub_id int64 [n], ratio float64 [n])`` sorted by (cell, organelle) bit patterns -- the
        ``mapping_{organelle}_ids`` / ``mapping_{organelle}_ratios`` of the cell supervoxels this rank owns."""
        dev, L = self.dev, self.dev._lib.load()
        out, bucketed, cnts = [], [], []
        for c, p in enumerate(final_pairs):
            p = p.clone() if p.shape[0] else p
            t = self._final.get(f"sub{c}")
            assert t is not None, "invert_mapping needs the final organelle tables of reduce_on_device"
            if p.shape[0]:
                dev.check(L.syk_pairs_attach_size(p.data_ptr(), p.shape[0], t.h, dev._stream_ptr()))
            self.launches += 1
            if self.world > 1:
                b, cc = dev.bucket_pairs(p[:, [1, 0, 2, 3]].contiguous(), self.world)   # owner key = first column = cell id
                bucketed.append(b)
                cnts.append(cc)
                self.launches += 3
            else:
                bucketed.append(p[:, [1, 0, 2, 3]])
        got = exchange_logs(bucketed, cnts, self.world, self.group) if self.world > 1 else bucketed
        for p in got:
            p = p[p[:, 3] != 0]
            o1 = torch.sort(p[:, 1], stable=True).indices
            o2 = torch.sort(p[o1, 0], stable=True).indices
            p = p[o1[o2]]
            out.append((p[:, 0], p[:, 1], p[:, 2].to(torch.float64) / p[:, 3].to(torch.float64)))
        return out

