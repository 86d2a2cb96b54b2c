"""Multi-GPU parity check: the sharded pipeline + hash-owner all-to-all must give exactly the records of a
single-rank run over all chunks (the reference's reduce, syconn/proc/sd_proc.py:511-556, 1248-1322).

  torchrun --nproc-per-node N tools/mgpu_check.py        (prints "MGPU_CHECK PASS world N")

``sharded_equals_single`` is also called by bench.py after its timed region at every N > 1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syconn_b200 import device as dev  # noqa: E402
from syconn_b200._lib import GEOM_DTYPE  # noqa: E402
from syconn_b200.chunked import ChunkPlan, ExtractionPipeline, cs_halo_geometry  # noqa: E402


def _canon(t):
    a = t.cpu().numpy()
    return a[np.lexsort(a.T[::-1])] if len(a) else a


def sharded_equals_single(rank, world, E=96, st=(13, 13, 7), nsub=3, grid=(2, 2, 2), log=print):
    """Run the pipeline sharded over `world` ranks and, on rank 0, once more over ALL chunks with world = 1; compare the
    final records and overlap pairs row by row.  Returns True/False on every rank (broadcast from rank 0)."""
    plan = ChunkPlan(tuple(g * E for g in grid), (E, E, E))
    geoms = {"cell": np.zeros(len(plan), GEOM_DTYPE), "cs": np.zeros(len(plan), GEOM_DTYPE)}
    for s in range(len(plan)):
        geoms["cell"][s] = geoms["cs"][s] = (plan.offsets[s], plan.sizes[s])

    def run(pipe, seqs):
        pipe.reset()
        for s in seqs:
            off, size = plan.offsets[s], plan.sizes[s]
            lo, ls, _, _ = cs_halo_geometry(off, size, st)
            cell = dev.synth_labels(size, off, (24, 20, 12), 4, 7, 0, order="F")
            subs = torch.empty((nsub,) + tuple(size[::-1]), dtype=torch.int64, device="cuda").permute(0, 3, 2, 1)
            for c in range(nsub):
                dev.synth_labels(size, off, (9, 8, 5), 4, 7, 1 + c, 2, out=subs[c])
            halo = dev.synth_labels(ls, lo, (24, 20, 12), 4, 7, 0, dtype=torch.int32, order="F")
            pipe.process_chunk(s, off, cell, subs, halo)
        owned, owned_pairs = pipe.finish()
        return pipe.reduce_on_device(owned, owned_pairs, geoms)

    mine = plan.chunks_of_rank(rank, world)
    final, final_pairs = run(ExtractionPipeline(nsub, st, 1 << 14, 1 << 16, 1 << 16, rank=rank, world=world), mine)
    gathered, gathered_pairs = {}, []
    for k in sorted(final):
        parts = [None] * world
        dist.all_gather_object(parts, _canon(final[k]))
        gathered[k] = np.concatenate([p for p in parts if len(p)]) if any(len(p) for p in parts) else parts[0]
    for c in range(nsub):
        parts = [None] * world
        dist.all_gather_object(parts, _canon(final_pairs[c]))
        gathered_pairs.append(np.concatenate([p for p in parts if len(p)]) if any(len(p) for p in parts) else parts[0])
    ok = True
    if rank == 0:
        ref, ref_pairs = run(ExtractionPipeline(nsub, st, 1 << 14, 1 << 16, 1 << 16, rank=0, world=1), list(range(len(plan))))
        for k in sorted(ref):
            a, b = gathered[k], _canon(ref[k])
            a = a[np.lexsort(a.T[::-1])] if len(a) else a
            same = a.shape == b.shape and np.array_equal(a, b)
            log(f"{k}: {len(b)} objects, sharded == single: {same}")
            ok &= bool(same)
        for c in range(nsub):
            a, b = gathered_pairs[c], _canon(ref_pairs[c])
            a = a[np.lexsort(a.T[::-1])] if len(a) else a
            same = a.shape == b.shape and np.array_equal(a, b)
            log(f"pairs{c}: {len(b)} pairs, sharded == single: {same}")
            ok &= bool(same)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = sharded_equals_single(rank, world)
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL", "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
