"""Multi-GPU parity check (run under torchrun): the sharded pipeline + hash-owner all-to-all must give exactly the
records of a single-rank run over all chunks.  torchrun --nproc-per-node N tools/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syconn_b200 import device as dev  # noqa: E402
from syconn_b200._lib import GEOM_DTYPE  # noqa: E402
from syconn_b200.chunked import ChunkPlan, ExtractionPipeline, cs_halo_geometry  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E, st, nsub = 96, (13, 13, 7), 3
plan = ChunkPlan((2 * E, 2 * E, 2 * E), (E, E, E))
geoms = {"cell": np.zeros(len(plan), GEOM_DTYPE), "cs": np.zeros(len(plan), GEOM_DTYPE)}
for s in range(len(plan)):
    lo, ls, oo, os_ = cs_halo_geometry(plan.offsets[s], plan.sizes[s], st)
    geoms["cell"][s], geoms["cs"][s] = (plan.offsets[s], plan.sizes[s]), (oo, os_)


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


def canon(t):
    a = t.cpu().numpy()
    return a[np.lexsort(a.T[::-1])] if len(a) else a


mine = plan.chunks_of_rank(rank, world)
final, final_pairs = run(ExtractionPipeline(nsub, st, 1 << 14, 1 << 16, 1 << 16, rank=rank, world=world), mine)
ok = True
# gather every rank's owned results on rank 0
for k in sorted(final):
    parts = [None] * world
    dist.all_gather_object(parts, canon(final[k]))
    if rank == 0:
        final[k] = np.concatenate([p for p in parts if len(p)])
pp = []
for c in range(nsub):
    parts = [None] * world
    dist.all_gather_object(parts, canon(final_pairs[c]))
    if rank == 0:
        pp.append(np.concatenate([p for p in parts if len(p)]))
if rank == 0:
    ref, ref_pairs = run(ExtractionPipeline(nsub, st, 1 << 14, 1 << 16, 1 << 16, rank=0, world=1), list(range(len(plan))))
    for k in sorted(ref):
        a, b = final[k], canon(ref[k])
        a = a[np.lexsort(a.T[::-1])]
        same = a.shape == b.shape and np.array_equal(a, b)
        print(f"{k}: {len(b)} objects, sharded == single: {same}")
        ok &= same
    for c in range(nsub):
        a, b = pp[c], canon(ref_pairs[c])
        a = a[np.lexsort(a.T[::-1])]
        same = a.shape == b.shape and np.array_equal(a, b)
        print(f"pairs{c}: {len(b)} pairs, sharded == single: {same}")
        ok &= same
    print("MGPU_CHECK", "PASS" if ok else "FAIL", "world", world)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
