import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syconn_b200 import device as dev
from syconn_b200.chunked import exchange_buckets
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
recs = torch.randint(1, 1 << 40, (400000, 8), dtype=torch.int64, device="cuda")
for it in range(4):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    b, counts = dev.bucket_records(recs, world)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    c = counts.tolist()
    t2 = time.perf_counter()
    out = exchange_buckets(b, c)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if rank == 0:
        print(f"iter {it}: bucket {1e3*(t1-t0):.2f} ms, tolist {1e3*(t2-t1):.2f} ms, exchange {1e3*(t3-t2):.2f} ms, rows {out.shape[0]}", flush=True)
dist.destroy_process_group()
