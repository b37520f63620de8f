#!/usr/bin/env python
"""Dev: only the strong-scaling sub-record of bench.py (C4 / C5 fixed 4K frame over the ranks), under torchrun.
TRAY_BENCH_STRONG_EXCHANGE=peer|peer-nccl|nccl|none picks how (whether) the shards reach rank 0's frame."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from tray_racing_b200 import cuda, host  # noqa: E402

world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
rec = bench.strong_scaling_record(torch, dist, cuda, host, rank, world, local_rank, stream, 40, which=which)
if rank == 0:
    print(json.dumps({k: v for k, v in rec.items() if k not in ("workload", "note")}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
for sc in bench._keep_alive:
    sc.close()
