"""Only the multi-GPU carving section of bench.py (carve_sharded_bench at 1024^3), for quick runs under torchrun:
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/probe_carve_multi.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = bench.carve_sharded_bench(int(os.environ.get("PROBE_N", "1024")), dev, world, rank, dist)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
