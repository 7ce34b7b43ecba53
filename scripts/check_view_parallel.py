"""Multi-rank check of the view-parallel exchange (sugar_b200/parallel.py) against a plain all-reduce:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_view_parallel.py
Every rank renders its own view of a small scene; the gradients summed inside the backward (SH factors gathered,
44-byte records reduced chunk by chunk) must equal the locally computed gradients summed with dist.all_reduce.
Runs bench.py's own `verify_exchange` for both SH modes and several chunk counts; prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import bench

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
from sugar_b200 import diff_gaussian_rasterization as mod, parallel
scenes = bench.load_scenes()
res = {}
os.environ.setdefault("SGR_PEER_WORLDS", str(world))   # "auto" = peer memory at this world size, NCCL if it cannot be mapped
fallback = None
for factors, peer in ((True, "auto"), (True, False), (False, False)):
    for chunks in (1, 4, 7):
        r = bench.verify_exchange(torch, dist, mod, parallel, scenes, dev, rank, world, 3, sh_factors=factors, chunks=chunks,
                                  peer=peer)
        res[f"factors={factors},peer={r['peer_memory']},chunks={chunks}"] = r["max_rel_err"]
        fallback = fallback or r.get("peer_fallback_reason")
if rank == 0:
    print(json.dumps({"world": world, "max_rel_err": res, "peer_fallback_reason": fallback}))
dist.destroy_process_group()
