"""C4 driver: the add-node search as a batch of what-if scenarios, sharded over ranks (one process per GPU).

  python tools/capacity_run.py                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/capacity_run.py

Every rank builds the same scenario set (seeded), runs scenarios sid % world == rank through simon_scenarios_run and
joins ONE all_reduce(MIN) of the packed key (k << 32 | scenario id) over NCCL.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "open-simulator_b200"), ROOT):
    sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=2000)
ap.add_argument("--workloads", type=int, default=50)
ap.add_argument("--replicas", type=int, default=100)
ap.add_argument("--ks", type=int, default=32, help="scenarios per spec: k = 1..ks (8 specs x ks scenarios)")
a = ap.parse_args()

import torch
from simon_b200 import capacity, synth

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
reduce_fn = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    reduce_fn = capacity.torch_all_reduce_min(f"cuda:{local}")

t0 = time.perf_counter()
cluster, apps, specs = synth.make_c4(n_nodes=a.nodes, n_workloads=a.workloads, replicas=a.replicas)
ss = capacity.build_scenarios(cluster, apps, specs, list(range(1, a.ks + 1)))
t_build = time.perf_counter() - t0
# warm-up pass: CUDA context, engine buffers and the first NCCL collective (communicator set-up) are not part of the search time
capacity.search(ss, capacity.gpu_runner(local), rank=rank, world=world, all_reduce_min=reduce_fn)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t1 = time.perf_counter()
best, local_res = capacity.search(ss, capacity.gpu_runner(local), rank=rank, world=world, all_reduce_min=reduce_fn)
torch.cuda.synchronize()
dt = time.perf_counter() - t1
if world > 1:
    tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
if rank == 0:
    dec = capacity.decode_key(best)
    pods_per_scen = int((ss.compiled.pods["pod_fixed_node"] == -1).sum())
    print(json.dumps({"scenarios": len(ss.scenarios), "n_gpus": world, "specs": len(specs), "best": dec,
                      "best_spec": (ss.scenarios[dec["scenario"]].spec if dec else None),
                      "search_s": dt, "host_build_s": t_build, "nodes": a.nodes, "scheduled_pods_per_scenario": pods_per_scen,
                      "decisions_per_s": len(ss.scenarios) * pods_per_scen / dt,
                      "collective": "one all_reduce(MIN) of int64 (NCCL)" if world > 1 else "none (single process)"}), flush=True)
if world > 1:
    dist.destroy_process_group()
