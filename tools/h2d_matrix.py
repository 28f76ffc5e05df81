#!/usr/bin/env python3
"""Host -> device contention matrix of one box: which GPUs share an uplink / a host-memory path?
Launch with torchrun (one rank per GPU, gloo rendezvous on 127.0.0.1).  Each rank binds to its GPU's NUMA node, allocates a pinned
256 MiB buffer and, phase by phase, the ranks of the phase's subset copy host -> device concurrently for ~0.4 s while the others
idle.  Rank 0 prints one JSON line per phase: the subset and each member's GB/s.  `--d2h` adds result traffic device -> host on a
second stream in the e2e leg's proportion (1 : 5)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import bind_to_gpu_numa_node

ap = argparse.ArgumentParser(); ap.add_argument("--d2h", action="store_true"); ap.add_argument("--seconds", type=float, default=0.4)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
note = bind_to_gpu_numa_node(lr)
torch.cuda.set_device(lr)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ.get("MASTER_PORT", "29541"), rank=rank, world_size=world)
N = 256 << 20
h = torch.empty(N, dtype=torch.uint8).pin_memory(); h.fill_(rank)
d = torch.empty(N, dtype=torch.uint8, device="cuda")
h2 = torch.empty(N // 5, dtype=torch.uint8).pin_memory(); d2 = torch.empty(N // 5, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
singles = [[i] for i in range(world)]
sets = singles + [s for s in ([0, 1], [0, 2], [0, 3], [0, 4], [0, 7], [2, 3], [4, 5], [6, 7], [0, 1, 2, 3], [4, 5, 6, 7], [0, 2, 4, 6], [0, 1, 4, 5], list(range(world))) if max(s) < world]
notes = [None] * world
dist.all_gather_object(notes, note)
if rank == 0:
    print(json.dumps({"numa": notes, "d2h": args.d2h}), flush=True)
for sub in sets:
    dist.barrier()
    gbs = 0.0
    if rank in sub:
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < args.seconds:
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
            if args.d2h:
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
            s1.synchronize(); n += 1
        torch.cuda.synchronize()
        gbs = n * N / (time.perf_counter() - t0) / 1e9
    out = [None] * world
    dist.all_gather_object(out, gbs)
    if rank == 0:
        print(json.dumps({"gpus": sub, "h2d_GBps": [round(out[i], 1) for i in sub], "sum": round(sum(out[i] for i in sub), 1)}), flush=True)
dist.barrier()
dist.destroy_process_group()
