#!/usr/bin/env python3
"""Latency of one halo exchange per level (torchrun, one rank per GPU): scripts/bench_halo.py [n]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pde_multigrid_b200 as mg  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    raw = (ctypes.c_ubyte * 128)()
    mg._lib.check(mg.lib().mg_comm_unique_id(raw))
    buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
dist.broadcast(buf, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
e = mg.MultiGrid3D(n, dtype=np.float64, residual_mode=mg.MG_CORRECTED, rank=rank, nranks=world, nccl_unique_id=bytes(buf.cpu().tolist()))
s = torch.cuda.ExternalStream(e.stream)
for level in range(3):
    for mask, up, down in ((2, 4, 4), (3, 2, 1), (3, 4, 4), (2, 1, 1)):
        e._call("halo_benchmark", ctypes.c_int(level), ctypes.c_int(mask), ctypes.c_int(up), ctypes.c_int(down), ctypes.c_int(5))
        e.sync()
        dist.barrier()
        reps = 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        e._call("halo_benchmark", ctypes.c_int(level), ctypes.c_int(mask), ctypes.c_int(up), ctypes.c_int(down), ctypes.c_int(reps))
        b.record(s)
        e.sync()
        if rank == 0:
            nl = e.level_size(level)
            print(json.dumps({"level": level, "n": nl, "colours": mask, "up": up, "down": down, "us_per_exchange": a.elapsed_time(b) * 1e3 / reps,
                              "bytes_per_direction": (1 if mask != 3 else 2) * max(up, down) * ((nl + 1) // 2 + 7) // 8 * 8 * nl * 8}), flush=True)
e.close()
dist.destroy_process_group()
