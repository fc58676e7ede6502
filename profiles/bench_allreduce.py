#!/usr/bin/env python
"""All-reduce latency of the payloads the sharded iteration exchanges (k x d = 2.8 MB and the
k x k statistics), inside a CUDA graph, for the NCCL settings given in the environment."""
import json
import os
import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
res = {}
for name, n in (('kxd_2.8MB', 8 * 44000), ('3kxk_1.5KB', 192), ('kxT_104KB', 8 * 1620)):
    t = torch.ones(n, dtype=torch.float64, device='cuda')
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g.capture_begin()
        for _ in range(20):
            dist.all_reduce(t)
        g.capture_end()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / 100 * 1e3, 2)
if rank == 0:
    print(json.dumps({'world': world, 'NCCL_ALGO': os.environ.get('NCCL_ALGO'),
                      'NCCL_PROTO': os.environ.get('NCCL_PROTO'), 'us_per_allreduce': res}))
dist.barrier()
dist.destroy_process_group()
