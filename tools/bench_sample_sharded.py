#!/usr/bin/env python
"""BASELINE config 4 (high-variance regime): 1 x 128x128, K=50, nb_samples=4096, the noise samples
sharded over the ranks with three all-reduces per forward+backward (pertrenderer_b200/dist.py).
Strong scaling: total work fixed.  Launch: python -m torch.distributed.run --nproc-per-node N tools/bench_sample_sharded.py
Prints one JSON line on rank 0 (device-timed with CUDA events, max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pertrenderer_b200 as pb  # noqa: E402
from pertrenderer_b200.dist import smooth_rgb_blend_sample_sharded  # noqa: E402

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
N, HW, K, S = 1, 128, 50, int(os.environ.get("S", "4096"))
steps, warmup = 20, 5
fr, col = pb.synthetic_fragments(N, HW, HW, K, kind="realistic", sigma=1e-3, seed=0, device=dev)  # replicated inputs
G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
rast, agg = pb.GaussianRast(nb_samples=S, sigma=1e-3), pb.GaussianAgg(nb_samples=S, gamma=1e-2)
blend = pb.BlendParams(background_color=(1.0, 1.0, 1.0))
zn, zf = torch.ones(N, device=dev), torch.full((N,), 100.0, device=dev)


def step():
    d = fr.dists.detach().requires_grad_(True)
    z = fr.zbuf.detach().requires_grad_(True)
    c = col.detach().requires_grad_(True)
    frag = pb.Fragments(fr.pix_to_face, z, None, d)
    if world > 1:
        img = smooth_rgb_blend_sample_sharded(c, frag, rast, agg, blend, znear=zn, zfar=zf)
    else:
        img = pb.smooth_rgb_blend(c, frag, rast, agg, blend, znear=zn, zfar=zf)
    (img * G).sum().backward()
    return d.grad


for _ in range(warmup):
    step()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    g = step()
e1.record()
torch.cuda.synchronize()
dist.barrier()
t = torch.tensor([e0.elapsed_time(e1)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = t.item() / steps
if rank == 0:
    print(json.dumps({"metric": "perturbed shader fwd+bwd pixel·face·samples/sec", "value": N * HW * HW * K * S / (ms * 1e-3),
                      "unit": "pixel·face·samples/s", "n_gpus": world, "ms_per_step": ms, "scaling": "strong",
                      "config": {"workload": f"BASELINE config 4: {N}x{HW}x{HW}, K={K}, nb_samples={S}, noise samples sharded over "
                                             f"{world} ranks, 3 all-reduces per fwd+bwd (public autograd API, host-side scalar reads included)"}}))
dist.destroy_process_group()
