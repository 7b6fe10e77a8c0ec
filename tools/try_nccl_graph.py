"""Debug driver: the graph-captured sample-sharded step on 2 GPUs, with progress prints (run under `timeout`)."""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pertrenderer_b200 as pb
from pertrenderer_b200 import dist as pdist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
def log(*a):
    print(f"[rank {rank} {time.strftime('%H:%M:%S')}]", *a, flush=True)
N, H, W, K, S = 1, 32, 32, 50, 64
fr, col = pb.synthetic_fragments(N, H, W, K, kind="realistic", sigma=1e-3, seed=3, device=dev)
G = torch.randn((N, H, W, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(5))
log("building step")
step = pdist.GraphedSampleShardedStep(fr.pix_to_face, fr.zbuf.contiguous(), fr.dists.contiguous(), col, G, sigma=1e-3, gamma=1e-2,
                                      S_rast=S, S_agg=S, seed=4242)
log("captured")
for i in range(3):
    out = step.replay()
    torch.cuda.synchronize(dev)
    log("replay", i, float(out[0].sum()), float(out[1].sum()))
step.close()
dist.barrier()
log("done")
dist.destroy_process_group()
