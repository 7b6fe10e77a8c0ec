"""Multi-GPU paths on real devices (skipped with fewer than 2 GPUs): one process per GPU over NCCL.

Sample sharding (BASELINE config 4): every rank ends with the single-GPU result —
indices exactly, gradients to rounding.  Batch sharding (config 3): the union of the ranks' images and
gradients is bit-identical to the single-GPU job; only d/d(sigma, gamma, alpha) are all-reduced."""

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 CUDA devices")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(device):
    import pertrenderer_b200 as pb
    N, H, W, K, S = 4, 16, 16, 50, 64
    fr, col = pb.synthetic_fragments(N, H, W, K, kind="realistic", sigma=1e-3, seed=3, device="cpu")
    G = torch.randn((N, H, W, 4), generator=torch.Generator().manual_seed(5))
    return N, S, fr, col, G


def _render(pb, fr, col, G, dev, S, sharded_group=None, pixel_offset=0):
    from pertrenderer_b200 import _cabi, ops
    rast = pb.GaussianRast(nb_samples=S, sigma=1e-3)
    agg = pb.GaussianAgg(nb_samples=S, gamma=1e-2, alpha=1.1)
    d = fr.dists.to(dev).requires_grad_(True)
    z = fr.zbuf.to(dev).requires_grad_(True)
    c = col.to(dev).requires_grad_(True)
    frag = pb.Fragments(fr.pix_to_face.to(dev), z, None, d)
    blend = pb.BlendParams(background_color=(0.2, 0.5, 0.8))
    N = d.shape[0]
    zn, zf = torch.full((N,), 1.0, device=dev), torch.full((N,), 100.0, device=dev)
    torch.manual_seed(77)  # every rank / the single-GPU run draw the same two seeds
    with ops.kernel_flags(_cabi.F_PER_SAMPLE_NOISE):  # sample-path comparable across shardings
        if sharded_group is not None:
            from pertrenderer_b200.dist import smooth_rgb_blend_sample_sharded
            img = smooth_rgb_blend_sample_sharded(c, frag, rast, agg, blend, znear=zn, zfar=zf, group=sharded_group,
                                                  sync_seeds=True)
        else:
            from pertrenderer_b200.random_rasterizer import _PerturbedShade, _background_tuple
            cfg = dict(background=_background_tuple(blend), eps=agg.eps, S_rast=S, S_agg=S, fixed_noise=False,
                       pixel_offset=pixel_offset)
            img = _PerturbedShade.apply(c, d, z, rast.sigma, agg.gamma, agg.alpha, frag.pix_to_face, zn, zf, cfg)
        (img * G.to(dev)).sum().backward()
    return dict(image=img.detach().cpu(), gd=d.grad.cpu(), gz=z.grad.cpu(), gc=c.grad.cpu(),
                scal=torch.stack([rast.sigma.grad, agg.gamma.grad, agg.alpha.grad])), (rast, agg)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import dist as pdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        N, S, fr, col, G = _case(dev)
        # ---- noise-sample sharding: inputs replicated, samples split, three all-reduces ----------------
        out_s, _ = _render(pb, fr, col, G, dev, S, sharded_group=dist.group.WORLD)
        # ---- batch sharding: views split, global pixel offsets, 3-float all-reduce of the scalar grads ---
        b0, b1 = pdist.batch_range(N, world, rank)
        sl = slice(b0, b1)
        frs = pb.Fragments(fr.pix_to_face[sl], fr.zbuf[sl], None, fr.dists[sl])
        out_b, (rast, agg) = _render(pb, frs, col[sl], G[sl], dev, S, pixel_offset=b0 * fr.zbuf.shape[1] * fr.zbuf.shape[2])
        pdist.all_reduce_scalar_grads([rast.sigma, agg.gamma, agg.alpha], device=dev)
        out_b["scal"] = torch.stack([rast.sigma.grad, agg.gamma.grad, agg.alpha.grad])
        torch.save(dict(sample=out_s, batch=out_b, span=(b0, b1)), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_nccl_sample_and_batch_sharding_world2(tmp_path):
    import torch.multiprocessing as mp
    import pertrenderer_b200 as pb
    from conftest import rel_err
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    N, S, fr, col, G = _case("cuda:0")
    whole, _ = _render(pb, fr, col, G, torch.device("cuda", 0), S)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    for o in outs:
        s = o["sample"]
        assert (s["image"] - whole["image"]).abs().max() <= 1e-6
        assert rel_err(s["gd"], whole["gd"]) <= 1e-5 and rel_err(s["gz"], whole["gz"]) <= 1e-5
        assert rel_err(s["gc"], whole["gc"]) <= 1e-6
        assert torch.allclose(s["scal"], whole["scal"], rtol=1e-4, atol=1e-6)
    for k in ("image", "gd", "gz", "gc"):
        assert torch.equal(outs[0]["sample"][k], outs[1]["sample"][k]), k  # ranks agree exactly
        cat = torch.cat([o["batch"][k] for o in outs])
        assert torch.equal(cat, whole[k]), k  # batch shards: bit-identical union
    assert torch.allclose(outs[0]["batch"]["scal"], whole["scal"], rtol=1e-5, atol=1e-7)
    assert torch.equal(outs[0]["batch"]["scal"], outs[1]["batch"]["scal"])


def _graph_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import dist as pdist
    from pertrenderer_b200 import ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        N, S, fr, col, G = _case(dev)
        p2f, z, d, c, Gd = (t.to(dev).contiguous() for t in (fr.pix_to_face, fr.zbuf, fr.dists, col, G))
        step = pdist.GraphedSampleShardedStep(p2f, z, d, c, Gd, sigma=1e-3, gamma=1e-2, alpha=1.1, S_rast=S, S_agg=S,
                                              background=(0.2, 0.5, 0.8), seed=4242, flags=ops.F_PER_SAMPLE_NOISE)
        out1 = [t.clone().cpu() for t in step.replay()]
        seeds1 = step.seed_device.clone().cpu()
        out2 = [t.clone().cpu() for t in step.replay()]
        step.close()  # before destroy_process_group
        torch.save(dict(out1=out1, out2=out2, seeds1=seeds1), os.path.join(out_dir, f"g{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_nccl_graph_captured_sample_sharded_step_world2(tmp_path):
    """The sample-sharded step with its three NCCL all-reduces captured in one CUDA graph per rank: ranks agree exactly,
    the result is the single-GPU job with the same effective seeds (per-sample flag: same sample path), and the next
    replay draws fresh noise."""
    import torch.multiprocessing as mp
    from conftest import rel_err
    from pertrenderer_b200 import ops
    world = 2
    mp.spawn(_graph_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"g{r}.pt") for r in range(world)]
    for a, b in zip(outs[0]["out1"], outs[1]["out1"]):
        assert torch.equal(a, b)
    assert torch.equal(outs[0]["seeds1"], outs[1]["seeds1"])
    assert not torch.equal(outs[0]["out1"][0], outs[0]["out2"][0])
    N, S, fr, col, G = _case("cuda:0")
    dev = torch.device("cuda", 0)
    seeds = outs[0]["seeds1"]
    pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face.to(dev), zbuf=fr.zbuf.to(dev), dists=fr.dists.to(dev), colors=col.to(dev),
                          znear=1.0, zfar=100.0, background=(0.2, 0.5, 0.8), sigma=1e-3, gamma=1e-2, alpha=1.1, eps=1e-10,
                          S_rast=S, S_agg=S, seed_rast=int(seeds[0]), seed_agg=int(seeds[1]), flags=ops.F_PER_SAMPLE_NOISE)
    image, saved = ops.shade_forward(pr)
    gd, gz, gc, scal = ops.shade_backward(pr, saved, G.to(dev))
    o = outs[0]["out1"]
    assert (o[0] - image.cpu()).abs().max() <= 1e-6
    assert rel_err(o[1], gd.cpu()) <= 1e-5 and rel_err(o[2], gz.cpu()) <= 1e-5 and rel_err(o[3], gc.cpu()) <= 1e-6
    assert torch.allclose(o[4], scal.cpu(), rtol=1e-4, atol=1e-6)
