"""Host-side logic of the multi-GPU paths on CPU: partitioning helpers, and the sample-sharded
orchestration (three all-reduces) driven with world_size=2 over gloo.  The CUDA
phases are replaced by the oracle's per-shard sums (test infrastructure), so what is under test is
the sequence of collectives and what is exchanged, against the unsharded oracle."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, rel_err
from oracle import pert_oracle as O
from pertrenderer_b200 import dist as pdist


def test_batch_range_partitions():
    for n, w in [(64, 8), (8, 8), (10, 4), (3, 4), (1, 2)]:
        spans = [pdist.batch_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pdist.batch_range(4, 2, 2)


def test_sample_range_is_quad_aligned():
    for S, w in [(4096, 8), (64, 2), (16, 4), (10, 2), (6, 4), (256, 3)]:
        spans = [pdist.sample_range(S, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == S
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(a % 4 == 0 for a, b in spans if b > a)
    assert pdist.sample_range(6, 4, 3) == (6, 6)  # empty shard past the last quad


class OracleStages:
    """Same interface as pertrenderer_b200.dist.CudaStages, computed by the CPU oracle on this
    rank's slice of explicit noise."""

    def __init__(self, g, sr, sa):
        self.g, self.sr, self.sa = g, sr, sa
        self.U, self.V = g["U"][sr[0]:sr[1]], g["V"][sa[0]:sa[1]]

    def rast(self):
        counts, self.rsum_local = O.rast_shard_sums(self.g["dists"], self.U, self.g["sigma"])
        return counts

    def agg(self, counts):
        g = self.g
        self.zeta, self.prob, self.aux = O.logits_from_counts(
            g["pix_to_face"], g["zbuf"], counts, int(g["S_r"]), g["znear_t"], g["zfar_t"], g["gamma"], g["alpha"], g["eps"])
        hist, self.a_s, self.a_0 = O.argmax_shard(self.zeta, self.V, g["gamma"])
        return hist

    def blend(self, hist):
        g = self.g
        image, self.weights = O.blend_from_hist(hist, int(g["S_a"]), self.prob, g["colors"], g["background"])
        return image

    def bwd_sample(self, grad_image):
        # one flat buffer for the one gradient all-reduce: [argmax score sums | per-pixel sums | coverage score sums]
        g = self.g
        grad_w = O._grad_weights(g["colors"], grad_image, g["background"])
        self.packed_shape = tuple(grad_w.shape[:-1]) + (grad_w.shape[-1] + 2,)
        packed = O.argmax_score_sums(grad_w, self.a_s, self.a_0, self.V)
        return torch.cat((packed.flatten(), self.rsum_local.flatten()))

    def bwd_finish(self, grad_image, flat, need_colors=True):
        g = self.g
        n = 1
        for d in self.packed_shape:
            n *= d
        packed, rsum = flat[:n].view(self.packed_shape), flat[n:].view(self.rsum_local.shape)
        aux = dict(self.aux, bg=g["background"])
        gr = O.shade_backward_from_sums(self.prob, self.weights, aux, grad_image, g["zbuf"], g["colors"], g["sigma"],
                                        g["gamma"], g["alpha"], g["eps"], int(g["S_r"]), int(g["S_a"]), rsum, packed)
        return gr["dists"], gr["zbuf"], gr["colors"], torch.stack((gr["sigma"], gr["gamma"], gr["alpha"]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("shade_" + case)
        g["znear_t"], g["zfar_t"] = g["znear"].reshape(-1, 1, 1, 1), g["zfar"].reshape(-1, 1, 1, 1)
        sr = pdist.sample_range(int(g["S_r"]), world, rank)
        sa = pdist.sample_range(int(g["S_a"]), world, rank)
        stages = OracleStages(g, sr, sa)
        image = pdist.sharded_forward(stages)
        gd, gz, gc, scal = pdist.sharded_backward(stages, g["grad_image"])
        # batch-sharding helper: sum of three CPU scalar leaves
        leaves = [torch.tensor(float(rank + 1), requires_grad=True) for _ in range(3)]
        for i, t in enumerate(leaves):
            t.grad = torch.tensor(float(10 * i + rank))
        pdist.all_reduce_scalar_grads(leaves)
        torch.save(dict(image=image, gd=gd, gz=gz, gc=gc, scal=scal, leaf=[t.grad for t in leaves], sr=sr, sa=sa),
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["small", "uneven"])
def test_sample_sharded_orchestration_gloo_world2(case, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    g = load_golden("shade_" + case)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert outs[0]["sr"][1] == outs[1]["sr"][0] and outs[1]["sr"][1] == int(g["S_r"])
    for o in outs:
        # forward: the exchanged integer state reproduces the unsharded image (to fp rounding of
        # the blend sum); backward within the gradient tolerance
        assert (o["image"] - g["image"]).abs().max() <= 1e-6
        assert rel_err(o["gd"], g["grad_dists"]) <= 1e-5
        assert rel_err(o["gz"], g["grad_zbuf"]) <= 1e-5
        assert rel_err(o["gc"], g["grad_colors"]) <= 1e-6
        for i, k in enumerate(("sigma", "gamma", "alpha")):
            ref = g["grad_" + k]
            assert abs(o["scal"][i].item() - ref) <= 2e-5 * max(abs(ref), 1e-12) + 1e-8
        assert [v.item() for v in o["leaf"]] == [1.0, 21.0, 41.0]
    # every rank ends with identical results
    for k in ("image", "gd", "gz", "gc", "scal"):
        assert torch.equal(outs[0][k], outs[1][k]), k


def _small_sample_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # nb_samples = 6 -> two quads for four ranks: sample_range gives ranks 2 and 3 an empty shard.  The check must
        # raise on EVERY rank (it is decided on data all ranks share); before, only the empty ranks raised and the
        # others blocked in the first all-reduce
        try:
            pdist.check_sample_sharding(6, 6, world)
            raised = False
        except ValueError:
            raised = True
        empty = pdist.sample_range(6, world, rank)
        flag = torch.tensor([1.0 if raised else 0.0])
        dist.all_reduce(flag)  # every rank reaches the collective: nobody raised alone before it
        torch.save(dict(raised=raised, empty=empty[0] == empty[1], total=flag.item()), os.path.join(out_dir, f"s{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_too_few_samples_raise_on_every_rank(tmp_path):
    world = 4
    mp.spawn(_small_sample_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"s{r}.pt") for r in range(world)]
    assert all(o["raised"] for o in outs) and outs[0]["total"] == world
    assert [o["empty"] for o in outs] == [False, False, True, True]
    pdist.check_sample_sharding(16, 8, 2)  # enough quads: no error
    with pytest.raises(ValueError):
        pdist.check_sample_sharding(16, 4, 2)
