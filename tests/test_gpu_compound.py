"""Statistical suite of the compound coverage sampler (csrc/tile.cuh rast_compound_list), the default
coverage stage with in-kernel noise.

The sampler draws, per fragment entry, the NUMBER of flipped coverage samples from Binomial(S, Phi(-|x|/sigma))
and that many tail normals, instead of S per-sample draws (randomras/smoothrast.py:21,32-36,46).  It is exact in
law, not per sample path, so it is checked (1) against the closed forms of that law, (2) against the per-sample
path of the same kernels (PERT_F_PER_SAMPLE_NOISE, which IS oracle-checked sample by sample in
test_gpu_parity.py) as a two-sample test, and (3) through the rendered image and its gradients.
"""

import math

import pytest
import torch

from oracle import pert_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


SIGMA = 1e-3
# |x| / sigma of the test entries: both sides of the hand-over threshold (0.84 for S <= 240) and out to the bound
T_VALUES = [0.0, 0.3, 0.8, 0.85, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0, 5.6]


def _coverage_run(S, flags, seed, signs=(1.0, -1.0), H=96, W=96, s_range=None):
    """Coverage phase only (PERT_PH_RAST) on a (1,H,W,K) problem whose entry k of every pixel has
    x = sign * T_VALUES[k % nT] * sigma.  Returns counts (P,K) int64, rsum (P,K) float64, x (K,)."""
    from pertrenderer_b200 import _cabi, ops
    dev = "cuda"
    K = len(T_VALUES) * len(signs)
    x = torch.tensor([s * t * SIGMA for s in signs for t in T_VALUES], dtype=torch.float32)
    dists = (-x).reshape(1, 1, 1, K).expand(1, H, W, K).contiguous().to(dev)
    p2f = torch.zeros((1, H, W, K), dtype=torch.int64, device=dev)
    z = torch.linspace(2.0, 3.0, K).reshape(1, 1, 1, K).expand(1, H, W, K).contiguous().to(dev)
    col = torch.zeros((1, H, W, K, 3), device=dev)
    pr = ops.ShadeProblem(pix_to_face=p2f, zbuf=z, dists=dists, colors=col, znear=1.0, zfar=100.0,
                          background=(1.0, 1.0, 1.0), sigma=SIGMA, gamma=1e-2, alpha=1.0, eps=1e-10, S_rast=S, S_agg=4,
                          seed_rast=seed, seed_agg=seed + 1, flags=flags, s_rast=s_range)
    _, saved = ops.shade_forward(pr, phases=_cabi.PH_RAST)
    torch.cuda.synchronize()
    cnt = (saved.counts.to(torch.int64) & 0xFFFF).reshape(-1, K).cpu()
    rs = saved.rsum.double().reshape(-1, K).cpu()
    return cnt, rs, x.double()


def _binom_pmf(n, p):
    k = torch.arange(n + 1, dtype=torch.float64)
    logc = torch.lgamma(torch.tensor(n + 1.0, dtype=torch.float64)) - torch.lgamma(k + 1) - torch.lgamma(n - k + 1)
    if p <= 0:
        out = torch.zeros(n + 1, dtype=torch.float64)
        out[0] = 1
        return out
    return torch.exp(logc + k * math.log(p) + (n - k) * math.log1p(-p))


def _chi2_sf(x, dof):
    """Upper tail of the chi-square law (Wilson-Hilferty): good to a few percent, ample for a 1e-6 gate."""
    if dof <= 0:
        return 1.0
    z = ((x / dof) ** (1.0 / 3.0) - (1 - 2.0 / (9 * dof))) / math.sqrt(2.0 / (9 * dof))
    return 0.5 * math.erfc(z / math.sqrt(2))


def _tail_moments(t):
    """Mean, variance and central fourth moment of T ~ N(0,1) | T > t."""
    p = 0.5 * math.erfc(t / math.sqrt(2))
    lam = math.exp(-0.5 * t * t) / math.sqrt(2 * math.pi) / p
    m1, m2, m3, m4 = lam, 1 + t * lam, (t * t + 2) * lam, 3 + (t ** 3 + 3 * t) * lam
    return lam, m2 - m1 * m1, m4 - 4 * m1 * m3 + 6 * m1 * m1 * m2 - 3 * m1 ** 4


@pytest.mark.parametrize("S", [64, 16, 256, 61])
def test_flip_count_and_score_sum_have_the_closed_form_law(S):
    """counts ~ h0 ? S - F : F with F ~ Binomial(S, Phi(-t)) (chi-square on the histogram); rsum | F is a sum of F tail
    normals (mean and variance, pooled over F); E[rsum] = S phi(t) (smoothrast.py:46)."""
    cnt, rs, x = _coverage_run(S, 0, seed=12345 + S)
    P = cnt.shape[0]
    worst = 1.0
    for k in range(x.numel()):
        t = abs(x[k].item()) / SIGMA
        p = 0.5 * math.erfc(t / math.sqrt(2))
        F = cnt[:, k] if x[k] < 0 else S - cnt[:, k]
        assert F.min() >= 0 and F.max() <= S
        pmf = _binom_pmf(S, p)
        obs = torch.bincount(F, minlength=S + 1).double()
        exp = pmf * P
        # merge the bins with a small expectation into one
        big = exp >= 8
        o = torch.cat([obs[big], obs[~big].sum()[None]])
        e = torch.cat([exp[big], exp[~big].sum()[None]])
        keep = e > 1e-3
        chi2 = (((o - e) ** 2)[keep] / e[keep]).sum().item()
        pv = _chi2_sf(chi2, int(keep.sum().item()) - 1)
        worst = min(worst, pv)
        assert pv > 1e-6, (S, t, chi2, int(keep.sum()))
        # the score sum: sign, mean, variance given F
        assert (rs[:, k] >= 0).all()
        assert (rs[:, k][F == 0] == 0).all()
        lam, var1, mu4 = _tail_moments(t)
        nF = F.sum().item()
        if nF >= 50:
            Fd = F.double()
            resid = rs[:, k] - Fd * lam  # given F: mean 0, variance F var1, Var(resid^2) = F mu4 + (2 F^2 - 3 F) var1^2
            zmean = resid.sum().item() / math.sqrt(nF * var1)
            assert abs(zmean) < 5.5, (S, t, zmean)
            v2 = (Fd * mu4 + (2 * Fd * Fd - 3 * Fd) * var1 * var1).clamp(min=0).sum().item()
            zvar = ((resid ** 2).sum().item() - nF * var1) / math.sqrt(v2 + 1e-30)
            assert abs(zvar) < 6.0, (S, t, zvar)
            # every flipped draw is beyond the threshold: rsum >= F t
            assert (rs[:, k] >= F.double() * t * (1 - 1e-5) - 1e-6).all()
        # unconditional: E[rsum] = S phi(t)
        phi = math.exp(-0.5 * t * t) / math.sqrt(2 * math.pi)
        if P * S * p >= 30:  # enough flips for a normal confidence interval
            sd = rs[:, k].std().item() / math.sqrt(P) + 1e-12
            assert abs(rs[:, k].mean().item() - S * phi) < 5.5 * sd + 1e-9, (S, t)
        else:  # rare flips: the total count is Poisson(P S p)
            lam_tot = P * S * p
            assert F.sum().item() <= lam_tot + 6 * math.sqrt(lam_tot) + 6, (S, t, F.sum().item(), lam_tot)
    assert worst > 1e-6


def test_compound_and_per_sample_paths_agree_in_law():
    """Two-sample comparison with the per-sample path of the same kernel (bit-checked against the oracle elsewhere):
    means, variances and the count / score-sum covariance of every entry class."""
    from pertrenderer_b200 import _cabi
    S = 64
    a_c, a_r, x = _coverage_run(S, 0, seed=777)
    b_c, b_r, _ = _coverage_run(S, _cabi.F_PER_SAMPLE_NOISE, seed=778)
    P = a_c.shape[0]
    # the two paths differ exactly where the sampler takes over (not a statement about law: a sanity check that the
    # default really runs the compound sampler and the flag really restores the per-sample path)
    c_c, c_r, _ = _coverage_run(S, _cabi.F_PER_SAMPLE_NOISE, seed=777)
    t_all = x.abs() / SIGMA
    direct = t_all < 0.84
    assert torch.equal(a_c[:, direct], c_c[:, direct]) and torch.equal(a_r[:, direct], c_r[:, direct])
    assert not torch.equal(a_c[:, ~direct], c_c[:, ~direct])
    for k in range(x.numel()):
        for u, v in ((a_c[:, k].double(), b_c[:, k].double()), (a_r[:, k], b_r[:, k]),
                     (a_c[:, k].double() * a_r[:, k], b_c[:, k].double() * b_r[:, k]),
                     (a_r[:, k] ** 2, b_r[:, k] ** 2)):
            se = math.sqrt(u.var().item() / P + v.var().item() / P)
            assert abs(u.mean().item() - v.mean().item()) <= 5.5 * se + 1e-12, (k, x[k].item())


def test_sample_shards_add_up_to_the_whole_law():
    """Noise-sample sharding: each shard draws Binomial(n_r, p) flips from its own counters; the sum over shards has the law
    of the whole job (means / variances against the closed form), and shards do not repeat each other's draws."""
    S = 64
    parts = [_coverage_run(S, 0, seed=4242, s_range=(s0, s0 + 16)) for s0 in range(0, S, 16)]
    x = parts[0][2]
    cnt = sum(p[0] for p in parts)
    rs = sum(p[1] for p in parts)
    P = cnt.shape[0]
    assert not torch.equal(parts[0][0], parts[1][0])
    for k in range(x.numel()):
        t = abs(x[k].item()) / SIGMA
        p = 0.5 * math.erfc(t / math.sqrt(2))
        F = cnt[:, k] if x[k] < 0 else S - cnt[:, k]
        se = math.sqrt(S * p * (1 - p) / P) + 1e-9
        assert abs(F.double().mean().item() - S * p) < 5.5 * se
        var = F.double().var().item()
        assert abs(var - S * p * (1 - p)) < 6 * math.sqrt(2.0 / P) * S * p * (1 - p) + 6 * math.sqrt(S * p / P) + 1e-6
        phi = math.exp(-0.5 * t * t) / math.sqrt(2 * math.pi)
        if P * S * p >= 30:
            assert abs(rs[:, k].mean().item() - S * phi) < 5.5 * rs[:, k].std().item() / math.sqrt(P) + 1e-9
    # correlation between two shards' flips of the same entry is that of independent draws
    k = T_VALUES.index(1.0)
    a, b = parts[0][0][:, k].double(), parts[1][0][:, k].double()
    assert abs(torch.corrcoef(torch.stack((a, b)))[0, 1].item()) < 5.5 / math.sqrt(P)


def test_image_and_gradients_keep_the_law_of_the_per_sample_path():
    """End to end through the fused kernels: mean image, mean gradients and their variances over many seeds, default
    (compound coverage draws, once-per-logit dead noise) against PERT_F_PER_SAMPLE_NOISE."""
    from gpu_util import problem_from_case, run_cuda, synthetic_case
    from pertrenderer_b200 import _cabi
    N, H, W, K, S = 1, 8, 8, 12, 16
    g = synthetic_case(N, H, W, K, S, S, kind="dense", seed=5)
    reps = 400
    runs = {0: {}, 1: {}}
    for m, f in ((0, 0), (1, _cabi.F_PER_SAMPLE_NOISE)):
        acc = {k: [] for k in ("image", "grad_dists", "grad_zbuf", "scalars", "counts", "rsum")}
        for r in range(reps):
            out = run_cuda(problem_from_case(g, explicit=False, seed_rast=31 * r + 7 + 100000 * m, seed_agg=17 * r + 3 + 200000 * m,
                                             flags=f), g["grad_image"])
            for k in acc:
                acc[k].append(out[k].double())
        runs[m] = {k: torch.stack(v) for k, v in acc.items()}
    for k in runs[0]:
        a, b = runs[0][k], runs[1][k]
        se = (a.var(0) / reps + b.var(0) / reps).sqrt()
        z = (a.mean(0) - b.mean(0)).abs() / (se + 1e-3 * se.max() + 1e-30)
        assert z.max().item() < 5.5, (k, z.max().item())
        # variances: z-score with the standard error of a sample variance, sqrt((m4 - var^2) / reps), estimated from the
        # samples themselves (gradients of rarely flipping entries are heavy tailed: a plain ratio test fails between
        # two runs of the SAME mode, tools/diag_law.py)
        va, vb = a.var(0), b.var(0)
        m4a, m4b = ((a - a.mean(0)) ** 4).mean(0), ((b - b.mean(0)) ** 4).mean(0)
        sev = ((m4a - va * va).clamp(min=0) / reps + (m4b - vb * vb).clamp(min=0) / reps).sqrt()
        zv = (va - vb).abs() / (sev + 1e-3 * sev.max() + 1e-30)
        assert zv.max().item() < 6.0, (k, "var", zv.max().item())
