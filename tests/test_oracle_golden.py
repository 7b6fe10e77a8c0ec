"""Pin the CPU oracle (oracle/pert_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), and against the closed forms of SURVEY.md Appendix A.4."""

import pytest
import torch

from conftest import load_golden, rel_err
from oracle import pert_oracle as O


def _run(g):
    return O.shade_fwd_bwd(g["pix_to_face"], g["zbuf"], g["dists"], g["colors"], g["background"],
                           g["znear_t"], g["zfar_t"], g["sigma"], g["gamma"], g["alpha"], g["eps"],
                           g["U"], g["V"], g["grad_image"])


def test_forward_bit_exact(shade_case):
    g = shade_case
    st, _ = _run(g)
    assert torch.equal(st.counts, g["counts"])
    assert torch.equal(st.prob, g["prob"])
    assert torch.equal(st.zeta, g["zeta"])
    assert torch.equal(st.a_s, g["a_s"])
    assert torch.equal(st.a_0, g["a_0"])
    assert torch.equal(st.weights, g["weights"])
    assert torch.equal(st.image, g["image"])


def test_backward_matches_reference_autograd(shade_case):
    g = shade_case
    _, gr = _run(g)
    assert rel_err(gr["colors"], g["grad_colors"]) <= 1e-6
    # grad_zbuf contains -sum_j grad_zeta_j (a cancelling sum of zero-mean noise terms), so its
    # rounding depends on summation order: north_star's 1e-5 is the bar there
    assert rel_err(gr["zbuf"], g["grad_zbuf"]) <= 1e-5
    assert rel_err(gr["dists"], g["grad_dists"]) <= 1e-6
    for k in ("sigma", "gamma", "alpha"):
        ref = g["grad_" + k]
        assert abs(gr[k].item() - ref) <= 2e-5 * max(abs(ref), 1e-12) + 1e-9, k


def test_reference_quirks(shade_case):
    """Appendix B: padded entries get zero gradients; empty pixels render the background with
    alpha 0; weights sum to one."""
    g = shade_case
    st, gr = _run(g)
    pad = g["pix_to_face"] < 0
    assert (gr["dists"][pad] == 0).all() and (gr["zbuf"][pad] == 0).all()
    empty = pad.all(dim=-1)
    if empty.any():
        assert torch.equal(st.image[empty][:, :3], g["background"].expand(int(empty.sum()), 3))
        assert (st.image[empty][:, 3] == 0).all()
    assert torch.allclose(st.weights.sum(-1), torch.ones_like(st.weights[..., 0]), atol=1e-6)


def test_standalone_ops():
    g = load_golden("ops_small")
    prob, h, h0 = O.random_heaviside_fwd(g["x"], g["U"], g["sigma"])
    assert torch.equal(prob, g["prob"])
    gx, gs = O.random_heaviside_bwd(g["grad_l"], h, h0, g["U"], g["sigma"])
    assert rel_err(gx, g["grad_x"]) <= 1e-6
    assert abs(gs.item() - g["grad_sigma"]) <= 1e-5 * abs(g["grad_sigma"])
    w, a_s, a_0 = O.random_argmax_fwd(g["z"], g["V"], g["gamma"])
    assert torch.equal(w, g["weights"])
    gz, gg = O.random_argmax_bwd(g["grad_w"], a_s, a_0, g["V"], g["gamma"])
    assert rel_err(gz, g["grad_z"]) <= 1e-6
    assert abs(gg.item() - g["grad_gamma"]) <= 1e-5 * abs(g["grad_gamma"])


def test_standalone_ops_without_control_variates():
    """randomHeaviside_wovr / randomArgmax_wovr (smoothrast.py:61-108, smoothagg.py:75-141): the reference's own
    outputs with recorded noise."""
    g = load_golden("ops_wovr")
    prob, h, h0 = O.random_heaviside_fwd(g["x"], g["U"], g["sigma"])
    assert torch.equal(prob, g["prob"])
    gx, gs = O.random_heaviside_bwd(g["grad_l"], h, h0, g["U"], g["sigma"], control_variate=False)
    assert rel_err(gx, g["grad_x"]) <= 1e-6
    assert abs(gs.item() - g["grad_sigma"]) <= 1e-5 * abs(g["grad_sigma"])
    # the control variate matters: the far-inside entry has a non-zero score sum without it
    gx_vr, _ = O.random_heaviside_bwd(g["grad_l"], h, h0, g["U"], g["sigma"])
    assert (gx_vr[..., -1] == 0).all() and (gx[..., -1] != 0).any()
    w, a_s, a_0 = O.random_argmax_fwd(g["z"], g["V"], g["gamma"])
    assert torch.equal(w, g["weights"])
    gz, gg = O.random_argmax_bwd(g["grad_w"], a_s, a_0, g["V"], g["gamma"], control_variate=False)
    assert rel_err(gz, g["grad_z"]) <= 1e-6
    assert abs(gg.item() - g["grad_gamma"]) <= 1e-5 * abs(g["grad_gamma"])


def test_closed_forms_monte_carlo():
    """Appendix A.4 with many samples: E[p_hat] = Phi(x/sigma), E[(h-h0)U]/sigma = phi(x/sigma)/sigma,
    two-way argmax weight = Phi(dzeta / (gamma sqrt 2))."""
    gen = torch.Generator().manual_seed(0)
    sigma, S = 1e-3, 20000
    d = torch.linspace(-2.5e-3, 2.5e-3, 11).reshape(1, 1, 11, 1)
    U = torch.normal(torch.zeros((S, 1, 1, 11, 1)), 1.0, generator=gen)
    prob, h, h0 = O.random_heaviside_fwd(-d, U, sigma)
    expect = O.expected_coverage(d, sigma)
    se = (expect * (1 - expect) / S).sqrt() + 1e-9
    assert ((prob.double() - expect).abs() <= 5 * se).all()
    gx, _ = O.random_heaviside_bwd(torch.ones_like(d), h, h0, U, sigma)
    expect_g = O.expected_coverage_score(d, sigma)
    assert ((gx.double() - expect_g).abs() <= 5 * (1.0 / sigma) / S ** 0.5).all()
    gamma = 1e-2
    zeta = torch.tensor([0.0, -0.7e-2]).reshape(1, 1, 1, 2)
    V = torch.normal(torch.zeros((S, 1, 1, 1, 2)), 1.0, generator=gen)
    w, _, _ = O.random_argmax_fwd(zeta, V, gamma)
    e = O.expected_two_way_weight(0.0, -0.7e-2, gamma).item()
    assert abs(w[0, 0, 0, 0].item() - e) <= 5 * (e * (1 - e) / S) ** 0.5


@pytest.mark.parametrize("case", ["small", "k50", "empty"])
def test_soft_oracle_matches_reference_golden(case):
    """SoftRast + SoftAgg (the shaders' default operators): the oracle's forward and hand-written backward
    against vectors produced by the unmodified reference (tests/golden/make_golden.py, soft_*.npz)."""
    g = load_golden("soft_" + case)
    zn, zf = g["znear"].reshape(-1, 1, 1, 1), g["zfar"].reshape(-1, 1, 1, 1)
    image, prob, weights, gr = O.soft_shade_fwd_bwd(g["pix_to_face"], g["zbuf"], g["dists"], g["colors"], g["background"], zn, zf,
                                                    g["sigma"], g["gamma"], g["alpha"], g["eps"], g["grad_image"])
    assert torch.equal(prob, g["prob"])
    assert (weights - g["weights"]).abs().max() <= 1e-6
    assert (image - g["image"]).abs().max() <= 1e-6
    assert rel_err(gr["colors"], g["grad_colors"]) <= 1e-6
    assert rel_err(gr["dists"], g["grad_dists"]) <= 1e-5
    assert rel_err(gr["zbuf"], g["grad_zbuf"]) <= 1e-5
    for k in ("sigma", "gamma", "alpha"):
        ref = float(g["grad_" + k])
        assert abs(gr[k].item() - ref) <= 2e-5 * max(abs(ref), 1e-12) + 1e-7, (k, gr[k].item(), ref)
