"""CUDA augmenter forward (mmidas_b200.Augmenter_smartseq, through the C ABI: mvae_fold_affine / mvae_linear_act /
mvae_fma_rows) against the outputs of the unmodified reference class (tests/golden/aug_*.npz) and the fp64 oracle.

Tolerances (relative L2 per output tensor, against the fp64 oracle):
  precision "tf32x3" (default, error-compensated): 2e-5 on the golden cases, 1e-4 at the production widths (the split
                     truncates: each product carries ~2^-21 relative error, and 14 layers up to 5032 wide compound it;
                     the reference's own fp32 result sits at ~1e-6)
  precision "tf32"   (single pass):                 1e-2
"""
import numpy as np
import pytest
import torch

from oracle import augmenter_oracle as AO
from test_augmenter_oracle import GOLD, load_case, rel_l2

TOL = {"tf32x3": 2e-5, "tf32": 1e-2}


def _module(sd, c, precision):
    from mmidas_b200 import Augmenter_smartseq
    net = Augmenter_smartseq(noise_dim=c["noise_dim"], latent_dim=c["latent_dim"], input_dim=c["input_dim"], n_dim=c["n_dim"],
                             precision=precision)
    net.load_state_dict(sd)            # strict: the reference's keys and shapes
    return net


def test_state_dict_keys_are_the_reference_ones():
    g, sd, c = load_case(GOLD[0])
    net = _module(sd, c, "tf32x3")
    assert list(net.state_dict().keys()) == list(sd.keys())
    with pytest.raises((RuntimeError, NotImplementedError)):
        net.eval()(torch.zeros(4, c["input_dim"]), False)          # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("path", GOLD, ids=[p.split("aug_")[-1][:-4] for p in GOLD])
def test_matches_reference_goldens(path, precision):
    g, sd, c = load_case(path)
    net = _module(sd, c, precision).cuda().eval()
    x = torch.from_numpy(g["x"]).cuda()
    noise = {"z": torch.from_numpy(g["z"]).cuda(), "eps": torch.from_numpy(g["eps"]).cuda()}
    xin = x.expand(c["A"], -1, -1) if c["A"] else x
    s, xa = net(xin, bool(c["A"]), c["scale"], noise=noise)
    torch.cuda.synchronize()
    assert s.shape == g["s"].shape and xa.shape == g["x_aug"].shape
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    s64, xa64 = AO.forward(sd64, torch.from_numpy(g["x"]).expand(c["A"], -1, -1) if c["A"] else torch.from_numpy(g["x"]),
                           torch.from_numpy(g["z"]), torch.from_numpy(g["eps"]), c["scale"])
    tol = TOL[precision]
    assert rel_l2(s.cpu().numpy(), s64.numpy()) <= tol, rel_l2(s.cpu().numpy(), s64.numpy())
    assert rel_l2(xa.cpu().numpy(), xa64.numpy()) <= tol, rel_l2(xa.cpu().numpy(), xa64.numpy())
    assert rel_l2(xa.cpu().numpy(), g["x_aug"]) <= 2 * tol
    assert float(xa.min()) >= 0.0                                   # relu(fc11(.)), udagan.py:329
    if c["A"] > 1:
        # arms given as separate copies of x take the per-(arm, cell) encoder path: same result as the shared-encoder path
        xs = x.unsqueeze(0).repeat(c["A"], 1, 1)
        s2, xa2 = net(xs, True, c["scale"], noise=noise)
        assert rel_l2(xa2.cpu().numpy(), xa.cpu().numpy()) <= 1e-6 and rel_l2(s2.cpu().numpy(), s.cpu().numpy()) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("tf32x3", 1e-4), ("tf32", 1e-2)])
def test_full_size_layers_against_fp64_oracle(precision, tol):
    """The production shapes (5032 genes -> 1006 -> 1006 -> 500 -> 500 | 50 -> 100 -> 10 -> ... -> 5032) on 300
    cells x 2 arms: exercises the 16-byte pitch padding of the 1006- and 550-wide layers."""
    sd = AO.random_state_dict(50, 10, 5032, 500, seed=3)
    c = dict(noise_dim=50, latent_dim=10, input_dim=5032, n_dim=500)
    net = _module(sd, c, precision).cuda().eval()      # "tf32" runs the wide layers on 256-column tiles
    g = torch.Generator().manual_seed(5)
    x = AO_synth(300, 5032, g)
    z, eps = torch.randn(2, 300, 50, generator=g), torch.randn(2, 300, 10, generator=g)
    s, xa = net(x.cuda().expand(2, -1, -1), True, 0.1, noise={"z": z.cuda(), "eps": eps.cuda()})
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    s64, xa64 = AO.forward(sd64, x.expand(2, -1, -1), z, eps, 0.1)
    assert rel_l2(s.cpu().numpy(), s64.numpy()) <= tol, rel_l2(s.cpu().numpy(), s64.numpy())
    assert rel_l2(xa.cpu().numpy(), xa64.numpy()) <= tol, rel_l2(xa.cpu().numpy(), xa64.numpy())
    # without injected noise the draws differ between calls and between arms
    s_a, xa_a = net(x.cuda().expand(2, -1, -1), True, 0.1)
    s_b, _ = net(x.cuda().expand(2, -1, -1), True, 0.1)
    assert not torch.equal(s_a, s_b) and not torch.equal(xa_a[0], xa_a[1])


def AO_synth(B, D, g):
    return torch.where(torch.rand(B, D, generator=g) < 0.35, torch.log1p(torch.exp(3.5 + 1.5 * torch.randn(B, D, generator=g))),
                       torch.zeros(()))
