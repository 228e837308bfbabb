"""The augmenter oracle (oracle/augmenter_oracle.py) against outputs of the unmodified reference class
(tests/golden/aug_*.npz, written by tests/golden/make_golden_aug.py); CPU only.  The CUDA path is compared with both in
tests/test_gpu_augmenter.py."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import augmenter_oracle as AO

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "aug_*.npz")))


def load_case(path):
    g = np.load(path)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    nz, nl, D, nd, A, B = (int(v) for v in g["cfg"])
    return g, sd, dict(noise_dim=nz, latent_dim=nl, input_dim=D, n_dim=nd, A=A, B=B, scale=float(g["scale"]))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_goldens_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[4:-4] for p in GOLD])
def test_oracle_matches_reference_outputs(path):
    g, sd, c = load_case(path)
    x = torch.from_numpy(g["x"])
    xin = x.expand(c["A"], -1, -1) if c["A"] else x
    s, xa = AO.forward(sd, xin, torch.from_numpy(g["z"]), torch.from_numpy(g["eps"]), c["scale"])
    assert s.shape == g["s"].shape and xa.shape == g["x_aug"].shape
    assert rel_l2(s.numpy(), g["s"]) < 2e-6
    assert rel_l2(xa.numpy(), g["x_aug"]) < 2e-6
    # the ReLU pattern of the output is the reference's except where the pre-activation is rounding noise
    flips = np.mean((xa.numpy() > 0) != (g["x_aug"] > 0))
    assert flips < 1e-4
    # fp64 evaluation of the same graph: the yardstick the GPU tolerances are stated against
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    s64, xa64 = AO.forward(sd64, xin, torch.from_numpy(g["z"]), torch.from_numpy(g["eps"]), c["scale"])
    assert rel_l2(g["x_aug"], xa64.numpy()) < 5e-6


def test_random_state_dict_has_reference_shapes():
    sd = AO.random_state_dict(50, 10, 5032, 500, 0)
    assert sd["fc1.weight"].shape == (1006, 5032) and sd["fc5.weight"].shape == (100, 550)
    assert sd["fc11.weight"].shape == (5032, 1006) and sd["bnz.weight"].shape == (50,)
    assert sum(v.numel() for k, v in sd.items() if k.endswith(("weight", "bias"))) == 13776332
