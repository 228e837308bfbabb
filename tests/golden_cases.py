"""Shared description of the golden cases (must mirror tests/golden/make_golden.py CASES)."""
import os

import numpy as np
import torch

from oracle import mixvae_oracle as O

SEED = 546
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "tiny": (dict(input_dim=64, n_categories=12, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0), 48, 3, 0.35, 2),
    "a3_hard": (dict(input_dim=96, fc_dim=48, lowD_dim=6, n_categories=7, state_dim=3, n_arm=3, x_drop=0.25,
                     s_drop=0.2, hard=True, lam=2.0, beta=0.5, temp=0.7, tau=0.01), 40, 2, 0.35, 2),
    "mid": (dict(input_dim=520, n_categories=100, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0), 300, 2, 0.35, 1),
    "cfg1": (dict(input_dim=5032, n_categories=92, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0), 1000, 1, 0.35, 0),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sample_idx(numel, k=97):
    g = np.random.default_rng(12345 + numel)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))


def case_inputs(name):
    """Regenerate the inputs of a golden case: hp, x, [noise per step], eval noise."""
    kw, B, n_steps, density, detail = CASES[name]
    hp = O.HP(**kw)
    gen = torch.Generator().manual_seed(SEED)
    x = O.synth_x(B, hp.input_dim, gen, density)
    noises = [O.synth_noise(hp, B, gen) for _ in range(n_steps)]
    eval_noise = O.synth_noise(hp, B, gen)
    return hp, x, noises, eval_noise, detail


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d
