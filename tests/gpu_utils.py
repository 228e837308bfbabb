"""Helpers for the -m gpu parity tests: build the CUDA model from the oracle's initial state,
run oracle and CUDA on the same seeded inputs."""
import numpy as np
import torch

from oracle import mixvae_oracle as O


def build_model(hp: O.HP, precision="fp32_simt", seed=546, device="cuda"):
    from mmidas_b200 import mixVAE_model
    m = mixVAE_model(input_dim=hp.input_dim, fc_dim=hp.fc_dim, n_categories=hp.n_categories, state_dim=hp.state_dim,
                     lowD_dim=hp.lowD_dim, x_drop=hp.x_drop, s_drop=hp.s_drop, n_arm=hp.n_arm, lam=hp.lam, lam_pc=1,
                     tau=hp.tau, beta=hp.beta, hard=hp.hard, variational=True, device=device, eps=hp.eps,
                     momentum=hp.momentum, ref_prior=False, loss_mode="MSE", precision=precision)
    m.load_state_dict(O.init_state_dict(hp, seed))
    return m.to(device)


def to_dev_noise(noise, device="cuda"):
    return {k: v.to(device) for k, v in noise.items()}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / n if n > 0 else np.linalg.norm(a - b)


def oracle_step(hp, sd, x, noise, dtype):
    st = O.TrainState(hp, O.cast_state_dict(sd, dtype))
    out = O.train_step(st, [x.to(dtype)] * hp.n_arm, noise, return_grads=True)
    return st, out


def loss_vector(ls):
    return np.array([float(ls["total"]), float(ls["joint"]), float(ls["ent"]), float(ls["dist"]), float(ls["l2"])])


def cuda_grads(model):
    return {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
