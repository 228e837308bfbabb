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


def load_oracle_state(model, opt, st):
    """Teacher forcing: put the oracle's parameters, BN buffers and Adam moments into the CUDA model / FusedAdam."""
    model.load_state_dict(st.sd)
    names = O.param_names(st.hp)
    if st.step > 0:
        state = {i: {"step": torch.tensor(float(st.step)), "exp_avg": st.m[n], "exp_avg_sq": st.v[n]} for i, n in enumerate(names)}
    else:
        state = {}
    opt.load_state_dict({"state": state, "param_groups": opt.state_dict()["param_groups"]})
    opt.step_count = st.step


def shard_model(hp, sd_full, a0, a1, precision, device="cuda"):
    """The arm shard [a0, a1) of a model with hp.n_arm arms, as mmidas_b200.parallel.ShardedTrainer builds it."""
    import dataclasses
    from mmidas_b200 import mixVAE_model
    from mmidas_b200.parallel import slice_arm_state
    m = mixVAE_model(input_dim=hp.input_dim, fc_dim=hp.fc_dim, n_categories=hp.n_categories, state_dim=hp.state_dim,
                     lowD_dim=hp.lowD_dim, x_drop=hp.x_drop, s_drop=hp.s_drop, n_arm=a1 - a0, lam=hp.lam, lam_pc=1,
                     tau=hp.tau, beta=hp.beta, hard=hp.hard, variational=True, device=device, eps=hp.eps,
                     momentum=hp.momentum, ref_prior=False, loss_mode="MSE", precision=precision)
    m.load_state_dict(slice_arm_state(sd_full, a0, a1))
    m = m.to(device)
    m.n_arm_total = hp.n_arm
    m.arm_offset = a0
    return m
