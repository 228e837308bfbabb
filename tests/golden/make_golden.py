"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE (build container only).

Run:  python tests/golden/make_golden.py          (needs /root/reference; CPU, ~1 min)

Imports ``mmidas.nn_model.mixVAE_model`` from /root/reference and executes its forward / loss /
backward / torch.optim.Adam step on seeded synthetic inputs with injected noise (the reference has
no noise hook, so ``sample_gumbel``, ``reparameterize`` and the two ``nn.Dropout`` members are
overridden on the INSTANCE; no reference file is edited or copied).  Outputs go to
``tests/golden/*.npz``; inputs are not stored — tests regenerate them from the same seeds through
``oracle.mixvae_oracle.synth_x / synth_noise``.

/root/reference does not exist on the GPU box, which is why the vectors are committed.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from mmidas.nn_model import mixVAE_model  # noqa: E402  (the reference)
from oracle import mixvae_oracle as O  # noqa: E402  (only for HP + synthetic input generators)

CASES = {
    # name: (HP kwargs, B, n_steps, density, detail)   detail 2 = every tensor, 1 = forward tensors + sampled
    # gradients/params, 0 = samples only
    "tiny": (dict(input_dim=64, n_categories=12, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0), 48, 3, 0.35, 2),
    "a3_hard": (dict(input_dim=96, fc_dim=48, lowD_dim=6, n_categories=7, state_dim=3, n_arm=3, x_drop=0.25, s_drop=0.2, hard=True,
                     lam=2.0, beta=0.5, temp=0.7, tau=0.01), 40, 2, 0.35, 2),
    "mid": (dict(input_dim=520, n_categories=100, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0), 300, 2, 0.35, 1),
    "cfg1": (dict(input_dim=5032, n_categories=92, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0), 1000, 1, 0.35, 0),
}
SEED = 546


class _InjectedDropout(torch.nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = p
        self.queue = []

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        keep = self.queue.pop(0)
        return x * (keep.to(x.dtype) / (1.0 - self.p))


def build_reference(hp: O.HP):
    torch.manual_seed(SEED)
    m = mixVAE_model(input_dim=hp.input_dim, fc_dim=hp.fc_dim, n_categories=hp.n_categories,
                     state_dim=hp.state_dim, lowD_dim=hp.lowD_dim, x_drop=hp.x_drop, s_drop=hp.s_drop,
                     n_arm=hp.n_arm, lam=hp.lam, lam_pc=1, tau=hp.tau, beta=hp.beta, hard=hp.hard,
                     variational=True, device="cpu", eps=hp.eps, momentum=hp.momentum, ref_prior=False,
                     loss_mode="MSE")
    m.x_dp = _InjectedDropout(hp.x_drop)
    m.s_dp = _InjectedDropout(hp.s_drop)
    return m


def inject(m, noise, hp):
    """Queue this step's noise in the reference's draw order (arms 0..A-1)."""
    A = hp.n_arm
    m.x_dp.queue = [noise["keep_x"][a] for a in range(A)]
    m.s_dp.queue = [noise["keep_s"][a] for a in range(A)]
    Us = [noise["U"][a] for a in range(A)]
    Es = [noise["E"][a] for a in range(A)]
    eps = hp.eps
    m.sample_gumbel = lambda shape: -torch.log(-torch.log(Us.pop(0).view(shape) + eps) + eps)
    m.reparameterize = lambda mu, lv: Es.pop(0) * lv.exp().sqrt() + mu


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def sample_idx(numel, k=97):
    g = np.random.default_rng(12345 + numel)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))


def run_case(name, hp_kw, B, n_steps, density, full):
    hp = O.HP(**hp_kw)
    m = build_reference(hp)
    out = {}
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    out["init_sha"] = np.array("".join(sha(sd0[k]) for k in sorted(sd0) if sd0[k].is_floating_point())[:4096])
    out["init_probe"] = np.array([float(sd0[f"fc1.0.weight"][0, 0]), float(sd0[f"fc11.{hp.n_arm-1}.bias"][-1]),
                                  float(sd0["fcc.0.weight"].sum())])
    opt = torch.optim.Adam(m.parameters(), lr=hp.lr)
    gen = torch.Generator().manual_seed(SEED)
    x = O.synth_x(B, hp.input_dim, gen, density)
    names = [n for n, _ in m.named_parameters()]
    for step in range(n_steps):
        noise = O.synth_noise(hp, B, gen)
        inject(m, noise, hp)
        m.train()
        xs = x.expand(hp.n_arm, -1, -1)
        opt.zero_grad()
        x_recs, _, _, x_lows, cs, s_smps, c_smps, s_means, s_logvars, c_probs = m(xs, hp.temp, 0.0)
        tot, rec, joint, ent, dist, l2, kls, _, lls = m.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)
        tot.backward()
        pre = f"s{step}_"
        out[pre + "losses"] = np.array([tot.item(), joint.item(), ent.item(), dist.item(), l2.item()], dtype=np.float64)
        out[pre + "rec"] = rec.numpy().astype(np.float64)
        out[pre + "kl"] = np.array([k.item() for k in kls])
        out[pre + "ll"] = np.array([k.item() for k in lls])
        out[pre + "argmax_qc"] = np.stack([c.argmax(-1).numpy() for c in cs]).astype(np.int16)
        out[pre + "argmax_csmp"] = np.stack([c.argmax(-1).numpy() for c in c_smps]).astype(np.int16)
        grads = dict((n, p.grad.detach().clone()) for n, p in m.named_parameters())
        out[pre + "grad_norm"] = np.array([grads[n].double().norm().item() for n in names])
        if full >= 1:
            for key, lst in (("qc", cs), ("c_smp", c_smps), ("s_mean", s_means), ("s_logvar", s_logvars),
                             ("x_low", x_lows), ("s_smp", s_smps), ("c_prob", c_probs)) + (
                                 (("x_rec", x_recs),) if full == 2 else ()):
                out[pre + key] = torch.stack([t.detach() for t in lst]).numpy()
        if full == 2 and step in (0, n_steps - 1):
            for n in names:
                out[pre + "grad/" + n] = grads[n].numpy()
        if full == 0:
            for key, lst in (("qc", cs), ("s_mean", s_means), ("s_logvar", s_logvars), ("x_low", x_lows)):
                t = torch.stack([t.detach() for t in lst]).reshape(-1)
                idx = sample_idx(t.numel(), 4001)
                out[pre + key + "_samp"] = t[idx].numpy()
        if full < 2:
            for n in names:
                g = grads[n].reshape(-1)
                out[pre + "gsamp/" + n] = g[sample_idx(g.numel())].numpy()
        opt.step()
    sd = m.state_dict()
    out["param_names"] = np.array(names)
    ost = opt.state_dict()["state"]
    out["adam_step"] = np.array(float(ost[0]["step"]))
    if full == 2:
        for k, v in sd.items():
            out["final/" + k] = v.numpy()
        for i, n in enumerate(names):
            out["adam_m/" + n] = ost[i]["exp_avg"].numpy()
            out["adam_v/" + n] = ost[i]["exp_avg_sq"].numpy()
    else:
        for k, v in sd.items():
            t = v.reshape(-1)
            out["fsamp/" + k] = t[sample_idx(t.numel())].numpy() if v.is_floating_point() else v.numpy()
        for i, n in enumerate(names):
            t = ost[i]["exp_avg"].reshape(-1)
            out["adam_m_samp/" + n] = t[sample_idx(t.numel())].numpy()
            t = ost[i]["exp_avg_sq"].reshape(-1)
            out["adam_v_samp/" + n] = t[sample_idx(t.numel())].numpy()
    if full >= 1:
        # eval-mode forward + loss after training (running-stat BN, no Gumbel noise, hard sample)
        noise = O.synth_noise(hp, B, gen)
        inject(m, noise, hp)
        m.eval()
        with torch.no_grad():
            xs = [x for _ in range(hp.n_arm)]
            x_recs, _, _, x_lows, cs, s_smps, c_smps, s_means, s_logvars, c_probs = m(x=xs, temp=hp.temp, prior_c=0.0, eval=True)
            tot, rec, joint, ent, dist, l2, kls, _, lls = m.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)
        out["eval_losses"] = np.array([tot.item(), joint.item(), ent.item(), dist.item(), l2.item()], dtype=np.float64)
        out["eval_rec"] = rec.numpy().astype(np.float64)
        for key, lst in (("qc", cs), ("c_smp", c_smps), ("s_mean", s_means), ("s_logvar", s_logvars)) + (
                (("x_rec", x_recs),) if full == 2 else ()):
            out["eval_" + key] = torch.stack([t.detach() for t in lst]).numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path)/1e3:.0f} kB", "total loss step0", out["s0_losses"][0])


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name, spec in CASES.items():
        run_case(name, *spec)
