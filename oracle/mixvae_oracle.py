"""CPU oracle for the coupled mixture-VAE (cpl-mixVAE / MMIDAS) training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product (``distributed-vae_b200/``) never imports it and has
no CPU path.

What this is: a functional restatement (plain tensors in dicts, no ``nn.Module``) of the
arithmetic of the reference's hot path, on the CPU, in fp32 or fp64:

* forward             mmidas/nn_model.py:263-287 (layers), :297-368 (forward),
                      :413-493 (reparameterisation, Gumbel-softmax)
* loss                mmidas/nn_model.py:39-86 (helpers), :495-598 (loss)
* step order          mmidas/cpl_mixvae.py:434-463 (zero_grad, forward, loss, backward, Adam)
* optimiser           torch.optim.Adam defaults as constructed at mmidas/cpl_mixvae.py:274
* parameter init      nn.Linear default init in the construction order of
                      mmidas/nn_model.py:184-208 (layer-major, arm-minor)

The arithmetic itself lives in a third-party dependency of the reference, PyTorch
(pinned by the reference only in dist/environment_312.yml:225 as pytorch=2.4.0; this
image has 2.11.0).  The oracle calls the same published primitives (linear, batch-norm
statistics, softmax, log, sigmoid) on the CPU and differentiates with autograd.

Pinning: the reference's own tests hold no golden vector for this path (SURVEY.md §4,
"parity unpinned" by the reference).  The oracle is therefore pinned against OUTPUTS OF
THE REFERENCE ITSELF, produced in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference/mmidas/nn_model.py`` with injected noise) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the oracle against them.

Noise is always injected (the reference has no hook; the golden script overrides
``sample_gumbel`` / ``reparameterize`` / the two ``nn.Dropout`` on the instance):
    U   [A,B,C] uniform [0,1)   Gumbel uniforms      (nn_model.py:440)
    E   [A,B,S] uniform [0,1)   state noise, uniform, NOT normal (nn_model.py:427)
    keep_x [A,B,D] bool         input-dropout keep mask (nn_model.py:264)
    keep_s [A,B,S] bool         state-dropout keep mask (nn_model.py:278)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

LAYERS = ("fc1", "fc2", "fc3", "fc4", "fc5", "fcc", "fc_mu", "fc_sigma",
          "fc6", "fc7", "fc8", "fc9", "fc10", "fc11")
BN_LAYERS = ("batch_l1", "batch_l2", "batch_l3", "batch_l4", "batch_l5", "batch_s")


@dataclass
class HP:
    """Hyper-parameters of mixVAE_model (nn_model.py:112-134); defaults = train.py:174-266."""
    input_dim: int = 5032
    fc_dim: int = 100
    n_categories: int = 100
    state_dim: int = 2
    lowD_dim: int = 10
    x_drop: float = 0.5
    s_drop: float = 0.0
    n_arm: int = 2
    lam: float = 1.0
    tau: float = 0.005
    beta: float = 1.0
    hard: bool = False
    eps: float = 1e-8
    momentum: float = 0.01
    temp: float = 1.0
    lr: float = 1e-3
    betas: tuple = (0.9, 0.999)
    adam_eps: float = 1e-8


def layer_shapes(hp: HP) -> Dict[str, tuple]:
    """(out_features, in_features) per layer — nn_model.py:184-208."""
    D, H, L, C, S = hp.input_dim, hp.fc_dim, hp.lowD_dim, hp.n_categories, hp.state_dim
    return {
        "fc1": (H, D), "fc2": (H, H), "fc3": (H, H), "fc4": (H, H), "fc5": (L, H),
        "fcc": (C, L), "fc_mu": (S, L + C), "fc_sigma": (S, L + C),
        "fc6": (L, S + C), "fc7": (H, L), "fc8": (H, H), "fc9": (H, H), "fc10": (H, H),
        "fc11": (D, H),
    }


def bn_features(hp: HP) -> Dict[str, int]:
    H, L, S = hp.fc_dim, hp.lowD_dim, hp.state_dim
    return {"batch_l1": H, "batch_l2": H, "batch_l3": H, "batch_l4": H, "batch_l5": L, "batch_s": S}


def init_state_dict(hp: HP, seed: int = 546, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reproduce the reference's random init bit-for-bit: ``torch.manual_seed(seed)`` then
    nn.Linear's default init (weight U(-1/sqrt(in), 1/sqrt(in)) via kaiming_uniform(a=sqrt 5),
    then bias U(-1/sqrt(in), 1/sqrt(in))), drawn layer-major / arm-minor exactly in the
    construction order of nn_model.py:184-208.  Keys follow the reference state_dict."""
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, (o, i) in layer_shapes(hp).items():
        for a in range(hp.n_arm):
            w = torch.empty(o, i)
            # kaiming_uniform_(a=sqrt(5)): gain = sqrt(2/(1+5)), bound = gain*sqrt(3/fan_in)
            gain = math.sqrt(2.0 / (1.0 + 5.0))
            bound_w = math.sqrt(3.0) * gain / math.sqrt(i)
            w.uniform_(-bound_w, bound_w)
            b = torch.empty(o)
            bound_b = 1.0 / math.sqrt(i)
            b.uniform_(-bound_b, bound_b)
            sd[f"{name}.{a}.weight"] = w.to(dtype)
            sd[f"{name}.{a}.bias"] = b.to(dtype)
    for name, n in bn_features(hp).items():
        for a in range(hp.n_arm):
            sd[f"{name}.{a}.running_mean"] = torch.zeros(n, dtype=dtype)
            sd[f"{name}.{a}.running_var"] = torch.ones(n, dtype=dtype)
            sd[f"{name}.{a}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def param_names(hp: HP) -> List[str]:
    """model.parameters() order of the reference: layer-major, arm-minor."""
    out = []
    for name in LAYERS:
        for a in range(hp.n_arm):
            out += [f"{name}.{a}.weight", f"{name}.{a}.bias"]
    return out


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synth_x(B: int, D: int, gen: torch.Generator, density: float = 0.35) -> torch.Tensor:
    """Smart-seq-shaped (density .35) / 10x-shaped (.08) log1p-CPM-like non-negative matrix."""
    u = torch.rand(B, D, generator=gen)
    v = torch.log1p(torch.exp(3.5 + 1.5 * torch.randn(B, D, generator=gen)))
    return torch.where(u < density, v, torch.zeros(())).to(torch.float32)


def synth_noise(hp: HP, B: int, gen: torch.Generator) -> Dict[str, torch.Tensor]:
    """Noise in the reference's draw order per arm: dropout mask, U, E."""
    A, D, C, S = hp.n_arm, hp.input_dim, hp.n_categories, hp.state_dim
    keep_x, U, E, keep_s = [], [], [], []
    for _ in range(A):
        keep_x.append(torch.rand(B, D, generator=gen) >= hp.x_drop)
        U.append(torch.rand(B, C, generator=gen))
        E.append(torch.rand(B, S, generator=gen))
        keep_s.append(torch.rand(B, S, generator=gen) >= hp.s_drop)
    return {"keep_x": torch.stack(keep_x), "U": torch.stack(U), "E": torch.stack(E),
            "keep_s": torch.stack(keep_s)}


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def _bn(h, sd, key, hp: HP, train: bool, new_buffers: Optional[dict]):
    """BatchNorm1d(affine=False, eps=hp.eps, momentum=hp.momentum) placed AFTER the ReLU
    (nn_model.py:264).  Training: biased batch variance normalises, unbiased variance goes
    into running_var; eval: running statistics."""
    rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
    if train:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(h, rm2, rv2, None, None, True, hp.momentum, hp.eps)
        if new_buffers is not None:
            new_buffers[key + ".running_mean"] = rm2
            new_buffers[key + ".running_var"] = rv2
            new_buffers[key + ".num_batches_tracked"] = sd[key + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(h, rm, rv, None, None, False, hp.momentum, hp.eps)


def _dropout(x, keep, p):
    if keep is None or p == 0.0:
        return x
    scale = (keep.to(x.dtype) / (1.0 - p))
    return x * scale


def forward(sd: Dict[str, torch.Tensor], xs: Sequence[torch.Tensor], noise: Dict[str, torch.Tensor],
            hp: HP, train: bool = True, new_buffers: Optional[dict] = None, mask=None) -> Dict[str, List[torch.Tensor]]:
    """mixVAE_model.forward (nn_model.py:297-368).  ``train=False`` is the reference's
    ``eval=True`` on a module in ``.eval()`` mode: running-stat BN, no dropout, no Gumbel noise,
    straight-through one-hot sample; the state noise E is still applied (nn_model.py:351).
    ``mask`` (indices of kept categories): the pruning path of nn_model.py:332-335."""
    out = {k: [] for k in ("x_rec", "x_low", "qc", "s_smp", "c_smp", "s_mean", "s_logvar", "c_prob",
                           "h_dec")}
    eps = hp.eps
    for a in range(hp.n_arm):
        W = lambda n: sd[f"{n}.{a}.weight"]
        b = lambda n: sd[f"{n}.{a}.bias"]
        x = xs[a]
        dt = x.dtype
        h = _dropout(x, noise["keep_x"][a] if train else None, hp.x_drop)
        for i in (1, 2, 3, 4, 5):
            h = _bn(F.relu(F.linear(h, W(f"fc{i}"), b(f"fc{i}"))), sd, f"batch_l{i}.{a}", hp, train, new_buffers)
        x_low = h
        c_prob = F.softmax(F.linear(x_low, W("fcc"), b("fcc")), dim=-1)          # :269
        if mask is not None:                                                     # :332-335
            idx = torch.as_tensor(mask, dtype=torch.long)
            qc = torch.zeros_like(c_prob).index_copy(1, idx, F.softmax(c_prob[:, idx] / hp.tau, dim=-1))
        else:
            qc = F.softmax(c_prob / hp.tau, dim=-1)                              # :337
        if train:
            U = noise["U"][a].to(dt)
            g = -torch.log(-torch.log(U + eps) + eps)                            # :440-441
            y = F.softmax(((qc + eps).log() + g) / hp.temp, dim=-1)              # :454-455
            hard = hp.hard
        else:
            y = qc
            hard = True
        if hard:                                                                 # :486-493
            ind = y.argmax(dim=-1, keepdim=True)
            y_hard = torch.zeros_like(y).scatter_(1, ind, 1.0)
            c_smp = (y_hard - y).detach() + y
        else:
            c_smp = y
        yy = torch.cat((x_low, c_smp), dim=1)                                    # :347
        s_mean = F.linear(yy, W("fc_mu"), b("fc_mu"))
        s_var = torch.sigmoid(F.linear(yy, W("fc_sigma"), b("fc_sigma")))
        s_logvar = (s_var + eps).log()                                           # :350
        s_smp = noise["E"][a].to(dt) * s_logvar.exp().sqrt() + s_mean            # :426-428
        s_in = _dropout(s_smp, noise["keep_s"][a] if train else None, hp.s_drop)
        z = torch.cat((c_smp, s_in), dim=1)                                      # :279
        h = z
        for i in (6, 7, 8, 9, 10):
            h = F.relu(F.linear(h, W(f"fc{i}"), b(f"fc{i}")))
        x_rec = F.relu(F.linear(h, W("fc11"), b("fc11")))                        # :287
        for k, v in (("x_rec", x_rec), ("x_low", x_low), ("qc", qc), ("s_smp", s_smp), ("c_smp", c_smp),
                     ("s_mean", s_mean), ("s_logvar", s_logvar), ("c_prob", c_prob), ("h_dec", h)):
            out[k].append(v)
    return out


# ----------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------
def loss(fw: Dict[str, List[torch.Tensor]], xs: Sequence[torch.Tensor], hp: HP) -> Dict[str, object]:
    """mixVAE_model.loss (nn_model.py:495-598), MSE mode, no reference prior."""
    A, C, eps = hp.n_arm, hp.n_categories, hp.eps
    B = xs[0].shape[0]
    rec, kls, lls, inds = [], [], [], []
    ents, l2s, dists = [], [], []
    logq = [torch.log(q + eps) for q in fw["qc"]]
    # inv_var (nn_model.py:75-77): 1/sqrt(unbiased batch variance + eps), per category
    w = [(1.0 / (q.var(0) + eps)).sqrt() for q in fw["qc"]]
    for a in range(A):
        x, xr = xs[a], fw["x_rec"][a]
        sse = ((xr - x) ** 2).sum()
        lls.append(sse / x.numel() + B * math.log(2 * math.pi))                     # :542
        mism = ((xr > 0.1) != (x > 0.1)).to(x.dtype).sum()
        # BCE on {0,1} inputs with PyTorch's log clamp at -100: 100 per mismatching element
        bce = 100.0 * mism / x.numel()
        rec.append(0.5 * sse / B + 0.5 * bce)                                       # :544-546
        mu, lv = fw["s_mean"][a], fw["s_logvar"][a]
        kls.append((-0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp(), dim=0)).sum())  # :43-44
        inds.append(rec[-1] + hp.beta * kls[-1])
        for b in range(a + 1, A):
            ents.append((fw["qc"][a] * logq[a]).sum(-1).mean() + (fw["qc"][b] * logq[b]).sum(-1).mean())
            l2s.append(((fw["c_smp"][a] - fw["c_smp"][b]) ** 2).sum(-1).mean())
            dists.append(((logq[a] * w[a] - logq[b] * w[b]) ** 2).sum(-1).mean())
    n_pairs = max(A * (A - 1) / 2, 1)
    joint = hp.lam * sum(dists) + sum(ents) + n_pairs * ((C / 2) * math.log(2 * math.pi) - 0.5 * math.log(2 * hp.lam))
    total = max(A - 1, 1) * sum(inds) + joint
    return {"total": total, "rec": torch.stack([r.detach() for r in rec]), "joint": joint,
            "ent": sum(ents) / len(ents), "dist": sum(dists) / len(dists), "l2": sum(l2s) / len(l2s),
            "kl": kls, "ll": lls}


# ----------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults; single-tensor formulation)
# ----------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, step: int, lr, betas, eps):
    b1, b2 = betas
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


@dataclass
class TrainState:
    hp: HP
    sd: Dict[str, torch.Tensor]
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)
    step: int = 0


def train_step(st: TrainState, xs: Sequence[torch.Tensor], noise: Dict[str, torch.Tensor],
               return_grads: bool = False, mask=None) -> Dict[str, object]:
    """One optimiser step in the reference's order (cpl_mixvae.py:434-463)."""
    hp = st.hp
    names = param_names(hp)
    leaves = {}
    sd = dict(st.sd)
    for n in names:
        leaves[n] = st.sd[n].detach().clone().requires_grad_(True)
        sd[n] = leaves[n]
    new_buffers: dict = {}
    fw = forward(sd, xs, noise, hp, train=True, new_buffers=new_buffers, mask=mask)
    ls = loss(fw, xs, hp)
    grads = torch.autograd.grad(ls["total"], [leaves[n] for n in names])
    st.step += 1
    with torch.no_grad():
        for n, g in zip(names, grads):
            if n not in st.m:
                st.m[n] = torch.zeros_like(st.sd[n])
                st.v[n] = torch.zeros_like(st.sd[n])
            adam_update(st.sd[n], g, st.m[n], st.v[n], st.step, hp.lr, hp.betas, hp.adam_eps)
        st.sd.update(new_buffers)
    out = {"loss": {k: (v.detach() if torch.is_tensor(v) else [t.detach() for t in v]) for k, v in ls.items()},
           "fw": {k: [t.detach() for t in v] for k, v in fw.items()}}
    if return_grads:
        out["grads"] = {n: g for n, g in zip(names, grads)}
    return out


def cast_state_dict(sd, dtype):
    """Always a deep copy: train_step updates the state in place."""
    return {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
