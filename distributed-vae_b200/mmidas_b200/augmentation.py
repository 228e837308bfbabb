"""B200-native forward of the pre-trained VAE-GAN augmenter that precedes every training step when ``aug_file`` is set
(SURVEY §8 f1).  Mirrors ``mmidas/augmentation/udagan.py:217-329`` (``Augmenter_smartseq``: same constructor, same
``state_dict`` keys and shapes, so reference checkpoints load) and ``mk_augmenter`` (``mmidas/cpl_mixvae.py:128-149``).

The training loop only ever runs the augmenter in eval mode (``netA.to(device).eval()``, cpl_mixvae.py:184), where each
``relu(batch_fcN(fcN(x)))`` is a Linear followed by a per-feature affine and an activation.  Here every layer is ONE
``mvae_linear_act`` call: a TMA-fed tcgen05 GEMM (error-compensated 3xTF32 by default, fp32-accurate) with the folded
BatchNorm/bias affine and the activation in its epilogue.  Two structural savings over the reference:

* ``x.expand(A, -1, -1)`` (cpl_mixvae.py:423) gives every arm the same cells and eval mode has no dropout, so fc1..fc4 are
  evaluated once per cell, not once per (arm, cell); the arms differ from the noise concatenation (fc5) onwards;
* nothing is recorded for autograd: the reference back-propagates through the augmenter for nothing (its ``no_grad`` is
  commented out, cpl_mixvae.py:421) while no optimiser holds the augmenter's parameters.

torch is used for storage and for drawing the two standard-normal noise tensors (as the reference does with
``torch.randn``); they can be injected (``noise={"z": [A,B,noise_dim], "eps": [A,B,latent_dim]}``) for parity tests.
There is no CPU path: forward raises on a CPU-resident module or in training mode.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Mapping, Optional

import torch
import torch.nn as nn

from . import _lib

ACT_NONE, ACT_RELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class Augmenter_smartseq(nn.Module):
    """Same members, in the same construction order, as the reference (udagan.py:218-283): the seeded default
    initialisation is therefore identical, and ``load_state_dict`` takes reference checkpoints (``aug_model["netA"]``)."""

    def __init__(self, noise_dim, latent_dim, input_dim=5000, n_dim=500, p_drop=0.5, precision="tf32x3"):
        super().__init__()
        moment = 0.01
        self.noise_dim, self.latent_dim, self.input_dim, self.n_dim = noise_dim, latent_dim, input_dim, n_dim
        if precision not in ("tf32x3", "tf32"):
            raise ValueError("precision must be 'tf32x3' (fp32-accurate, default) or 'tf32'")
        self.precision = precision
        self.dp = nn.Dropout(p_drop)
        self.noise = nn.Linear(noise_dim, noise_dim, bias=False)
        self.bnz = nn.BatchNorm1d(self.noise.out_features)

        def bn(n):
            return nn.BatchNorm1d(num_features=n, eps=1e-10, momentum=moment, affine=False)

        self.fc1 = nn.Linear(input_dim, input_dim // 5)
        self.batch_fc1 = bn(self.fc1.out_features)
        self.fc2 = nn.Linear(self.fc1.out_features, self.fc1.out_features)
        self.batch_fc2 = bn(self.fc2.out_features)
        self.fc3 = nn.Linear(self.fc2.out_features, n_dim)
        self.batch_fc3 = bn(self.fc3.out_features)
        self.fc4 = nn.Linear(n_dim, n_dim)
        self.batch_fc4 = bn(self.fc4.out_features)
        self.fc5 = nn.Linear(n_dim + noise_dim, n_dim // 5)
        self.batch_fc5 = bn(self.fc5.out_features)
        self.fc_mu = nn.Linear(self.fc5.out_features, latent_dim)
        self.fc_sigma = nn.Linear(self.fc5.out_features, latent_dim)
        self.batch_fc_mu = bn(self.fc_mu.out_features)
        self.fc6 = nn.Linear(self.fc_mu.out_features, n_dim // 5)
        self.batch_fc6 = bn(self.fc6.out_features)
        self.fc7 = nn.Linear(self.fc6.out_features, n_dim)
        self.batch_fc7 = bn(self.fc7.out_features)
        self.fc8 = nn.Linear(n_dim, n_dim)
        self.batch_fc8 = bn(self.fc8.out_features)
        self.fc9 = nn.Linear(n_dim, input_dim // 5)
        self.batch_fc9 = bn(self.fc9.out_features)
        self.fc10 = nn.Linear(self.fc9.out_features, self.fc9.out_features)
        self.batch_fc10 = bn(self.fc10.out_features)
        self.fc11 = nn.Linear(self.fc10.out_features, input_dim)
        self._plan = None          # folded epilogues + padded weights, rebuilt when a parameter or buffer changes
        self._plan_key = None
        self._bufs = {}
        self.launches = 0          # library launches issued by this module (bench bookkeeping)

    # (linear, batch norm or None, activation) of every layer, udagan.py:285-329
    _LAYERS = (("noise", "bnz", ACT_ELU), ("fc1", "batch_fc1", ACT_RELU), ("fc2", "batch_fc2", ACT_RELU),
               ("fc3", "batch_fc3", ACT_RELU), ("fc4", "batch_fc4", ACT_RELU), ("fc5", "batch_fc5", ACT_RELU),
               ("fc_mu", "batch_fc_mu", ACT_NONE), ("fc_sigma", None, ACT_SIGMOID), ("fc6", "batch_fc6", ACT_RELU),
               ("fc7", "batch_fc7", ACT_RELU), ("fc8", "batch_fc8", ACT_RELU), ("fc9", "batch_fc9", ACT_RELU),
               ("fc10", "batch_fc10", ACT_RELU), ("fc11", None, ACT_RELU))

    # ------------------------------------------------------------------------------------------
    def _stream(self, dev):
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def _build_plan(self, dev):
        lib = _lib.load()
        plan = {}
        for lin_name, bn_name, act in self._LAYERS:
            lin = getattr(self, lin_name)
            bnm = getattr(self, bn_name) if bn_name else None
            n_out, k = lin.out_features, lin.in_features
            w = lin.weight.detach()
            if k % 4 != 0 or w.data_ptr() % 16 != 0 or not w.is_contiguous():
                wp = torch.zeros(n_out, _pad4(k), dtype=torch.float32, device=dev)      # TMA needs 16-byte row pitches
                wp[:, :k].copy_(w)
                w = wp
            scale = torch.empty(n_out, dtype=torch.float32, device=dev)
            shift = torch.empty(n_out, dtype=torch.float32, device=dev)
            ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            bias = lin.bias.detach() if lin.bias is not None else None
            mean = bnm.running_mean if bnm is not None else None
            var = bnm.running_var if bnm is not None else None
            gamma = bnm.weight.detach() if bnm is not None and bnm.affine else None
            beta = bnm.bias.detach() if bnm is not None and bnm.affine else None
            _lib.check(lib.mvae_fold_affine(ptr(bias), ptr(mean), ptr(var), ptr(gamma), ptr(beta),
                                            C.c_float(bnm.eps if bnm is not None else 0.0), n_out, ptr(scale), ptr(shift),
                                            self._stream(dev)), "mvae_fold_affine")
            self.launches += 1
            plan[lin_name] = (w, w.shape[1], scale, shift, act, n_out, k)
        return plan

    def _get_plan(self, dev):
        key = tuple(t._version for t in list(self.parameters()) + list(self.buffers())) + (str(dev),)
        if self._plan is None or key != self._plan_key:
            self._plan, self._plan_key = self._build_plan(dev), key
        return self._plan

    def _buf(self, name, rows, cols, dev):
        """Zero-initialised [rows, pad4(cols)] activation buffer, cached per shape (pad columns stay zero)."""
        k = (name, rows, cols, str(dev))
        b = self._bufs.get(k)
        if b is None:
            b = torch.zeros(rows, _pad4(cols), dtype=torch.float32, device=dev)
            self._bufs[k] = b
        return b

    def _linear(self, name, x2d, y2d):
        """y2d[:, :n_out] = act(affine(x2d[:, :k] @ W^T)) through the library (x2d / y2d: 2-D views with unit column stride)."""
        w, w_pitch, scale, shift, act, n_out, k = self._plan[name]
        rows = x2d.shape[0]
        assert x2d.stride(1) == 1 and y2d.stride(1) == 1 and x2d.stride(0) % 4 == 0 and x2d.data_ptr() % 16 == 0
        _lib.check(_lib.load().mvae_linear_act(
            C.c_void_p(x2d.data_ptr()), x2d.stride(0), C.c_void_p(w.data_ptr()), w_pitch, C.c_void_p(y2d.data_ptr()),
            y2d.stride(0), rows, n_out, k, C.c_void_p(scale.data_ptr()), C.c_void_p(shift.data_ptr()), act,
            1 if self.precision == "tf32x3" else 0, self._stream(x2d.device)), f"mvae_linear_act({name})")
        self.launches += 1

    def _fma(self, a, b, c, out, n, a_scale=1.0):
        rows = a.shape[0]
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        s = lambda t: t.stride(0) if t is not None else 0
        _lib.check(_lib.load().mvae_fma_rows(p(a), s(a), p(b), s(b), p(c), s(c), p(out), s(out), rows, n, C.c_float(a_scale),
                                             self._stream(a.device)), "mvae_fma_rows")
        self.launches += 1

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, batched, scale=1.0, noise: Optional[Mapping[str, torch.Tensor]] = None):
        """``(s, x_aug)`` like the reference: batched ``x`` [A, B, D] -> s [A, B, latent], x_aug [A, B, D];
        otherwise ``x`` [B, D] -> s [B, latent], x_aug [B, D]."""
        if self.training:
            raise NotImplementedError("the B200 augmenter implements the eval-mode forward the training loop uses "
                                      "(cpl_mixvae.py:184); augmenter training is outside the hot path")
        if not x.is_cuda:
            raise RuntimeError("Augmenter_smartseq (B200) needs CUDA tensors; there is no CPU path")
        dev = x.device
        self._get_plan(dev)
        if batched:
            A, B, D = x.shape
            shared = A == 1 or x.stride(0) == 0          # every arm sees the same cells (x.expand, cpl_mixvae.py:423)
        else:
            (B, D), A, shared = x.shape, 1, True
        assert D == self.input_dim
        rows = A * B
        x2 = (x[0] if batched else x) if shared else x.reshape(rows, D)
        if x2.stride(1) != 1 or x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0 or x2.dtype != torch.float32:
            xp = self._buf("x", x2.shape[0], D, dev)
            xp[:, :D].copy_(x2)
            x2 = xp
        enc_rows = x2.shape[0]
        nz, nl, nd, F1 = self.noise_dim, self.latent_dim, self.n_dim, self.fc1.out_features
        z_raw = noise["z"] if noise is not None else torch.randn(A, B, nz, device=dev)
        eps = noise["eps"] if noise is not None else torch.randn(A, B, nl, device=dev)
        z_raw = z_raw.to(device=dev, dtype=torch.float32).reshape(rows, nz).contiguous()
        eps = eps.to(device=dev, dtype=torch.float32).reshape(rows, nl).contiguous()

        cat = self._buf("cat", rows, nd + nz, dev)                     # [x4 | z], udagan.py:299
        zin = self._buf("zin", rows, nz, dev)
        self._fma(z_raw, None, None, zin, nz, a_scale=float(scale))    # scale * randn, :287-293
        self._linear("noise", zin, cat[:, nd:])                        # elu(bnz(noise(z))), :294
        h1 = self._buf("h1", enc_rows, F1, dev)
        h2 = self._buf("h2", enc_rows, F1, dev)
        h3 = self._buf("h3", enc_rows, nd, dev)
        self._linear("fc1", x2, h1)                                    # :295-297 (dropout is the identity in eval mode)
        self._linear("fc2", h1, h2)
        self._linear("fc3", h2, h3)
        if shared and A > 1:
            h4 = self._buf("h4", enc_rows, nd, dev)
            self._linear("fc4", h3, h4)
            for a in range(A):                                         # the same encoding under every arm's noise
                self._fma(h4, None, None, cat[a * B:(a + 1) * B], nd)
        else:
            self._linear("fc4", h3, cat)
        h5 = self._buf("h5", rows, self.fc5.out_features, dev)
        mu = self._buf("mu", rows, nl, dev)
        sg = self._buf("sigma", rows, nl, dev)
        s = self._buf("s", rows, nl, dev)
        self._linear("fc5", cat, h5)
        self._linear("fc_mu", h5, mu)                                  # batch_fc_mu(fc_mu(x)), :302
        self._linear("fc_sigma", h5, sg)                               # sigmoid(fc_sigma(x)), :303
        self._fma(eps, sg, mu, s, nl)                                  # reparam_trick, aug_utils.py:51-65
        h6 = self._buf("h6", rows, self.fc6.out_features, dev)
        h7 = self._buf("h7", rows, nd, dev)
        h8 = self._buf("h8", rows, nd, dev)
        h9 = self._buf("h9", rows, F1, dev)
        h10 = self._buf("h10", rows, F1, dev)
        self._linear("fc6", s, h6)
        self._linear("fc7", h6, h7)
        self._linear("fc8", h7, h8)
        self._linear("fc9", h8, h9)
        self._linear("fc10", h9, h10)
        out = torch.empty(rows, D, dtype=torch.float32, device=dev)
        self._linear("fc11", h10, out)                                 # relu(fc11(x)), :329
        s_out = s[:, :nl].clone()
        if batched:
            return s_out.reshape(A, B, nl), out.reshape(A, B, D)
        return s_out, out


def mk_augmenter(pretrained: str, load_weights: bool, precision: str = "tf32x3") -> tuple[Mapping[Any, Any], Mapping[Any, Any], nn.Module]:
    """mmidas/cpl_mixvae.py:128-149: checkpoint dict with ``parameters`` (num_n, num_z, n_features) and ``netA``.
    ``precision`` (added): "tf32x3" (fp32-accurate, default) or "tf32" (single pass, ~1.7x faster)."""
    aug_model = torch.load(pretrained, map_location="cpu")
    aug_param = aug_model["parameters"]
    if not load_weights:
        raise NotImplementedError("load_weights=False builds the legacy `Augmenter` class in the reference "
                                  "(cpl_mixvae.py:142-149), which the training path does not use")
    print("loading augmenter weights")
    netA = Augmenter_smartseq(noise_dim=aug_param["num_n"], latent_dim=aug_param["num_z"],
                              input_dim=aug_param["n_features"], precision=precision)
    netA.load_state_dict(aug_model["netA"])
    return aug_model, aug_param, netA
