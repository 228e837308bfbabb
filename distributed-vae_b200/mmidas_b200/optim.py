"""Fused Adam over the flat parameter buffer of ``mixVAE_model`` (replaces ``torch.optim.Adam`` as
constructed at mmidas/cpl_mixvae.py:274 and re-constructed at train.py:144-147).

One kernel over ``[n_arm * arm_stride]`` floats instead of a multi-tensor foreach over 28*A tensors.
``state_dict()`` / ``load_state_dict()`` speak the layout of ``torch.optim.Adam`` (state[i] =
{step, exp_avg, exp_avg_sq}, i in ``model.parameters()`` order: layer-major, arm-minor), so optimizer
checkpoints written by the reference load here and vice versa.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import PARAM_ORDER


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, model=None,
                 decoupled_weight_decay=False):
        if model is None:
            raise ValueError("FusedAdam needs model= (the mixVAE_model whose flat buffer it updates)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=decoupled_weight_decay)
        super().__init__(params, defaults)
        self.model = model
        plist = [p for g in self.param_groups for p in g["params"]]
        mlist = list(model.parameters())
        if len(plist) != len(mlist) or any(a is not b for a, b in zip(plist, mlist)):
            raise ValueError("FusedAdam must be given model.parameters() of the bound model, in order")
        self.step_count = 0
        self._graph_counters = None      # set while nn_model.StepGraph captures: Adam's step index lives on the device
        self._m = None
        self._v = None

    def flat_state(self):
        flat = self.model.flat_parameters()
        if self._m is None or self._m.shape != flat.shape or self._m.device != flat.device:
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
        return self._m, self._v

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        model = self.model
        flat = model.flat_parameters()
        if flat.device.type != "cuda":
            raise RuntimeError("FusedAdam has no CPU path")
        m, v = self.flat_state()
        g = self.param_groups[0]
        self.step_count += 1
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        _lib.check(_lib.load().mvae_adam(flat.data_ptr(), model.flat_grads().data_ptr(), m.data_ptr(), v.data_ptr(),
                                         flat.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                         float(g["eps"]), float(g["weight_decay"]),
                                         int(bool(g.get("decoupled_weight_decay", False))), self.step_count,
                                         self._graph_counters[1:].data_ptr() if self._graph_counters is not None else None,
                                         C.c_void_p(stream)), "mvae_adam")
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # gradients are written (never accumulated) by the backward kernels: nothing to clear.
        if set_to_none:
            for p in self.model.parameters():
                p.grad = None

    # ---- torch.optim.Adam-compatible (de)serialisation ------------------------------------------
    def _slots(self):
        """flat (arm, offset, numel, shape) for every parameter in model.parameters() order."""
        lay = self.model._layout
        out = []
        for li, name in enumerate(PARAM_ORDER):
            ml = getattr(self.model, name)
            for a in range(self.model.n_arm):
                out.append((a, lay.offset[2 * li], lay.numel[2 * li], ml[a].weight.shape))
                out.append((a, lay.offset[2 * li + 1], lay.numel[2 * li + 1], ml[a].bias.shape))
        return out

    def state_dict(self):
        m, v = self.flat_state()
        state = {}
        if self.step_count > 0:
            for i, (a, off, n, shape) in enumerate(self._slots()):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": m[a, off:off + n].view(shape).clone(),
                            "exp_avg_sq": v[a, off:off + n].view(shape).clone()}
        groups = []
        start = 0
        for g in self.param_groups:
            d = {k: val for k, val in g.items() if k != "params"}
            d["params"] = list(range(start, start + len(g["params"])))
            start += len(g["params"])
            groups.append(d)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state_dict):
        m, v = self.flat_state()
        slots = self._slots()
        st = state_dict["state"]
        m.zero_()
        v.zero_()
        step = 0
        for i, (a, off, n, shape) in enumerate(slots):
            s = st.get(i, st.get(str(i)))
            if s is None:
                continue
            m[a, off:off + n].view(shape).copy_(s["exp_avg"])
            v[a, off:off + n].view(shape).copy_(s["exp_avg_sq"])
            step = int(float(s["step"]))
        self.step_count = step
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            for k, val in sg.items():
                if k != "params":
                    g[k] = val
