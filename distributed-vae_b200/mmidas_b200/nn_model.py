"""Drop-in mirror of ``mmidas/nn_model.py`` (reference: AllenInstitute/distributed-vae) whose
arithmetic runs in hand-written sm_100a CUDA kernels (``libmixvae_b200.so``).

Kept from the reference: the ``mixVAE_model`` constructor signature (nn_model.py:112-134), the
``state_dict`` keys/shapes (``fc1.{a}.weight`` ... ``batch_s.{a}.num_batches_tracked``), the 10-tuple
returned by ``forward`` (:368), the 9-tuple returned by ``loss`` (:588-598), ``VAEConfig`` (:14) and
``mk_vae`` (:679).  Not kept: the arithmetic.  Parameters are views into one flat fp32 buffer
``[n_arm, arm_stride]`` so that the fused kernels and the fused Adam see contiguous memory; the
gradient buffer has the same layout and ``p.grad`` are views into it.

There is no CPU path: a model can be constructed and (de)serialised on the CPU (host logic), but
``forward`` raises unless the parameters live on an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
from torch import nn
from torch.nn import ModuleList as mdl

from . import _lib
from ._lib import BN_ORDER, PARAM_ORDER


@dataclass
class VAEConfig:
    """Same fields and defaults as the reference dataclass (nn_model.py:14-36)."""
    n_categories: int = 92
    state_dim: int = 2
    input_dim: int = 5032
    fc_dim: int = 100
    lowD_dim: int = 10
    x_drop: float = 0.5
    s_drop: float = 0.2
    lr: float = 0.001
    lam: float = 1
    lam_pc: float = 1
    n_arm: int = 2
    temp: float = 1.0
    tau: float = 0.005
    beta: float = 1.0
    hard: bool = False
    variational: bool = True
    ref_prior: bool = False
    trained_model: Optional[str] = None
    n_pr: int = 0
    momentum: float = 0.01
    mode: str = "MSE"


class _StepContext:
    """Everything one forward leaves behind for loss()/backward(): C structs + the tensors that keep
    the borrowed device pointers alive."""

    def __init__(self):
        self.gen = 0
        self.training = False
        self.keep = []          # tensors referenced by raw pointers
        self.dims = None
        self.hp = None
        self.state = None
        self.inputs = None
        self.outputs = None
        self.out_tensors = {}
        self.loss_vec = None
        self.loss_done = False


class _LossFn(torch.autograd.Function):
    """Autograd anchor: makes ``total.backward()`` run the backward kernels (cpl_mixvae.py:462)."""

    @staticmethod
    def forward(ctx, anchor, total, model, gen):
        ctx.model = model
        ctx.gen = gen
        return total.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.model._run_backward(ctx.gen, grad_out)
        return None, None, None, None


class mixVAE_model(nn.Module):
    """Coupled mixture-VAE with ``n_arm`` arms; constructor arguments as in the reference."""

    def __init__(self, input_dim, fc_dim, n_categories, state_dim, lowD_dim, x_drop, s_drop, n_arm, lam,
                 lam_pc, tau, beta, hard, variational, device, eps, momentum, ref_prior, loss_mode,
                 norm="batch", precision: str = "tf32x3_fc1"):
        super().__init__()
        if loss_mode != "MSE":
            raise NotImplementedError("only loss_mode='MSE' is live in the reference (nn_model.py:315 asserts not ZINB)")
        if not variational:
            raise NotImplementedError("non-variational mode is disabled in the reference (nn_model.py:316)")
        if ref_prior:
            raise NotImplementedError("ref_prior is disabled in the reference (nn_model.py:578)")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        self.input_dim = input_dim
        self.fc_dim = fc_dim
        self.lowD_dim = lowD_dim
        self.state_dim = state_dim
        self.n_categories = n_categories
        self.x_drop = float(x_drop)
        self.s_drop = float(s_drop)
        self.x_dp = nn.Dropout(x_drop)      # kept for attribute parity; dropout runs inside the fc1 kernels
        self.s_dp = nn.Dropout(s_drop)
        self.hard = hard
        self.n_arm = n_arm
        self.lam = lam
        self.lam_pc = lam_pc
        self.tau = tau
        self.beta = beta
        self.varitional = variational       # (sic) attribute name of the reference, nn_model.py:174
        self.eps = eps
        self.ref_prior = ref_prior
        self.momentum = momentum
        self.device = device
        self.loss_mode = loss_mode
        self.precision = precision
        # arm sharding (set by mmidas_b200.parallel): this rank owns arms [arm_offset, arm_offset+n_arm)
        self.n_arm_total = n_arm
        self.arm_offset = 0
        self.seed_salt = 0      # xor-ed into the noise seed; data-parallel replicas get distinct salts
        # forward() returns the materialised reconstruction by default (reference behaviour);
        # the fused trainer switches it off: x_hat then never touches HBM.
        self.materialize_recon = True

        D, H, L, Cc, S = input_dim, fc_dim, lowD_dim, n_categories, state_dim
        shapes = {"fc1": (D, H), "fc2": (H, H), "fc3": (H, H), "fc4": (H, H), "fc5": (H, L), "fcc": (L, Cc),
                  "fc_mu": (L + Cc, S), "fc_sigma": (L + Cc, S), "fc6": (S + Cc, L), "fc7": (L, H), "fc8": (H, H),
                  "fc9": (H, H), "fc10": (H, H), "fc11": (H, D)}
        # nn.Linear default init, constructed in the reference's order (nn_model.py:184-208) so that the
        # same torch seed yields bit-identical initial weights.
        for name in PARAM_ORDER:
            i, o = shapes[name]
            setattr(self, name, mdl([nn.Linear(i, o) for _ in range(n_arm)]))
        bnf = {"batch_l1": H, "batch_l2": H, "batch_l3": H, "batch_l4": H, "batch_l5": L, "batch_s": S}
        for name in BN_ORDER:
            setattr(self, name, mdl([nn.BatchNorm1d(num_features=bnf[name], eps=eps, momentum=momentum, affine=False)
                                     for _ in range(n_arm)]))

        self._dims0 = _lib.Dims(n_arm, 2, D, H, L, Cc, S, n_arm, 0)
        self._layout = _lib.compute_layout(self._dims0)     # validates the shapes (raises on unsupported)
        self._flat_params = None
        self._flat_grads = None
        self._flat_bn = None
        self._flat_nbt = None
        self._work = None
        self._ctx = _StepContext()
        self._gen = 0
        self._step_counter = 0
        self._grad_anchor = None
        self._flat_alloc = None
        self._graph_counters = None
        self._flatten()

    # ------------------------------------------------------------------------------------------
    # flat storage
    # ------------------------------------------------------------------------------------------
    def _named_slots(self):
        """(tensor index t, arm a, parameter) in flat-layout order."""
        for li, name in enumerate(PARAM_ORDER):
            ml = getattr(self, name)
            for a in range(self.n_arm):
                yield 2 * li, a, ml[a].weight
                yield 2 * li + 1, a, ml[a].bias

    def _flatten(self):
        """(Re)allocate the flat buffers on the parameters' current device and re-point every
        parameter / BN buffer at its view.  Parameter objects keep their identity."""
        lay = self._layout
        p0 = self.fc1[0].weight
        dev, A = p0.device, self.n_arm
        # (mmidas_b200.parallel swaps in a peer-accessible allocator for the two buffers replicas exchange)
        alloc = self._flat_alloc or (lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev))
        flat = alloc(A, lay.arm_stride)
        grads = alloc(A, lay.arm_stride)
        with torch.no_grad():
            for t, a, p in self._named_slots():
                off, n = lay.offset[t], lay.numel[t]
                view = flat[a, off:off + n].view(p.shape)
                view.copy_(p.data.to(dev, torch.float32))
                p.data = view
                p.grad = None
        bn = torch.zeros(A, lay.bn_stride, dtype=torch.float32, device=dev)
        nbt = torch.zeros(A, 6, dtype=torch.int64, device=dev)
        with torch.no_grad():
            for bi, name in enumerate(BN_ORDER):
                ml = getattr(self, name)
                for a in range(A):
                    m = ml[a]
                    n = m.num_features
                    off = lay.bn_offset[bi]
                    rm = bn[a, off:off + n]
                    rv = bn[a, off + n:off + 2 * n]
                    rm.copy_(m.running_mean.to(dev, torch.float32))
                    rv.copy_(m.running_var.to(dev, torch.float32))
                    t = nbt[a, bi]
                    t.copy_(m.num_batches_tracked.to(dev))
                    m._buffers["running_mean"] = rm
                    m._buffers["running_var"] = rv
                    m._buffers["num_batches_tracked"] = t
        self._flat_params, self._flat_grads, self._flat_bn, self._flat_nbt = flat, grads, bn, nbt
        self._work = None
        self._grad_anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._ctx = _StepContext()

    def _apply(self, fn, recurse=True):
        super()._apply(fn)
        self._flatten()
        p0 = self.fc1[0].weight
        if p0.dtype != torch.float32:
            raise TypeError("mixVAE_model (B200) stores fp32 parameters like the reference")
        return self

    def ctor_kwargs(self) -> dict:
        """The constructor arguments of this model (mmidas_b200.parallel rebuilds arm shards from them)."""
        return dict(input_dim=self.input_dim, fc_dim=self.fc_dim, n_categories=self.n_categories, state_dim=self.state_dim,
                    lowD_dim=self.lowD_dim, x_drop=self.x_drop, s_drop=self.s_drop, n_arm=self.n_arm, lam=self.lam,
                    lam_pc=self.lam_pc, tau=self.tau, beta=self.beta, hard=self.hard, variational=self.varitional,
                    device=self.device, eps=self.eps, momentum=self.momentum, ref_prior=self.ref_prior,
                    loss_mode=self.loss_mode, precision=self.precision)

    def flat_parameters(self) -> torch.Tensor:
        """[n_arm, arm_stride] fp32: every parameter of every local arm (padding is zero)."""
        return self._flat_params

    def flat_grads(self) -> torch.Tensor:
        return self._flat_grads

    def bind_grads(self):
        """Point every ``p.grad`` at its view of the flat gradient buffer."""
        lay = self._layout
        for t, a, p in self._named_slots():
            if p.grad is None or p.grad.data_ptr() != self._flat_grads[a, lay.offset[t]].data_ptr():
                off, n = lay.offset[t], lay.numel[t]
                p.grad = self._flat_grads[a, off:off + n].view(p.shape)

    # ------------------------------------------------------------------------------------------
    # C-ABI plumbing
    # ------------------------------------------------------------------------------------------
    def _require_cuda(self):
        dev = self._flat_params.device
        if dev.type != "cuda":
            raise RuntimeError("mixVAE_model (B200) has no CPU path: move the model to an sm_100 GPU "
                               "(the reference's CPU implementation is /root/reference/mmidas/nn_model.py)")
        return dev

    def _hparams(self, temp) -> _lib.HParams:
        return _lib.HParams(float(self.tau), float(temp), float(self.beta), float(self.lam), float(self.eps),
                            float(self.momentum), self.x_drop, self.s_drop, int(bool(self.hard)),
                            _lib.PRECISIONS[self.precision])

    def _dims(self, B) -> _lib.Dims:
        return _lib.Dims(self.n_arm, int(B), self.input_dim, self.fc_dim, self.lowD_dim, self.n_categories,
                         self.state_dim, self.n_arm_total, self.arm_offset)

    def _workspace(self, dims: _lib.Dims):
        # one workspace per (batch, arm placement), kept for the life of the model: captured CUDA graphs hold raw
        # pointers into it, so a workspace is never freed or resized behind their back
        key = (dims.batch, dims.n_arm_total, dims.arm_offset)
        if self._work is None:
            self._work = {}
        w = self._work.get(key)
        if w is None:
            lay = _lib.compute_layout(dims)
            w = self._work[key] = torch.zeros(lay.work_floats, dtype=torch.float32, device=self._flat_params.device)
        return w

    def _state(self, dims, adam_m=None, adam_v=None) -> _lib.State:
        work = self._workspace(dims)
        return _lib.State(self._flat_params.data_ptr(), self._flat_grads.data_ptr(),
                          adam_m.data_ptr() if adam_m is not None else None,
                          adam_v.data_ptr() if adam_v is not None else None,
                          self._flat_bn.data_ptr(), self._flat_nbt.data_ptr(), work.data_ptr())

    def _inputs(self, xt, x_arm_stride, x_row_stride, U, E, keep_x, keep_s, training, counters=None, cat_mask=None) -> _lib.Inputs:
        """Per-step inputs.  In-kernel noise (dropout, Gumbel, state) is keyed on (torch.initial_seed() ^ seed_salt, step
        counter, GLOBAL arm index): ``seed_salt`` separates data-parallel replicas (mmidas_b200.parallel), the arm index
        separates arms wherever they live.  ``counters`` (device int64[2]) moves the step counter onto the device."""
        ptr = lambda t: t.data_ptr() if t is not None else None
        seed = (torch.initial_seed() ^ self.seed_salt) & 0xFFFFFFFFFFFFFFFF
        if counters is None:
            counters = self._graph_counters       # set while a StepGraph captures: the step index lives on the device
        return _lib.Inputs(xt.data_ptr(), x_arm_stride, x_row_stride, ptr(U), ptr(E), ptr(keep_x), ptr(keep_s), ptr(cat_mask),
                           C.c_uint64(seed), self._step_counter, ptr(counters), int(training))

    def _prep_x(self, x):
        """Accept the reference's input forms: a list of A [B,D] tensors, or an [A,B,D] tensor
        (``x.expand(A,-1,-1)`` at cpl_mixvae.py:425 has stride 0 over arms -> read x once per arm
        from the same memory)."""
        A = self.n_arm
        dev = self._flat_params.device
        if torch.is_tensor(x):
            if x.dim() != 3 or x.size(0) != A:
                raise ValueError(f"x must be [n_arm={A}, B, D]")
            if x.stride(0) == 0:
                base = x[0]
            else:
                base = None
                xs = x
        else:
            if len(x) != A:
                raise ValueError(f"len(x)={len(x)} != n_arm={A}")   # reference: assert len(x) == self.n_arm
            if all(xi is x[0] or (xi.data_ptr() == x[0].data_ptr() and xi.shape == x[0].shape) for xi in x):
                base = x[0]
            else:
                base = None
                xs = torch.stack(list(x))
        if base is not None:
            base = base.detach()
            if base.device != dev or base.dtype != torch.float32 or base.stride(-1) != 1:
                base = base.to(dev, torch.float32).contiguous()
            if base.size(-1) != self.input_dim:
                raise ValueError("x has the wrong number of genes")
            return base, 0, base.stride(0), base.size(0)
        xs = xs.detach().to(dev, torch.float32).contiguous()
        if xs.size(-1) != self.input_dim:
            raise ValueError("x has the wrong number of genes")
        return xs, xs.stride(0), xs.stride(1), xs.size(1)

    def _prep_noise(self, noise, B, training):
        A, Cc, S, D = self.n_arm, self.n_categories, self.state_dim, self.input_dim
        dev = self._flat_params.device
        noise = noise or {}

        def get(key, shape, dtype, make):
            t = noise.get(key)
            if t is None:
                return make()
            t = t.to(dev)
            if dtype == torch.uint8:
                t = t.to(torch.uint8)
            else:
                t = t.to(dtype)
            t = t.reshape(shape).contiguous()
            return t
        # draw order per arm in the reference: dropout mask, Gumbel uniforms, state noise;
        # here the streams are independent device draws (the reference has no noise-injection hook,
        # parity tests always inject).
        # U / E not injected: None -> the library draws them in-kernel (counter-based, keyed by torch.initial_seed()
        # and the step counter; the backward regenerates the same E)
        U = get("U", (A, B, Cc), torch.float32, lambda: None) if training else None
        E = get("E", (A, B, S), torch.float32, lambda: None)
        keep_x = None
        if training and self.x_drop > 0 and noise.get("keep_x") is not None:
            keep_x = get("keep_x", (A, B, D), torch.uint8, None)
        keep_s = None
        if training and self.s_drop > 0:
            keep_s = get("keep_s", (A, B, S), torch.uint8,
                         lambda: (torch.rand(A, B, S, device=dev) >= self.s_drop).to(torch.uint8))
        return U, E, keep_x, keep_s

    def _category_mask(self, mask):
        """forward(mask=...) (nn_model.py:332-335): indices of the categories that are kept -> uint8 [C] on the device."""
        if mask is None:
            return None
        idx = torch.as_tensor(mask, dtype=torch.long, device="cpu").reshape(-1)
        if idx.numel() == 0 or int(idx.min()) < 0 or int(idx.max()) >= self.n_categories:
            raise ValueError("mask must hold category indices in [0, n_categories)")
        keep = torch.zeros(self.n_categories, dtype=torch.uint8)
        keep[idx] = 1
        return keep.to(self._flat_params.device)

    def _launch_forward(self, x, temp, eval, noise, materialize, mask=None):
        dev = self._require_cuda()
        lib = _lib.load()
        training = bool(self.training and not eval)
        if self.training == bool(eval):
            raise NotImplementedError(
                "forward(eval=...) must agree with the module mode: model.train() with eval=False, or "
                "model.eval() with eval=True (the only combinations the reference trainer uses)")
        xt, x_arm_stride, x_row_stride, B = self._prep_x(x)
        A, Cc, S, L, D = self.n_arm, self.n_categories, self.state_dim, self.lowD_dim, self.input_dim
        U, E, keep_x, keep_s = self._prep_noise(noise, B, training)
        dims = self._dims(B)
        hp = self._hparams(temp)
        st = self._state(dims)
        self._step_counter += 1
        cat_mask = self._category_mask(mask)
        inp = self._inputs(xt, x_arm_stride, x_row_stride, U, E, keep_x, keep_s, training, cat_mask=cat_mask)
        mk = lambda n: torch.empty(A, B, n, dtype=torch.float32, device=dev)
        ot = {"x_low": mk(L), "c_prob": mk(Cc), "qc": mk(Cc), "c_smp": mk(Cc), "s_mean": mk(S), "s_logvar": mk(S),
              "s_smp": mk(S), "x_rec": mk(D) if materialize else None}
        out = _lib.Outputs(*[ot[k].data_ptr() if ot[k] is not None else None
                             for k in ("x_low", "c_prob", "qc", "c_smp", "s_mean", "s_logvar", "s_smp", "x_rec")])
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mvae_forward(C.byref(dims), C.byref(hp), C.byref(st), C.byref(inp), C.byref(out),
                                    C.c_void_p(stream)), "mvae_forward")
        ctx = _StepContext()
        self._gen += 1
        ctx.gen = self._gen
        ctx.training = training
        ctx.keep = [xt, U, E, keep_x, keep_s, cat_mask]
        ctx.dims, ctx.hp, ctx.state, ctx.inputs, ctx.outputs, ctx.out_tensors = dims, hp, st, inp, out, ot
        self._ctx = ctx
        return ctx

    # ------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------
    def forward(self, x, temp, prior_c=[], eval=False, mask=None, noise=None):
        """mixVAE_model.forward (nn_model.py:297-368).  Returns
        ``(x_recs, [], [], x_lows, cs, s_smps, c_smps, s_means, s_logvars, c_probs)``, lists over arms.
        ``mask``: indices of the categories kept by the pruning path (:332-335, eval_model passes the non-zero
        entries of ``fcc[0].bias``): q is a softmax over those only and exactly 0 elsewhere.
        ``noise`` (not in the reference) injects {"U","E","keep_x","keep_s"} for parity tests."""
        ctx = self._launch_forward(x, temp, eval, noise, self.materialize_recon or eval, mask=mask)
        ot = ctx.out_tensors
        A = self.n_arm
        split = lambda t: [t[a] for a in range(A)]
        x_recs = split(ot["x_rec"]) if ot["x_rec"] is not None else [None] * A
        return (x_recs, [], [], split(ot["x_low"]), split(ot["qc"]), split(ot["s_smp"]), split(ot["c_smp"]),
                split(ot["s_mean"]), split(ot["s_logvar"]), split(ot["c_prob"]))

    def loss(self, recon_x, p_x, r_x, x, mu, log_sigma, qc, c, prior_c=[], qc_all=None, c_smp_all=None):
        """mixVAE_model.loss (nn_model.py:495-598) on the outputs of the LAST forward call.
        Returns ``(total, rec[A], joint, neg_joint_entropy, qc_distance, c_distance, [kl_a], [], [ll_a])``.
        ``qc_all`` / ``c_smp_all`` ([n_arm_total,B,C], only with sharded arms) carry every arm's posteriors."""
        ctx = self._ctx
        if ctx.dims is None:
            raise RuntimeError("loss() called before forward()")
        ot = ctx.out_tensors
        if len(qc) != self.n_arm or qc[0].data_ptr() != ot["qc"].data_ptr() or c[0].data_ptr() != ot["c_smp"].data_ptr():
            raise RuntimeError("loss() must be given the tensors returned by the most recent forward() "
                               "(activations of earlier calls are overwritten)")
        lib = _lib.load()
        dev = self._flat_params.device
        At = self.n_arm_total
        want_grad = int(torch.is_grad_enabled() and ctx.training)
        if At != self.n_arm:
            if qc_all is None or c_smp_all is None:
                raise RuntimeError("sharded arms: pass qc_all / c_smp_all gathered over the arm axis")
            qa, ca = qc_all.contiguous(), c_smp_all.contiguous()
        else:
            qa, ca = ot["qc"], ot["c_smp"]
        loss_vec = torch.empty(5 + 3 * At, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mvae_loss(C.byref(ctx.dims), C.byref(ctx.hp), C.byref(ctx.state), C.byref(ctx.inputs),
                                 C.byref(ctx.outputs), qa.data_ptr(), ca.data_ptr(), loss_vec.data_ptr(), want_grad,
                                 C.c_void_p(stream)), "mvae_loss")
        ctx.keep += [qa, ca]
        ctx.loss_vec = loss_vec
        ctx.loss_done = bool(want_grad)
        total = loss_vec[0]
        if want_grad:
            total = _LossFn.apply(self._grad_anchor, total, self, ctx.gen)
        A0 = self.arm_offset
        rec = loss_vec[5 + A0:5 + A0 + self.n_arm]
        kls = [loss_vec[5 + At + A0 + a] for a in range(self.n_arm)]
        lls = [loss_vec[5 + 2 * At + A0 + a] for a in range(self.n_arm)]
        return total, rec, loss_vec[1], loss_vec[2], loss_vec[3], loss_vec[4], kls, [], lls

    def _run_backward(self, gen, grad_out):
        ctx = self._ctx
        if ctx.gen != gen or not ctx.loss_done:
            raise RuntimeError("backward() of a stale loss: forward() was called again before backward()")
        lib = _lib.load()
        dev = self._flat_params.device
        g = None if grad_out is None else grad_out.detach().to(dev, torch.float32).contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mvae_backward(C.byref(ctx.dims), C.byref(ctx.hp), C.byref(ctx.state), C.byref(ctx.inputs),
                                     C.byref(ctx.outputs), g.data_ptr() if g is not None else None,
                                     C.c_void_p(stream)), "mvae_backward")
        ctx.keep.append(g)
        ctx.loss_done = False
        self.bind_grads()

    # ------------------------------------------------------------------------------------------
    # fused step (cpl_mixvae.py:434-463 in one C call)
    # ------------------------------------------------------------------------------------------
    def fused_train_step(self, x, temp, optimizer, noise=None):
        """zero_grad + forward + loss + backward + Adam.  Returns the device loss vector
        (layout: include/mixvae_b200.h MVAE_LOSS_FLOATS).  ``optimizer`` must be ``FusedAdam``."""
        from .optim import FusedAdam
        if not isinstance(optimizer, FusedAdam) or optimizer.model is not self:
            raise TypeError("fused_train_step needs the FusedAdam bound to this model")
        if not self.training:
            raise RuntimeError("fused_train_step needs model.train()")
        if self.n_arm != self.n_arm_total:
            raise RuntimeError("sharded arms use mmidas_b200.parallel.ShardedTrainer")
        dev = self._require_cuda()
        lib = _lib.load()
        xt, x_arm_stride, x_row_stride, B = self._prep_x(x)
        U, E, keep_x, keep_s = self._prep_noise(noise, B, True)
        dims = self._dims(B)
        hp = self._hparams(temp)
        m, v = optimizer.flat_state()
        st = self._state(dims, m, v)
        self._step_counter += 1
        inp = self._inputs(xt, x_arm_stride, x_row_stride, U, E, keep_x, keep_s, True)
        ot = self._static_outputs(B)
        out = _lib.Outputs(*[ot[k].data_ptr() for k in ("x_low", "c_prob", "qc", "c_smp", "s_mean", "s_logvar", "s_smp")],
                           None)
        loss_vec = torch.empty(5 + 3 * self.n_arm, dtype=torch.float32, device=dev)
        g = optimizer.param_groups[0]
        optimizer.step_count += 1
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mvae_train_step(C.byref(dims), C.byref(hp), C.byref(st), C.byref(inp), C.byref(out),
                                       loss_vec.data_ptr(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                       float(g["eps"]), optimizer.step_count, C.c_void_p(stream)), "mvae_train_step")
        self._gen += 1
        ctx = _StepContext()
        ctx.gen = self._gen
        ctx.keep = [xt, U, E, keep_x, keep_s]
        ctx.out_tensors = ot
        self._ctx = ctx
        return loss_vec

    def fused_grad_step(self, x, temp, noise=None):
        """zero_grad + forward + loss + backward in one C call (``mvae_grad_step``), no optimiser step: what a data-parallel
        replica runs before its gradients are averaged.  Returns the device loss vector; gradients are in ``flat_grads()``."""
        if not self.training:
            raise RuntimeError("fused_grad_step needs model.train()")
        if self.n_arm != self.n_arm_total:
            raise RuntimeError("sharded arms run forward / loss / backward around the all-gather")
        dev = self._require_cuda()
        lib = _lib.load()
        xt, x_arm_stride, x_row_stride, B = self._prep_x(x)
        U, E, keep_x, keep_s = self._prep_noise(noise, B, True)
        dims = self._dims(B)
        hp = self._hparams(temp)
        st = self._state(dims)
        self._step_counter += 1
        inp = self._inputs(xt, x_arm_stride, x_row_stride, U, E, keep_x, keep_s, True)
        ot = self._static_outputs(B)
        out = _lib.Outputs(*[ot[k].data_ptr() for k in ("x_low", "c_prob", "qc", "c_smp", "s_mean", "s_logvar", "s_smp")],
                           None)
        loss_vec = torch.empty(5 + 3 * self.n_arm, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mvae_grad_step(C.byref(dims), C.byref(hp), C.byref(st), C.byref(inp), C.byref(out),
                                      loss_vec.data_ptr(), C.c_void_p(stream)), "mvae_grad_step")
        self._gen += 1
        ctx = _StepContext()
        ctx.gen = self._gen
        ctx.keep = [xt, U, E, keep_x, keep_s]
        ctx.out_tensors = ot
        ctx.loss_vec = loss_vec
        self._ctx = ctx
        self.bind_grads()
        return loss_vec

    def _static_outputs(self, B):
        key = ("static", B)
        cache = getattr(self, "_static_out", None)
        if cache is None or cache[0] != key:
            dev = self._flat_params.device
            A, Cc, S, L = self.n_arm, self.n_categories, self.state_dim, self.lowD_dim
            mk = lambda n: torch.empty(A, B, n, dtype=torch.float32, device=dev)
            cache = (key, {"x_low": mk(L), "c_prob": mk(Cc), "qc": mk(Cc), "c_smp": mk(Cc), "s_mean": mk(S),
                           "s_logvar": mk(S), "s_smp": mk(S)})
            self._static_out = cache
        return cache[1]

    def last_outputs(self):
        """Tensors written by the most recent forward / fused step ([n_arm, B, .])."""
        return self._ctx.out_tensors

    def argmax_labels(self, q: torch.Tensor) -> torch.Tensor:
        """Device-side ``classify`` (mmidas/_utils.py:78): argmax over categories -> int32."""
        self._require_cuda()
        q = q.contiguous()
        rows = q.numel() // q.size(-1)
        out = torch.empty(q.shape[:-1], dtype=torch.int32, device=q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(_lib.load().mvae_argmax(q.data_ptr(), out.data_ptr(), rows, q.size(-1), C.c_void_p(stream)),
                   "mvae_argmax")
        return out


_LIVE_GRAPHS = weakref.WeakSet()


def release_all_graphs():
    """Reset every captured step graph of this process.  A graph that captured NCCL collectives pins their communicator:
    ``destroy_process_group`` blocks until such graphs are gone (measured: the teardown of a 2-rank run hung for good), so
    ``_dist_utils.destroy_dist_env`` calls this first."""
    for g in list(_LIVE_GRAPHS):
        g.release()


class StepGraph:
    """One training step captured in a CUDA graph and replayed with ONE launch.

    What makes the step replayable: the library enqueues everything on the caller's stream without host
    synchronisation or allocation, tensor maps are baked into the kernel arguments at capture time, and the two
    per-step scalars — the noise step index and Adam's step — live on the device (``mvae_inputs.counters``), bumped by
    the first kernel of the step.  ``step_fn`` is any closure that runs one step on the model (the fused C call, or
    the sharded forward / all-gather / loss / backward / all-reduce / Adam sequence incl. its NCCL calls) and returns
    the device loss vector; it must read its input from the same device buffer on every replay.
    """

    def __init__(self, model: "mixVAE_model", optimizer, step_fn):
        dev = model._require_cuda()
        self.model, self.optimizer = model, optimizer
        self.counters = torch.zeros(2, dtype=torch.int64, device=dev)
        self._sync_counters()
        self.graph = torch.cuda.CUDAGraph()
        n0 = int(_lib.load().mvae_launch_count())
        model._graph_counters = self.counters
        optimizer._graph_counters = self.counters
        try:
            with torch.cuda.graph(self.graph):
                self.loss_vec = step_fn()
        finally:
            model._graph_counters = None
            optimizer._graph_counters = None
        # capture enqueues nothing: the host mirrors were advanced by step_fn's bookkeeping, undo that
        model._step_counter = self._held[0]
        optimizer.step_count = self._held[1]
        self.n_launches = int(_lib.load().mvae_launch_count()) - n0      # kernels of the library per replay
        _lib.note_capture(self.n_launches)
        self.ctx = model._ctx               # its output tensors live in the graph's pool: rewritten by every replay
        _LIVE_GRAPHS.add(self)

    def release(self):
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def _sync_counters(self):
        want = (self.model._step_counter, self.optimizer.step_count)
        if getattr(self, "_held", None) != want:
            self.counters[0].fill_(want[0])
            self.counters[1].fill_(want[1])
            self._held = want

    def replay(self) -> torch.Tensor:
        """Run the captured step; returns the (static) device loss vector of this replay."""
        if self.graph is None:
            raise RuntimeError("this step graph was released (process group torn down)")
        self._sync_counters()          # eager steps / eval forwards in between moved the host-side counters
        self.graph.replay()
        _lib.note_replay(self.n_launches)
        self.model._step_counter += 1
        self.optimizer.step_count += 1
        self._held = (self.model._step_counter, self.optimizer.step_count)
        self.model._gen += 1            # outstanding forward contexts are stale now
        self.ctx.gen, self.ctx.loss_done = -1, False
        self.model._ctx = self.ctx      # last_outputs() = what this replay wrote
        return self.loss_vec


def mk_vae(C, state_dim, input_dim, device, eps=1e-8, fc_dim=100, latent_dim=10, x_drop=0.5, s_drop=0.2, lr=0.001,
           lam=1, lam_pc=1, A=2, tau=0.005, beta=1.0, hard=False, variational=True, ref_prior=False, momentum=0.01,
           mode="MSE") -> nn.Module:
    """Same signature as the reference helper (nn_model.py:679-721)."""
    return mixVAE_model(input_dim=input_dim, fc_dim=fc_dim, n_categories=C, state_dim=state_dim, lowD_dim=latent_dim,
                        x_drop=x_drop, s_drop=s_drop, n_arm=A, lam=lam, lam_pc=lam_pc, tau=tau, beta=beta, hard=hard,
                        variational=variational, device=device, eps=eps, ref_prior=ref_prior, momentum=momentum,
                        loss_mode=mode).to(device)
