"""Trainer mirror of ``mmidas/cpl_mixvae.py`` (class cpl_mixVAE) for the B200 path.

Kept: ``cpl_mixVAE(saving_folder, aug_file, device, eps, save_flag, load_weights)`` (:153-161),
``init_model(...)`` (:193-216), ``train(...)`` signature (:323-337), ``load_model`` (:317), ``eval_model`` and the
dictionary it returns (:1450-1619), ``save_file`` / ``load_file`` (:1621-1650), the loss names printed/logged
(``train/total-loss`` ... ``val/consensus``, :536-560, :765-775) and the checkpoint files/dict keys (:777-788,
:851-865, :947-967).  ``.model`` and ``.optimizer`` stay public and re-assignable (train.py:141,145 re-creates the
optimizer).

Changed: the batch-loop body (:415-478) is one fused C call per step, replayed from a CUDA graph when the batch
arrives in a buffer seen before; the per-step host synchronisations of the reference (``_loss.item()`` :469,
``to_np(cs[a])`` :476) are gone — losses are summed on the device and read once per epoch, labels are computed by a
device argmax and only arm-pair confusion counts are copied once per epoch; the next batch's H2D copy (dense, or
row-packed and expanded on the device) runs on a side stream while the current step computes.  ``train(ws > 1)`` —
which the reference gates off (train.py:274-275) — runs on the (arm x dp) mesh of ``mmidas_b200.parallel``.
"""
from __future__ import annotations

import os
import pickle
import time
from typing import Iterable, Optional

import numpy as np
import torch

from ._utils import confmat_device, consensus, consensus_from_counts
from .dataloader import PackedBatch
from .nn_model import StepGraph, VAEConfig, mixVAE_model
from .optim import FusedAdam


def is_master(rank):
    return rank == 0 or rank in ("mps", "cpu", "cuda") or (isinstance(rank, str) and rank.startswith("cuda"))


class HostBatchFeeder:
    """Iterates over host batches and yields device tensors, copying batch i+1 on a side stream while batch i is being
    consumed (the reference does a blocking ``x.to(rank)`` at :416).

    Items may be dense CPU tensors (pinned on the fly when they are not), tuples ``(x, idx)`` as a DataLoader yields,
    device tensors (passed through) or ``PackedBatch`` objects.  Dense batches land in a small ring of device buffers
    per batch shape.  Packed batches cross PCIe in row-packed form into a ring of staging buffers and are expanded
    (bit-exactly) on the COMPUTE stream right before they are handed out, into one dense buffer per shape: the copy
    stream then carries nothing but the copy, which is what bounds an end-to-end step.  Either way consecutive steps
    see the same few device pointers — what lets the trainer replay a captured CUDA graph.  A ring slot is refilled
    only after the work that read it has finished (event edge to the copy stream).  Rings are pooled per
    (device, shape) and handed from an exhausted feeder to the next one, so the pointers — and the graphs captured for
    them — survive across epochs.
    """

    _ring_pool = {}       # (device, kind, shape, depth) -> [ring, ...] not in use by a live feeder

    def __init__(self, batches: Iterable, device, pin: bool = True, depth: int = 3):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.pin = pin
        self.depth = max(2, int(depth))
        self.stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.batches_copied = 0
        self._rings = {}          # (kind, shape) -> [[buffers], next slot]
        self._events = []         # event k: everything enqueued on the compute stream before the k-th hand-over
        self._next = None
        self._mark()
        self._preload()

    def _release(self):
        for key, ring in self._rings.items():
            HostBatchFeeder._ring_pool.setdefault((self.device, self.depth) + key, []).append(ring)
        self._rings = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _slot(self, kind, shape, make, n=None):
        key = (kind, shape)
        ring = self._rings.get(key)
        if ring is None:
            free = HostBatchFeeder._ring_pool.get((self.device, self.depth) + key)
            ring = free.pop() if free else [[make() for _ in range(n or self.depth)], 0]
            self._rings[key] = ring
        buf = ring[0][ring[1]]
        ring[1] = (ring[1] + 1) % len(ring[0])
        return buf

    def _mark(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._events.append(ev)
        if len(self._events) > self.depth:
            self._events.pop(0)

    def _preload(self):
        try:
            item = next(self.it)
        except StopIteration:
            self._next = None
            return
        x = item[0] if isinstance(item, (tuple, list)) else item
        # the slot about to be refilled held the batch handed out `depth` calls ago; whatever read it was enqueued before
        # the call after that one, i.e. before event[-(depth - 1)]
        ev = self._events[-(self.depth - 1)] if len(self._events) >= self.depth - 1 else None
        if ev is not None:
            self.stream.wait_event(ev)
        packed = None
        with torch.cuda.stream(self.stream):
            if isinstance(x, PackedBatch):
                cap = (x.nbytes * 5 // 4 + 4095) // 4096 * 4096
                st = self._slot("packed", tuple(x.shape), lambda: torch.empty(cap, dtype=torch.uint8, device=self.device))
                if st.numel() < x.nbytes:          # a denser batch than the ring was sized for: one-off buffer
                    st = torch.empty(x.nbytes, dtype=torch.uint8, device=self.device)
                st[:x.nbytes].copy_(x.buffer, non_blocking=True)
                packed, xd = (x, st), None
                self.h2d_bytes += x.nbytes
                self.batches_copied += 1
            elif x.device.type == "cpu":
                if x.dtype != torch.float32:
                    x = x.float()
                if self.pin and not x.is_pinned():
                    x = x.pin_memory()
                shape = tuple(x.shape)
                xd = self._slot("dense", shape, lambda: torch.empty(shape, dtype=torch.float32, device=self.device))
                xd.copy_(x, non_blocking=True)
                self.h2d_bytes += x.numel() * x.element_size()
                self.batches_copied += 1
            else:
                xd = x
        self._next = (xd, item, packed)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            self._release()
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        xd, item, packed = self._next
        if packed is not None:       # expand on the compute stream: ordered before the step that reads it, after the last one
            pb, st = packed
            shape = tuple(pb.shape)
            out = self._slot("unpacked", shape, lambda: torch.empty(shape, dtype=torch.float32, device=self.device), n=1)
            xd = pb.expand_from(st, out)
            st.record_stream(cur)
        self._mark()
        self._preload()
        return xd, item


class cpl_mixVAE:
    def __init__(self, saving_folder="", aug_file="", device=None, eps=1e-8, save_flag=True, load_weights=True,
                 aug_precision="tf32x3"):
        self.eps = eps
        self.save = save_flag
        self.folder = saving_folder
        self.aug_file = aug_file
        self.models = []
        if device is None or device == "cpu" or device == "mps":
            raise RuntimeError("cpl_mixVAE (B200) needs a CUDA device; there is no CPU path")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if aug_file:     # cpl_mixvae.py:182-186: pre-trained VAE-GAN generator, eval mode
            from .augmentation import mk_augmenter
            self.aug_model, self.aug_param, netA = mk_augmenter(aug_file, load_weights, precision=aug_precision)
            self.netA = netA.to(self.device).eval()
        else:
            self.aug_model, self.aug_param, self.netA = None, None, None
        self.precision = "tf32x3_fc1"
        self.use_cuda_graph = True      # replay the fused step from a CUDA graph when the batch buffer repeats
        self.mesh_mode = "dp"           # train(ws > 1): "dp", "arm" or "auto" (mmidas_b200.parallel.plan_mesh)
        self._graphs = {}
        self._graph_misses = 0

    # ------------------------------------------------------------------------------------------
    def init_model(self, n_categories, state_dim, input_dim, fc_dim=100, lowD_dim=10, x_drop=0.5, s_drop=0.2,
                   lr=0.001, lam=1, lam_pc=1, n_arm=2, temp=1.0, tau=0.005, beta=1.0, hard=False, variational=True,
                   ref_prior=False, trained_model="", n_pr=0, momentum=0.01, mode="MSE"):
        self.lowD_dim = lowD_dim
        self.n_categories = n_categories
        self.state_dim = state_dim
        self.input_dim = input_dim
        self.temp = temp
        self.n_arm = n_arm
        self.fc_dim = fc_dim
        self.ref_prior = ref_prior
        self.model = mixVAE_model(input_dim=input_dim, fc_dim=fc_dim, n_categories=n_categories, state_dim=state_dim,
                                  lowD_dim=lowD_dim, x_drop=x_drop, s_drop=s_drop, n_arm=n_arm, lam=lam, lam_pc=lam_pc,
                                  tau=tau, beta=beta, hard=hard, variational=variational, device=self.device,
                                  eps=self.eps, ref_prior=ref_prior, momentum=momentum, loss_mode=mode,
                                  precision=self.precision)
        self.model = self.model.to(self.device)
        self.optimizer = FusedAdam(self.model.parameters(), lr=lr, model=self.model)
        self._graphs = {}
        self._graph_misses = 0
        if len(trained_model) > 0:
            print("Load the pre-trained model")
            loaded_file = torch.load(trained_model, map_location="cpu")
            self.model.load_state_dict(loaded_file["model_state_dict"])
            self.optimizer.load_state_dict(loaded_file["optimizer_state_dict"])
            self.init = False
            self.n_pr = n_pr
        else:
            self.init = True
            self.n_pr = 0

    def append(self, c: VAEConfig):
        model = mixVAE_model(input_dim=c.input_dim, fc_dim=c.fc_dim, n_categories=c.n_categories, state_dim=c.state_dim,
                             lowD_dim=c.lowD_dim, x_drop=c.x_drop, s_drop=c.s_drop, n_arm=c.n_arm, lam=c.lam,
                             lam_pc=c.lam_pc, tau=c.tau, beta=c.beta, hard=c.hard, variational=c.variational,
                             device=self.device, eps=self.eps, ref_prior=c.ref_prior, momentum=c.momentum,
                             loss_mode=c.mode).to(self.device)
        optimizer = FusedAdam(model.parameters(), lr=c.lr, model=model)
        if c.trained_model:
            loaded_file = torch.load(c.trained_model, map_location="cpu")
            model.load_state_dict(loaded_file["model_state_dict"])
            optimizer.load_state_dict(loaded_file["optimizer_state_dict"])
        self.models.append({"model": model, "opt": optimizer})

    def load_model(self, trained_model):
        loaded_file = torch.load(trained_model, map_location="cpu")
        self.model.load_state_dict(loaded_file["model_state_dict"])
        self.current_time = time.strftime("%Y-%m-%d-%H-%M-%S")

    def _save(self, path, state=None):
        print(f"saving model to: {path}")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        msd, osd = state if state is not None else (self.model.state_dict(), self.optimizer.state_dict())
        torch.save({"model_state_dict": msd, "optimizer_state_dict": osd}, path)

    def save_file(self, fname, **kwargs):
        """cpl_mixvae.py:1621-1635: pickle the keyword arguments to ``fname + '.p'`` (protocol 4)."""
        with open(fname + ".p", "wb") as f:
            pickle.dump(dict(kwargs), f, protocol=4)

    def load_file(self, fname):
        """cpl_mixvae.py:1637-1650."""
        with open(fname + ".p", "rb") as f:
            return pickle.load(f)

    # ------------------------------------------------------------------------------------------
    # one optimiser step (cpl_mixvae.py:434-463)
    # ------------------------------------------------------------------------------------------
    def train_batch(self, x: torch.Tensor, noise=None, aug_noise=None) -> torch.Tensor:
        """zero_grad -> forward -> loss -> backward -> Adam on one batch ``x`` [B, D] already on the device.
        Returns the device loss vector (total, joint, entropy, distance, l2, rec[A], kl[A], ll[A]); nothing
        is synchronised.  The vector is overwritten by the next call."""
        A = self.n_arm
        model, opt = self.model, self.optimizer
        fused = isinstance(opt, FusedAdam) and opt.model is model
        graphable = (fused and self.use_cuda_graph and self.netA is None and noise is None and aug_noise is None
                     and model.s_drop == 0.0 and model.training and x.is_cuda and x.dtype == torch.float32
                     and x.dim() == 2 and x.stride(-1) == 1 and self._graph_misses < 32)
        if graphable:
            g0 = opt.param_groups[0]
            key = (x.data_ptr(), tuple(x.shape), x.stride(0), float(self.temp), float(g0["lr"]), tuple(g0["betas"]), float(g0["eps"]),
                   id(model), id(opt))
            g = self._graphs.get(key)
            if g is not None and g.graph is not None:
                self._graph_misses = 1
                return g.replay()
            if self._graph_misses > 0 or self._graphs:        # (the very first step runs eagerly: lazy allocations)
                if len(self._graphs) >= 8:
                    self._graphs.pop(next(iter(self._graphs)))
                torch.cuda.synchronize(self.device)
                try:
                    g = StepGraph(model, opt, lambda: model.fused_train_step(x.expand(A, -1, -1), self.temp, opt))
                except Exception as err:           # capture refused: stay on the eager path
                    print(f"cpl_mixVAE: CUDA graph capture failed ({err}); continuing without graphs")
                    self.use_cuda_graph = False
                    g = None
                if g is not None:
                    self._graphs[key] = g
                    self._graph_misses += 1
                    return g.replay()
            self._graph_misses += 1
        xs = x.expand(A, -1, -1)
        if self.netA is not None:      # cpl_mixvae.py:422-423: every arm trains on its own augmented copy of the batch
            xs = self.netA(xs, True, 0.1, noise=aug_noise)[1]
        if fused:
            return model.fused_train_step(xs, self.temp, opt, noise=noise)
        # a caller replaced .optimizer (train.py:144-147): reference-shaped sequence on the same kernels
        opt.zero_grad()
        x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = model(xs, self.temp, 0.0, noise=noise)
        model.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)[0].backward()
        opt.step()
        return model._ctx.loss_vec

    # ------------------------------------------------------------------------------------------
    def train(self, train_loader, test_loader, n_epoch, n_epoch_p, c_p=0, c_onehot=0, min_con=0.5, max_prun_it=0,
              rank=None, run=None, ws=1, good_enuf_consensus=0.75):
        """Epoch loop with the reference's bookkeeping (cpl_mixvae.py:323-967): per epoch one training
        pass, one eval-mode pass over the training set, one validation pass; consensus between arms;
        checkpoints every 10 epochs, at the consensus threshold and at the end.  Returns the curves.
        ``ws > 1`` (one process per GPU, process group initialised by ``init_dist_env``): the step runs on the
        (arm x dp) mesh ``self.mesh_mode``; every rank iterates ITS loader, and the epoch sums are all-reduced as
        at :480-483."""
        if rank is None:
            rank = self.device
        if n_epoch_p > 0 or max_prun_it > 0:
            raise NotImplementedError("pruning is forcibly disabled in the reference (cpl_mixvae.py:1007-1008)")
        A, C, E, D = self.n_arm, self.n_categories, n_epoch, self.input_dim
        Bs, Bs_val = len(train_loader), len(test_loader)
        B_val = test_loader.batch_size
        self.current_time = time.strftime("%Y-%m-%d-%H-%M-%S")
        dev = self.device
        dist_on = ws > 1
        st = None
        if dist_on:
            import torch.distributed as dist
            from .parallel import ShardedTrainer
            if not dist.is_initialized():
                raise RuntimeError("train(ws > 1) needs an initialised process group (mmidas_b200._dist_utils.init_dist_env)")
            if self.netA is not None:
                raise NotImplementedError("the augmenter runs on the single-GPU path only")
            st = ShardedTrainer(model=self.model, lr=self.optimizer.param_groups[0]["lr"], mode=self.mesh_mode, temp=self.temp,
                                use_cuda_graph=self.use_cuda_graph)
            model = st.model
        else:
            model = self.model
        model.materialize_recon = False
        losses, loss_joints, c_ents, c_l2_dists, c_dists = [], [], [], [], []
        loss_recs = [[] for _ in range(A)]
        consensus_train, consensus_aug, consensus_val = [], [], []
        validation_loss = np.zeros(E)
        validation_rec_loss = np.zeros(E)
        epoch_times = []
        if not getattr(self, "init", True):
            return {"skipped": True}     # reference: a loaded model skips the loop (:397)
        print("training started")
        master = (not dist_on) or dist.get_rank() == 0

        def save(path):
            if dist_on:
                state = st.full_state_dicts()        # collective over the arm axis: every rank calls it
                if master:
                    self._save(path, state)
            else:
                self._save(path)

        for e in range(E):
            t0 = time.time()
            model.train()
            loss_sum = torch.zeros(5 + 3 * A, device=dev)
            cm_aug = None                     # [pairs, C, C] co-assignment counts, accumulated on the device
            nb = 0
            for x, _item in HostBatchFeeder(train_loader, dev):
                if dist_on:
                    lv = st.step(x)
                    labels = st.all_labels(model.last_outputs()["qc"])
                else:
                    lv = self.train_batch(x)
                    labels = model.argmax_labels(model.last_outputs()["qc"])
                loss_sum += lv
                cm_aug = confmat_device(labels, C, cm_aug)
                nb += 1
            cnt = torch.tensor([float(nb)], device=dev)
            if dist_on:      # cpl_mixvae.py:480-483 (loss carries [sum, count]; rec and distance sums ride in the same vector)
                dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM)
                dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            ls = loss_sum.cpu().numpy() / max(float(cnt.item()), 1.0)          # the only D2H of the training pass
            losses.append(float(ls[0]))
            loss_joints.append(float(ls[1]))
            c_ents.append(float(ls[2]))
            c_dists.append(float(ls[3]))
            c_l2_dists.append(float(ls[4]))
            for a in range(A):
                loss_recs[a].append(float(ls[5 + a]) / D)
            consensus_aug.append(consensus_from_counts(cm_aug) if cm_aug is not None else float("nan"))
            _time = time.time() - t0
            print(f"epoch {e} | loss: {losses[-1]:.2f} | rec: {loss_recs[0][-1]:.2f} | distance: {c_dists[-1]:.2f} | "
                  f"l2 distance: {c_l2_dists[-1]:.2f} | aug-cns: {consensus_aug[-1]:.2f} | time: {_time:.2f} | "
                  f"avg time: {np.mean(epoch_times) if epoch_times else float('nan'):.2f} | ", end="")
            if run:
                run.log({"train/total-loss": losses[-1], "train/joint-loss": loss_joints[-1],
                         "train/negative-joint-entropy": c_ents[-1], "train/simplex-distance": c_dists[-1],
                         "train/l2-distance": c_l2_dists[-1], "train/time": _time,
                         "train/mem": torch.cuda.memory_allocated() / 1e6, "train/consensus_aug": consensus_aug[-1],
                         **{f"train/rec-loss{a}": loss_recs[a][-1] for a in range(A)}})

            # ---- eval-mode pass over the training set (:563-663)
            model.eval()
            cm_noaug = self._eval_confmat(train_loader if B_val > 1 else [train_loader.dataset.tensors], st)
            consensus_train.append(consensus_from_counts(cm_noaug))
            if run:
                run.log({"train/consensus": consensus_train[-1]})

            # ---- validation (:666-775); test batch_size==1 means "whole test set as one batch"
            val_batches = test_loader if B_val > 1 else [test_loader.dataset.tensors]
            val_loss, val_rec, lab_val, nvb = self._eval_loss(val_batches, st)
            consensus_val.append(consensus([lab_val[a] for a in range(A)], C))
            denom = Bs_val if B_val > 1 else 1
            validation_rec_loss[e] = val_rec / denom / A
            validation_loss[e] = val_loss / denom
            print(f"val-loss {validation_loss[e]:.2f} | rec-loss {validation_rec_loss[e]:.2f} | val-cns {consensus_val[-1]:.2f}")
            if run:
                run.log({"val/total-loss": validation_loss[e], "val/rec-loss": validation_rec_loss[e],
                         "val/consensus": consensus_val[-1]})
            if self.save and e > 0 and e % 10 == 0:
                save(self.folder + f"/model/cpl_mixVAE_model_epoch_{e}.pth")
            stop = consensus_train[-1] >= good_enuf_consensus or e == E - 1
            if dist_on:      # every rank must take the same branch (the checkpoint gather is collective)
                flag = torch.tensor([1.0 if stop else 0.0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                stop = bool(flag.item() > 0)
            if stop:
                if self.save:
                    save(self.folder + f"/model/cns_cpl_mixVAE_model_before_pruning_A{A}_" + self.current_time + ".pth")
                epoch_times.append(time.time() - t0)
                break
            epoch_times.append(time.time() - t0)
        if self.save and n_epoch > 0:
            save(self.folder + f"/model/cpl_mixVAE_model_before_pruning_A{A}_" + self.current_time + ".pth")
        if dist_on:      # hand the trained weights back to the public full model (every rank)
            msd, osd = st.full_state_dicts()
            self.model.load_state_dict(msd)
            self.optimizer.load_state_dict(osd)
        return {"losses": losses, "loss_joints": loss_joints, "loss_recs": loss_recs, "c_ents": c_ents,
                "c_dists": c_dists, "c_l2_dists": c_l2_dists, "consensus_aug": consensus_aug,
                "consensus_train": consensus_train, "consensus_val": consensus_val,
                "validation_loss": validation_loss, "validation_rec_loss": validation_rec_loss,
                "epoch_times": epoch_times}

    # ------------------------------------------------------------------------------------------
    def _eval_confmat(self, batches, st=None):
        """Eval-mode pass (cpl_mixvae.py:563-663): argmax labels and their co-assignment counts stay on the device;
        only [pairs, C, C] integers come back."""
        model, A = (st.model if st is not None else self.model), self.n_arm
        cm = None
        with torch.no_grad():
            for x, _ in HostBatchFeeder(batches, self.device):
                if st is not None:
                    _, labels = st.eval_batch(x)
                else:
                    xs = [x for _ in range(A)]
                    model.materialize_recon = False
                    ctx = model._launch_forward(xs, self.temp, True, None, False)
                    labels = model.argmax_labels(ctx.out_tensors["qc"])
                cm = confmat_device(labels, self.n_categories, cm)
        return cm

    def _eval_loss(self, batches, st=None):
        model, A, D = (st.model if st is not None else self.model), self.n_arm, self.input_dim
        tot = torch.zeros((), device=self.device)
        rec = torch.zeros((), device=self.device)
        labs = []
        n = 0
        with torch.no_grad():
            for x, _ in HostBatchFeeder(batches, self.device):
                if st is not None:
                    lv, labels = st.eval_batch(x)
                    tot += lv[0]
                    rec += lv[5:5 + A].sum() / D
                    labs.append(labels)
                else:
                    xs = [x for _ in range(A)]
                    x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = model(x=xs, temp=self.temp, prior_c=0.0, eval=True)
                    loss, loss_rec, *_ = model.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)
                    tot += loss
                    rec += loss_rec.sum() / D
                    labs.append(model.argmax_labels(torch.stack(cs)))
                n += 1
        return tot.item(), rec.item(), torch.cat(labs, dim=1).cpu().numpy().astype(np.int64), n

    def eval_model(self, data_loader, c_p=0, c_onehot=0, noise=None):
        """Inference summary with the reference's contract (cpl_mixvae.py:1450-1619): eval-mode forward with the
        category mask taken from the non-zero entries of ``fcc[0].bias`` (:1475-1477, :1534), loss per batch, and the
        dictionary of :1590-1619 (``predicted_label`` / ``state_cat`` 1-based, float64 arrays).  ``noise`` (not in the
        reference): ``{"E": [n_batches, A, B, S]}`` injects the state noise that forward draws even in eval mode
        (nn_model.py:351), for parity tests.  A loader with ``batch_size == 1`` is read as "the whole set as one batch",
        the convention of the reference's training loop (:722-748) — per-cell batches would make ``inv_var``'s batch
        variance (nn_model.py:75) undefined."""
        model, A, C = self.model, self.n_arm, self.n_categories
        N = len(data_loader.dataset)
        D, D_low, S = self.input_dim, self.lowD_dim, self.state_dim
        model.eval()
        bias = model.fcc[0].bias.detach().cpu().numpy()
        pruning_mask = np.where(bias != 0.0)[0]
        prune_indx = np.where(bias == 0.0)[0]
        whole = getattr(data_loader, "batch_size", None) == 1
        batches = [data_loader.dataset.tensors] if whole else data_loader
        keys = ("s_mean", "s_logvar", "qc", "c_smp", "x_low", "x_rec")
        outs = {k: [] for k in keys}
        idxs, losses, c_dists, c_l2_dists, loss_recs, lls, labels = [], [], [], [], [], [], []
        mat = model.materialize_recon
        model.materialize_recon = True
        with torch.no_grad():
            for i, (x, item) in enumerate(HostBatchFeeder(batches, self.device)):
                xs = [x for _ in range(A)]
                nz = None if noise is None else {"E": noise["E"][i]}
                x_recs, _, _, x_lows, cs, s_smps, c_smps, s_means, s_logvars, _ = model(
                    xs, self.temp, prior_c=0.0, eval=True, mask=pruning_mask if len(prune_indx) else None, noise=nz)
                ls = model.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)
                lv = model._ctx.loss_vec
                losses.append(lv[0:1])
                c_dists.append(lv[3:4])
                c_l2_dists.append(lv[4:5])
                loss_recs.append(ls[1])
                lls.append(torch.stack(list(ls[8])))
                for k, v in zip(keys, (s_means, s_logvars, cs, c_smps, x_lows, x_recs)):
                    outs[k].append(torch.stack(v).cpu())
                labels.append(model.argmax_labels(torch.stack(cs)))
                idx = item[1] if isinstance(item, (tuple, list)) and len(item) > 1 else torch.arange(x.shape[0])
                idxs.append(torch.as_tensor(idx).reshape(-1).cpu().to(torch.float64))
        model.materialize_recon = mat
        cat = lambda k: torch.cat(outs[k], dim=1).numpy().astype(np.float64)
        cs_np = cat("qc")
        lab = torch.cat(labels, dim=1).cpu().numpy().astype(np.int64)
        mean_of = lambda lst: torch.stack([t.reshape(-1) for t in lst]).double().mean(0).cpu().numpy()
        return {
            "state_mu": cat("s_mean"),
            "state_var": cat("s_logvar"),
            "state_cat": (lab + 1).astype(np.float64),
            "prob_cat": cs_np.max(-1),
            "total_loss_rec": mean_of(loss_recs),
            "total_likelihood": mean_of(lls),
            "total_dist_z": float(mean_of(c_dists)[0]),
            "total_dist_qz": float(mean_of(c_l2_dists)[0]),
            "mean_test_rec": np.zeros(A),
            "predicted_label": (lab + 1).astype(np.float64),
            "data_indx": torch.cat(idxs).numpy()[:N],
            "z_prob": cs_np,
            "z_sample": cat("c_smp"),
            "x_low": cat("x_low"),
            "recon_c": cat("x_rec"),
            "prune_indx": prune_indx,
            "cnss": consensus([lab[a] for a in range(A)], C),
            "total_loss": float(mean_of(losses)[0]),
        }
