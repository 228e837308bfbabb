"""Trainer mirror of ``mmidas/cpl_mixvae.py`` (class cpl_mixVAE) for the B200 path.

Kept: ``cpl_mixVAE(saving_folder, aug_file, device, eps, save_flag, load_weights)`` (:153-161),
``init_model(...)`` (:193-216), ``train(...)`` signature (:323-337), ``load_model`` (:317), the loss
names printed/logged (``train/total-loss`` ... ``val/consensus``, :536-560, :765-775) and the
checkpoint files/dict keys (:777-788, :851-865, :947-967).  ``.model`` and ``.optimizer`` stay public
and re-assignable (train.py:141,145 re-creates the optimizer).

Changed: the batch-loop body (:415-478) is one fused C call per step; the per-step host
synchronisations of the reference (``_loss.item()`` :469, ``to_np(cs[a])`` :476) are gone — losses are
summed on the device and read once per epoch, labels are computed by a device argmax and copied once
per epoch; the next batch's H2D copy runs on a side stream while the current step computes.
"""
from __future__ import annotations

import os
import time
from typing import Iterable, Optional

import numpy as np
import torch

from ._utils import confmat_device, consensus, consensus_from_counts
from .nn_model import VAEConfig, mixVAE_model
from .optim import FusedAdam


def is_master(rank):
    return rank == 0 or rank in ("mps", "cpu", "cuda") or (isinstance(rank, str) and rank.startswith("cuda"))


class HostBatchFeeder:
    """Iterates over host batches and yields device tensors, copying batch i+1 on a side stream
    while batch i is being consumed (the reference does a blocking ``x.to(rank)`` at :416)."""

    def __init__(self, batches: Iterable, device, pin: bool = True):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.pin = pin
        self.stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self._next = None
        self._preload()

    def _preload(self):
        try:
            item = next(self.it)
        except StopIteration:
            self._next = None
            return
        x = item[0] if isinstance(item, (tuple, list)) else item
        if x.device.type == "cpu":
            if self.pin and not x.is_pinned():
                x = x.pin_memory()
            self.h2d_bytes += x.numel() * x.element_size()
        with torch.cuda.stream(self.stream):
            xd = x.to(self.device, non_blocking=True)
        self._next = (xd, item)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        xd, item = self._next
        xd.record_stream(torch.cuda.current_stream(self.device))
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

    def _save(self, path):
        print(f"saving model to: {path}")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"model_state_dict": self.model.state_dict(),
                    "optimizer_state_dict": self.optimizer.state_dict()}, path)

    # ------------------------------------------------------------------------------------------
    # one optimiser step (cpl_mixvae.py:434-463)
    # ------------------------------------------------------------------------------------------
    def train_batch(self, x: torch.Tensor, noise=None, aug_noise=None) -> torch.Tensor:
        """zero_grad -> forward -> loss -> backward -> Adam on one batch ``x`` [B, D] already on the device.
        Returns the device loss vector (total, joint, entropy, distance, l2, rec[A], kl[A], ll[A]); nothing
        is synchronised."""
        A = self.n_arm
        xs = x.expand(A, -1, -1)
        if self.netA is not None:      # cpl_mixvae.py:422-423: every arm trains on its own augmented copy of the batch
            xs = self.netA(xs, True, 0.1, noise=aug_noise)[1]
        if isinstance(self.optimizer, FusedAdam) and self.optimizer.model is self.model:
            return self.model.fused_train_step(xs, self.temp, self.optimizer, noise=noise)
        # a caller replaced .optimizer (train.py:144-147): reference-shaped sequence on the same kernels
        self.optimizer.zero_grad()
        x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = self.model(xs, self.temp, 0.0, noise=noise)
        self.model.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)[0].backward()
        self.optimizer.step()
        return self.model._ctx.loss_vec

    # ------------------------------------------------------------------------------------------
    def train(self, train_loader, test_loader, n_epoch, n_epoch_p, c_p=0, c_onehot=0, min_con=0.5, max_prun_it=0,
              rank=None, run=None, ws=1, good_enuf_consensus=0.75):
        """Epoch loop with the reference's bookkeeping (cpl_mixvae.py:323-967): per epoch one training
        pass, one eval-mode pass over the training set, one validation pass; consensus between arms;
        checkpoints every 10 epochs, at the consensus threshold and at the end.  Returns the curves."""
        if rank is None:
            rank = self.device
        if ws > 1:
            raise NotImplementedError("use mmidas_b200.parallel.ShardedTrainer for multi-GPU runs "
                                      "(the reference itself raises for ws > 1, train.py:274-275)")
        if n_epoch_p > 0 or max_prun_it > 0:
            raise NotImplementedError("pruning is forcibly disabled in the reference (cpl_mixvae.py:1007-1008)")
        A, C, E, D = self.n_arm, self.n_categories, n_epoch, self.input_dim
        Bs, Bs_val = len(train_loader), len(test_loader)
        B_val = test_loader.batch_size
        self.current_time = time.strftime("%Y-%m-%d-%H-%M-%S")
        model = self.model
        model.materialize_recon = False
        dev = self.device
        losses, loss_joints, c_ents, c_l2_dists, c_dists = [], [], [], [], []
        loss_recs = [[] for _ in range(A)]
        consensus_train, consensus_aug, consensus_val = [], [], []
        validation_loss = np.zeros(E)
        validation_rec_loss = np.zeros(E)
        epoch_times = []
        if not getattr(self, "init", True):
            return {"skipped": True}     # reference: a loaded model skips the loop (:397)
        print("training started")
        for e in range(E):
            t0 = time.time()
            model.train()
            loss_sum = torch.zeros(5 + 3 * A, device=dev)
            cm_aug = None                     # [pairs, C, C] co-assignment counts, accumulated on the device
            nb = 0
            for x, _item in HostBatchFeeder(train_loader, dev):
                lv = self.train_batch(x)
                loss_sum += lv
                cm_aug = confmat_device(model.argmax_labels(model.last_outputs()["qc"]), C, cm_aug)
                nb += 1
            ls = loss_sum.cpu().numpy() / max(nb, 1)          # the only D2H of the training pass
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
            cm_noaug = self._eval_confmat(train_loader if B_val > 1 else [train_loader.dataset.tensors])
            consensus_train.append(consensus_from_counts(cm_noaug))
            if run:
                run.log({"train/consensus": consensus_train[-1]})

            # ---- validation (:666-775); test batch_size==1 means "whole test set as one batch"
            val_batches = test_loader if B_val > 1 else [test_loader.dataset.tensors]
            val_loss, val_rec, lab_val, nvb = self._eval_loss(val_batches)
            consensus_val.append(consensus([lab_val[a] for a in range(A)], C))
            denom = Bs_val if B_val > 1 else 1
            validation_rec_loss[e] = val_rec / denom / A
            validation_loss[e] = val_loss / denom
            print(f"val-loss {validation_loss[e]:.2f} | rec-loss {validation_rec_loss[e]:.2f} | val-cns {consensus_val[-1]:.2f}")
            if run:
                run.log({"val/total-loss": validation_loss[e], "val/rec-loss": validation_rec_loss[e],
                         "val/consensus": consensus_val[-1]})
            if self.save and e > 0 and e % 10 == 0:
                self._save(self.folder + f"/model/cpl_mixVAE_model_epoch_{e}.pth")
            if consensus_train[-1] >= good_enuf_consensus or e == E - 1:
                if self.save:
                    self._save(self.folder + f"/model/cns_cpl_mixVAE_model_before_pruning_A{A}_" + self.current_time + ".pth")
                epoch_times.append(time.time() - t0)
                break
            epoch_times.append(time.time() - t0)
        if self.save and n_epoch > 0:
            self._save(self.folder + f"/model/cpl_mixVAE_model_before_pruning_A{A}_" + self.current_time + ".pth")
        return {"losses": losses, "loss_joints": loss_joints, "loss_recs": loss_recs, "c_ents": c_ents,
                "c_dists": c_dists, "c_l2_dists": c_l2_dists, "consensus_aug": consensus_aug,
                "consensus_train": consensus_train, "consensus_val": consensus_val,
                "validation_loss": validation_loss, "validation_rec_loss": validation_rec_loss,
                "epoch_times": epoch_times}

    # ------------------------------------------------------------------------------------------
    def _eval_confmat(self, batches):
        """Eval-mode pass (cpl_mixvae.py:563-663): argmax labels and their co-assignment counts stay on the device;
        only [pairs, C, C] integers come back."""
        model, A = self.model, self.n_arm
        cm = None
        with torch.no_grad():
            for x, _ in HostBatchFeeder(batches, self.device):
                xs = [x for _ in range(A)]
                model.materialize_recon = False
                ctx = model._launch_forward(xs, self.temp, True, None, False)
                cm = confmat_device(model.argmax_labels(ctx.out_tensors["qc"]), self.n_categories, cm)
        return cm

    def _eval_labels(self, batches):
        model, A = self.model, self.n_arm
        labs = []
        with torch.no_grad():
            for x, _ in HostBatchFeeder(batches, self.device):
                xs = [x for _ in range(A)]
                model.materialize_recon = False
                ctx = model._launch_forward(xs, self.temp, True, None, False)
                labs.append(model.argmax_labels(ctx.out_tensors["qc"]))
        return torch.cat(labs, dim=1).cpu().numpy().astype(np.int64)

    def _eval_loss(self, batches):
        model, A, D = self.model, self.n_arm, self.input_dim
        tot = torch.zeros((), device=self.device)
        rec = torch.zeros((), device=self.device)
        labs = []
        n = 0
        with torch.no_grad():
            for x, _ in HostBatchFeeder(batches, self.device):
                xs = [x for _ in range(A)]
                x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = model(x=xs, temp=self.temp, prior_c=0.0, eval=True)
                loss, loss_rec, *_ = model.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)
                tot += loss
                rec += loss_rec.sum() / D
                labs.append(model.argmax_labels(torch.stack(cs)))
                n += 1
        return tot.item(), rec.item(), torch.cat(labs, dim=1).cpu().numpy().astype(np.int64), n

    def eval_model(self, data_loader, c_p=0, c_onehot=0):
        """Inference summary (cpl_mixvae.py:1450-1619), reduced to what the hot path produces: per-arm
        categorical posteriors, argmax labels, state means/samples, low-D representation, losses."""
        model, A = self.model, self.n_arm
        model.eval()
        outs = {k: [] for k in ("c_prob", "qc", "c_smp", "s_mean", "s_logvar", "s_smp", "x_low", "labels")}
        tot = []
        # a loader with batch_size 1 means "the whole set as one batch" (the reference's convention, :722-748)
        batches = [data_loader.dataset.tensors] if getattr(data_loader, "batch_size", None) == 1 else data_loader
        with torch.no_grad():
            for x, _ in HostBatchFeeder(batches, self.device):
                xs = [x for _ in range(A)]
                x_recs, _, _, x_lows, cs, s_smps, c_smps, s_means, s_logvars, c_probs = model(x=xs, temp=self.temp, prior_c=0.0, eval=True)
                ls = model.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0)
                tot.append(ls[0])
                for k, v in (("c_prob", c_probs), ("qc", cs), ("c_smp", c_smps), ("s_mean", s_means),
                             ("s_logvar", s_logvars), ("s_smp", s_smps), ("x_low", x_lows)):
                    outs[k].append(torch.stack(v))
                outs["labels"].append(model.argmax_labels(torch.stack(cs)))
        res = {k: torch.cat(v, dim=1).cpu().numpy() for k, v in outs.items()}
        res["total_loss"] = float(torch.stack(tot).mean().item()) if tot else float("nan")
        return res
