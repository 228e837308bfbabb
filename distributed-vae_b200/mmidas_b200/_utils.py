"""Consensus helpers used by the trainer's epoch bookkeeping (mirror of the functions the reference's
train loop imports from mmidas/_utils.py:64-128).  Host-side numpy; the only device-side piece is
``mixVAE_model.argmax_labels`` which replaces ``classify`` on the per-step hot path."""
from __future__ import annotations

import numpy as np
import torch


def to_np(x):
    """mmidas/_utils.py:64"""
    return x.cpu().detach().numpy()


def classify(probs):
    """mmidas/_utils.py:78 — argmax over the category axis."""
    return np.argmax(probs, axis=-1)


def compute_confmat(labels1, labels2, K=None):
    """mmidas/_utils.py:83 — K x K co-assignment counts of two label vectors."""
    labels1 = np.asarray(labels1)
    labels2 = np.asarray(labels2)
    if len(labels1) != len(labels2):
        raise ValueError("label vectors differ in length")
    if labels1.ndim != 1 or labels2.ndim != 1:
        raise ValueError("labels must be 1-d")
    labels1 = labels1.astype(np.int64)
    labels2 = labels2.astype(np.int64)
    if K is None:
        K = max(len(np.unique(labels1)), len(np.unique(labels2)))
    flat = np.bincount(labels1 * K + labels2, minlength=K * K)
    return flat.reshape(K, K).astype(np.float64)


def confmat_normalize(cm):
    """mmidas/_utils.py:96 — divide column k by max(row-sum_k, col-sum_k); empty categories give 0."""
    cm = np.asarray(cm, dtype=np.float64)
    maxes = np.maximum(cm.sum(axis=0), cm.sum(axis=1))
    out = np.zeros_like(cm)
    np.divide(cm, maxes, out=out, where=maxes != 0)
    return out


def confmat_mean(cm):
    """mmidas/_utils.py:127 — mean of the diagonal."""
    return float(np.mean(np.diag(cm)))


def consensus(labels, n_categories):
    """Mean over arm pairs of confmat_mean(confmat_normalize(confmat(a, b))) — the quantity the
    reference logs as aug-cns / train-cns / val-cns (cpl_mixvae.py:512-523)."""
    A = len(labels)
    vals = []
    for a in range(A):
        for b in range(a + 1, A):
            vals.append(confmat_mean(confmat_normalize(compute_confmat(labels[a], labels[b], n_categories))))
    return float(np.mean(vals)) if vals else float("nan")


def confmat_device(labels: torch.Tensor, n_categories: int, counts: torch.Tensor = None) -> torch.Tensor:
    """Device-side ``compute_confmat`` for every arm pair (mmidas/_utils.py:83): ``labels`` int32 [A, n] on the GPU
    (``mixVAE_model.argmax_labels``) -> int32 counts [n_pairs, K, K], pairs (a < b) in order.  ``counts`` accumulates
    across calls (one call per batch; the labels of an epoch never leave the device)."""
    import ctypes as C
    from . import _lib
    if labels.dtype != torch.int32 or labels.dim() != 2 or not labels.is_cuda:
        raise ValueError("labels must be a CUDA int32 tensor [n_arm, n_cells]")
    A, n = labels.shape
    if A < 2:
        raise ValueError("the consensus needs at least two arms")
    n_pairs = A * (A - 1) // 2
    if counts is None:
        counts = torch.zeros(n_pairs, n_categories, n_categories, dtype=torch.int32, device=labels.device)
    labels = labels.contiguous()
    stream = torch.cuda.current_stream(labels.device).cuda_stream
    _lib.check(_lib.load().mvae_confmat(labels.data_ptr(), n, A, n_categories, counts.data_ptr(), C.c_void_p(stream)),
               "mvae_confmat")
    return counts


def consensus_from_counts(counts) -> float:
    """Mean over arm pairs of confmat_mean(confmat_normalize(cm)) from accumulated counts [n_pairs, K, K]."""
    cms = counts.cpu().numpy() if torch.is_tensor(counts) else np.asarray(counts)
    vals = [confmat_mean(confmat_normalize(cm.astype(np.float64))) for cm in cms]
    return float(np.mean(vals)) if vals else float("nan")


def ecdf(labels):
    """mmidas/_utils.py:280: empirical distribution of integer labels."""
    labels = np.asarray(labels)
    assert len(labels.shape) == 1
    return np.bincount(labels) / len(labels)


def set_seeds(seed):
    """mmidas/_utils.py:34"""
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
