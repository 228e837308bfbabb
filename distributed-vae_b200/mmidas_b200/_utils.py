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


def set_seeds(seed):
    """mmidas/_utils.py:34"""
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
