"""Device-resident replacement for the training DataLoader of the reference.

The reference feeds the step from ``DataLoader(TensorDataset(x, idx), batch_size, shuffle=True, drop_last=True, pin_memory=True,
num_workers=2)`` (mmidas/utils/dataloader.py:123-131) and copies every batch to the GPU (cpl_mixvae.py:415-417): 100 MB per
step at the reference sizes, which makes the B200 step PCIe-bound (bench.py ``e2e``).  The whole Smart-seq matrix is 450 MB:
``ResidentLoader`` keeps the tensors on the device once and only the shuffled indices are drawn per epoch.

The permutation stream is the DataLoader's own: per epoch ``RandomSampler`` seeds a fresh generator from the global torch RNG
and calls ``torch.randperm`` (torch/utils/data/sampler.py:160-183), after the iterator has drawn its ``_base_seed`` from the
same RNG (torch/utils/data/dataloader.py, ``_BaseDataLoaderIter.__init__``).  ``ResidentLoader`` makes the same two draws in
the same order, so under the same ``torch.manual_seed`` it yields the same cells in the same batches as the reference loader
(tests/test_dataloader.py checks this against a real DataLoader).
"""
from __future__ import annotations

import torch


class _Tensors:
    """Stands in for ``TensorDataset`` where callers read ``loader.dataset.tensors`` (the reference's "batch_size == 1 means
    the whole set" convention, cpl_mixvae.py:722-748)."""

    def __init__(self, tensors):
        self.tensors = tuple(tensors)

    def __len__(self):
        return self.tensors[0].shape[0]

    def __getitem__(self, i):
        return tuple(t[i] for t in self.tensors)


class ResidentLoader:
    def __init__(self, *tensors, batch_size, device, shuffle=True, drop_last=True, generator=None):
        if not tensors or any(t.shape[0] != tensors[0].shape[0] for t in tensors):
            raise ValueError("tensors must share their first dimension")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ResidentLoader keeps the data set on a CUDA device")
        self.dataset = _Tensors(t.to(self.device) for t in tensors)       # the one H2D copy of the data set
        self.batch_size, self.shuffle, self.drop_last, self.generator = int(batch_size), shuffle, drop_last, generator
        self.h2d_bytes_per_epoch = 0

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    @staticmethod
    def epoch_permutation(n, shuffle=True, generator=None):
        """The index order a ``DataLoader(shuffle=shuffle)`` iterator would use for its next epoch (CPU int64 tensor)."""
        # _BaseDataLoaderIter.__init__: base seed for the workers, drawn even with num_workers == 0
        torch.empty((), dtype=torch.int64).random_(generator=generator)
        if not shuffle:
            return torch.arange(n)
        if generator is None:                                             # RandomSampler.__iter__
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
        else:
            g = generator
        return torch.randperm(n, generator=g)

    def __iter__(self):
        n, B = len(self.dataset), self.batch_size
        perm = self.epoch_permutation(n, self.shuffle, self.generator)
        idx = perm.to(self.device, non_blocking=True)                      # 8 bytes per cell per epoch
        self.h2d_bytes_per_epoch = idx.numel() * idx.element_size()
        for b in range(len(self)):
            sel = idx[b * B:(b + 1) * B]
            yield tuple(t.index_select(0, sel) for t in self.dataset.tensors)
