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
    def __init__(self, *tensors, batch_size, device, shuffle=True, drop_last=True, generator=None, reuse_buffers=False):
        if not tensors or any(t.shape[0] != tensors[0].shape[0] for t in tensors):
            raise ValueError("tensors must share their first dimension")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ResidentLoader keeps the data set on a CUDA device")
        self.dataset = _Tensors(t.to(self.device) for t in tensors)       # the one H2D copy of the data set
        self.batch_size, self.shuffle, self.drop_last, self.generator = int(batch_size), shuffle, drop_last, generator
        self.h2d_bytes_per_epoch = 0
        # reuse_buffers: full batches are gathered into two alternating sets of buffers, so the step sees the same device
        # pointers again and again (CUDA-graph replay); a batch is then only valid until the one after next is drawn
        self.reuse_buffers = bool(reuse_buffers)
        self._ring = None

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
        if self.reuse_buffers and self._ring is None:
            self._ring = [[torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                           for t in self.dataset.tensors] for _ in range(2)]
        for b in range(len(self)):
            sel = idx[b * B:(b + 1) * B]
            if self.reuse_buffers and sel.numel() == B:      # (producer and consumer share one stream)
                bufs = self._ring[b & 1]
                yield tuple(torch.index_select(t, 0, sel, out=o) for t, o in zip(self.dataset.tensors, bufs))
            else:
                yield tuple(t.index_select(0, sel) for t in self.dataset.tensors)


class PackedBatch:
    """A host batch in row-packed form (bitmap + non-zero values + row offsets) in pinned memory.

    Expression matrices are sparse (Smart-seq-shaped data ~35 % non-zeros, 10x-shaped ~8 %), and the host->device copy
    of the dense fp32 batch (`x.to(rank)`, cpl_mixvae.py:416: 100.6 MB at B=5000, D=5032) is what bounds an end-to-end
    step on a B200 (PCIe ~57 GB/s vs a 0.6 ms device step).  Packing is lossless: ``unpack`` restores the dense matrix bit
    for bit (``mvae_unpack_rows``; zeros come back as +0.0).  ``HostBatchFeeder`` accepts PackedBatch items and expands
    them on its copy stream into the staging buffer the step reads.
    """

    def __init__(self, x: torch.Tensor, pin: bool = True):
        if x.dim() != 2 or x.dtype != torch.float32 or x.device.type != "cpu":
            raise ValueError("PackedBatch packs a CPU fp32 matrix [cells, genes]")
        B, D = x.shape
        W = (D + 31) // 32
        nz = x != 0
        pad = W * 32 - D
        bits = torch.nn.functional.pad(nz, (0, pad)).view(B, W, 32).to(torch.int64)
        words = (bits << torch.arange(32, dtype=torch.int64)).sum(-1)               # bit j of word w = gene 32 w + j
        bitmap = words.to(torch.uint32).view(torch.int32).contiguous()
        values = x[nz].contiguous()                                                 # row-major order of the non-zeros
        rp = torch.zeros(B + 1, dtype=torch.int64)
        rp[1:] = nz.sum(1).cumsum(0)
        # ONE contiguous (pinned) buffer [row_ptr | bitmap | values] -> one host->device copy per batch
        n_rp, n_bm, n_va = rp.numel() * 8, bitmap.numel() * 4, values.numel() * 4
        self._off = (0, n_rp, n_rp + n_bm, n_rp + n_bm + n_va)
        buf = torch.empty(self._off[3], dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            buf = buf.pin_memory()
        buf[:n_rp].view(torch.int64).copy_(rp)
        buf[n_rp:n_rp + n_bm].view(torch.int32).copy_(bitmap.reshape(-1))
        buf[n_rp + n_bm:].view(torch.float32).copy_(values)
        self.buffer = buf
        self.shape = (B, D)
        self.words = W

    def views(self, buf):
        """(bitmap [B, W] int32, values [nnz] fp32, row_ptr [B + 1] int64) as views of a byte buffer laid out like ours."""
        o = self._off
        B = self.shape[0]
        return (buf[o[1]:o[2]].view(torch.int32).view(B, self.words), buf[o[2]:o[3]].view(torch.float32),
                buf[o[0]:o[1]].view(torch.int64))

    bitmap = property(lambda self: self.views(self.buffer)[0])
    values = property(lambda self: self.views(self.buffer)[1])
    row_ptr = property(lambda self: self.views(self.buffer)[2])

    @property
    def nbytes(self) -> int:
        """Bytes that cross PCIe for this batch."""
        return self.buffer.numel()

    def expand_from(self, dbuf: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Expand a device copy ``dbuf`` of ``self.buffer`` into the dense matrix ``out`` [B, D] on the current stream."""
        import ctypes as C
        from . import _lib
        B, D = self.shape
        bm, va, rp = self.views(dbuf)
        stream = torch.cuda.current_stream(out.device).cuda_stream
        _lib.check(_lib.load().mvae_unpack_rows(bm.data_ptr(), va.data_ptr() if va.numel() else None, rp.data_ptr(), B, D,
                                                out.data_ptr(), out.stride(0), C.c_void_p(stream)), "mvae_unpack_rows")
        return out

    def unpack(self, device, out: torch.Tensor = None, staging: torch.Tensor = None) -> torch.Tensor:
        """Copy the packed buffer to ``device`` (one async copy on the current stream) and expand it into ``out`` [B, D]
        (allocated when None).  ``staging``: optional device byte buffer (>= nbytes, 16-byte aligned) to reuse."""
        B, D = self.shape
        device = torch.device(device)
        if staging is None:
            dbuf = self.buffer.to(device, non_blocking=True)
        else:
            dbuf = staging[:self.nbytes]
            dbuf.copy_(self.buffer, non_blocking=True)
        if out is None:
            out = torch.empty(B, D, dtype=torch.float32, device=device)
        self.expand_from(dbuf, out)
        if staging is None:
            dbuf.record_stream(torch.cuda.current_stream(device))
        return out
