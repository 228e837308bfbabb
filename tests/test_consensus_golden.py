"""Known-answer tests for the consensus helpers: the vectors are the ones the REFERENCE's own tests hold
(`/root/reference/tests/test_utils.py:18-113`, expecttest snapshots of compute_confmat / confmat_normalize /
confmat_mean / classify), restated here as arrays.  CPU part: the numpy mirror `mmidas_b200._utils`.  GPU part: the
device kernels behind `argmax_labels` / `confmat_device` must reproduce them bit for bit."""
import numpy as np
import pytest
import torch

from mmidas_b200._utils import classify, compute_confmat, confmat_mean, confmat_normalize, consensus, ecdf

L1, L2 = np.array([1, 0, 2, 3, 0, 3]), np.array([1, 0, 2, 3, 1, 3])
CM_ID = np.eye(4)
CM_12 = np.array([[1, 1, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 2]], dtype=float)     # test_utils.py:27-37
CMN_12 = np.array([[0.5, 0.5, 0, 0], [0, 0.5, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])            # test_utils.py:66-80
P3 = np.array([[0.7, 0.2, 0.1], [0.4, 0.1, 0.5], [0.3, 0.6, 0.1], [0.1, 0.1, 0.8]])         # test_utils.py:106-110


def test_reference_goldens_numpy_mirror():
    np.testing.assert_array_equal(compute_confmat(np.array([1, 0, 2, 3]), np.array([1, 0, 2, 3])), CM_ID)
    np.testing.assert_array_equal(compute_confmat(L1, L2), CM_12)
    np.testing.assert_array_equal(confmat_normalize(CM_ID), CM_ID)
    np.testing.assert_array_equal(confmat_normalize(compute_confmat(L1, L2)), CMN_12)
    assert confmat_mean(CM_ID) == 1.0
    assert confmat_mean(compute_confmat(L1, L2)) == 1.25                                     # test_utils.py:94-103
    np.testing.assert_array_equal(classify(P3[:3]), [0, 2, 1])
    np.testing.assert_array_equal(classify(P3), [0, 2, 1, 2])


def test_ecdf_reference_goldens():
    # test_utils.py:39-49
    np.testing.assert_array_equal(ecdf(np.array([1, 0, 2, 3])), [0.25, 0.25, 0.25, 0.25])
    np.testing.assert_allclose(ecdf(np.array([1, 0, 2, 3, 0, 3])), [1 / 3, 1 / 6, 1 / 6, 1 / 3], rtol=0, atol=1e-15)


def test_vectorised_equals_naive_loop():
    # the identity the reference checks in test_confmat_vectorize_correctness (test_utils.py:113-124)
    rng = np.random.default_rng(0)
    K = 10
    a, b = rng.integers(0, K, 100), rng.integers(0, K, 100)
    naive = np.zeros((K, K))
    for i in range(100):
        naive[a[i], b[i]] += 1
    np.testing.assert_array_equal(compute_confmat(a, b, K), naive)
    maxes = np.array([max(naive[k, :].sum(), naive[:, k].sum()) for k in range(K)])
    want = np.divide(naive, maxes, out=np.zeros_like(naive), where=maxes != 0)
    np.testing.assert_allclose(confmat_normalize(compute_confmat(a, b, K)), want)


@pytest.mark.gpu
def test_device_confmat_and_argmax_match_reference_goldens():
    from mmidas_b200._utils import confmat_device, consensus_from_counts
    from mmidas_b200 import _lib
    import ctypes as C
    dev = torch.device("cuda", 0)
    lab = torch.tensor(np.stack([L1, L2]), dtype=torch.int32, device=dev)
    cm = confmat_device(lab, 4)
    np.testing.assert_array_equal(cm[0].cpu().numpy(), CM_12.astype(np.int32))
    assert consensus_from_counts(cm) == confmat_mean(CMN_12)
    # device classify on the reference's vector
    q = torch.tensor(P3, dtype=torch.float32, device=dev)
    out = torch.empty(4, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().mvae_argmax(q.data_ptr(), out.data_ptr(), 4, 3, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "mvae_argmax")
    np.testing.assert_array_equal(out.cpu().numpy(), [0, 2, 1, 2])
    # random labels at the reference's sizes (K = 92, 5000 cells, test_utils.py:127-130), 3 arms, accumulated over 2 batches
    rng = np.random.default_rng(1)
    K, n, A = 92, 5000, 3
    labs = [rng.integers(0, K, (A, n)) for _ in range(2)]
    counts = None
    for l in labs:
        counts = confmat_device(torch.tensor(l, dtype=torch.int32, device=dev), K, counts)
    allab = np.concatenate(labs, axis=1)
    pair = 0
    for a in range(A):
        for b in range(a + 1, A):
            np.testing.assert_array_equal(counts[pair].cpu().numpy(), compute_confmat(allab[a], allab[b], K).astype(np.int32))
            pair += 1
    assert abs(consensus_from_counts(counts) - consensus([allab[a] for a in range(A)], K)) < 1e-12
