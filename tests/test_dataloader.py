"""ResidentLoader draws the reference DataLoader's permutation stream: same cells, same batches, same epochs
under the same torch.manual_seed (CPU part); the device gather returns those rows (GPU part)."""
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from mmidas_b200.dataloader import ResidentLoader


@pytest.mark.parametrize("n,B", [(320, 128), (1000, 250), (77, 10)])
def test_permutation_stream_equals_dataloader(n, B):
    x = torch.arange(n, dtype=torch.float32).unsqueeze(1).repeat(1, 3)
    idx = torch.arange(n, dtype=torch.float32)
    torch.manual_seed(546)
    ref = DataLoader(TensorDataset(x, idx), batch_size=B, shuffle=True, drop_last=True)      # mmidas/utils/dataloader.py:123
    want = [[b[1].long() for b in ref] for _ in range(3)]                                    # three epochs
    torch.manual_seed(546)
    got = []
    for _ in range(3):
        perm = ResidentLoader.epoch_permutation(n)
        got.append([perm[i * B:(i + 1) * B] for i in range(n // B)])
    assert len(want[0]) == n // B
    for e in range(3):
        for a, b in zip(want[e], got[e]):
            assert torch.equal(a, b)
    # an explicit generator (DataLoader(generator=g)) is honoured the same way
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    ref = DataLoader(TensorDataset(x, idx), batch_size=B, shuffle=True, drop_last=True, generator=g1)
    a = torch.cat([b[1].long() for b in ref])
    b = ResidentLoader.epoch_permutation(n, True, g2)[:(n // B) * B]
    assert torch.equal(a, b)


def test_cpu_device_is_refused():
    with pytest.raises(RuntimeError):
        ResidentLoader(torch.zeros(4, 2), batch_size=2, device="cpu")


@pytest.mark.gpu
def test_device_batches_are_the_dataloader_batches():
    n, B, D = 700, 128, 64
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, D, generator=g)
    idx = torch.arange(n, dtype=torch.float32)
    torch.manual_seed(3)
    ref = [(bx.clone(), bi.clone()) for bx, bi in DataLoader(TensorDataset(x, idx), batch_size=B, shuffle=True, drop_last=True)]
    torch.manual_seed(3)
    rl = ResidentLoader(x, idx, batch_size=B, device="cuda")
    assert len(rl) == len(ref) == n // B and rl.dataset.tensors[0].is_cuda
    got = list(rl)
    for (rx, ri), (gx, gi) in zip(ref, got):
        assert gx.is_cuda and torch.equal(gx.cpu(), rx) and torch.equal(gi.cpu(), ri)
    assert rl.h2d_bytes_per_epoch == n * 8


@pytest.mark.gpu
def test_trainer_runs_from_a_resident_loader(tmp_path):
    from mmidas_b200.cpl_mixvae import cpl_mixVAE
    from oracle import mixvae_oracle as O
    gen = torch.Generator().manual_seed(546)
    x = O.synth_x(384, 128, gen)
    idx = torch.arange(384, dtype=torch.float32)
    train = ResidentLoader(x[:320], idx[:320], batch_size=128, device="cuda")
    test = ResidentLoader(x[320:], idx[320:], batch_size=1, device="cuda", shuffle=False, drop_last=False)
    t = cpl_mixVAE(saving_folder=str(tmp_path), aug_file="", device="cuda", save_flag=False)
    t.init_model(n_categories=9, state_dim=2, input_dim=128, x_drop=0.5, s_drop=0.0, n_arm=2, lr=1e-3)
    out = t.train(train, test, n_epoch=2, n_epoch_p=0, rank="cuda", good_enuf_consensus=2.0)
    assert len(out["losses"]) == 2 and all(l == l for l in out["losses"])
