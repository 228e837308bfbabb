"""Trainer mirror (mmidas_b200.cpl_mixvae.cpl_mixVAE) on the GPU: the epoch loop, logged names,
checkpoint files in the reference's layout, resume, and the sharded trainer's single-process path."""
import glob
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from oracle import mixvae_oracle as O

pytestmark = pytest.mark.gpu


class _Run:
    def __init__(self):
        self.logs = []

    def log(self, d):
        self.logs.append(d)


def _loaders(N=384, D=128, B=128):
    gen = torch.Generator().manual_seed(546)
    x = O.synth_x(N, D, gen)
    idx = torch.arange(N, dtype=torch.float32)
    train = DataLoader(TensorDataset(x[:320], idx[:320]), batch_size=B, shuffle=True, drop_last=True)
    test = DataLoader(TensorDataset(x[320:], idx[320:]), batch_size=1)     # reference: test batch_size == 1
    return train, test


def test_train_loop_checkpoints_and_resume(tmp_path):
    from mmidas_b200.cpl_mixvae import cpl_mixVAE
    train, test = _loaders()
    folder = str(tmp_path / "run")
    os.makedirs(folder + "/model", exist_ok=True)
    torch.manual_seed(546)
    t = cpl_mixVAE(saving_folder=folder, aug_file="", device="cuda")
    t.init_model(n_categories=9, state_dim=2, input_dim=128, x_drop=0.5, s_drop=0.0, n_arm=2, lr=1e-3)
    run = _Run()
    out = t.train(train, test, n_epoch=3, n_epoch_p=0, run=run, rank="cuda", good_enuf_consensus=2.0)
    assert len(out["losses"]) == 3 and all(np.isfinite(out["losses"]))
    assert all(0.0 <= c <= 1.0 for c in out["consensus_train"] + out["consensus_val"] + out["consensus_aug"])
    keys = set().union(*[set(d) for d in run.logs])
    for k in ("train/total-loss", "train/joint-loss", "train/negative-joint-entropy", "train/simplex-distance",
              "train/l2-distance", "train/time", "train/mem", "train/consensus_aug", "train/rec-loss0", "train/rec-loss1",
              "train/consensus", "val/total-loss", "val/rec-loss", "val/consensus"):
        assert k in keys, k                                     # names of cpl_mixvae.py:541-560, :658-663, :768-775
    files = sorted(glob.glob(folder + "/model/*.pth"))
    assert any("cns_cpl_mixVAE_model_before_pruning_A2_" in f for f in files)
    assert any(os.path.basename(f).startswith("cpl_mixVAE_model_before_pruning_A2_") for f in files)
    ck = torch.load(files[0], map_location="cpu")
    assert set(ck) == {"model_state_dict", "optimizer_state_dict"}
    sd = ck["model_state_dict"]
    assert len(sd) == 46 * 2 and sd["fc1.0.weight"].shape == (100, 128) and sd["batch_l5.1.running_var"].shape == (10,)
    assert int(sd["batch_l1.0.num_batches_tracked"]) == 3 * 2 and int(sd["batch_s.0.num_batches_tracked"]) == 0
    osd = ck["optimizer_state_dict"]
    assert len(osd["state"]) == 56 and set(osd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    # a torch.optim.Adam built on an equally shaped reference-style model accepts the optimizer state
    ref_params = [torch.nn.Parameter(torch.zeros_like(osd["state"][i]["exp_avg"])) for i in range(56)]
    torch.optim.Adam(ref_params, lr=1e-3).load_state_dict(osd)
    # resume: trained_model= loads model + optimizer and (like the reference, :397) skips the loop
    t2 = cpl_mixVAE(saving_folder=folder, aug_file="", device="cuda")
    t2.init_model(n_categories=9, state_dim=2, input_dim=128, x_drop=0.5, s_drop=0.0, n_arm=2, trained_model=files[0])
    assert t2.init is False and t2.optimizer.step_count == 6
    for (k, a), (_, b) in zip(t.model.state_dict().items(), t2.model.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k
    res = t2.eval_model(test)        # the reference's dictionary (cpl_mixvae.py:1590-1619)
    assert res["z_prob"].shape == (2, 64, 9) and res["predicted_label"].shape == (2, 64) and np.isfinite(res["total_loss"])
    assert res["predicted_label"].min() >= 1 and res["predicted_label"].max() <= 9 and res["recon_c"].shape == (2, 64, 128)
    assert np.array_equal(res["data_indx"], np.arange(320, 384)) and 0.0 <= res["cnss"] <= 1.0


def test_caller_replaced_optimizer_still_trains():
    """train.py:144-147 re-creates the optimizer with torch.optim.Adam(model.parameters()): parameters and
    .grad are views of the flat buffers, so the stock optimizer drives the same kernels."""
    from mmidas_b200.cpl_mixvae import cpl_mixVAE
    torch.manual_seed(546)
    t = cpl_mixVAE(saving_folder="", aug_file="", device="cuda", save_flag=False)
    t.init_model(n_categories=9, state_dim=2, input_dim=128, x_drop=0.5, s_drop=0.0, n_arm=2)
    t.optimizer = torch.optim.Adam(t.model.parameters(), lr=1e-3)
    t.model.train()
    gen = torch.Generator().manual_seed(1)
    x = O.synth_x(128, 128, gen).cuda()
    p0 = t.model.flat_parameters().clone()
    l0 = t.train_batch(x)[0].item()
    for _ in range(5):
        lv = t.train_batch(x)
    assert np.isfinite(lv[0].item()) and not torch.equal(p0, t.model.flat_parameters())
    assert t.model.fc1[0].weight.data_ptr() == t.model.flat_parameters().data_ptr()


def test_sharded_trainer_single_process_matches_fused_step():
    """ShardedTrainer on a one-rank world is the fused step (eager and graph-replayed): same losses and parameters as
    cpl_mixVAE.train_batch from the same seed."""
    import torch.distributed as dist
    from mmidas_b200 import _dist_utils as D
    from mmidas_b200.cpl_mixvae import cpl_mixVAE
    from mmidas_b200.parallel import ShardedTrainer
    kw = dict(input_dim=128, fc_dim=100, n_categories=9, state_dim=2, lowD_dim=10, x_drop=0.5, s_drop=0.0, n_arm=2, lam=1,
              lam_pc=1, tau=0.005, beta=1.0, hard=False, variational=True, device="cuda", eps=1e-8, momentum=0.01,
              ref_prior=False, loss_mode="MSE")
    gen = torch.Generator().manual_seed(1)
    xs = [O.synth_x(128, 128, gen).cuda() for _ in range(2)]
    D.init_dist_env(0, 1, "127.0.0.1", str(D.find_port("127.0.0.1")), backend="nccl")
    try:
        torch.manual_seed(546)
        st = ShardedTrainer(kw, lr=1e-3, mode="auto", seed=546)
        assert st.plan.arm_ranks == 1 and st.plan.dp_ranks == 1 and st.active
        torch.manual_seed(546)                      # noise seed of the steps (the constructor re-seeded the device generator)
        got = [st.step(xs[i % 2])[0].item() for i in range(6)]
        assert len(st._graphs) == 2                 # two input buffers -> two captured graphs
        p_sharded = st.model.flat_parameters().clone()
    finally:
        D.destroy_dist_env()
    torch.manual_seed(546)
    t = cpl_mixVAE(saving_folder="", aug_file="", device="cuda", save_flag=False)
    t.use_cuda_graph = False
    t.init_model(n_categories=9, state_dim=2, input_dim=128, x_drop=0.5, s_drop=0.0, n_arm=2)
    t.model.train()
    torch.manual_seed(546)
    want = [t.train_batch(xs[i % 2])[0].item() for i in range(6)]
    np.testing.assert_allclose(got, want, rtol=1e-5)
    d = (p_sharded - t.model.flat_parameters()).abs()
    assert float((d > 1e-6).float().mean()) < 1e-2


def test_training_step_on_augmented_input_matches_oracle(tmp_path):
    """aug_file set (cpl_mixvae.py:182-186, :422-423): the step trains on netA(x.expand(A,-1,-1), True, 0.1)[1].  With the
    augmenter's and the VAE's noise injected, the loss equals the oracle's loss on the oracle-augmented cells."""
    from mmidas_b200.cpl_mixvae import cpl_mixVAE
    from oracle import augmenter_oracle as AO
    D, B, A, C = 260, 96, 2, 9
    # mk_augmenter builds Augmenter_smartseq with its default n_dim = 500 (cpl_mixvae.py:135-139)
    sd_aug = AO.random_state_dict(50, 10, D, 500, seed=21)
    path = str(tmp_path / "aug.pth")
    torch.save({"parameters": {"num_n": 50, "num_z": 10, "n_features": D}, "netA": sd_aug}, path)
    hp = O.HP(input_dim=D, n_categories=C, state_dim=2, n_arm=A, x_drop=0.0, s_drop=0.0)
    t = cpl_mixVAE(saving_folder="", aug_file=path, device="cuda", save_flag=False)
    assert t.aug_param["n_features"] == D and t.netA is not None and not t.netA.training
    t.precision = "fp32_simt"
    t.init_model(n_categories=C, state_dim=2, input_dim=D, x_drop=0.0, s_drop=0.0, n_arm=A, lr=1e-3)
    sd0 = O.init_state_dict(hp, 546)
    t.model.load_state_dict(sd0)
    t.model.train()
    gen = torch.Generator().manual_seed(4)
    x = O.synth_x(B, D, gen)
    z, eps = torch.randn(A, B, 50, generator=gen), torch.randn(A, B, 10, generator=gen)
    noise = O.synth_noise(hp, B, gen)
    lv = t.train_batch(x.cuda(), noise={k: v.cuda() for k, v in noise.items()}, aug_noise={"z": z.cuda(), "eps": eps.cuda()})
    torch.cuda.synchronize()
    _, xa = AO.forward(sd_aug, x.expand(A, -1, -1), z, eps, 0.1)
    st = O.TrainState(hp, O.cast_state_dict(sd0, torch.float32))
    ref = O.train_step(st, [xa[a] for a in range(A)], noise)
    assert abs(lv[0].item() / float(ref["loss"]["total"]) - 1) < 1e-4, (lv[0].item(), float(ref["loss"]["total"]))
