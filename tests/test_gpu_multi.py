"""The multi-GPU step on real hardware: needs >= 2 B200s (skipped otherwise; `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`).  Each test spawns one process per GPU over NCCL and compares ShardedTrainer.step with
the CPU oracle emulating the ranks: the arm axis is exactly the single-process reference; the dp axis
follows the reference's FSDP semantics (local BatchNorm / inv_var statistics per replica, averaged gradients)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import mixvae_oracle as O
from gpu_utils import rel_l2

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]

HP = dict(input_dim=520, n_categories=40, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0)
B_LOCAL, STEPS = 384, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(world, mode):
    hp = O.HP(**HP)
    gen = torch.Generator().manual_seed(546)
    n_rep = 1 if mode == "arm" else world
    xs = [O.synth_x(B_LOCAL, hp.input_dim, gen) for _ in range(n_rep)]
    noises = [[O.synth_noise(hp, B_LOCAL, gen) for _ in range(n_rep)] for _ in range(STEPS)]
    return hp, xs, noises


def _worker(rank, world, port, mode, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "distributed-vae_b200"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from mmidas_b200 import _dist_utils as D
    from mmidas_b200.parallel import ShardedTrainer
    D.init_dist_env(rank, world, "127.0.0.1", str(port))
    try:
        hp, xs, noises = _inputs(world, mode)
        kw = dict(input_dim=hp.input_dim, fc_dim=hp.fc_dim, n_categories=hp.n_categories, state_dim=hp.state_dim,
                  lowD_dim=hp.lowD_dim, x_drop=hp.x_drop, s_drop=hp.s_drop, n_arm=hp.n_arm, lam=hp.lam, lam_pc=1, tau=hp.tau,
                  beta=hp.beta, hard=hp.hard, variational=True, device="cuda", eps=hp.eps, momentum=hp.momentum,
                  ref_prior=False, loss_mode="MSE", precision="fp32_simt")
        st = ShardedTrainer(kw, lr=hp.lr, mode=mode, seed=546, use_cuda_graph=False)
        a0, a1 = st.plan.local_arms(rank)
        rep = st.dp_coord
        x = xs[rep].cuda()
        losses = []
        for s in range(STEPS):
            nz = {k: v[a0:a1].cuda() for k, v in noises[s][rep].items()}
            losses.append(st.step(x, noise=nz).cpu().numpy().copy())
        torch.cuda.synchronize()
        msd, osd = st.full_state_dicts()
        res = {"losses": losses, "sd": {k: v.numpy() for k, v in msd.items()},
               "m": {i: s["exp_avg"].cpu().numpy() for i, s in osd["state"].items()}}
        # graph-replayed steps with in-kernel noise == the same steps launched eagerly (fresh trainers, same seeds)
        fin = []
        for use_graph in (False, True):
            torch.manual_seed(77)
            st2 = ShardedTrainer(dict(kw, precision="tf32x3_fc1"), lr=hp.lr, mode=mode, seed=546, use_cuda_graph=use_graph)
            torch.manual_seed(77)
            bufs = [x.clone(), x.clone()]
            tot = [st2.step(bufs[i % 2])[0].item() for i in range(6)]
            torch.cuda.synchronize()
            fin.append((tot, st2.model.flat_parameters().cpu().numpy().copy(), len(st2._graphs)))
        res["graph"] = fin
        out[rank] = res
    finally:
        D.destroy_dist_env()


def _oracle(world, mode):
    """Single-process emulation: arm mesh == the plain reference step; dp mesh == per-replica steps on local
    statistics, gradients averaged, one Adam."""
    hp, xs, noises = _inputs(world, mode)
    st = O.TrainState(hp, O.init_state_dict(hp, 546))
    names = O.param_names(hp)
    losses = []
    for s in range(STEPS):
        if mode == "arm":
            r = O.train_step(st, [xs[0]] * hp.n_arm, noises[s][0])
            losses.append([float(r["loss"]["total"])])
            continue
        grads, ls, bufs = [], [], None
        for rep in range(world):
            tmp = O.TrainState(hp, O.cast_state_dict(st.sd, torch.float32))
            r = O.train_step(tmp, [xs[rep]] * hp.n_arm, noises[s][rep], return_grads=True)
            grads.append(r["grads"])
            ls.append(float(r["loss"]["total"]))
            if rep == 0:
                bufs = {k: v for k, v in tmp.sd.items() if "batch_" in k}
        st.step += 1
        with torch.no_grad():
            for n in names:
                g = sum(gr[n] for gr in grads) / world
                if n not in st.m:
                    st.m[n] = torch.zeros_like(st.sd[n])
                    st.v[n] = torch.zeros_like(st.sd[n])
                O.adam_update(st.sd[n], g, st.m[n], st.v[n], st.step, hp.lr, hp.betas, hp.adam_eps)
        st.sd.update(bufs)            # rank 0's running statistics are the ones saved
        losses.append(ls)
    return hp, st, losses


@pytest.mark.parametrize("mode", ["arm", "dp"])
def test_sharded_step_matches_rank_emulating_oracle(mode):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    hp, st, want = _oracle(world, mode)
    names = O.param_names(hp)
    for rank in (0, 1):
        res = out[rank]
        for s in range(STEPS):
            ref = want[s][0] if mode == "arm" else want[s][rank]
            tol = 1e-4 if s == 0 else 5e-3
            assert abs(res["losses"][s][0] / ref - 1) < tol, (rank, s, res["losses"][s][0], ref)
        for i, n in enumerate(names):
            assert rel_l2(res["m"][i], st.m[n].numpy()) < 2e-2, (rank, n, rel_l2(res["m"][i], st.m[n].numpy()))
        for k, v in st.sd.items():
            got = res["sd"][k]
            if v.is_floating_point() and "running" not in k:
                d = np.abs(got.astype(np.float64) - v.numpy())
                assert d.max() <= 2 * hp.lr * STEPS + 1e-6, (k, d.max())
                assert (d > 1e-5).mean() <= 0.05, (k, (d > 1e-5).mean())
            elif "running" in k and (mode == "arm" or rank == 0):
                np.testing.assert_allclose(got, v.numpy(), rtol=1e-3, atol=1e-5, err_msg=k)
    # replicas / arm ranks agree on the full parameters after the steps
    for k in out[0]["sd"]:
        if "running" not in k and "num_batches" not in k:
            np.testing.assert_array_equal(out[0]["sd"][k], out[1]["sd"][k])
    # CUDA-graph replay of the whole sharded step (incl. the NCCL collectives) == eager launches
    for rank in (0, 1):
        (t0, p0, g0), (t1, p1, g1) = out[rank]["graph"]
        assert g0 == 0 and g1 == 2
        np.testing.assert_allclose(t1, t0, rtol=1e-5)
        d = np.abs(p0 - p1)
        assert (d > 1e-6).mean() < 1e-2


def _train_worker(rank, world, port, mode, folder, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "distributed-vae_b200"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from torch.utils.data import DataLoader, TensorDataset
    from mmidas_b200 import _dist_utils as D
    from mmidas_b200.cpl_mixvae import cpl_mixVAE
    D.init_dist_env(rank, world, "127.0.0.1", str(port))
    try:
        gen = torch.Generator().manual_seed(546)
        x = O.synth_x(384, 128, gen)
        idx = torch.arange(384, dtype=torch.float32)
        # the reference's loop gives every rank its own loader; in the arm mesh both ranks must see the same cells
        lo = 0 if mode == "arm" else rank * 160
        torch.manual_seed(1 + (0 if mode == "arm" else rank))
        train = DataLoader(TensorDataset(x[lo:lo + 160], idx[lo:lo + 160]), batch_size=80, shuffle=True, drop_last=True)
        test = DataLoader(TensorDataset(x[320:], idx[320:]), batch_size=1)
        torch.manual_seed(546)
        t = cpl_mixVAE(saving_folder=folder, aug_file="", device=f"cuda:{rank}")
        t.mesh_mode = mode
        t.init_model(n_categories=9, state_dim=2, input_dim=128, x_drop=0.5, s_drop=0.0, n_arm=2, lr=1e-3)
        p0 = t.model.flat_parameters().clone()
        res = t.train(train, test, n_epoch=2, n_epoch_p=0, rank=rank, ws=world, good_enuf_consensus=2.0)
        torch.cuda.synchronize()
        out[rank] = {"losses": res["losses"], "cns": res["consensus_train"] + res["consensus_val"] + res["consensus_aug"],
                     "moved": float((t.model.flat_parameters() - p0).abs().max()), "p": t.model.flat_parameters().cpu().numpy()}
    finally:
        D.destroy_dist_env()


@pytest.mark.parametrize("mode", ["dp", "arm"])
def test_train_loop_on_two_gpus(mode, tmp_path):
    """cpl_mixVAE.train(ws=2): the reference's epoch loop on the mesh (the reference raises for ws > 1, train.py:274):
    epoch all-reduces, eval passes, one reference-layout checkpoint written by rank 0."""
    import glob
    world = 2
    folder = str(tmp_path / "run")
    os.makedirs(folder + "/model", exist_ok=True)
    out = mp.Manager().dict()
    mp.spawn(_train_worker, args=(world, _free_port(), mode, folder, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    for r in (0, 1):
        assert len(out[r]["losses"]) == 2 and np.all(np.isfinite(out[r]["losses"])) and out[r]["moved"] > 0
        assert all(0.0 <= c <= 1.0 for c in out[r]["cns"])
    np.testing.assert_allclose(out[0]["losses"], out[1]["losses"], rtol=1e-6)       # all-reduced epoch means
    np.testing.assert_array_equal(out[0]["p"], out[1]["p"])                          # same full model on every rank
    files = sorted(glob.glob(folder + "/model/*.pth"))
    assert len(files) == 2                                                           # cns_... and final, written once
    ck = torch.load(files[0], map_location="cpu")
    sd, osd = ck["model_state_dict"], ck["optimizer_state_dict"]
    assert len(sd) == 92 and list(sd)[:4] == ["fc1.0.weight", "fc1.0.bias", "fc1.1.weight", "fc1.1.bias"]
    assert len(osd["state"]) == 56 and float(osd["state"][0]["step"]) == 4.0
    ref_params = [torch.nn.Parameter(torch.zeros_like(osd["state"][i]["exp_avg"])) for i in range(56)]
    torch.optim.Adam(ref_params, lr=1e-3).load_state_dict(osd)
