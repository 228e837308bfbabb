"""Host-side logic of the multi-GPU path (mmidas_b200/parallel.py), exercised with world_size-2
gloo process groups on the CPU.  The arithmetic of each rank is played by the CPU oracle (tests may use
it as the checker); what is under test is the plumbing: mesh planning, arm all-gather, loss-vector
fix-up, bucketed gradient averaging, and that arm-sharded / data-parallel results equal the
single-process reference semantics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mixvae_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_mesh():
    from mmidas_b200.parallel import plan_mesh
    p = plan_mesh(3, 3, "auto")
    assert (p.arm_ranks, p.dp_ranks) == (3, 1) and p.arm_ranges == [(0, 1), (1, 2), (2, 3)]
    p = plan_mesh(8, 2, "auto")
    assert (p.arm_ranks, p.dp_ranks) == (2, 4)
    assert p.arm_group_ranks(5) == [4, 5] and p.dp_group_ranks(5) == [1, 3, 5, 7] and p.local_arms(5) == (1, 2)
    p = plan_mesh(8, 5, "auto")          # 5 arms do not divide 8 ranks: all arms everywhere, pure DP
    assert (p.arm_ranks, p.dp_ranks) == (1, 8) and p.local_arms(3) == (0, 5)
    p = plan_mesh(4, 4, "dp")
    assert (p.arm_ranks, p.dp_ranks) == (1, 4)
    p = plan_mesh(8, 4, "arm")
    assert (p.arm_ranks, p.dp_ranks) == (4, 2) and p.arm_ranges[3] == (3, 4)
    with pytest.raises(ValueError):
        plan_mesh(8, 5, "arm")
    with pytest.raises(ValueError):
        plan_mesh(0, 2)


def test_consensus_helpers_known_answers():
    from mmidas_b200._utils import classify, compute_confmat, confmat_mean, confmat_normalize, consensus
    probs = np.array([[0.7, 0.2, 0.1], [0.5, 0.4, 0.1], [0.3, 0.3, 0.4]])
    assert classify(probs).tolist() == [0, 0, 2]            # example in mmidas/cpl_mixvae.py:188-191
    a = np.array([0, 0, 1, 2, 2, 2])
    b = np.array([0, 1, 1, 2, 2, 0])
    cm = compute_confmat(a, b, 3)
    assert cm.tolist() == [[1, 1, 0], [0, 1, 0], [1, 0, 2]]
    naive = np.zeros((3, 3))
    for i, j in zip(a, b):
        naive[i, j] += 1
    assert np.array_equal(cm, naive)
    nm = confmat_normalize(cm)
    maxes = np.maximum(cm.sum(0), cm.sum(1))
    assert np.allclose(nm, cm / maxes)
    assert np.isclose(confmat_mean(nm), np.mean(np.diag(cm / maxes)))
    assert np.isclose(consensus([a, a], 3), 1.0)
    z = confmat_normalize(np.zeros((2, 2)))
    assert np.array_equal(z, np.zeros((2, 2)))             # empty categories give 0, not NaN


def _worker(rank, world, port, mode, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from mmidas_b200 import _dist_utils as D
    from mmidas_b200.parallel import (all_gather_arms, allreduce_mean, fixup_loss_vector, grad_buckets, make_groups,
                                      plan_mesh)
    D.init_dist_env(rank, world, "127.0.0.1", str(port), backend="gloo")
    try:
        torch.set_num_threads(2)
        hp = O.HP(input_dim=64, n_categories=12, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0)
        B = 48
        gen = torch.Generator().manual_seed(546)
        x = O.synth_x(B, hp.input_dim, gen)
        noise = O.synth_noise(hp, B, gen)
        sd = O.init_state_dict(hp, 546)
        plan = plan_mesh(world, hp.n_arm, mode)
        arm_group, dp_group, mesh_group = make_groups(plan, rank)
        assert mesh_group is not None
        if mode == "arm":
            # each rank owns one arm; q(c|x) and the samples are all-gathered, the coupling terms are
            # evaluated on the full set, rec/KL only for the local arm, then fixed up.
            a0, a1 = plan.local_arms(rank)
            fw = O.forward(sd, [x] * hp.n_arm, noise, hp, train=True)     # (oracle computes all arms; keep ours)
            q_local = torch.stack(fw["qc"][a0:a1]).detach()
            c_local = torch.stack(fw["c_smp"][a0:a1]).detach()
            q_all = all_gather_arms(q_local, plan, arm_group)
            c_all = all_gather_arms(c_local, plan, arm_group)
            assert torch.equal(q_all, torch.stack(fw["qc"]).detach())
            assert torch.equal(c_all, torch.stack(fw["c_smp"]).detach())
            ls = O.loss(fw, [x] * hp.n_arm, hp)
            At = hp.n_arm
            vec = torch.zeros(5 + 3 * At)
            vec[1], vec[2], vec[3], vec[4] = ls["joint"].detach(), ls["ent"].detach(), ls["dist"].detach(), ls["l2"].detach()
            for a in range(a0, a1):                                        # what mvae_loss writes on this rank
                vec[5 + a] = ls["rec"][a]
                vec[5 + At + a] = ls["kl"][a].detach()
                vec[5 + 2 * At + a] = ls["ll"][a].detach()
            vec[0] = max(At - 1, 1) * sum(vec[5 + a] + hp.beta * vec[5 + At + a] for a in range(a0, a1)) + vec[1]
            fixed = fixup_loss_vector(vec, plan, arm_group, hp.beta)
            out[rank] = (fixed.numpy().copy(), float(ls["total"]))
        else:
            # data parallel: each rank takes half of the cells, local BatchNorm statistics, gradients
            # averaged in two buckets (fc11 | rest) over the flat layout [A, arm_stride].
            half = B // world
            sl = slice(rank * half, (rank + 1) * half)
            nz = {k: v[:, sl] for k, v in noise.items()}
            st = O.TrainState(hp, O.cast_state_dict(sd, torch.float32))
            res = O.train_step(st, [x[sl]] * hp.n_arm, nz, return_grads=True)
            names = O.param_names(hp)
            # lay the gradients out like the flat buffer: per arm, layer order, fc11 last
            per_arm = []
            for a in range(hp.n_arm):
                per_arm.append(torch.cat([res["grads"][n].reshape(-1) for n in names if int(n.split(".")[1]) == a]))
            flat = torch.stack(per_arm)
            n11 = res["grads"]["fc11.0.weight"].numel() + res["grads"]["fc11.0.bias"].numel()
            off11 = flat.shape[1] - n11
            local = flat.clone()
            late, early = grad_buckets(flat, off11)
            allreduce_mean(late, dp_group, plan.dp_ranks)
            allreduce_mean(early, dp_group, plan.dp_ranks)
            out[rank] = (local.numpy().copy(), flat.numpy().copy())
    finally:
        D.destroy_dist_env()


@pytest.mark.parametrize("mode", ["arm", "dp"])
def test_two_rank_gloo(mode):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, mode, out), nprocs=world, join=True)
    assert set(out.keys()) == {0, 1}
    if mode == "arm":
        for r in (0, 1):
            fixed, total = out[r]
            assert abs(fixed[0] / total - 1) < 1e-6           # total rebuilt from the all-reduced per-arm terms
            assert np.all(fixed[5:] != 0)                      # every arm's rec / KL / ll present on every rank
        np.testing.assert_array_equal(out[0][0], out[1][0])
    else:
        l0, avg0 = out[0]
        l1, avg1 = out[1]
        np.testing.assert_array_equal(avg0, avg1)              # replicas agree after the all-reduce
        np.testing.assert_allclose(avg0, (l0.astype(np.float64) + l1) / 2, rtol=1e-6, atol=1e-30)


def test_dist_utils_surface():
    from mmidas_b200 import _dist_utils as D
    for name in ("init_dist_env", "destroy_dist_env", "set_print", "find_addr", "find_port"):
        assert callable(getattr(D, name))                      # mmidas/_dist_utils.py:12,20,54,58,62
    port = D.find_port("127.0.0.1")
    assert 1024 < port < 65536


def test_arm_shard_state_round_trip():
    """slice_arm_state / merge_arm_states: a full reference-layout checkpoint cut into arm shards and merged back is the
    same checkpoint (model keys in the reference's order, optimizer state in model.parameters() order)."""
    from mmidas_b200 import FusedAdam, mixVAE_model
    from mmidas_b200.parallel import merge_arm_states, plan_mesh, slice_arm_state
    kw = dict(input_dim=64, fc_dim=32, n_categories=12, state_dim=2, lowD_dim=6, x_drop=0.5, s_drop=0.0, lam=1, lam_pc=1,
              tau=0.005, beta=1.0, hard=False, variational=True, device="cpu", eps=1e-8, momentum=0.01, ref_prior=False,
              loss_mode="MSE")
    torch.manual_seed(546)
    full = mixVAE_model(n_arm=4, **kw)
    assert full.ctor_kwargs()["n_arm"] == 4 and full.ctor_kwargs()["fc_dim"] == 32
    opt = FusedAdam(full.parameters(), model=full)
    m, v = opt.flat_state()
    m.copy_(torch.randn_like(m)); v.copy_(torch.rand_like(v)); opt.step_count = 3
    sd, osd = full.state_dict(), opt.state_dict()
    plan = plan_mesh(2, 4, "arm")
    parts = []
    for r in range(2):
        a0, a1 = plan.arm_ranges[r]
        local = mixVAE_model(n_arm=a1 - a0, **kw)
        local.load_state_dict(slice_arm_state(sd, a0, a1))
        lopt = FusedAdam(local.parameters(), model=local)
        # the shard's optimizer state = the full state's entries of its arms, renumbered (layer-major, local-arm-minor)
        lstate = {}
        for li in range(14):
            for la in range(a1 - a0):
                for wb in range(2):
                    lstate[(li * (a1 - a0) + la) * 2 + wb] = osd["state"][(li * 4 + a0 + la) * 2 + wb]
        lopt.load_state_dict({"state": lstate, "param_groups": lopt.state_dict()["param_groups"]})
        parts.append((local.state_dict(), lopt.state_dict()["state"]))
    msd, state = merge_arm_states(parts, plan)
    assert list(msd.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(msd[k], sd[k]), k
    assert set(state) == set(osd["state"])
    for i in osd["state"]:
        assert torch.equal(state[i]["exp_avg"], osd["state"][i]["exp_avg"]), i
        assert torch.equal(state[i]["exp_avg_sq"], osd["state"][i]["exp_avg_sq"]), i


def _subset_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from mmidas_b200 import _dist_utils as D
    from mmidas_b200.parallel import make_groups, plan_mesh
    D.init_dist_env(rank, world, "127.0.0.1", str(port), backend="gloo")
    try:
        plan = plan_mesh(2, 2, "arm")                      # a 2-rank mesh on ranks [2, 0] of a 3-rank world
        arm_g, dp_g, mesh_g = make_groups(plan, rank, ranks=[2, 0])
        if rank in (0, 2):
            t = torch.tensor([float(rank)])
            dist.all_reduce(t, group=arm_g)
            out[rank] = (float(t), mesh_g is not None, dp_g is None)
        else:
            out[rank] = (None, mesh_g is not None, dp_g is None)
    finally:
        D.destroy_dist_env()


def test_mesh_on_a_subset_of_ranks():
    """BASELINE config 3 runs A=3 on 3 of the 4 GPUs gpurun hands out: groups must form on a subset of the world."""
    world = 3
    out = mp.Manager().dict()
    mp.spawn(_subset_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == (2.0, True, True) and out[2] == (2.0, True, True) and out[1] == (None, False, True)
