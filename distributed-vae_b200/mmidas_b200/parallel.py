"""Multi-GPU replacement for the reference's (disabled) FSDP wrapper (train.py:140-143, gated off at
train.py:274-275): a 2-D mesh (arm axis x data-parallel axis), one process per GPU.

* arm axis: each rank owns a contiguous range of arms (parameters, Adam state, BN buffers). The only
  exchange is an all-gather of the categorical posteriors q(c|x) and samples ([B, C] fp32 per arm)
  before the coupling loss; every rank then evaluates all pair terms and the gradient of its own
  arms (remote arms are constants, exactly what autograd gives each arm).
* data-parallel axis: the cell batch is split; replicas keep LOCAL BatchNorm / inv_var statistics
  (the semantics of the reference's FSDP wrap) and average gradients with an NCCL
  all-reduce.

The whole step — forward, all-gather, loss, backward, all-reduce, Adam — is captured once per input
buffer in a CUDA graph (nn_model.StepGraph; NCCL collectives are capturable) and replayed with one
launch, so the Python dispatch of the six calls and the host-side tensor-map encoding disappear from
the step.  Everything here is plumbing around torch.distributed; the arithmetic stays in
libmixvae_b200.so.  The helpers take plain tensors so that the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass
class MeshPlan:
    world_size: int
    n_arm: int
    arm_ranks: int          # size of the arm axis
    dp_ranks: int           # size of the data-parallel axis
    arm_ranges: List[Tuple[int, int]]   # per arm-rank [start, end)

    def coords(self, rank: int) -> Tuple[int, int]:
        """rank -> (arm coordinate, dp coordinate); arm axis is the fast one."""
        return rank % self.arm_ranks, rank // self.arm_ranks

    def arm_group_ranks(self, rank: int) -> List[int]:
        _, d = self.coords(rank)
        return [d * self.arm_ranks + a for a in range(self.arm_ranks)]

    def dp_group_ranks(self, rank: int) -> List[int]:
        a, _ = self.coords(rank)
        return [d * self.arm_ranks + a for d in range(self.dp_ranks)]

    def local_arms(self, rank: int) -> Tuple[int, int]:
        return self.arm_ranges[self.coords(rank)[0]]


def plan_mesh(world_size: int, n_arm: int, mode: str = "auto") -> MeshPlan:
    """mode: "dp" (all arms on every GPU, pure data parallel), "arm" (arm axis as large as divides
    the world), "auto" = "arm" when n_arm ranks divide the world evenly, else "dp".
    BASELINE configs: A=3 on 3 GPUs -> 3x1; A=2 on 8 -> 2x4; A=5 on 8 -> 1x8 (5 does not divide 8)."""
    if world_size < 1 or n_arm < 1:
        raise ValueError("world_size and n_arm must be positive")
    arm_ranks = 1
    if mode in ("arm", "auto"):
        for cand in range(min(n_arm, world_size), 0, -1):
            if world_size % cand == 0 and n_arm % cand == 0:
                arm_ranks = cand
                break
        if mode == "arm" and arm_ranks == 1 and world_size > 1 and n_arm > 1:
            raise ValueError(f"no arm sharding of {n_arm} arms divides {world_size} ranks evenly")
    elif mode != "dp":
        raise ValueError("mode must be dp, arm or auto")
    per = n_arm // arm_ranks
    ranges = [(i * per, (i + 1) * per) for i in range(arm_ranks)]
    return MeshPlan(world_size, n_arm, arm_ranks, world_size // arm_ranks, ranges)


def make_groups(plan: MeshPlan, rank: int, ranks: Optional[Sequence[int]] = None):
    """Create the arm-axis and dp-axis process groups.  EVERY rank of the default group must call this (group
    creation is collective), also ranks outside ``ranks``.  ``ranks``: the global ranks that form the mesh, in mesh
    order (default: all of them); ``rank`` is the caller's global rank.  Returns (arm_group, dp_group, mesh_group);
    all None on a rank outside the mesh."""
    ranks = list(range(plan.world_size)) if ranks is None else list(ranks)
    if len(ranks) != plan.world_size:
        raise ValueError("len(ranks) must equal the mesh size")
    arm_group = dp_group = None
    for d in range(plan.dp_ranks):
        members = [ranks[d * plan.arm_ranks + a] for a in range(plan.arm_ranks)]
        g = dist.new_group(members) if plan.arm_ranks > 1 else None
        if rank in members:
            arm_group = g
    for a in range(plan.arm_ranks):
        members = [ranks[d * plan.arm_ranks + a] for d in range(plan.dp_ranks)]
        g = dist.new_group(members) if plan.dp_ranks > 1 else None
        if rank in members:
            dp_group = g
    mesh_group = None
    if len(ranks) != dist.get_world_size():
        g = dist.new_group(ranks)
        mesh_group = g if rank in ranks else None
    elif rank in ranks:
        mesh_group = dist.group.WORLD
    return arm_group, dp_group, mesh_group


def all_gather_arms(local: torch.Tensor, plan: MeshPlan, group) -> torch.Tensor:
    """[A_local, B, C] on every arm rank -> [A_total, B, C] in global arm order."""
    if plan.arm_ranks == 1:
        return local
    out = torch.empty((plan.n_arm,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def fixup_loss_vector(loss_vec: torch.Tensor, plan: MeshPlan, arm_group, beta: float) -> torch.Tensor:
    """mvae_loss on an arm-sharded rank fills rec/kl/ll only for its own arms and sums only their
    reconstruction+KL into `total`.  Sum the per-arm entries over the arm axis and rebuild total
    (nn_model.py:587): total = max(A-1,1) * sum_a(rec_a + beta*kl_a) + joint."""
    At = plan.n_arm
    if plan.arm_ranks > 1:
        per_arm = loss_vec[5:5 + 3 * At].clone()
        dist.all_reduce(per_arm, op=dist.ReduceOp.SUM, group=arm_group)
        loss_vec = loss_vec.clone()
        loss_vec[5:5 + 3 * At] = per_arm
        rec, kl = per_arm[:At], per_arm[At:2 * At]
        loss_vec[0] = max(At - 1, 1) * (rec + beta * kl).sum() + loss_vec[1]
    return loss_vec


def grad_buckets(flat_grads: torch.Tensor, off_fc11: int) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """Split the flat gradient buffer [A, arm_stride] into the two all-reduce buckets: fc11 (weight +
    bias, the tail of every arm's range) and everything before it."""
    A = flat_grads.shape[0]
    late = [flat_grads[a, off_fc11:] for a in range(A)]
    early = [flat_grads[a, :off_fc11] for a in range(A)]
    return late, early


def allreduce_mean(tensors: Sequence[torch.Tensor], group, world: int, async_op: bool = False):
    """Average `tensors` in place over `group`.  NCCL: one AVG all-reduce per tensor (no scaling kernel);
    gloo (CPU tests): SUM then scale."""
    works = []
    nccl = dist.get_backend(group) == "nccl"
    for t in tensors:
        op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        w = dist.all_reduce(t, op=op, group=group, async_op=async_op)
        works.append(w)
    if not async_op and not nccl:
        for t in tensors:
            t.div_(world)
    return works


def slice_arm_state(sd: Dict[str, torch.Tensor], a0: int, a1: int) -> Dict[str, torch.Tensor]:
    """Entries of a full reference-layout state dict (``fc1.{a}.weight`` ...) for arms [a0, a1), renumbered from 0."""
    out = {}
    for k, v in sd.items():
        name, a, rest = k.split(".", 2)
        a = int(a)
        if a0 <= a < a1:
            out[f"{name}.{a - a0}.{rest}"] = v
    return out


def merge_arm_states(parts, plan: MeshPlan, n_layers: int = 14):
    """Inverse of the arm sharding for checkpoints.  ``parts[r] = (model_state_dict, optimizer_state)`` of arm rank r
    (keys / indices local to that rank) -> reference-layout dicts of the whole model: model keys ``name.{a}.rest`` in
    the reference's key order, optimizer state indexed in ``model.parameters()`` order (layer-major, arm-minor)."""
    per = plan.n_arm // plan.arm_ranks
    full = {}
    for r, (sd_r, _) in enumerate(parts):
        a0 = plan.arm_ranges[r][0]
        for k, v in sd_r.items():
            name, a, rest = k.split(".", 2)
            full[(name, int(a) + a0, rest)] = v
    layout = {}                                     # module name -> its per-arm entries, both in the reference's order
    for k in parts[0][0].keys():
        name, _, rest = k.split(".", 2)
        rests = layout.setdefault(name, [])
        if rest not in rests:
            rests.append(rest)
    ordered = {f"{name}.{a}.{rest}": full[(name, a, rest)]
               for name, rests in layout.items() for a in range(plan.n_arm) for rest in rests}
    state = {}
    for li in range(n_layers):
        for a in range(plan.n_arm):
            r, la = a // per, a % per
            for wb in range(2):
                src = parts[r][1].get((li * per + la) * 2 + wb)
                if src is not None:
                    state[(li * plan.n_arm + a) * 2 + wb] = src
    return ordered, state


class ShardedTrainer:
    """One training step on a (arm x dp) mesh.  Each rank constructs it after init_dist_env().

    ``model_kwargs``: constructor arguments of the FULL model (all arms), built under ``torch.manual_seed(seed)`` so
    that arm a has the same initial weights on every rank (and the same as the single-GPU model); or ``model``: an
    existing full model whose state is taken over.  ``ranks``: global ranks forming the mesh (default: the whole
    world); ranks outside it get ``self.active == False`` and must not call ``step``.
    """

    def __init__(self, model_kwargs: Optional[dict] = None, lr: float = 1e-3, mode: str = "auto", temp: float = 1.0,
                 seed: int = 546, rank: Optional[int] = None, world_size: Optional[int] = None,
                 overlap: Optional[bool] = None, model=None, ranks: Optional[Sequence[int]] = None,
                 use_cuda_graph: bool = True, peer_adam: Optional[bool] = None):
        from .nn_model import mixVAE_model
        from .optim import FusedAdam
        self.rank = dist.get_rank() if rank is None else rank
        self.ranks = list(range(dist.get_world_size() if world_size is None else world_size)) if ranks is None else list(ranks)
        self.world = len(self.ranks)
        self.active = self.rank in self.ranks
        if model is not None:
            model_kwargs = model.ctor_kwargs()
        n_arm = model_kwargs["n_arm"]
        self.plan = plan_mesh(self.world, n_arm, mode)
        self.arm_group, self.dp_group, self.mesh_group = make_groups(self.plan, self.rank, self.ranks)
        self.temp = temp
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self._graph_misses = 0
        if not self.active:
            self.model = self.optimizer = None
            return
        self.mesh_rank = self.ranks.index(self.rank)
        a0, a1 = self.plan.local_arms(self.mesh_rank)
        dev = torch.device("cuda", torch.cuda.current_device())
        if model is not None:
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        else:
            torch.manual_seed(seed)
            sd = mixVAE_model(**dict(model_kwargs, device="cpu")).state_dict()
        local = mixVAE_model(**dict(model_kwargs, n_arm=a1 - a0, device=dev))
        local.load_state_dict(slice_arm_state(sd, a0, a1))
        self.model = local.to(dev)
        self.model.n_arm_total = n_arm
        self.model.arm_offset = a0
        self.model.materialize_recon = False
        # noise: arms are told apart by their global index inside the generators; data-parallel replicas by a seed salt,
        # and torch's own device generator (keep_s masks) is re-seeded per rank after the shared-seed weight init
        _, dp_coord = self.plan.coords(self.mesh_rank)
        self.dp_coord = dp_coord
        self.model.seed_salt = (dp_coord * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        torch.cuda.manual_seed(seed + 7919 * (self.rank + 1))
        self.optimizer = FusedAdam(self.model.parameters(), lr=lr, model=self.model)
        self.comm_stream = torch.cuda.Stream(dev)
        self.off_fc11 = int(self.model._layout.offset[26])
        self.device = dev
        # Gradient exchange over the dp axis.  overlap=True: two buckets per arm, fc11 (final right after the fused
        # loss+grad kernels) reduced on a side stream under the whole backward chain.  overlap=False: ONE all-reduce of
        # the whole flat gradient buffer after backward.  The buffers are small (4.3 MB per arm), so the exchange is
        # latency-bound: below ~32 MB of gradients one call beats 2*A overlapped ones (measured at N=2: 0.99 -> 0.8x ms).
        grad_bytes = self.model.flat_grads().numel() * 4
        self.overlap = (grad_bytes > 32 * 2 ** 20) if overlap is None else bool(overlap)
        # peer_adam: the gradient exchange fused with Adam over NVLink peer memory (mvae_adam_peer) instead of an NCCL
        # all-reduce followed by a local Adam; falls back to NCCL when symmetric memory cannot be set up
        self.peer = None
        want_peer = (self.plan.dp_ranks > 1) if peer_adam is None else (bool(peer_adam) and self.plan.dp_ranks > 1)
        if want_peer and dist.get_backend(self.dp_group) == "nccl":
            self._setup_peer_adam()

    # ------------------------------------------------------------------------------------------
    def _setup_peer_adam(self):
        """Move the flat parameter / gradient buffers into symmetric memory shared by the dp replicas (every rank of the dp
        group calls this; the ranks agree on success through an all-reduce, so either all use the fused kernel or none)."""
        ok, err, hp, hg = 1, None, None, None
        m = self.model
        try:
            import torch.distributed._symmetric_memory as symm
            gname = self.dp_group.group_name
            if hasattr(symm, "enable_symm_mem_for_group"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    symm.enable_symm_mem_for_group(gname)
            dev = self.device
            m._flat_alloc = lambda *shape: symm.empty(*shape, dtype=torch.float32, device=dev).zero_()
            m._flatten()
            torch.cuda.synchronize(dev)
        except Exception as e:          # allocation refused on this rank
            ok, err = 0, e
        flag = torch.tensor([float(ok)], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.dp_group)
        if flag.item() > 0:
            try:
                hp = symm.rendezvous(m.flat_parameters(), self.dp_group)
                hg = symm.rendezvous(m.flat_grads(), self.dp_group)
            except Exception as e:
                ok, err = 0, e
            flag = torch.tensor([float(ok)], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.dp_group)
        if flag.item() <= 0:
            if err is not None and self.rank == self.ranks[0]:
                print(f"ShardedTrainer: peer-memory Adam unavailable ({err}); using the NCCL all-reduce")
            if m._flat_alloc is not None:
                m._flat_alloc = None
                m._flatten()
            self._reset_optimizer()
            return
        W = hp.world_size
        self.peer = {"hp": hp, "hg": hg, "rank": hp.rank, "world": W,
                     "pp": (C.c_void_p * W)(*[int(x) for x in hp.buffer_ptrs]),
                     "pg": (C.c_void_p * W)(*[int(x) for x in hg.buffer_ptrs])}
        self._reset_optimizer()

    def _reset_optimizer(self):
        from .optim import FusedAdam
        lr = self.optimizer.param_groups[0]["lr"]
        self.optimizer = FusedAdam(self.model.parameters(), lr=lr, model=self.model)

    def _owned_range(self):
        """[start, end) in floats of the flat buffer whose Adam moments live on this rank (peer_adam), else everything."""
        n = self.model.flat_parameters().numel()
        if self.peer is None:
            return 0, n
        n4, r, W = n // 4, self.peer["rank"], self.peer["world"]
        return 4 * (n4 * r // W), 4 * (n4 * (r + 1) // W)

    def _peer_adam_step(self):
        from . import _lib
        m, opt, pr = self.model, self.optimizer, self.peer
        mm, vv = opt.flat_state()
        g = opt.param_groups[0]
        opt.step_count += 1
        ctr = opt._graph_counters[1:].data_ptr() if opt._graph_counters is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pr["hg"].barrier(channel=0)            # every replica's gradients are final
        _lib.check(_lib.load().mvae_adam_peer(pr["pp"], pr["pg"], mm.data_ptr(), vv.data_ptr(), m.flat_parameters().numel(),
                                              pr["rank"], pr["world"], float(g["lr"]), float(g["betas"][0]),
                                              float(g["betas"][1]), float(g["eps"]), opt.step_count, ctr,
                                              C.c_void_p(stream)), "mvae_adam_peer")
        pr["hp"].barrier(channel=1)            # every replica's parameters are final

    # ------------------------------------------------------------------------------------------
    def _step_eager(self, x_local: torch.Tensor, noise=None) -> torch.Tensor:
        m, plan = self.model, self.plan
        if plan.world_size == 1:
            return m.fused_train_step(x_local.expand(m.n_arm, -1, -1), self.temp, self.optimizer, noise=noise)
        xs = x_local.expand(m.n_arm, -1, -1)
        if plan.arm_ranks == 1 and not self.overlap:
            # pure data parallel: the replica's forward + loss + backward is one C call (mvae_grad_step), then the exchange
            lv = m.fused_grad_step(xs, self.temp, noise=noise)
            if self.peer is not None:
                self._peer_adam_step()
            else:
                allreduce_mean([m.flat_grads()], self.dp_group, plan.dp_ranks)
                self.optimizer.step()
            return lv
        x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = m(xs, self.temp, 0.0, noise=noise)
        ot = m.last_outputs()
        qc_all = all_gather_arms(ot["qc"], plan, self.arm_group)
        cs_all = all_gather_arms(ot["c_smp"], plan, self.arm_group)
        m.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0, qc_all=qc_all, c_smp_all=cs_all)
        cur = torch.cuda.current_stream(self.device)
        if plan.dp_ranks > 1 and self.overlap:
            late, early = grad_buckets(m.flat_grads(), self.off_fc11)
            # fc11 gradients are final after the fused loss+grad kernel: reduce them under the backward
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                allreduce_mean(late, self.dp_group, plan.dp_ranks)
            m._run_backward(m._ctx.gen, None)
            allreduce_mean(early, self.dp_group, plan.dp_ranks)
            cur.wait_stream(self.comm_stream)
        elif self.peer is not None:
            m._run_backward(m._ctx.gen, None)
            self._peer_adam_step()
            return fixup_loss_vector(m._ctx.loss_vec, plan, self.arm_group, float(m.beta))
        else:
            m._run_backward(m._ctx.gen, None)
            if plan.dp_ranks > 1:
                allreduce_mean([m.flat_grads()], self.dp_group, plan.dp_ranks)
        self.optimizer.step()
        return fixup_loss_vector(m._ctx.loss_vec, plan, self.arm_group, float(m.beta))

    def step(self, x_local: torch.Tensor, noise=None) -> torch.Tensor:
        """x_local: this dp-replica's cells [B_local, D] (identical on the ranks of one arm group).
        Returns the loss vector of the local replica (global over arms)."""
        if not self.active:
            raise RuntimeError("this rank is not part of the mesh")
        m = self.model
        m.train()
        graphable = (self.use_cuda_graph and noise is None and m.s_drop == 0.0 and x_local.is_cuda
                     and x_local.dtype == torch.float32 and x_local.stride(-1) == 1 and self._graph_misses < 32)
        if not graphable:
            return self._step_eager(x_local, noise)
        from .nn_model import StepGraph
        key = (x_local.data_ptr(), tuple(x_local.shape), x_local.stride(0), float(self.optimizer.param_groups[0]["lr"]))
        g = self._graphs.get(key)
        if g is not None and g.graph is None:        # released (process group torn down and re-created)
            g = None
        if g is None:
            if not self._graphs and self._graph_misses == 0:
                self._graph_misses += 1          # very first step: eager (lazy allocations, communicator warm-up)
                return self._step_eager(x_local, noise)
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self._graph_misses += 1
            torch.cuda.synchronize(self.device)
            g = StepGraph(m, self.optimizer, lambda: self._step_eager(x_local, None))
            self._graphs[key] = g
        else:
            self._graph_misses = 1
        return g.replay()

    # ------------------------------------------------------------------------------------------
    def eval_batch(self, x_local: torch.Tensor):
        """Eval-mode forward + loss of one batch (cpl_mixvae.py:563-775 on a mesh): returns (loss vector over all arms,
        int32 labels [A_total, B] of every arm), both on the device."""
        m, plan = self.model, self.plan
        m.eval()
        with torch.no_grad():
            xs = [x_local for _ in range(m.n_arm)]
            x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = m(x=xs, temp=self.temp, prior_c=0.0, eval=True)
            ot = m.last_outputs()
            qc_all = all_gather_arms(ot["qc"], plan, self.arm_group)
            cs_all = all_gather_arms(ot["c_smp"], plan, self.arm_group)
            m.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0, qc_all=qc_all, c_smp_all=cs_all)
            lv = fixup_loss_vector(m._ctx.loss_vec, plan, self.arm_group, float(m.beta))
            labels = m.argmax_labels(qc_all)
        return lv, labels

    def all_labels(self, q_local: torch.Tensor) -> torch.Tensor:
        """argmax labels of every arm of the model for the cells of this replica: int32 [A_total, B]."""
        return self.model.argmax_labels(all_gather_arms(q_local, self.plan, self.arm_group))

    def full_state_dicts(self):
        """Reference-layout model and optimizer state dicts of the WHOLE model (cpl_mixvae.py:782-788), gathered over
        the arm axis; every rank of the mesh must call this, every rank gets the result.  dp replicas hold identical
        parameters (averaged gradients, same Adam); BN running buffers are the local replica's, as they would be under
        the reference's FSDP wrap."""
        m, plan = self.model, self.plan
        msd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        if self.peer is not None:
            # the Adam moments are sharded over the dp replicas (each owns the range it updates): sum the owned ranges
            mm, vv = self.optimizer.flat_state()
            a, b = self._owned_range()
            keep = (mm.clone(), vv.clone())
            for t in (mm, vv):
                flat = t.view(-1)
                flat[:a] = 0
                flat[b:] = 0
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.dp_group)
            osd = self.optimizer.state_dict()
            mm.copy_(keep[0])
            vv.copy_(keep[1])
        else:
            osd = self.optimizer.state_dict()
        if plan.arm_ranks == 1:
            return msd, osd
        mine = (msd, {i: {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in s.items()} for i, s in osd["state"].items()})
        parts = [None] * plan.arm_ranks
        dist.all_gather_object(parts, mine, group=self.arm_group)
        ordered, state = merge_arm_states(parts, plan)
        groups = [dict(g, params=list(range(28 * plan.n_arm))) for g in osd["param_groups"]]
        return ordered, {"state": state, "param_groups": groups}
