"""Multi-GPU replacement for the reference's (disabled) FSDP wrapper (train.py:140-143, gated off at
train.py:274-275): a 2-D mesh (arm axis x data-parallel axis), one process per GPU.

* arm axis: each rank owns a contiguous range of arms (parameters, Adam state, BN buffers). The only
  exchange is an all-gather of the categorical posteriors q(c|x) and samples ([B, C] fp32 per arm)
  before the coupling loss; every rank then evaluates all pair terms and the gradient of its own
  arms (remote arms are constants, exactly what autograd gives each arm).
* data-parallel axis: the cell batch is split; replicas keep LOCAL BatchNorm / inv_var statistics
  (the semantics of the reference's FSDP wrap) and average gradients with an NCCL
  all-reduce in two buckets: fc11 (final right after the fused loss+grad kernel, overlapping the whole
  backward chain) and the rest.

Everything here is plumbing around torch.distributed; the arithmetic stays in libmixvae_b200.so.
The helpers take plain tensors so that the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

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


def make_groups(plan: MeshPlan, rank: int):
    """Create the arm-axis and dp-axis process groups (every rank must call this)."""
    arm_group = dp_group = None
    for d in range(plan.dp_ranks):
        ranks = [d * plan.arm_ranks + a for a in range(plan.arm_ranks)]
        g = dist.new_group(ranks) if plan.arm_ranks > 1 else None
        if rank in ranks:
            arm_group = g
    for a in range(plan.arm_ranks):
        ranks = [d * plan.arm_ranks + a for d in range(plan.dp_ranks)]
        g = dist.new_group(ranks) if plan.dp_ranks > 1 else None
        if rank in ranks:
            dp_group = g
    return arm_group, dp_group


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


class ShardedTrainer:
    """One training step on a (arm x dp) mesh.  Each rank constructs it after init_dist_env()."""

    def __init__(self, model_kwargs: dict, lr: float = 1e-3, mode: str = "auto", temp: float = 1.0,
                 seed: int = 546, rank: Optional[int] = None, world_size: Optional[int] = None,
                 overlap: Optional[bool] = None):
        from .nn_model import mixVAE_model
        from .optim import FusedAdam
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world_size is None else world_size
        n_arm = model_kwargs["n_arm"]
        self.plan = plan_mesh(self.world, n_arm, mode)
        self.arm_group, self.dp_group = make_groups(self.plan, self.rank)
        a0, a1 = self.plan.local_arms(self.rank)
        dev = torch.device("cuda", torch.cuda.current_device())
        # construct every arm with the shared seed so that arm a has the same initial weights on
        # every rank (and the same as the single-GPU model), then keep only the local arms
        torch.manual_seed(seed)
        full = mixVAE_model(**dict(model_kwargs, device="cpu"))
        kw = dict(model_kwargs, n_arm=a1 - a0, device=dev)
        local = mixVAE_model(**kw)
        sd = full.state_dict()
        lsd = {}
        for k, v in sd.items():
            name, a, rest = k.split(".", 2)
            a = int(a)
            if a0 <= a < a1:
                lsd[f"{name}.{a - a0}.{rest}"] = v
        local.load_state_dict(lsd)
        local.n_arm_total = n_arm
        local.arm_offset = a0
        self.model = local.to(dev)
        self.model.n_arm_total = n_arm
        self.model.arm_offset = a0
        self.model.materialize_recon = False
        self.optimizer = FusedAdam(self.model.parameters(), lr=lr, model=self.model)
        self.temp = temp
        self.comm_stream = torch.cuda.Stream(dev)
        self.off_fc11 = int(self.model._layout.offset[26])
        self.device = dev
        # Gradient exchange over the dp axis.  overlap=True: two buckets per arm, fc11 (final right after the fused
        # loss+grad kernels) reduced on a side stream under the whole backward chain.  overlap=False: ONE all-reduce of
        # the whole flat gradient buffer after backward.  The buffers are small (4.3 MB per arm), so the exchange is
        # latency-bound: below ~32 MB of gradients one call beats 2*A overlapped ones (measured at N=2: 0.99 -> 0.8x ms).
        grad_bytes = self.model.flat_grads().numel() * 4
        self.overlap = (grad_bytes > 32 * 2 ** 20) if overlap is None else bool(overlap)

    def step(self, x_local: torch.Tensor, noise=None) -> torch.Tensor:
        """x_local: this dp-replica's cells [B_local, D] (identical on the ranks of one arm group).
        Returns the loss vector of the local replica (global over arms)."""
        m, plan = self.model, self.plan
        if plan.world_size == 1:
            return m.fused_train_step(x_local.expand(m.n_arm, -1, -1), self.temp, self.optimizer, noise=noise)
        m.train()
        xs = x_local.expand(m.n_arm, -1, -1)
        x_recs, _, _, _, cs, _, c_smps, s_means, s_logvars, _ = m(xs, self.temp, 0.0, noise=noise)
        ot = m.last_outputs()
        qc_all = all_gather_arms(ot["qc"], plan, self.arm_group)
        cs_all = all_gather_arms(ot["c_smp"], plan, self.arm_group)
        ls = m.loss(x_recs, [], [], xs, s_means, s_logvars, cs, c_smps, 0.0, qc_all=qc_all, c_smp_all=cs_all)
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
        else:
            m._run_backward(m._ctx.gen, None)
            if plan.dp_ranks > 1:
                allreduce_mean([m.flat_grads()], self.dp_group, plan.dp_ranks)
        self.optimizer.step()
        return fixup_loss_vector(m._ctx.loss_vec, plan, self.arm_group, float(m.beta))
