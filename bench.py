#!/usr/bin/env python
"""bench.py — train cells/sec of the coupled mixture-VAE (cpl-mixVAE / MMIDAS) training step.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

One "step" = zero_grad + forward + loss + backward + Adam over one batch of synthetic cells
(mmidas/cpl_mixvae.py:434-463).  Workload at N=1: BASELINE.json configs[1] (A=2 arms, B=5000 cells,
D=5032 genes, C=100 categories, S=2).  N>1: weak scaling, every GPU a data-parallel replica with
its own 5000 cells and an NCCL gradient all-reduce (the reference has no runnable multi-GPU path).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: the trainer's public
per-batch call with pinned HOST batches, H2D copy and loss read-back inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "distributed-vae_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (A, B, D, C, S, density)
    "cfg2": dict(A=2, B=5000, D=5032, C=100, S=2, density=0.35,
                 desc="cpl_mixvae A=2, 5032 genes x batch 5000, 100 categories, state_dim 2 (BASELINE configs[1])"),
    "cfg5": dict(A=2, B=16384, D=30000, C=100, S=2, density=0.08,
                 desc="cpl_mixvae A=2, 30000 genes x batch 16384 (BASELINE configs[4], 10x-shaped)"),
    "cfg1": dict(A=2, B=1000, D=5032, C=92, S=2, density=0.35,
                 desc="cpl_mixvae A=2, 5032 genes x batch 1000, 92 categories (BASELINE configs[0])"),
    "a5": dict(A=5, B=5000, D=5032, C=100, S=2, density=0.35, desc="cpl_mixvae A=5, 5032 genes x batch 5000"),
}
N_ROTATING_BATCHES = 4


def p_arm(D, H, L, C, S):
    return 2 * D * H + 6 * H * H + 2 * H * L + L * C + 2 * (L + C) * S + (C + S) * L + D + 8 * H + 2 * L + C + 2 * S


def algorithmic_bytes(w):
    """A * (16*B*D + 28*P_arm) — four fp32 passes over the gene matrix per arm + Adam."""
    return w["A"] * (16 * w["B"] * w["D"] + 28 * p_arm(w["D"], 100, 10, w["C"], w["S"]))


def synth_x_device(B, D, density, gen, device):
    u = torch.rand(B, D, generator=gen, device=device)
    v = torch.log1p(torch.exp(3.5 + 1.5 * torch.randn(B, D, generator=gen, device=device)))
    return torch.where(u < density, v, torch.zeros((), device=device)).float()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference algorithm (oracle port) on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_reference(w, steps, warmup, max_seconds=25.0, device="cpu"):
    """Times oracle.train_step (the CPU restatement of mmidas/nn_model.py + Adam, pinned to the
    reference by tests/test_oracle_golden.py) on a bounded sample: same shapes, full batch.
    device="cuda" (--ref-device cuda, informational only) runs the same eager torch ops on the GPU:
    the stock-PyTorch path the reference would take there, including its loss.item() sync per step."""
    from oracle import mixvae_oracle as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    hp = O.HP(input_dim=w["D"], n_categories=w["C"], state_dim=w["S"], n_arm=w["A"], x_drop=0.5, s_drop=0.0)
    gen = torch.Generator().manual_seed(546)
    x = O.synth_x(w["B"], w["D"], gen, w["density"]).to(device)
    st = O.TrainState(hp, {k: v.to(device) for k, v in O.init_state_dict(hp, 546).items()})
    xs = [x] * hp.n_arm
    noise = {k: (v.to(device) if torch.is_tensor(v) else [t.to(device) for t in v])
             for k, v in O.synth_noise(hp, w["B"], gen).items()}
    on_gpu = device != "cpu"
    for _ in range(max(1, warmup)):
        O.train_step(st, xs, noise)
    if on_gpu:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        out = O.train_step(st, xs, noise)
        if on_gpu:
            float(out["loss"]["total"].item())           # cpl_mixvae.py:469
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    if on_gpu:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    where = f"torch {torch.__version__} eager on {torch.cuda.get_device_name(0)}" if on_gpu else f"torch {torch.__version__} CPU"
    return {"value": w["B"] * done / dt, "unit": "cells/s", "cores": cores, "kind": "port",
            "sample": f"{done} full steps of the same workload (B={w['B']}, D={w['D']}, A={w['A']}), fp32, "
                      f"{where}, {dt / done * 1e3:.1f} ms/step"}, dt / done


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 20)
    cb, ms = cpu_reference(w, steps, min(args.warmup, 2), max_seconds=60.0, device=args.ref_device)
    line = {"impl": "reference", "metric": "train cells/sec", "value": cb["value"], "unit": "cells/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": ms * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "note": "reference algorithm (oracle port of mmidas/nn_model.py + "
                       "torch Adam) on the host CPU cores; the reference has no runnable multi-GPU path"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "cells/s", "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="tf32x3_fc1", choices=["tf32x3_fc1", "tf32x3", "tf32", "fp32_simt"])
    ap.add_argument("--mesh", default="dp", choices=["dp", "arm", "auto"])
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu (the contract) or cuda (informational: eager torch on the GPU)")
    ap.add_argument("--augment", default="off", choices=["off", "tf32x3", "tf32"],
                    help="single GPU only: run the VAE-GAN augmenter forward (SURVEY f1, production default "
                         "--augmentation True) in front of every step, random-init weights; adds an `augmenter` object")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch.distributed as dist
    from mmidas_b200 import FusedAdam, _lib, mixVAE_model
    from mmidas_b200.cpl_mixvae import HostBatchFeeder, cpl_mixVAE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    A, B, D, C, S = w["A"], w["B"], w["D"], w["C"], w["S"]
    model_kwargs = dict(input_dim=D, fc_dim=100, n_categories=C, state_dim=S, lowD_dim=10, x_drop=0.5, s_drop=0.0,
                        n_arm=A, lam=1, lam_pc=1, tau=0.005, beta=1.0, hard=False, variational=True, device=dev,
                        eps=1e-8, momentum=0.01, ref_prior=False, loss_mode="MSE", precision=args.precision)
    gen = torch.Generator(device=dev).manual_seed(546 + rank)
    batches = [synth_x_device(B, D, w["density"], gen, dev) for _ in range(N_ROTATING_BATCHES)]

    if world == 1:
        trainer = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
        trainer.precision = args.precision
        torch.manual_seed(546)
        trainer.init_model(n_categories=C, state_dim=S, input_dim=D, x_drop=0.5, s_drop=0.0, n_arm=A)
        trainer.model.train()
        step_fn = trainer.train_batch
        parallelism = "single GPU"
        dp_ranks = 1
        if args.augment != "off":
            from mmidas_b200 import Augmenter_smartseq
            torch.manual_seed(547)
            netA = Augmenter_smartseq(noise_dim=50, latent_dim=10, input_dim=D, precision=args.augment)
            for m in netA.modules():                       # non-trivial running statistics, as after training
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.running_mean.normal_(0.0, 0.2)
                    m.running_var.uniform_(0.05, 0.55)
            trainer.netA = netA.to(dev).eval()
    else:
        from mmidas_b200.parallel import ShardedTrainer
        st = ShardedTrainer(model_kwargs, lr=1e-3, mode=args.mesh)
        step_fn = st.step
        parallelism = f"mesh arm{st.plan.arm_ranks} x dp{st.plan.dp_ranks}, NCCL"
        dp_ranks = st.plan.dp_ranks          # ranks of one arm group see the SAME cells: count them once
        if st.plan.arm_ranks > 1:
            # every rank of an arm group must be fed the same batch
            gen = torch.Generator(device=dev).manual_seed(546 + rank // st.plan.arm_ranks)
            batches = [synth_x_device(B, D, w["density"], gen, dev) for _ in range(N_ROTATING_BATCHES)]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput --------------------------------------------------------------
    for i in range(args.warmup):
        lv = step_fn(batches[i % N_ROTATING_BATCHES])
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        lv = step_fn(batches[i % N_ROTATING_BATCHES])
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    last_loss = float(lv[0].item())
    if not (last_loss == last_loss) or abs(last_loss) == float("inf"):
        raise SystemExit(f"non-finite loss {last_loss}")
    ms_per_step = ms / args.steps
    value = dp_ranks * B * args.steps / (ms / 1e3)

    # ---- per-kernel-group device time (CUDA events inside the library, on the launching stream) --
    roofline = None
    groups = None
    if rank == 0:
        peak, peak_src = measured_peak()
    _lib.timing_enable(True)
    for i in range(args.steps):
        step_fn(batches[i % N_ROTATING_BATCHES])
    torch.cuda.synchronize()
    tim = _lib.timing_read()
    _lib.timing_enable(False)
    if rank == 0:
        groups = {g: {"ms_per_step": v[0] / args.steps, "spans": v[1]} for g, v in tim.items() if v[1]}
        gbytes = {"fc1_fwd": A * 4 * B * D, "fc11_loss_grad": A * 8 * B * D, "fc1_wgrad": A * 4 * B * D,
                  "adam": A * 28 * p_arm(D, 100, 10, C, S)}
        dom = max((g for g in groups if g in gbytes), key=lambda g: groups[g]["ms_per_step"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload, {}).get(dom)
            except Exception:
                traffic = None
        ach = gbytes[dom] / (groups[dom]["ms_per_step"] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": gbytes[dom],
                    "step": {"algorithmic_bytes": algorithmic_bytes(w),
                             "achieved": algorithmic_bytes(w) / (ms_per_step / 1e3) / 1e9,
                             "frac": algorithmic_bytes(w) / (ms_per_step / 1e3) / 1e9 / peak},
                    "groups_ms_per_step": {g: round(v["ms_per_step"], 4) for g, v in groups.items()}}

    # ---- the augmenter forward alone (when enabled): device time and achieved tensor throughput ----
    aug = None
    if world == 1 and args.augment != "off" and rank == 0:
        netA = trainer.netA
        F1, nd, nz, nl = D // 5, 500, 50, 10
        enc = D * F1 + F1 * F1 + F1 * nd + nd * nd
        dec = nz * nz + (nd + nz) * (nd // 5) + 2 * (nd // 5) * nl + nl * (nd // 5) + (nd // 5) * nd + nd * nd + nd * F1 + F1 * F1 + F1 * D
        flops = 2.0 * (B * enc + A * B * dec)             # fc1..fc4 once per cell, the rest per (arm, cell)
        xs = batches[0].expand(A, -1, -1)
        for _ in range(3):
            netA(xs, True, 0.1)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        a0.record()
        for i in range(args.steps):
            netA(batches[i % N_ROTATING_BATCHES].expand(A, -1, -1), True, 0.1)
        a1.record()
        torch.cuda.synchronize()
        ams = a0.elapsed_time(a1) / args.steps
        aug_cpu = None
        if not args.no_cpu_baseline:
            # the reference's own path for this row on the host cores: the oracle restatement of Augmenter_smartseq.forward
            # (eval mode, pinned to the reference class by tests/test_augmenter_oracle.py) on a bounded sample of the cells
            from oracle import augmenter_oracle as AO
            cores = len(os.sched_getaffinity(0))
            torch.set_num_threads(cores)
            Bs = min(B, 500)
            sd_cpu = {k: v.detach().cpu() for k, v in netA.state_dict().items()}
            xc = batches[0][:Bs].cpu().expand(A, -1, -1)
            zc, ec = torch.randn(A, Bs, nz), torch.randn(A, Bs, nl)
            with torch.no_grad():
                AO.forward(sd_cpu, xc, zc, ec, 0.1)
                t0 = time.perf_counter()
                reps = 0
                while reps < 5 and time.perf_counter() - t0 < 15.0:
                    AO.forward(sd_cpu, xc, zc, ec, 0.1)
                    reps += 1
            dtc = (time.perf_counter() - t0) / reps
            aug_cpu = {"value": Bs / dtc, "unit": "cells/s", "cores": cores, "kind": "port",
                       "sample": f"{reps} forwards of {Bs} cells x {A} arms (fp32 torch CPU, {dtc * 1e3:.0f} ms each; the reference "
                                 "evaluates fc1..fc4 per arm)"}
        aug = {"ms_per_step": ams, "precision": args.augment, "gflop_per_step": flops / 1e9, "cells_per_s": B / (ams / 1e3),
               "cpu_baseline": aug_cpu,
               "achieved_tflops": flops / (ams / 1e3) / 1e12, "launches_per_step": (_lib.launch_count() - l0) / args.steps,
               "note": "Augmenter_smartseq eval forward (udagan.py:285-329), x.expand over arms: fc1..fc4 evaluated once per cell; "
                       "3xTF32 issues 3 MMAs per product (achieved_tflops counts the fp32-equivalent product once)"}

    # ---- end to end: pinned host batches -> H2D -> step -> loss read-back -------------------------
    e2e = None
    if not args.no_e2e:
        host = [b.cpu().pin_memory() for b in batches]
        n_e2e = args.steps

        def host_iter():
            for i in range(args.warmup + n_e2e):
                yield host[i % N_ROTATING_BATCHES]
        feeder = HostBatchFeeder(host_iter(), dev)
        it = iter(feeder)
        for _ in range(args.warmup):
            xd, _ = next(it)
            float(step_fn(xd)[0].item())
        sync()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            xd, _ = next(it)
            lv = step_fn(xd)
            float(lv[0].item())                     # D2H of the step's loss, every step
        sync()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": dp_ranks * B * n_e2e / dt, "unit": "cells/s", "h2d_bytes_per_step": B * D * 4,
               "d2h_bytes_per_step": 4, "ms_per_step": dt / n_e2e * 1e3,
               "h2d_gbs_per_gpu": B * D * 4 / (dt / n_e2e) / 1e9,   # ~56 GB/s = the PCIe ceiling: e2e is copy-bound
               "api": "cpl_mixVAE.train_batch via HostBatchFeeder (pinned host batch -> side-stream H2D -> fused step -> loss.item())"}

    # the sampler ran from before the timed region through the per-group and end-to-end passes (the same steps, same load):
    # a 20 ms timed region alone would see a single 200 ms sample
    clocks = sampler.stop() if rank == 0 else None
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference(w, 20, 2)

    if rank == 0:
        line = {"metric": "train cells/sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 storage; " + {"tf32x3_fc1": "fc1 3xTF32, other gene GEMMs TF32, rest fp32",
                                                                   "tf32x3": "gene GEMMs 3xTF32, rest fp32",
                                                                   "tf32": "gene GEMMs TF32, rest fp32",
                                                                   "fp32_simt": "fp32 SIMT everywhere"}[args.precision],
                "data": "synthetic",
                "config": {"workload": w["desc"], "cells_per_step_per_gpu": B, "parallelism": parallelism,
                           "precision": args.precision, "dropout": "x_drop=0.5 in-kernel counter-based generator",
                           "l2": f"inputs rotate over {N_ROTATING_BATCHES} distinct batches "
                                 f"({N_ROTATING_BATCHES * B * D * 4 / 1e6:.0f} MB > 126 MB L2)",
                           "last_total_loss": last_loss,
                           **({"augmenter": f"Augmenter_smartseq forward ({args.augment}) in front of every step"} if aug else {})},
                **({"augmenter": aug} if aug else {}),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_base}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
