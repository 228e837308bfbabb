"""Where does an end-to-end step go?  Times (a) raw pinned->device copies of the packed / dense batch, (b) the unpack kernel,
(c) the feeder + graph-replayed step loop at several ring depths.  Run on a B200: python profiles/tools/e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-vae_b200")]
import torch
from bench import synth_x_device
from mmidas_b200.cpl_mixvae import HostBatchFeeder, cpl_mixVAE
from mmidas_b200.dataloader import PackedBatch

from mmidas_b200 import _lib
if len(sys.argv) > 1:
    _lib.pdl_enable(bool(int(sys.argv[1])))
    print("pdl", sys.argv[1])
dev = torch.device("cuda", 0)
B, D = 5000, 5032
gen = torch.Generator(device=dev).manual_seed(1)
xs = [synth_x_device(B, D, 0.35, gen, dev) for _ in range(4)]
dense = [x.cpu().pin_memory() for x in xs]
packed = [PackedBatch(x.cpu()) for x in xs]
print("packed bytes", packed[0].nbytes, "pinned", packed[0].buffer.is_pinned())


def ev_time(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dst = torch.empty(B, D, device=dev)
st = torch.empty(packed[0].nbytes + 4096, dtype=torch.uint8, device=dev)
t = ev_time(lambda: dst.copy_(dense[0], non_blocking=True)); print(f"dense H2D   {t:.3f} ms  {B*D*4/t/1e6:.1f} GB/s")
t = ev_time(lambda: st[:packed[0].nbytes].copy_(packed[0].buffer, non_blocking=True)); print(f"packed H2D  {t:.3f} ms  {packed[0].nbytes/t/1e6:.1f} GB/s")
t = ev_time(lambda: packed[0].unpack(dev, out=dst, staging=st)); print(f"packed H2D + unpack {t:.3f} ms")

tr = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
torch.manual_seed(546)
tr.init_model(n_categories=100, state_dim=2, input_dim=D, x_drop=0.5, s_drop=0.0, n_arm=2)
tr.model.train()
for host, name in ((packed, "packed"), (dense, "dense")):
    for depth in (2, 3):
        for sync in (True, False):
            def it(n):
                for i in range(n):
                    yield host[i % 4]
            for xd, _ in HostBatchFeeder(it(8), dev, depth=depth):
                tr.train_batch(xd)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 40
            for xd, _ in HostBatchFeeder(it(n), dev, depth=depth):
                lv = tr.train_batch(xd)
                if sync:
                    float(lv[0].item())
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            print(f"{name:7s} depth {depth} item-sync {sync}: {dt*1e3:.3f} ms/step  {B/dt/1e6:.2f} M cells/s")
