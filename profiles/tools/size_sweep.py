"""Per-group step times at several batch sizes (A=2, D=5032): separates each kernel group's fixed cost from its per-cell
cost.  python profiles/tools/size_sweep.py [B ...]   (eager launches, library event timing per group)"""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, os.environ.get("MVAE_PKG_DIR", "distributed-vae_b200"))]
import bench
from mmidas_b200 import _lib
from mmidas_b200.cpl_mixvae import cpl_mixVAE
Bs = [int(v) for v in sys.argv[1:]] or [1250, 2500, 5000, 10000, 20000]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
gen = torch.Generator(device=dev).manual_seed(546)
for B in Bs:
    D = int(os.environ.get("SWEEP_D", "5032"))
    batches = [bench.synth_x_device(B, D, 0.35, gen, dev) for _ in range(4)]
    tr = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
    tr.use_cuda_graph = False
    torch.manual_seed(546)
    tr.init_model(n_categories=100, state_dim=2, input_dim=D, x_drop=0.5, s_drop=0.0, n_arm=int(os.environ.get("SWEEP_A", "2")))
    tr.model.train()
    for i in range(5): tr.train_batch(batches[i % 4])
    torch.cuda.synchronize()
    _lib.timing_enable(True)
    n = 20
    for i in range(n): tr.train_batch(batches[i % 4])
    torch.cuda.synchronize()
    tim = _lib.timing_read()
    _lib.timing_enable(False)
    g = {k: round(v[0] / n * 1000, 1) for k, v in tim.items() if v[1]}
    print(json.dumps({"B": B, "D": D, "us": g, "sum_us": round(sum(g.values()), 1)}), flush=True)
    del tr, batches
    torch.cuda.empty_cache()
