import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-vae_b200"))
import bench
from mmidas_b200 import _lib
from mmidas_b200.cpl_mixvae import cpl_mixVAE
w = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
gen = torch.Generator(device=dev).manual_seed(546)
batches = [bench.synth_x_device(w["B"], w["D"], w["density"], gen, dev) for _ in range(4)]
tr = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
torch.manual_seed(546)
tr.init_model(n_categories=w["C"], state_dim=w["S"], input_dim=w["D"], x_drop=0.5, s_drop=0.0, n_arm=w["A"])
tr.model.train()
tr.use_cuda_graph = False          # per-group timing brackets eager launches
for i in range(5): tr.train_batch(batches[i % 4])
torch.cuda.synchronize()
_lib.timing_enable(True)
n = 20
for i in range(n): tr.train_batch(batches[i % 4])
torch.cuda.synchronize()
tim = _lib.timing_read()
print({g: round(v[0] / n, 4) for g, v in tim.items() if v[1]})
