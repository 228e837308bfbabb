"""A few eager (no CUDA graph) training steps of BASELINE configs[1] for ncu: python profiles/tools/step_eager.py [steps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-vae_b200")]
import bench
from mmidas_b200.cpl_mixvae import cpl_mixVAE
w = bench.WORKLOADS[os.environ.get("MVAE_WORKLOAD", "cfg2")]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
gen = torch.Generator(device=dev).manual_seed(546)
batches = [bench.synth_x_device(w["B"], w["D"], w["density"], gen, dev) for _ in range(2)]
tr = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
tr.use_cuda_graph = False
torch.manual_seed(546)
tr.init_model(n_categories=w["C"], state_dim=w["S"], input_dim=w["D"], x_drop=0.5, s_drop=0.0, n_arm=w["A"])
tr.model.train()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for i in range(n):
    lv = tr.train_batch(batches[i % 2])
torch.cuda.synchronize()
print("loss", float(lv[0]))
