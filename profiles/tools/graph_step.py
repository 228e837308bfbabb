"""Device time of the CUDA-graph-replayed cfg2 step (what bench.py's `value` measures), several repetitions.
usage: python profiles/tools/graph_step.py [workload] [reps] [steps] [pdl: 1|0]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "distributed-vae_b200"))
import bench
from mmidas_b200.cpl_mixvae import cpl_mixVAE
from mmidas_b200 import _lib
if os.environ.get("MVAE_LIB"):            # (tool only: time an experimental build of the library)
    _lib.LIB_PATH = os.environ["MVAE_LIB"]
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
pdl = int(sys.argv[4]) if len(sys.argv) > 4 else 1
_lib.pdl_enable(bool(pdl))
w = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
gen = torch.Generator(device=dev).manual_seed(546)
batches = [bench.synth_x_device(w["B"], w["D"], w["density"], gen, dev) for _ in range(4)]
tr = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
tr.use_cuda_graph = True
torch.manual_seed(546)
tr.init_model(n_categories=w["C"], state_dim=w["S"], input_dim=w["D"], x_drop=0.5, s_drop=0.0, n_arm=w["A"])
tr.model.train()
out = []
for r in range(reps):
    ms, lv = bench.time_steps(tr.train_batch, batches, steps, 8 if r == 0 else 2, dev)
    out.append(round(ms / steps, 4))
print(wl, "pdl=%d" % pdl, "ms/step", out, "loss", float(lv[0].item()))
