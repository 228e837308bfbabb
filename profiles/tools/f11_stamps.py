"""clock64 stamps of the fc11 kernels' protocol events (development build with -DF11_STAMPS, see fc11_ts.cu).
usage: MVAE_LIB=<lib built with -DF11_STAMPS> python profiles/tools/f11_stamps.py  -> gpurun_out/f11_stamps.npy"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-vae_b200")]
import bench
from mmidas_b200 import _lib
_lib.LIB_PATH = os.environ["MVAE_LIB"]
from mmidas_b200.cpl_mixvae import cpl_mixVAE
w = bench.WORKLOADS["cfg2"]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
gen = torch.Generator(device=dev).manual_seed(546)
batches = [bench.synth_x_device(w["B"], w["D"], w["density"], gen, dev) for _ in range(2)]
tr = cpl_mixVAE(saving_folder="", aug_file="", device=dev, save_flag=False)
tr.use_cuda_graph = False
torch.manual_seed(546)
tr.init_model(n_categories=w["C"], state_dim=w["S"], input_dim=w["D"], x_drop=0.5, s_drop=0.0, n_arm=w["A"])
tr.model.train()
for i in range(4):
    tr.train_batch(batches[i % 2])
torch.cuda.synchronize()
buf = np.zeros((2, 2, 10, 128), dtype=np.int64)
lib = _lib.load()
rc = lib.mvae_debug_f11_stamps(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "f11_stamps.npy"), buf)
print("saved", buf.shape, int(buf.max() - buf[buf > 0].min()))
