"""clock64 stamps of the encoder chain kernels phase boundaries (development build with -DCHAIN_STAMPS, see fc11_ts.cu).
usage: MVAE_LIB=<lib built with -DCHAIN_STAMPS> python profiles/tools/chain_stamps.py  -> gpurun_out/chain_stamps.npy"""
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
buf = np.zeros((2, 64), dtype=np.int64)
lib = _lib.load()
rc = lib.mvae_debug_chain_stamps(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "chain_stamps.npy"), buf)
print("saved", buf.shape, int(buf.max() - buf[buf > 0].min()))
names_f = ["start", "fc1 fix-up + barrier"] + [f"L{l}:{x}" for l in range(4) for x in ("stats+weights", "normalise", "gemm", "epilogue+sums", "atomics+barrier")]
names_b = ["start"] + [f"it{it}:{x}" for it in range(5) for x in ("consts+operands", "bn/relu bwd", "gemm", "epilogue+sums", "atomics+barrier")] + ["end"]
for d, names in ((0, names_f), (1, names_b)):
    a = buf[d]; idx = [k for k in range(64) if a[k] > 0]
    t0 = a[idx[0]]
    print("fwd" if d == 0 else "bwd", "total cycles", int(a[idx[-1]] - t0))
    prev = t0
    for k in idx[1:]:
        print(f"  {names[k] if k < len(names) else k:28s} {int(a[k] - prev):7d}")
        prev = a[k]
