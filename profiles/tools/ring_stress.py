"""run-to-run agreement of 30 fused steps at B=15000, D=5032 (the ring-protocol stress test) with an experimental library"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-vae_b200"), os.path.join(ROOT, "tests")]
from mmidas_b200 import _lib
if os.environ.get("MVAE_LIB"):
    _lib.LIB_PATH = os.environ["MVAE_LIB"]
from oracle import mixvae_oracle as O
from gpu_utils import build_model
from mmidas_b200 import FusedAdam
B, D = 15000, 5032
hp = O.HP(input_dim=D, n_categories=100, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0)
gen = torch.Generator().manual_seed(7)
x = O.synth_x(B, D, gen).cuda()
runs = []
for rep in range(3):
    model = build_model(hp, "tf32x3_fc1")
    opt = FusedAdam(model.parameters(), lr=hp.lr, model=model)
    model.train()
    torch.manual_seed(11)
    losses = []
    for step in range(30):
        lv = model.fused_train_step(x.expand(hp.n_arm, -1, -1), hp.temp, opt)
        losses.append(lv[0:1].clone())
    torch.cuda.synchronize()
    runs.append(torch.cat(losses))
d = [(runs[0] - r).abs().max().item() / runs[0].abs().max().item() for r in runs[1:]]
print(os.environ.get("MVAE_LIB", "default"), "max rel diff between runs", d, "first diverging step", [int(((runs[0] - r).abs() > 0).nonzero()[0]) if ((runs[0] - r).abs() > 0).any() else -1 for r in runs[1:]])
