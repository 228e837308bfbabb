"""The oracle against the extra goldens of the unmodified reference (tests/golden/make_golden_extra.py):
the category-mask path of forward (nn_model.py:332-335) and a checkpoint written by the reference trainer
together with what the reference's eval_model returned for it.  CPU only."""
import os

import numpy as np
import torch

from oracle import mixvae_oracle as O
from golden_cases import GOLDEN, SEED, rel_l2, sample_idx

MASK_HP = dict(input_dim=64, n_categories=12, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0)
MASK_B = 48
CKPT = dict(n_categories=7, state_dim=2, input_dim=64, fc_dim=32, lowD_dim=6, x_drop=0.5, s_drop=0.0, n_arm=2)
CKPT_N, CKPT_NTEST = 192, 48


def mask_inputs():
    hp = O.HP(**MASK_HP)
    gen = torch.Generator().manual_seed(SEED)
    x = O.synth_x(MASK_B, hp.input_dim, gen, 0.35)
    return hp, x, O.synth_noise(hp, MASK_B, gen), O.synth_noise(hp, MASK_B, gen)


def ckpt_inputs():
    gen = torch.Generator().manual_seed(SEED + 1)
    x = O.synth_x(CKPT_N + CKPT_NTEST, CKPT["input_dim"], gen, 0.35)
    return x[CKPT_N:], torch.arange(CKPT_N, CKPT_N + CKPT_NTEST, dtype=torch.float32)


def test_masked_forward_loss_grads_match_reference():
    g = np.load(os.path.join(GOLDEN, "mask.npz"))
    hp, x, n_train, n_eval = mask_inputs()
    mask = g["mask"]
    st = O.TrainState(hp, O.init_state_dict(hp, 546))
    out = O.train_step(st, [x] * hp.n_arm, n_train, return_grads=True, mask=mask)
    ls = out["loss"]
    got = np.array([float(ls["total"]), float(ls["joint"]), float(ls["ent"]), float(ls["dist"]), float(ls["l2"])])
    np.testing.assert_allclose(got, g["train_losses"], rtol=5e-5)   # tiny ill-conditioned batch: BLAS thread count moves fc1 by an ulp, tau amplifies
    for key in ("qc", "c_smp", "s_mean", "x_rec"):
        np.testing.assert_allclose(torch.stack(out["fw"][key]).numpy(), g["train_" + key], rtol=1e-5, atol=1e-6, err_msg=key)
    q = torch.stack(out["fw"]["qc"]).numpy()
    dropped = np.setdiff1d(np.arange(hp.n_categories), mask)
    assert (q[..., dropped] == 0).all()                                   # exactly zero outside the mask
    names = O.param_names(hp)
    gn = np.array([out["grads"][n].double().norm().item() for n in names])
    np.testing.assert_allclose(gn, g["train_grad_norm"], rtol=1e-5)
    for n in names:
        gg = out["grads"][n].reshape(-1)
        want = g["train_gsamp/" + n]
        np.testing.assert_allclose(gg[sample_idx(gg.numel())].numpy(), want, rtol=1e-4, atol=1e-6 * max(1.0, np.abs(want).max()),
                                   err_msg=n)
    # eval-mode forward with the mask, on the reference's UNTRAINED weights (the golden script did not step)
    sd0 = O.init_state_dict(hp, 546)
    fw = O.forward(sd0, [x] * hp.n_arm, n_eval, hp, train=False, mask=mask)
    lo = O.loss(fw, [x] * hp.n_arm, hp)
    # the reference model went through one training-mode forward first: its BN running statistics moved
    nb = {}
    O.forward(sd0, [x] * hp.n_arm, n_train, hp, train=True, new_buffers=nb, mask=mask)
    sd1 = dict(sd0)
    sd1.update(nb)
    fw = O.forward(sd1, [x] * hp.n_arm, n_eval, hp, train=False, mask=mask)
    lo = O.loss(fw, [x] * hp.n_arm, hp)
    for key in ("qc", "c_smp", "s_mean", "s_logvar", "x_rec"):
        np.testing.assert_allclose(torch.stack(fw[key]).numpy(), g["eval_" + key], rtol=1e-5, atol=1e-6, err_msg=key)
    got = np.array([float(lo["total"]), float(lo["joint"]), float(lo["ent"]), float(lo["dist"]), float(lo["l2"])])
    np.testing.assert_allclose(got, g["eval_losses"], rtol=5e-6)


def test_reference_checkpoint_and_eval_model_goldens():
    """A checkpoint written by the reference trainer: the oracle's eval forward on its weights reproduces what the
    reference's eval_model returned (mmidas/cpl_mixvae.py:1590-1619)."""
    ck = torch.load(os.path.join(GOLDEN, "ref_ckpt.pth"), map_location="cpu")
    assert set(ck) == {"model_state_dict", "optimizer_state_dict"}
    sd = ck["model_state_dict"]
    assert len(sd) == 46 * 2
    g = np.load(os.path.join(GOLDEN, "ref_ckpt_eval.npz"))
    hp = O.HP(input_dim=64, fc_dim=32, lowD_dim=6, n_categories=7, state_dim=2, n_arm=2, x_drop=0.5, s_drop=0.0)
    x, idx = ckpt_inputs()
    E = torch.from_numpy(g["E"])               # [batches, A, 16, S]
    nb, B = E.shape[0], E.shape[2]
    mask = np.where(sd["fcc.0.bias"].numpy() != 0.0)[0]
    qs, mus = [], []
    for i in range(nb):
        xb = x[i * B:(i + 1) * B]
        fw = O.forward(sd, [xb] * 2, {"E": E[i]}, hp, train=False, mask=mask)
        qs.append(torch.stack(fw["qc"]))
        mus.append(torch.stack(fw["s_mean"]))
    q = torch.cat(qs, 1).numpy()
    np.testing.assert_allclose(q, g["z_prob"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(torch.cat(mus, 1).numpy(), g["state_mu"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(q.argmax(-1) + 1, g["predicted_label"])        # 1-based labels
    np.testing.assert_array_equal(g["data_indx"], idx.numpy())
