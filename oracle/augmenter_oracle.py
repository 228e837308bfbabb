"""CPU restatement of the augmenter forward that precedes the training step when ``aug_file`` is set.
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product.

Follows ``Augmenter_smartseq.forward`` (/root/reference/mmidas/augmentation/udagan.py:285-329) in the mode the training
loop runs it (``netA.eval()``, mmidas/cpl_mixvae.py:184; ``netA(x.expand(A,-1,-1), True, 0.1)[1]``, :423): BatchNorm1d uses
its running statistics, Dropout is the identity, and the two ``randn`` draws (udagan.py:287-293 and reparam_trick,
aug_utils.py:51-65) are INJECTED (``z`` [A,B,noise_dim] or [B,noise_dim], ``eps`` like the latent) so that results are
comparable.  Works on a reference ``state_dict`` in any float dtype (fp64 gives the yardstick for tolerances).

Pinned: tests/golden/aug_*.npz hold outputs of the UNMODIFIED reference class produced by tests/golden/make_golden_aug.py
(torch.randn / randn_like patched to return the injected draws); tests/test_augmenter_oracle.py checks this file against them.
"""
import torch
import torch.nn.functional as F


def _bn_eval(y, sd, name, eps, affine=False):
    # nn.BatchNorm1d in eval mode on the feature axis (the reference permutes [A,B,F] -> [B,F,A] and back, udagan.py:294-309:
    # with running statistics that is a per-feature affine on the last axis)
    out = (y - sd[name + ".running_mean"]) / torch.sqrt(sd[name + ".running_var"] + eps)
    if affine:
        out = out * sd[name + ".weight"] + sd[name + ".bias"]
    return out


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def forward(sd, x, z, eps, scale=1.0):
    """sd: state_dict of Augmenter_smartseq; x [..., D]; z [..., noise_dim] standard-normal draw; eps [..., latent] draw.
    Returns (s, x_aug) as the reference does (udagan.py:329)."""
    dt = sd["fc1.weight"].dtype
    x, z, eps = x.to(dt), z.to(dt), eps.to(dt)
    z = F.elu(_bn_eval(_lin(scale * z, sd, "noise"), sd, "bnz", 1e-5, affine=True))          # :287-294 (BatchNorm1d default eps)
    h = x                                                                                    # self.dp: identity in eval
    for i in (1, 2, 3, 4):                                                                   # :295-298
        h = F.relu(_bn_eval(_lin(h, sd, f"fc{i}"), sd, f"batch_fc{i}", 1e-10))
    h = torch.cat((h, z), dim=-1)                                                            # :299
    h = F.relu(_bn_eval(_lin(h, sd, "fc5"), sd, "batch_fc5", 1e-10))                         # :300
    mu = _bn_eval(_lin(h, sd, "fc_mu"), sd, "batch_fc_mu", 1e-10)                            # :302
    sigma = torch.sigmoid(_lin(h, sd, "fc_sigma"))                                           # :303
    s = eps * sigma + mu                                                                     # reparam_trick, aug_utils.py:64-65
    h = s
    for i in (6, 7, 8, 9, 10):                                                               # :305-309
        h = F.relu(_bn_eval(_lin(h, sd, f"fc{i}"), sd, f"batch_fc{i}", 1e-10))
    return s, F.relu(_lin(h, sd, "fc11"))                                                    # :329


def random_state_dict(noise_dim, latent_dim, input_dim, n_dim, seed, dtype=torch.float32):
    """A state_dict with the reference's shapes, non-trivial running statistics and bnz affine (a freshly constructed
    module has mean 0 / var 1, which would not exercise the BatchNorm folding)."""
    g = torch.Generator().manual_seed(seed)
    F1 = input_dim // 5
    shapes = [("noise", noise_dim, noise_dim, False), ("fc1", F1, input_dim, True), ("fc2", F1, F1, True),
              ("fc3", n_dim, F1, True), ("fc4", n_dim, n_dim, True), ("fc5", n_dim // 5, n_dim + noise_dim, True),
              ("fc_mu", latent_dim, n_dim // 5, True), ("fc_sigma", latent_dim, n_dim // 5, True),
              ("fc6", n_dim // 5, latent_dim, True), ("fc7", n_dim, n_dim // 5, True), ("fc8", n_dim, n_dim, True),
              ("fc9", F1, n_dim, True), ("fc10", F1, F1, True), ("fc11", input_dim, F1, True)]
    sd = {}
    for name, n_out, n_in, bias in shapes:
        bound = 1.0 / n_in ** 0.5
        sd[name + ".weight"] = ((torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound * 1.7).to(dtype)
        if bias:
            sd[name + ".bias"] = ((torch.rand(n_out, generator=g) * 2 - 1) * bound).to(dtype)
    bns = [("bnz", noise_dim), ("batch_fc1", F1), ("batch_fc2", F1), ("batch_fc3", n_dim), ("batch_fc4", n_dim),
           ("batch_fc5", n_dim // 5), ("batch_fc_mu", latent_dim), ("batch_fc6", n_dim // 5), ("batch_fc7", n_dim),
           ("batch_fc8", n_dim), ("batch_fc9", F1), ("batch_fc10", F1)]
    for name, n in bns:
        sd[name + ".running_mean"] = (torch.randn(n, generator=g) * 0.2).to(dtype)
        sd[name + ".running_var"] = (torch.rand(n, generator=g) * 0.5 + 0.05).to(dtype)
        sd[name + ".num_batches_tracked"] = torch.tensor(7, dtype=torch.int64)
    sd["bnz.weight"] = (torch.rand(noise_dim, generator=g) + 0.5).to(dtype)
    sd["bnz.bias"] = (torch.randn(noise_dim, generator=g) * 0.1).to(dtype)
    return sd
