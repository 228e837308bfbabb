"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol declared in
include/mixvae_b200.h; layout bookkeeping; the Python mirror's state_dict / optimizer layout."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import mixvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from mmidas_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_library_exports_every_declared_symbol():
    lib_mod = _lib()
    lib = lib_mod.load()
    hdr = open(os.path.join(ROOT, "include", "mixvae_b200.h")).read()
    declared = set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(lib_mod.EXPORTS) == declared
    assert lib.mvae_abi_version() == 2


def test_layout_matches_reference_parameter_count():
    lib_mod = _lib()
    d = lib_mod.Dims(2, 5000, 5032, 100, 10, 100, 2, 2, 0)
    lay = lib_mod.compute_layout(d)
    assert sum(lay.numel) == 1076816            # params per arm at D=5032, C=100
    offs = list(lay.offset)
    assert offs == sorted(offs) and all(o % 32 == 0 for o in offs)
    for t in range(27):
        assert offs[t] + lay.numel[t] <= offs[t + 1]
    assert lay.arm_stride % 256 == 0 and lay.arm_stride >= offs[27] + lay.numel[27]
    assert lay.work_floats > 0


def test_layout_rejects_unsupported_shapes():
    lib_mod = _lib()
    lay = lib_mod.Layout()
    for bad in (lib_mod.Dims(2, 100, 64, 200, 10, 12, 2, 2, 0),      # fc_dim > 128
                lib_mod.Dims(2, 100, 64, 100, 10, 300, 2, 2, 0),     # too many categories
                lib_mod.Dims(2, 100, 64, 100, 10, 118, 2, 2, 0),     # lowD_dim + n_categories = 128: no room for the bias column
                lib_mod.Dims(2, 0, 64, 100, 10, 12, 2, 2, 0),        # empty batch
                lib_mod.Dims(17, 100, 64, 100, 10, 12, 2, 17, 0)):
        rc = lib_mod.load().mvae_compute_layout(C.byref(bad), C.byref(lay))
        assert rc < 0
        assert len(lib_mod.load().mvae_last_error()) > 0


def test_model_state_dict_is_reference_layout():
    from mmidas_b200 import FusedAdam, mixVAE_model
    hp = O.HP(input_dim=64, n_categories=12)
    torch.manual_seed(546)
    m = mixVAE_model(input_dim=64, fc_dim=100, n_categories=12, state_dim=2, lowD_dim=10, x_drop=0.5, s_drop=0.0,
                     n_arm=2, lam=1, lam_pc=1, tau=0.005, beta=1.0, hard=False, variational=True, device="cpu",
                     eps=1e-8, momentum=0.01, ref_prior=False, loss_mode="MSE")
    ref = O.init_state_dict(hp, 546)          # pinned bit-exactly to the reference by test_oracle_golden
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys()) or set(sd) == set(ref)
    assert len(sd) == 46 * hp.n_arm
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k    # same seed -> same initial weights as the reference
    assert [n for n, _ in m.named_parameters()] == O.param_names(hp)
    # parameters are views of one flat buffer; load_state_dict / .to() keep that
    m.load_state_dict({k: v + 1 if v.is_floating_point() else v for k, v in ref.items()})
    assert m.fc1[0].weight.data_ptr() == m.flat_parameters().data_ptr()
    assert torch.equal(m.fc11[1].bias, ref["fc11.1.bias"] + 1)
    m2 = m.to("cpu")
    assert m2.fc1[0].weight.data_ptr() == m2.flat_parameters().data_ptr()
    opt = FusedAdam(m.parameters(), lr=1e-3, model=m)
    osd = opt.state_dict()
    want = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))]).state_dict()["param_groups"][0]
    assert set(osd["param_groups"][0]) == set(want)
    assert osd["param_groups"][0]["params"] == list(range(56))


def test_no_cpu_fallback():
    from mmidas_b200 import mixVAE_model
    m = mixVAE_model(input_dim=64, fc_dim=100, n_categories=12, state_dim=2, lowD_dim=10, x_drop=0.5, s_drop=0.0,
                     n_arm=2, lam=1, lam_pc=1, tau=0.005, beta=1.0, hard=False, variational=True, device="cpu",
                     eps=1e-8, momentum=0.01, ref_prior=False, loss_mode="MSE")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m([torch.zeros(4, 64)] * 2, 1.0)
    for kw in (dict(loss_mode="ZINB"), dict(variational=False), dict(ref_prior=True)):
        args = dict(input_dim=64, fc_dim=100, n_categories=12, state_dim=2, lowD_dim=10, x_drop=0.5, s_drop=0.0,
                    n_arm=2, lam=1, lam_pc=1, tau=0.005, beta=1.0, hard=False, variational=True, device="cpu",
                    eps=1e-8, momentum=0.01, ref_prior=False, loss_mode="MSE")
        args.update(kw)
        with pytest.raises(NotImplementedError):
            mixVAE_model(**args)
