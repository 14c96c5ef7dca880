"""CPU-side checks of the `wflib` drop-in surface: names and call signatures equal the reference's (fixture extracted
from /root/reference/wflib/IDEAL_model.py by oracle/gen_signatures.py), host-side functions match the reference's
vectors, and the operators refuse to run without a CUDA device instead of falling back."""
import inspect
import json
import os

import numpy as np
import pytest
import torch

import wflib as wf
from conftest import GOLDEN, assert_close
from idealgan import _lib as L
from idealgan import dlpack


def ref_signatures():
    return json.load(open(os.path.join(GOLDEN, "reference_signatures.json")))


def _sig(fn, drop_self=False):
    params = list(inspect.signature(fn).parameters.values())
    if drop_self:
        params = params[1:]
    return [[p.name, None if p.default is inspect.Parameter.empty else p.default] for p in params]


def _same(ours, ref, extra_ok=()):
    """Reference parameters appear first, in order, with equal defaults; extra keyword parameters must be declared."""
    assert len(ours) >= len(ref), (ours, ref)
    for (n, d), (rn, rd) in zip(ours, ref):
        assert n == rn, (n, rn)
        if rd is not None:
            assert d == eval(rd), (n, d, rd)        # defaults are literals such as 1.5, 200.0, None, False
        else:
            assert d is None
    assert [n for n, _ in ours[len(ref):]] == list(extra_ok), ours[len(ref):]


def test_names_and_signatures_match_reference():
    ref = ref_signatures()
    for name in ref["constants"]:
        assert hasattr(wf, name), name
    assert wf.fm_sc == 300.0 and wf.rho_sc == 1.4 and wf.ns == 2 and wf.species == ["water", "fat"]
    extra = {"acq_to_acq": ("only_mag", "legacy_single")}       # the 2-result form the reference's callers use
    for name, sig in ref["functions"].items():
        assert hasattr(wf, name), f"wflib.{name} missing"
        _same(_sig(getattr(wf, name)), sig, extra.get(name, ()))
    for cname, methods in ref["classes"].items():
        cls = getattr(wf, cname)
        _same(_sig(cls.__init__, drop_self=True), [m for m in methods["__init__"] if m[0] != "self"])
        _same(_sig(cls.call, drop_self=True), [m for m in methods["call"] if m[0] != "self"])
        assert callable(cls())


def test_constants_match_reference_values():
    f = np.array([0., -3.80, -3.40, -2.60, -1.94, -0.39, 0.60]) * 1e-6 * 42.58e6
    assert_close(np.asarray(wf.f_p)[0].real, f.astype(np.float32), 1e-7)
    assert np.asarray(wf.A_p).shape == (7, 2) and abs(np.asarray(wf.A_p)[1:, 1].real.sum() - 0.999) < 1e-6


def test_gen_tevar_matches_reference(golden):
    g = golden("tables")
    assert_close(wf.gen_TEvar(6, 2, orig=True).numpy(), g["te_orig6"], 1e-7)
    assert_close(wf.gen_TEvar(12, 1, orig=True).numpy(), g["te_orig12"], 1e-7)
    assert_close(wf.gen_TEvar(6, 2, TE_ini_min=0.879e-3, TE_ini_d=None, d_TE_min=0.6623e-3, d_TE_d=None).numpy(), g["te_3T"], 1e-7)
    np.random.seed(7)
    assert_close(wf.gen_TEvar(8, 3).numpy(), g["te_rand_seed7"], 1e-7)
    np.random.seed(7)
    assert_close(wf.gen_TEvar(6, 2, TE_ini_d=0.4e-3, d_TE_min=1.0e-3, d_TE_d=0.3e-3).numpy(), g["te_rand_bip_seed7"], 1e-7)


@pytest.mark.parametrize("case", ["orig6_1p5", "rand6_3p0", "rand12_1p5"])
def test_gen_M_and_gen_A_match_reference(golden, case):
    g = golden("tables")
    te, field = torch.from_numpy(g[case + "_te"]), float(g[case + "_field"])
    M, Mp = wf.gen_M(te, field=field)
    assert M.dtype == torch.complex64 and tuple(M.shape) == g[case + "_M"].shape
    assert_close(M.numpy(), g[case + "_M"], 2e-6)
    assert_close(Mp.numpy(), g[case + "_Mpinv"], 3e-6)
    _, P0, Mp2 = wf.gen_M(te, field=field, get_P0=True)
    assert_close(P0.numpy(), g[case + "_P0"], 5e-6)
    _, _, Hp = wf.gen_M(te, field=field, get_H=True)
    assert_close(Hp.numpy(), g[case + "_Hpinv"], 2e-6)
    assert torch.equal(wf.gen_M(te, field=field, get_Mpinv=False), M)
    assert wf.gen_M(te, get_Mpinv=False, get_P0=True) is None            # reference arity quirk (:70-77)
    A, Ap, AtAp = wf.gen_A(M, gen_AtA_pinv=True)
    assert_close(A.numpy(), g[case + "_A"], 2e-6)
    assert_close(Ap.numpy(), g[case + "_Apinv"], 2e-5)


def test_operators_need_a_cuda_device_and_validate_shapes():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(L.IdealGanError):
        wf.IDEAL_Layer()(torch.zeros(1, 3, 4, 4, 2))
    with pytest.raises(L.IdealGanError):
        wf.get_rho(torch.zeros(1, 6, 4, 4, 2), torch.zeros(1, 1, 4, 4, 2))
    with pytest.raises(ValueError):
        wf.acq_to_acq(torch.zeros(1, 6, 4, 4, 1), torch.zeros(1, 1, 4, 4, 1))      # magnitude input belongs to CSE_mag
    with pytest.raises(ValueError):
        wf.acq_to_acq(torch.zeros(1, 6, 4, 4, 2), torch.zeros(1, 1, 4, 4, 2), field=7.0)


def test_dlpack_capsule_inspection_is_zero_copy():
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    info = dlpack.tensor_info(t)
    assert info.ptr == t.data_ptr() and info.shape == (2, 3, 4) and info.strides == (12, 4, 1)
    assert (info.dtype_code, info.bits, info.device_type) == (dlpack.kDLFloat, 32, dlpack.kDLCPU) and info.contiguous
    view = t[:, :, ::2]
    vinfo = dlpack.tensor_info(view)
    assert vinfo.ptr == view.data_ptr() and not vinfo.contiguous
    off = dlpack.tensor_info(t[1])
    assert off.ptr == t[1].data_ptr()
    with pytest.raises(ValueError):
        dlpack.require_f32_cuda(info, "x")                               # CPU tensor: no CPU fallback
    with pytest.raises(ValueError):
        dlpack.capsule_info(object())
