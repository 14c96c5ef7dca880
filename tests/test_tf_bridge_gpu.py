"""The TensorFlow binding (idealgan/tf_ops.py) driven through a stand-in `tf` module with TensorFlow's calling conventions.

Real TensorFlow cannot be installed in this image, so what is checked here is the binding's own logic -- DLPack capsule
round trips of device memory, the tf.custom_gradient contract (forward returns (outputs, grad_fn), grad_fn maps upstream
tensors to one gradient per input), the tf.py_function hop with set_shape in graph mode, non-differentiable outputs --
against the torch operators.  What it cannot check is TensorFlow itself (its DLPack ownership rules, stream semantics)."""
import numpy as np
import pytest
import torch

import wflib as wf          # imported before any stand-in is installed: its layer classes stay plain callables
from conftest import assert_close
from idealgan import synth, tf_ops
from idealgan import torch_ops as TO

pytestmark = pytest.mark.gpu


class FakeTensor:
    """An opaque framework tensor: the bridge may only touch it through the fake `tf` API below."""

    def __init__(self, t):
        self._t = t
        self.shape = tuple(t.shape)
        self.static_shape = None

    def set_shape(self, s):
        assert tuple(s) == self.shape
        self.static_shape = tuple(s)


class FakeTF:
    float32 = "float32"
    Tensor = FakeTensor
    Variable = FakeTensor

    def __init__(self, eager=True):
        self._eager = eager
        self.recorded = []            # (outputs, grad_fn) of every custom_gradient call
        self.py_function_calls = 0
        outer = self

        class _DLPack:
            @staticmethod
            def to_dlpack(x):
                assert isinstance(x, FakeTensor)
                return torch.utils.dlpack.to_dlpack(x._t)

            @staticmethod
            def from_dlpack(capsule):
                return FakeTensor(torch.utils.dlpack.from_dlpack(capsule))

        class _Experimental:
            dlpack = _DLPack

        self.experimental = _Experimental

        def custom_gradient(f):
            def wrapped(*args):
                out, grad = f(*args)
                outer.recorded.append((out, grad))
                return out
            return wrapped

        self.custom_gradient = custom_gradient

    def convert_to_tensor(self, x, dtype=None):
        return x if isinstance(x, FakeTensor) else FakeTensor(torch.as_tensor(x).cuda())

    def executing_eagerly(self):
        return self._eager

    def zeros_like(self, x):
        return FakeTensor(torch.zeros_like(x._t))

    def py_function(self, func, inp, Tout):
        self.py_function_calls += 1
        was, self._eager = self._eager, True              # inside py_function TensorFlow executes eagerly
        try:
            res = func(*inp)
        finally:
            self._eager = was
        res = list(res) if isinstance(res, (tuple, list)) else [res]
        assert len(res) == len(Tout)
        return res


@pytest.fixture
def fake_tf(monkeypatch):
    fake = FakeTF()
    monkeypatch.setattr(tf_ops, "tf", fake)
    return fake


def _case(rng, nb=2, H=16, W=16, ne=6):
    from oracle import ideal_oracle as orc
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
    te = synth.te_random(nb, ne, rng)
    acqs = synth.add_noise(orc.IDEAL_model(torch.from_numpy(maps), [1.5, torch.from_numpy(te)]).numpy(), rng)
    pm = np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.9)
    return torch.from_numpy(acqs).cuda(), torch.from_numpy(pm).cuda(), torch.from_numpy(te).cuda()


def test_bridge_forward_and_custom_gradient_match_torch(fake_tf):
    acqs, pm, te = _case(np.random.default_rng(0))
    op = tf_ops.bridge(lambda a, p: TO.acq_to_acq(a, p, te))
    rho_tf, shat_tf = op(FakeTensor(acqs), FakeTensor(pm))
    a, p = acqs.clone().requires_grad_(True), pm.clone().requires_grad_(True)
    rho, shat = TO.acq_to_acq(a, p, te)
    assert torch.equal(rho_tf._t, rho) and torch.equal(shat_tf._t, shat)
    assert tf_ops.is_tf_tensor(rho_tf) and not tf_ops.is_tf_tensor(rho)
    # backward through the recorded custom_gradient closure, as TensorFlow's tape would call it
    (_, grad_fn), = fake_tf.recorded
    up_r, up_s = torch.randn_like(rho), torch.randn_like(shat)
    ga_tf, gp_tf = grad_fn(FakeTensor(up_r), FakeTensor(up_s))
    ga, gp = torch.autograd.grad([rho, shat], [a, p], [up_r, up_s])
    assert torch.equal(ga_tf._t, ga) and torch.equal(gp_tf._t, gp)
    ga2, _ = grad_fn(FakeTensor(up_r), FakeTensor(up_s))          # a persistent tape calls it again
    assert torch.equal(ga2._t, ga)


def test_bridge_shares_device_memory_both_ways(fake_tf):
    x = torch.arange(8, dtype=torch.float32, device="cuda")
    view = tf_ops.to_torch(FakeTensor(x))
    assert view.data_ptr() == x.data_ptr()
    back = tf_ops.to_tf(view)
    assert back._t.data_ptr() == x.data_ptr()


def test_bridge_fused_objective_and_unused_input_gradient(fake_tf):
    acqs, pm, te = _case(np.random.default_rng(1))
    op = tf_ops.bridge(lambda a, p: TO.physics_loss_a2a(a, p, te))
    loss_tf = op(FakeTensor(acqs), FakeTensor(pm))
    p = pm.clone().requires_grad_(True)
    loss = TO.physics_loss_a2a(acqs, p, te)
    # the ring kernel's block partials follow its dynamic tile schedule: the scalar repeats to ~1e-7, not bit for bit
    np.testing.assert_allclose(loss_tf._t.item(), loss.item(), rtol=1e-6)
    (_, grad_fn), = fake_tf.recorded
    ga_tf, gp_tf = grad_fn(FakeTensor(torch.ones((), device="cuda")))
    (gp,) = torch.autograd.grad(loss, [p])
    assert torch.equal(gp_tf._t, gp)
    assert not ga_tf._t.any() and ga_tf.shape == tuple(acqs.shape)      # data input: zeros, one gradient per input


def test_bridge_graph_mode_hops_through_py_function_and_restores_shapes(fake_tf):
    fake_tf._eager = False
    acqs, pm, te = _case(np.random.default_rng(2))
    shapes = [(2, 2, 16, 16, 2), tuple(acqs.shape)]
    op = tf_ops.bridge(lambda a, p: TO.acq_to_acq(a, p, te), out_shapes=shapes)
    rho_tf, shat_tf = op(FakeTensor(acqs), FakeTensor(pm))
    assert fake_tf.py_function_calls == 1
    assert rho_tf.static_shape == (2, 2, 16, 16, 2) and shat_tf.static_shape == tuple(acqs.shape)
    rho, shat = TO.acq_to_acq(acqs, pm, te)
    assert torch.equal(shat_tf._t, shat)


def test_wflib_surface_routes_framework_tensors_through_the_bridge(fake_tf):
    """`wf.acq_to_acq` with framework tensors -> frontend._dispatch -> tf_ops.bridge, results equal the torch route."""
    acqs, pm, te = _case(np.random.default_rng(3))
    rho_t, shat_t = wf.acq_to_acq(acqs, pm, te=te)
    rho_f, shat_f = wf.acq_to_acq(FakeTensor(acqs), FakeTensor(pm), te=te)
    assert isinstance(shat_f, FakeTensor)
    assert_close(shat_f._t.cpu().numpy(), shat_t.cpu().numpy(), 0.0)
    assert_close(rho_f._t.cpu().numpy(), rho_t.cpu().numpy(), 0.0)


class _Moments:
    """tfp-like distribution object: the operators only call .mean() / .variance() (IDEAL_model.py:640-655,733-745)."""

    def __init__(self, mean, var):
        self._m, self._v = mean, var

    def mean(self):
        return self._m

    def variance(self):
        return self._v


def _symbolic(fake_tf, *ts):
    """Graph mode: tensors are symbolic (FakeTensor has no .numpy(), so any host hop raises AttributeError)."""
    fake_tf._eager = False
    return [FakeTensor(t) for t in ts]


def test_graph_mode_symbolic_echo_times_travel_as_a_py_function_input(fake_tf):
    """inside @tf.function `te` is symbolic -- including the one gen_TEvar builds in-graph -- so the operators
    may not read it on the host; it rides through tf.py_function as an input and the table is built in the eager body."""
    acqs, pm, te = _case(np.random.default_rng(4))
    rho_t, shat_t = wf.acq_to_acq(acqs, pm, te=te)
    a_s, p_s, te_s = _symbolic(fake_tf, acqs, pm, te)
    rho_f, shat_f = wf.acq_to_acq(a_s, p_s, te=te_s)
    assert fake_tf.py_function_calls == 1
    assert shat_f.static_shape == tuple(acqs.shape) and rho_f.static_shape == (2, 2, 16, 16, 2)
    assert torch.equal(shat_f._t, shat_t) and torch.equal(rho_f._t, rho_t)
    (_, grad_fn), = fake_tf.recorded
    grads = grad_fn(FakeTensor(torch.randn_like(rho_t)), FakeTensor(torch.randn_like(shat_t)))
    assert len(grads) == 3 and grads[2] is None                    # no gradient for the echo times


def test_graph_mode_multi_output_operators_declare_every_result(fake_tf):
    """CSE_mag (5 results), eigenvals (2), PDFF_uncertainty (2) and acq_uncertainty (1) in graph mode."""
    rng = np.random.default_rng(5)
    acqs, pm, te = _case(rng)
    nb, ne, H, W, _ = acqs.shape
    mag = acqs.pow(2).sum(-1, keepdim=True).sqrt().contiguous()
    r2 = pm[..., 1:].contiguous()
    ref = wf.CSE_mag(mag, r2, [1.5, te], demod_signal=True, uncertainty=True)
    m_s, r_s, te_s = _symbolic(fake_tf, mag, r2, te)
    got = wf.CSE_mag(m_s, r_s, [1.5, te_s], demod_signal=True, uncertainty=True)
    assert len(got) == 4 and got[0].static_shape == (nb, 2, H, W, 1) and got[3].static_shape == (nb, 1, H, W, 1)
    for g, r in zip(got, ref):
        assert torch.equal(g._t, r)

    X = torch.rand((nb, H * W, 3), device="cuda")
    fake_tf._eager = True
    xy_r, ratio_r = wf.eigenvals(X)
    (X_s,) = _symbolic(fake_tf, X)
    xy, ratio = wf.eigenvals(X_s)
    assert xy.static_shape == (nb, H * W, 2) and ratio.static_shape == (nb, H * W, 1)
    assert torch.equal(xy._t, xy_r) and torch.equal(ratio._t, ratio_r)

    plane = lambda lo, hi: (lo + (hi - lo) * torch.rand((nb, 1, H, W, 1), device="cuda")).contiguous()      # noqa: E731
    phi_m, phi_v, r2_m, r2_v = plane(-0.5, 0.5), plane(1e-4, 1e-3), plane(0.0, 0.5), plane(1e-4, 1e-3)
    rho_r, cov_r = wf.PDFF_uncertainty(acqs, _Moments(phi_m, phi_v), _Moments(r2_m, r2_v), te=te)
    rho_map, _ = wf.acq_to_acq(acqs, pm, te=te)
    var_r = wf.acq_uncertainty(rho_map, _Moments(phi_m, phi_v), _Moments(r2_m, r2_v), ne=ne, te=te)
    a_s, pm_s, pv_s, rm_s, rv_s, te_s, rho_s = _symbolic(fake_tf, acqs, phi_m, phi_v, r2_m, r2_v, te, rho_map)
    rho, cov = wf.PDFF_uncertainty(a_s, _Moments(pm_s, pv_s), _Moments(rm_s, rv_s), te=te_s)
    assert rho.static_shape == (nb, 2, H, W, 2) and cov.static_shape == (nb, 4, H, W, 1)
    assert torch.equal(rho._t, rho_r) and torch.equal(cov._t, cov_r)
    var = wf.acq_uncertainty(rho_s, _Moments(pm_s, pv_s), _Moments(rm_s, rv_s), ne=ne, te=te_s)
    assert isinstance(var, FakeTensor) and var.static_shape == (nb, ne, H, W, 2)     # VarMeanSquaredError reads shape[-1] // 2
    assert torch.equal(var._t, var_r)


def test_graph_mode_without_static_shapes_is_refused(fake_tf):
    fake_tf._eager = False
    acqs, pm, te = _case(np.random.default_rng(6))
    with pytest.raises(ValueError, match="static output shapes"):
        tf_ops.bridge(lambda a, p: TO.acq_to_acq(a, p, te))(FakeTensor(acqs), FakeTensor(pm))


def test_stream_handoff_two_host_waits_per_direction_and_none_on_a_shared_stream(fake_tf, monkeypatch):
    """Default: one device-wide wait in + one wait on our stream out, per direction.  With the framework's stream handed
    over (use_stream) the kernels are launched on it and the bridge never waits on the host."""
    acqs, pm, te = _case(np.random.default_rng(7))
    op = tf_ops.bridge(lambda a, p: TO.acq_to_acq(a, p, te))
    before = tf_ops.sync_calls
    op(FakeTensor(acqs), FakeTensor(pm))
    assert tf_ops.sync_calls - before == 2
    (_, grad_fn), = fake_tf.recorded
    grad_fn(FakeTensor(torch.randn((2, 2, 16, 16, 2), device="cuda")), FakeTensor(torch.randn_like(acqs)))
    assert tf_ops.sync_calls - before == 4

    rho_ref, shat_ref = TO.acq_to_acq(acqs, pm, te)
    framework_stream = torch.cuda.Stream()
    launched_on = []
    from idealgan import ops
    real = ops._stream
    monkeypatch.setattr(ops, "_stream", lambda: launched_on.append(real()) or launched_on[-1])
    tf_ops.use_stream(framework_stream.cuda_stream)
    try:
        before = tf_ops.sync_calls
        with torch.cuda.stream(framework_stream):
            a2 = acqs.clone()                                   # produced on the framework's stream, consumed without a host wait
        rho_f, shat_f = op(FakeTensor(a2), FakeTensor(pm))
        assert tf_ops.sync_calls == before
        assert launched_on and all(s == framework_stream.cuda_stream for s in launched_on)
        framework_stream.synchronize()
        assert torch.equal(shat_f._t, shat_ref) and torch.equal(rho_f._t, rho_ref)
    finally:
        tf_ops.use_stream(None)


def test_table_cache_one_build_per_echo_train_and_no_host_copy_of_device_echo_times():
    """acq_to_acq + get_rho + acq_uncertainty on the same `te` build the per-sample table once; `te` on the
    device is never moved to the host; in-place edits and new echo trains miss."""
    acqs, pm, te = _case(np.random.default_rng(8))
    TO.table_cache.clear()
    h0, m0 = TO.table_cache.hits, TO.table_cache.misses
    rho, _ = wf.acq_to_acq(acqs, pm, te=te)
    wf.get_rho(acqs, pm, te=te)
    wf.acq_to_acq(acqs, pm, te=te)
    assert (TO.table_cache.misses - m0, TO.table_cache.hits - h0) == (1, 2)
    te_host = te.cpu()                                              # what gen_TEvar returns: keyed on content
    wf.get_rho(acqs, pm, te=te_host)
    wf.get_rho(acqs, pm, te=te_host.clone())
    assert (TO.table_cache.misses - m0, TO.table_cache.hits - h0) == (2, 3)
    ref = wf.get_rho(acqs, pm, te=te)
    te.mul_(1.01)                                                   # in-place edit bumps the version: rebuilt
    new = wf.get_rho(acqs, pm, te=te)
    assert TO.table_cache.misses - m0 == 3 and not torch.equal(ref, new)
