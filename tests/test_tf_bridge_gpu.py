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
    op = tf_ops.bridge(lambda a, p: TO.acq_to_acq(a, p, te), 2)
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
    op = tf_ops.bridge(lambda a, p: TO.physics_loss_a2a(a, p, te), 2)
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
    shapes = lambda sa, sp: [(sa[0], 2) + tuple(sa[2:]), tuple(sa)]      # noqa: E731
    op = tf_ops.bridge(lambda a, p: TO.acq_to_acq(a, p, te), 2, out_shapes=shapes)
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
