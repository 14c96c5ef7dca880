"""The TensorFlow binding against REAL TensorFlow: skipped (not absent) where `import tensorflow` fails, which includes the
build / bench image.  What runs on a machine that has it: the drop-in `wflib` fed tf.Tensors, eagerly and inside
@tf.function with a symbolic `te`, outputs and tf.GradientTape gradients against the CPU oracle at 1e-5."""
import numpy as np
import pytest
import torch

from idealgan import synth, tf_ops

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not tf_ops.available(), reason="TensorFlow is not installed")]


def _case(rng, nb=2, H=16, W=16, ne=6):
    from oracle import ideal_oracle as orc
    maps = synth.wfpm_maps(nb, H, W, rng, neg_r2_frac=0.0)
    te = synth.te_random(nb, ne, rng)
    acqs = synth.add_noise(orc.IDEAL_model(torch.from_numpy(maps), [1.5, torch.from_numpy(te)]).numpy(), rng)
    pm = np.ascontiguousarray(maps[:, 2:3]) * np.float32(0.9)
    return acqs, pm, te


@pytest.mark.parametrize("graph", [False, True])
def test_unsup_step_through_the_drop_in_matches_the_oracle(graph):
    import wflib as wf
    from oracle import ideal_oracle as orc
    tf = tf_ops.tf
    acqs, pm, te = _case(np.random.default_rng(0))

    def step(a, p, t):                                  # train-IDEAL-unsup.py:214-218,236,255
        with tf.GradientTape() as tape:
            tape.watch(p)
            rho, shat = wf.acq_to_acq(a, p, te=t)
            shat = tf.where(a != 0.0, shat, 0.0)
            loss = tf.reduce_mean(tf.square(a - shat))
        return loss, tape.gradient(loss, p), rho

    fn = tf.function(step) if graph else step
    with tf.device("/GPU:0"):
        loss, g, rho = fn(tf.constant(acqs), tf.constant(pm), tf.constant(te))
    p = torch.from_numpy(pm).requires_grad_(True)
    lref, rho_ref, _ = orc.physics_loss_a2a(torch.from_numpy(acqs), p, te=torch.from_numpy(te))
    (gref,) = torch.autograd.grad(lref, [p])
    assert abs(float(loss) - lref.item()) <= 1e-5 * lref.item()
    assert np.abs(g.numpy() - gref.numpy()).max() <= 1e-5 * np.abs(gref.numpy()).max()
    assert np.abs(rho.numpy() - rho_ref.detach().numpy()).max() <= 1e-5 * np.abs(rho_ref.detach().numpy()).max()
    assert tuple(rho.shape) == (2, 2, 16, 16, 2)


def test_dlpack_round_trip_shares_device_memory():
    tf = tf_ops.tf
    with tf.device("/GPU:0"):
        x = tf.random.uniform((1024,))
    t = tf_ops.to_torch(x)
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), x.numpy())
    assert tf_ops.to_torch(tf_ops.to_tf(t)).data_ptr() == t.data_ptr()
