"""TensorFlow binding: tf.Tensor <-> the kernels, zero-copy through DLPack, gradients via tf.custom_gradient.

Every operator is already a differentiable torch function (idealgan.torch_ops); `bridge(fn)` lifts such a function to
TensorFlow:

  forward : tf eager tensors --to_dlpack--> torch views of the same device memory --fn--> torch outputs
            --from_dlpack--> tf tensors (no copies in either direction)
  backward: registered with tf.custom_gradient; upstream tf gradients are viewed as torch tensors and pushed
            through torch.autograd.grad of the recorded forward, i.e. through the adjoint kernels.

Graph mode (`@tf.function`: every train step of the reference, and the Keras functional model of DLlib/module.py:431-432)
hops to eager through tf.py_function.  Everything the eager body needs arrives as a py_function INPUT -- the echo times
included, which are symbolic there -- and the static output shapes (the reference reads `.shape` as Python ints everywhere)
are restored with set_shape from the shapes the frontend computes from the static input shapes.

Stream ordering.  TensorFlow launches on its own non-blocking CUDA stream and does not expose it to Python, so by default a
hop costs one device-wide wait on the way in (the producers of the inputs must have finished before kernels on another
stream read them) and one wait for OUR stream on the way out: two host waits per direction where the first version had four
device-wide ones.  An integrator who has TensorFlow's stream handle (a 15-line custom op returns it) calls
`use_stream(handle)`: every launch then goes onto that stream (`stream` is an argument of every C-ABI entry point) and no
host synchronisation happens at all.

TensorFlow cannot be installed in the build / bench image: tests/test_tf_bridge_gpu.py drives this module through a stand-in
`tf` with TensorFlow's calling conventions (eager and graph mode, symbolic echo times, multi-output operators), and
tests/test_tf_real.py runs the same checks against real TensorFlow wherever `import tensorflow` works (skipped otherwise).
"""
import contextlib

import torch

try:
    import tensorflow as tf
    if not hasattr(tf, "custom_gradient"):         # oracle/tf_shim (a torch stand-in for the golden generator) is not TensorFlow
        tf = None
except Exception:
    tf = None

_external_stream = None
sync_calls = 0                                     # host waits issued by the bridge (tests and INTEGRATION.md count them)


def available():
    return tf is not None


def is_tf_tensor(x):
    return tf is not None and isinstance(x, (tf.Tensor, tf.Variable))


def use_stream(cuda_stream_handle):
    """Run every bridged operator on the given cudaStream_t (TensorFlow's compute stream, as an integer handle) and stop
    synchronising with the host.  None restores the default (own stream + host waits)."""
    global _external_stream
    _external_stream = None if cuda_stream_handle is None else torch.cuda.ExternalStream(int(cuda_stream_handle))


def _launch_stream():
    return torch.cuda.stream(_external_stream) if _external_stream is not None else contextlib.nullcontext()


def _hand_in():
    """Inputs produced on TensorFlow's stream become visible to the stream the kernels run on."""
    global sync_calls
    if _external_stream is None:
        sync_calls += 1
        torch.cuda.synchronize()


def _hand_out():
    """Results written on our stream are complete before TensorFlow's stream may read them."""
    global sync_calls
    if _external_stream is None:
        sync_calls += 1
        torch.cuda.current_stream().synchronize()


def to_torch(x):
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.convert_to_tensor(x)))


def to_tf(t):
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


def bridge(fn, out_shapes=None, n_const=0):
    """Lift `fn(*torch_tensors) -> torch tensor | tuple` to a TensorFlow function of the same number of tensors.  The last
    `n_const` arguments are constants of the operator (the echo times): they travel like the others but get no gradient.
    `out_shapes`: list of static output shapes, one per result -- required in graph mode, where it also fixes the number of
    results tf.py_function is declared with."""

    def eager(*tf_args):
        n_diff = len(tf_args) - n_const

        @tf.custom_gradient
        def op(*tensors):
            _hand_in()
            with _launch_stream():
                tin = [to_torch(a).requires_grad_(True) for a in tensors[:n_diff]] + [to_torch(a) for a in tensors[n_diff:]]
                with torch.enable_grad():
                    out = fn(*tin)
                single = not isinstance(out, (tuple, list))
                outs = [out] if single else list(out)
                _hand_out()
            tf_out = [to_tf(o.detach()) for o in outs]

            def grad(*ups):
                _hand_in()
                with _launch_stream():
                    gs = [to_torch(u) for u in ups]
                    live = [(o, g) for o, g in zip(outs, gs) if o.requires_grad]
                    grads = torch.autograd.grad([o for o, _ in live], tin[:n_diff], [g for _, g in live], allow_unused=True,
                                                retain_graph=True)
                    _hand_out()
                res = [to_tf(g) if g is not None else tf.zeros_like(a) for g, a in zip(grads, tensors)]
                return res + [None] * n_const

            return (tf_out[0] if single else tuple(tf_out)), grad

        return op(*tf_args)

    def call(*args):
        tensors = [tf.convert_to_tensor(a, dtype=tf.float32) for a in args]
        if tf.executing_eagerly():
            return eager(*tensors)
        if out_shapes is None:
            raise ValueError("bridge: graph mode needs the static output shapes (tf.py_function declares its results up front)")
        res = tf.py_function(lambda *a: eager(*a), tensors, [tf.float32] * len(out_shapes))
        res = list(res) if isinstance(res, (tuple, list)) else [res]
        for r, s in zip(res, out_shapes):
            r.set_shape(s)
        return res[0] if len(out_shapes) == 1 else tuple(res)

    return call
