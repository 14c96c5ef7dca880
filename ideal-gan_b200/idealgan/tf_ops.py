"""TensorFlow binding: tf.Tensor <-> the kernels, zero-copy through DLPack, gradients via tf.custom_gradient.

TensorFlow is not installable in the build/bench image, so this module is import-guarded and exercised only where
TF exists (see INTEGRATION.md for the procedure and caveats).  Design: every operator is already a differentiable
torch function (idealgan.torch_ops); `bridge(fn)` lifts such a function to TensorFlow:

  forward : tf eager tensors --to_dlpack--> torch views of the same device memory --fn--> torch outputs
            --from_dlpack--> tf tensors (no copies in either direction)
  backward: registered with tf.custom_gradient; upstream tf gradients are viewed as torch tensors and pushed
            through torch.autograd.grad of the recorded forward, i.e. through the adjoint kernels.

Inside @tf.function graphs (every train step of the reference) the call hops to eager through tf.py_function and
static shapes are restored with set_shape, as the reference reads `.shape` as Python ints everywhere.
Stream ordering: TF and torch use different CUDA streams; the bridge synchronises the device around each hop.
"""
import torch

try:                                               # pragma: no cover - TensorFlow is absent from the CI image
    import tensorflow as tf
except Exception:                                  # pragma: no cover
    tf = None


def available():
    return tf is not None


def is_tf_tensor(x):
    return tf is not None and isinstance(x, (tf.Tensor, tf.Variable))


def to_torch(x):                                   # pragma: no cover
    t = torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.convert_to_tensor(x)))
    return t


def to_tf(t):                                      # pragma: no cover
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


def bridge(fn, n_tensor_args, out_shapes=None):    # pragma: no cover
    """Lift `fn(*torch_tensors, **static) -> torch tensor | tuple` to a differentiable TensorFlow function of its first
    `n_tensor_args` positional arguments."""

    def eager(*tf_args, **static):
        @tf.custom_gradient
        def op(*tensors):
            torch.cuda.synchronize()
            tin = [to_torch(a).requires_grad_(True) for a in tensors]
            with torch.enable_grad():
                out = fn(*tin, **static)
            single = not isinstance(out, (tuple, list))
            outs = [out] if single else list(out)
            torch.cuda.synchronize()
            tf_out = [to_tf(o.detach()) for o in outs]

            def grad(*ups):
                torch.cuda.synchronize()
                gs = [to_torch(u) for u in ups]
                live = [(o, g) for o, g in zip(outs, gs) if o.requires_grad]
                grads = torch.autograd.grad([o for o, _ in live], tin, [g for _, g in live], allow_unused=True, retain_graph=True)
                torch.cuda.synchronize()
                return [to_tf(g) if g is not None else tf.zeros_like(a) for g, a in zip(grads, tensors)]

            return (tf_out[0] if single else tuple(tf_out)), grad

        return op(*tf_args)

    def call(*args, **static):
        tensors = [tf.convert_to_tensor(a, dtype=tf.float32) for a in args[:n_tensor_args]]
        if tf.executing_eagerly():
            return eager(*tensors, **static)
        shapes = out_shapes(*[t.shape for t in tensors], **static) if out_shapes else None
        n_out = len(shapes) if shapes is not None else 1
        res = tf.py_function(lambda *a: eager(*a, **static), tensors, [tf.float32] * n_out)
        if shapes is not None:
            for r, s in zip(res, shapes):
                r.set_shape(s)
        return res[0] if n_out == 1 else tuple(res)

    return call
