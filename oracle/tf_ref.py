"""The reference itself -- /root/reference/wflib under TensorFlow -- as oracle pin and as timed CPU baseline.
TEST / BENCH INFRASTRUCTURE ONLY: nothing under ideal-gan_b200/ imports this.

SURVEY.md §8c: "when TF *is* available, `sys.path.insert(0, '/root/reference'); import wflib` works without the rest of
the repo; use it to (a) validate the restatement once, (b) serve as the timed CPU baseline".  TensorFlow is not
installable in the build / bench image (no wheel, no network), so in this container the functions below run on
`oracle/tf_shim` (the torch stand-in the golden vectors were generated with: `tests/test_tf_ref.py` replays every fixture
through this module and requires bit-equality with the committed vectors, which pins the replay logic itself).  On any
machine where `import tensorflow` gives the real thing and the reference checkout is reachable, the same functions run
the reference on TensorFlow's own kernels:

    python oracle/tf_ref.py            # prints which implementation ran and the worst deviation per fixture

  * `check_goldens()`   every forward / solve / acq_to_acq / loss fixture of tests/golden recomputed by the reference,
                         outputs AND tf.GradientTape gradients, deviation relative to the tensor's max-norm.  With real
                         TensorFlow the bound is 1e-6 (only the reduction order inside matmul / QR differs);
                         `tests/test_tf_ref.py::test_real_tensorflow_pins_the_goldens` asserts it and is SKIPPED, not absent,
                         where TensorFlow is missing.
  * `c2_step_fn()`       the C2 step of train-IDEAL-unsup.py:214-218,236,255 on the reference's own `wflib`, as bench.py's
                         `--impl reference` arm and `cpu_baseline` leg time it (`kind: "tf"`) when TensorFlow is there.

Reference location: $IDEALGAN_REFERENCE, /root/reference, or baseline/_ref (a checkout an integrator drops there).
"""
import importlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def find_reference():
    for cand in (os.environ.get("IDEALGAN_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "wflib", "IDEAL_model.py")):
            return cand
    return None


def real_tensorflow():
    """The real TensorFlow module, or None.  (oracle/tf_shim also answers to `import tensorflow`; it has no `__file__`
    under site-packages and its version string ends in '-shim'.)"""
    try:
        tf = importlib.import_module("tensorflow")
    except Exception:
        return None
    return None if str(getattr(tf, "__version__", "")).endswith("-shim") else tf


class Reference:
    """The reference's `wflib` (and tf2gan/loss.py) imported from its checkout under whatever `tensorflow` resolves to."""

    def __init__(self, allow_shim=False):
        self.path = find_reference()
        if self.path is None:
            raise RuntimeError("reference checkout not found (IDEALGAN_REFERENCE, /root/reference, baseline/_ref)")
        self.tf = real_tensorflow()
        self.kind = "tf"
        if self.tf is None:
            if not allow_shim:
                raise RuntimeError("TensorFlow is not importable")
            shim = os.path.join(HERE, "tf_shim")
            if shim not in sys.path:
                sys.path.insert(0, shim)
            sys.modules.pop("tensorflow", None)
            self.tf = importlib.import_module("tensorflow")
            self.kind = "shim"
        # import under a private name: the drop-in package is also called `wflib`
        spec = importlib.util.spec_from_file_location("_reference_wflib_IDEAL_model", os.path.join(self.path, "wflib", "IDEAL_model.py"))
        self.wf = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(self.wf)
        spec = importlib.util.spec_from_file_location("_reference_tf2gan_loss", os.path.join(self.path, "tf2gan", "loss.py"))
        self.loss = importlib.util.module_from_spec(spec)
        try:
            spec.loader.exec_module(self.loss)
        except Exception:                       # tf2gan/loss.py needs tensorflow_probability for the Rician loss only
            self.loss = None

    # -- framework glue ------------------------------------------------------------------------------------
    def T(self, a):
        return self.tf.convert_to_tensor(np.ascontiguousarray(a))

    @staticmethod
    def N(t):
        if hasattr(t, "numpy") and not hasattr(t, "as_subclass"):
            return np.asarray(t.numpy())
        return t.detach().as_subclass(type(t).__mro__[1]).numpy().copy()           # shim tensor -> numpy

    def vjp(self, fn, inputs, ups):
        """outputs of fn(*inputs) and d sum_k <out_k, up_k> / d inputs through tf.GradientTape."""
        tf = self.tf
        xs = [self.T(x) for x in inputs]
        with tf.GradientTape() as tape:
            for x in xs:
                tape.watch(x)
            outs = fn(*xs)
            outs = list(outs) if isinstance(outs, (tuple, list)) else [outs]
            loss = 0.0
            for o, u in zip(outs, ups):
                loss = loss + tf.reduce_sum(o * self.T(u))
        grads = tape.gradient(loss, xs)
        return [self.N(o) for o in outs], [np.zeros_like(x) if g is None else self.N(g) for g, x in zip(grads, inputs)]

    def a2a(self, acqs, pm, **kw):
        """acq_to_acq's S_hat whichever arity the checked-out library has."""
        out = self.wf.acq_to_acq(acqs, pm, **kw)
        return out[-1] if isinstance(out, (tuple, list)) else out

    # -- the C2 step for the bench ---------------------------------------------------------------------------
    def c2_step_fn(self, field=1.5, r2_sc=200.0, graph=True):
        """train-IDEAL-unsup.py:214-218,236,255 with the generator removed: acq_to_acq -> where(A != 0) -> MSE -> d/dPM."""
        tf = self.tf

        def step(acqs, pm, te):
            with tf.GradientTape() as tape:
                tape.watch(pm)
                shat = self.a2a(acqs, pm, te=te, field=field, r2_sc=r2_sc)
                shat = tf.where(acqs != 0.0, shat, 0.0 * shat)
                loss = tf.reduce_mean(tf.square(acqs - shat))
            return loss, tape.gradient(loss, pm)

        return tf.function(step) if graph and self.kind == "tf" else step


def _rel(a, b):
    b = np.asarray(b)
    d = float(np.abs(np.asarray(a, dtype=b.dtype) - b).max())
    return d / max(float(np.abs(b).max()), 1e-30)


def check_goldens(ref=None):
    """{fixture: worst relative deviation} of the reference (as imported now) from tests/golden."""
    ref = ref or Reference(allow_shim=True)
    wf, T = ref.wf, ref.T
    res = {}
    g = np.load(os.path.join(GOLDEN, "tables.npz"))
    for k in ("orig6_1p5", "rand6_3p0", "rand12_1p5", "rand3_1p5"):
        M, Mp = wf.gen_M(T(g[f"{k}_te"]), field=float(g[f"{k}_field"]))
        res[f"tables/{k}"] = max(_rel(ref.N(M), g[f"{k}_M"]), _rel(ref.N(Mp), g[f"{k}_Mpinv"]))
    g = np.load(os.path.join(GOLDEN, "forward.npz"))
    for name in ("wfpm_orig6", "wfpm_bip_rand6", "wfpm_rand3", "wfpm_bip_rand12"):
        layer = wf.IDEAL_Layer(field=float(g[f"{name}_field"]), r2_sc=float(g[f"{name}_r2sc"]))
        te = T(g[f"{name}_te"])
        (y,), (gm,) = ref.vjp(lambda m: layer(m, te=te, training=False), [g[f"{name}_maps"]], [g[f"{name}_up"]])
        res[f"forward/{name}"] = max(_rel(y, g[f"{name}_out"]), _rel(gm, g[f"{name}_gmaps"]))
    for name, sep in (("ffpd_orig6", False), ("ffpd_rand5", False), ("magpha_rand6", True), ("magpha_orig4", True)):
        layer = wf.IDEAL_mag_Layer(field=float(g[f"{name}_field"]), sep_phase=sep)
        te = T(g[f"{name}_te"])
        (y,), (gm,) = ref.vjp(lambda m: layer(m, te, training=False), [g[f"{name}_maps"]], [g[f"{name}_up"]])
        res[f"forward/{name}"] = max(_rel(y, g[f"{name}_out"]), _rel(gm, g[f"{name}_gmaps"]))
    g = np.load(os.path.join(GOLDEN, "solve.npz"))
    for name in ("rho_orig6", "rho_rand6_pc", "rho_rand9"):
        kw = dict(field=float(g[f"{name}_field"]), te=T(g[f"{name}_te"]), r2_sc=float(g[f"{name}_r2sc"]),
                  phase_constraint=bool(g[f"{name}_pc"]), acq_demod=True)
        outs, grads = ref.vjp(lambda a, p: wf.get_rho(a, p, **kw), [g[f"{name}_acqs"], g[f"{name}_pm"]],
                              [g[f"{name}_up_rho"], g[f"{name}_up_demod"]])
        res[f"solve/{name}"] = max(_rel(outs[0], g[f"{name}_rho"]), _rel(outs[1], g[f"{name}_demod"]),
                                   _rel(grads[0], g[f"{name}_gacqs"]), _rel(grads[1], g[f"{name}_gpm"]))
    for name, explicit_te in (("a2a_orig6", False), ("a2a_3T", False), ("a2a_rand7", True)):
        kw = dict(field=float(g[f"{name}_field"]))
        if explicit_te:
            kw["te"] = T(g[f"{name}_te"])
        outs, grads = ref.vjp(lambda a, p: ref.a2a(a, p, **kw), [g[f"{name}_acqs"], g[f"{name}_pm"]], [g[f"{name}_up"]])
        dev = max(_rel(outs[0], g[f"{name}_out"]), _rel(grads[0], g[f"{name}_gacqs"]), _rel(grads[1], g[f"{name}_gpm"]))
        # the config-2 objective and its gradient, as the bench's reference arm computes them
        step = ref.c2_step_fn(field=kw["field"], graph=False)
        te = kw.get("te", T(g[f"{name}_te"]))
        loss, gl = step(T(g[f"{name}_acqs"]), T(g[f"{name}_pm"]), te)
        dev = max(dev, abs(float(ref.N(loss)) - float(g[f"{name}_loss"])) / float(g[f"{name}_loss"]), _rel(ref.N(gl), g[f"{name}_loss_gpm"]))
        res[f"solve/{name}"] = dev
    return res


if __name__ == "__main__":
    r = Reference(allow_shim=True)
    print(f"reference at {r.path}, tensorflow = {r.kind} ({getattr(r.tf, '__version__', '?')})")
    worst = check_goldens(r)
    for k, v in worst.items():
        print(f"  {k:28s} {v:.2e}")
    bound = 1e-6 if r.kind == "tf" else 0.0
    bad = {k: v for k, v in worst.items() if v > bound}
    print("PINNED" if not bad else f"DEVIATIONS above {bound:g}: {bad}")
    sys.exit(0 if not bad else 1)
