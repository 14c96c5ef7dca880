"""ctypes loader for libidealgan.so (the C ABI declared in include/idealgan.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C ideal-gan_b200/csrc`.  There is no
fallback: if the shared object is missing, or the device is not sm_100, every operator raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IDEALGAN_LIB: another build of the same library (kernel A/B experiments, `make SUFFIX=_x EXTRA=-D...`); the product is libidealgan.so
LIB_PATH = os.environ.get("IDEALGAN_LIB") or os.path.join(_HERE, "libidealgan.so")

MAX_NE = 16
REC_FLOATS = 16
TAB_AP_OFF = MAX_NE * REC_FLOATS
TAB_META_OFF = TAB_AP_OFF + 3 * MAX_NE
TAB_FLOATS = TAB_META_OFF + 16
(REC_TE, REC_KPHI, REC_NTE_L2E, REC_SGN, REC_C_RE, REC_C_IM, REC_PW_RE, REC_PW_IM, REC_PF_RE, REC_PF_IM,
 REC_TPW_RE, REC_TPW_IM, REC_TPF_RE, REC_TPF_IM) = range(14)


def unpack_table(tab, ne):
    """(nb, TAB_FLOATS) array/tensor -> dict of views: te, c, pw, pf (complex parts as (re, im) pairs), ap (nb,3,ne)."""
    nb = tab.shape[0]
    rec = tab[:, :TAB_AP_OFF].reshape(nb, MAX_NE, REC_FLOATS)[:, :ne]
    ap = tab[:, TAB_AP_OFF:TAB_META_OFF].reshape(nb, 3, MAX_NE)[:, :, :ne]
    return {"te": rec[:, :, REC_TE], "c": (rec[:, :, REC_C_RE], rec[:, :, REC_C_IM]),
            "pw": (rec[:, :, REC_PW_RE], rec[:, :, REC_PW_IM]), "pf": (rec[:, :, REC_PF_RE], rec[:, :, REC_PF_IM]),
            "ap": ap, "rec": rec}
MODEL_WFPM, MODEL_FFPD, MODEL_MAGPHA = 0, 1, 2
PEER_HANDLE_BYTES = 64
F_PHASE_CONSTRAINT, F_FLAT, F_ONLY_MAG, F_NO_RELU, F_NO_CLIP = 1, 2, 4, 8, 16
HOST_WRITE_COMBINED = 1

_fp = C.c_void_p          # device / host float pointers travel as integers
_i, _f, _l, _sz = C.c_int, C.c_float, C.c_long, C.c_size_t

# name -> (restype, argtypes); one entry per symbol of include/idealgan.h
SIGNATURES = {
    "ig_version": (_i, []),
    "ig_last_error": (C.c_char_p, []),
    "ig_device_ok": (_i, []),
    "ig_gen_tables": (_i, [_fp, _i, _i, _f, _fp, _fp]),
    "ig_gen_tables_ahead": (_i, [_fp, _i, _i, _f, _fp, _fp]),
    "ig_gen_tables_host": (_i, [_fp, _i, _i, _f, _fp]),
    "ig_loss_scratch_bytes": (_sz, [_i, _i]),
    "ig_ideal_fwd": (_i, [_i, _fp, _i, _fp, _i, _i, _i, _f, _i, _fp, _fp]),
    "ig_ideal_decode": (_i, [_i, _fp, _i, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp, _fp, _fp]),
    "ig_ideal_bwd": (_i, [_i, _fp, _i, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp]),
    "ig_ideal_loss": (_i, [_i, _fp, _i, _fp, _fp, _i, _i, _i, _f, _i, _f, _fp, _fp, _fp, _fp, _sz, _fp]),
    "ig_get_rho_fwd": (_i, [_fp, _fp, _l, _fp, _l, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp]),
    "ig_get_rho_maps": (_i, [_fp, _fp, _l, _fp, _l, _fp, _i, _i, _i, _f, _i, _i, _fp, _fp, _fp, _fp]),
    "ig_get_rho_bwd": (_i, [_fp, _fp, _l, _fp, _l, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp, _fp, _fp, _fp]),
    "ig_a2a_fwd": (_i, [_fp, _fp, _l, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp]),
    "ig_a2a_bwd": (_i, [_fp, _fp, _l, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp, _fp, _fp]),
    "ig_a2a_loss": (_i, [_fp, _fp, _l, _fp, _i, _i, _i, _f, _f, _fp, _fp, _fp, _fp, _fp, _sz, _fp]),
    "ig_a2a_uq_loss": (_i, [_fp, _fp, _l, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _f, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _sz, _fp]),
    "ig_a2a_rician_loss": (_i, [_fp, _fp, _l, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _f, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _sz, _fp]),
    "ig_eigenvals": (_i, [_fp, _l, _fp, _fp, _fp]),
    "ig_eigenvals_bwd": (_i, [_fp, _l, _fp, _fp, _fp, _fp]),
    "ig_cse_mag_fwd": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _f, _fp, _fp, _fp, _fp, _fp, _fp]),
    "ig_cse_mag_bwd": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _f, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "ig_acq_unc_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _i, _fp, _fp]),
    "ig_acq_unc_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _i, _fp, _fp, _fp, _fp, _fp]),
    "ig_pdff_unc": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _fp, _fp, _fp]),
    "ig_pdff_extract": (_i, [_fp, _i, _i, _i, _fp, _fp]),
    "ig_mag_regs_scratch_bytes": (_sz, [_i, _i, _i]),
    "ig_mag_regs": (_i, [_fp, _fp, _fp, _i, _i, _i, _i, _f, _f, _f, _f, _fp, _fp, _fp, _fp, _fp, _sz, _fp]),
    "ig_roi_maps": (_i, [_fp, _fp, _i, _i, _i, _fp, _fp]),
    "ig_acq_to_flat": (_i, [_fp, _i, _i, _i, _fp, _fp]),
    "ig_acq_from_flat": (_i, [_fp, _i, _i, _i, _fp, _fp]),
    "ig_maps_to_flat": (_i, [_fp, _i, _i, _i, _i, _f, _fp, _fp]),
    "ig_maps_from_flat": (_i, [_fp, _i, _i, _i, _fp, _fp]),
    "ig_peer_create": (_i, [_i, _i, C.POINTER(C.c_void_p)]),
    "ig_peer_handle": (_i, [C.c_void_p, _fp]),
    "ig_peer_connect": (_i, [C.c_void_p, _fp]),
    "ig_peer_connect_local": (_i, [C.POINTER(C.c_void_p), _i]),
    "ig_a2a_loss_peer": (_i, [_fp, _fp, _l, _fp, _i, _i, _i, _f, _f, _fp, _fp, _fp, _fp, _fp, _sz, C.c_void_p, C.c_uint, _i, _fp, _fp]),
    "ig_peer_publish": (_i, [C.c_void_p, C.c_uint, _i, _fp, _fp, _fp]),
    "ig_peer_reduce": (_i, [C.c_void_p, C.c_uint, _fp, _fp]),
    "ig_peer_destroy": (None, [C.c_void_p]),
    "ig_ctx_create": (_i, [_i, _i, _i, _i, C.POINTER(C.c_void_p)]),
    "ig_ctx_destroy": (None, [C.c_void_p]),
    "ig_a2a_loss_host": (_i, [C.c_void_p, _fp, _fp, _fp, _i, _f, _f, _f, _fp, _fp]),
    "ig_decode_ctx_create": (_i, [_i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_void_p)]),
    "ig_decode_ctx_destroy": (None, [C.c_void_p]),
    "ig_decode_host": (_i, [C.c_void_p, _fp, _fp, _i, _f, _f, _i, _fp, _fp, _fp, _fp]),
    "ig_host_alloc": (_i, [_sz, _i, _i, C.POINTER(C.c_void_p)]),
    "ig_host_free": (_i, [C.c_void_p]),
    "ig_host_numa_node": (_i, [_i]),
    "ig_copy_probe": (_i, [_fp, _fp, _fp, _fp, _sz, _i, _i, C.POINTER(C.c_double)]),
}

_lib = None


class IdealGanError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IdealGanError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ideal-gan_b200/csrc`. There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ig_last_error().decode("utf-8", "replace")
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise (ValueError if rc < 0 else IdealGanError)(f"{what}: {kind} {rc}: {msg}")
