"""CPU-side checks of the C ABI: the library loads, exports every symbol include/idealgan.h declares,
validates arguments without touching a GPU, and its host table builder (gen_M / gen_A replacement)
matches the reference's vectors."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_close
from idealgan import _lib as L


def header_symbols():
    text = open(os.path.join(ROOT, "include", "idealgan.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ig_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/idealgan.h but not exported"
    assert sorted(L.SIGNATURES) == syms, "ctypes signature table out of sync with the header"
    assert lib.ig_version() == 100


def test_argument_validation_without_gpu():
    lib = L.load()
    assert lib.ig_gen_tables(0, 1, 6, 1.5, 0, 0) == -1                       # null pointers
    assert b"null" in lib.ig_last_error()
    buf = (ctypes.c_float * 16)()
    p = ctypes.addressof(buf)
    assert lib.ig_gen_tables_host(p, 1, 17, 1.5, p) == -2                    # ne > IG_MAX_NE
    assert lib.ig_ideal_fwd(0, p, 2, p, 1, 6, 16, 200.0, 0, p, 0) == -1      # rows = 2 is no WF-PM tensor (>= 3: the last row is the bipolar one)
    assert lib.ig_ideal_fwd(1, p, 4, p, 1, 6, 16, 200.0, 0, p, 0) == -1      # ff/pd/phase maps have exactly 3 rows
    assert lib.ig_ideal_decode(2, p, 3, p, 1, 6, 16, 200.0, 0, 0, 0, 0, 0, 0) == -1   # no output requested
    assert lib.ig_a2a_loss(p, p, 0, p, 1, 6, 16, 200.0, 1.0, p, 0, 0, p, p, 8, 0) == -4   # scratch too small
    assert lib.ig_get_rho_fwd(p, p, 0, 0, 0, p, 1, 1, 16, 200.0, 0, p, 0, 0) == -2        # LS solve needs ne >= 2
    # uncertainty objectives: r2_mean without r2_var; too few echoes
    assert lib.ig_a2a_uq_loss(p, p, 0, p, p, 0, p, 1, 6, 16, 200.0, 1.0, p, p, 0, 0, 0, p, p, 1 << 20, 0) == -1
    assert b"together" in lib.ig_last_error()
    assert lib.ig_a2a_rician_loss(p, p, 0, p, 0, 0, p, 1, 1, 16, 200.0, 1.0, p, p, 0, 0, 0, p, p, 1 << 20, 0) == -2
    # layout adapters: echo count, mode
    assert lib.ig_acq_to_flat(p, 1, 0, 16, p, 0) == -2
    assert lib.ig_maps_from_flat(p, 1, 16, 7, p, 0) == -1
    assert lib.ig_maps_to_flat(p, 1, 16, 1, 2, 3.0, p, 0) == -1                # mag/phase needs >= 3 channels
    # script-level reductions: nothing to reduce, gradient of an absent input, scratch, mode without variance maps
    assert lib.ig_mag_regs(0, 0, 0, 1, 6, 4, 4, 0.0, 0.0, 0.0, 0.0, p, 0, 0, 0, p, 1 << 20, 0) == -1
    assert lib.ig_mag_regs(p, 0, 0, 1, 6, 4, 4, 0.0, 0.0, 0.0, 0.0, p, 0, p, 0, p, 1 << 20, 0) == -1
    assert lib.ig_mag_regs(p, 0, 0, 1, 6, 4, 4, 0.0, 0.0, 0.0, 0.0, p, 0, 0, 0, p, 8, 0) == -1
    assert lib.ig_mag_regs_scratch_bytes(2, 384, 384) == 16 + 16 * 2 * 576
    assert lib.ig_roi_maps(p, 0, 1, 16, 1, p, 0) == -1
    # peer exchange: rank outside the world, missing context
    assert lib.ig_peer_create(3, 2, ctypes.byref(ctypes.c_void_p())) == -1
    assert lib.ig_a2a_loss_peer(p, p, 0, p, 1, 6, 16, 200.0, 1.0, p, 0, 0, p, p, 1 << 20, None, 0, 1, 0, 0) == -1
    # the interleaved layout is a forward-only output option
    assert lib.ig_ideal_bwd(0, p, 3, p, 1, 6, 16, 200.0, L.F_FLAT, p, p, 0) == -5
    with pytest.raises(ValueError):
        L.check(-1, "x")


@pytest.mark.parametrize("case", ["orig6_1p5", "rand6_3p0", "rand12_1p5", "rand3_1p5"])
def test_host_tables_match_reference(golden, case):
    g = golden("tables")
    te = np.ascontiguousarray(g[case + "_te"][:, :, 0])
    nb, ne = te.shape
    tab = np.zeros((nb, L.TAB_FLOATS), np.float32)
    L.check(L.load().ig_gen_tables_host(te.ctypes.data, nb, ne, float(g[case + "_field"]), tab.ctypes.data), "tables")
    u = L.unpack_table(tab, ne)
    assert_close(u["te"], te, 0)
    c = u["c"][0] + 1j * u["c"][1]
    assert_close(c, g[case + "_M"][:, :, 1], 2e-6, "fat phasor")
    assert np.all(g[case + "_M"][:, :, 0] == 1)
    pw = u["pw"][0] + 1j * u["pw"][1]
    pf = u["pf"][0] + 1j * u["pf"][1]
    assert_close(np.stack([pw, pf], 1), g[case + "_Mpinv"], 3e-6, "M pinv")
    assert_close(u["ap"], g[case + "_Apinv"], 2e-5 if ne > 3 else 5e-3, "A pinv")     # ne == 3: square, ill-conditioned
    rec = u["rec"]
    assert_close(rec[:, :, L.REC_TPW_RE] + 1j * rec[:, :, L.REC_TPW_IM], te * pw, 2e-7)
    assert_close(rec[:, :, L.REC_KPHI], te * 300.0, 2e-7)
    assert_close(rec[:, :, L.REC_NTE_L2E], -te * np.log2(np.e), 2e-7)
    assert np.all(rec[:, :, L.REC_SGN] == np.where(np.arange(ne) % 2 == 0, -1.0, 1.0))
    full = tab[:, :L.TAB_AP_OFF].reshape(nb, L.MAX_NE, L.REC_FLOATS)
    assert not full[:, ne:].any()                                                # zero padding beyond ne
    assert np.all(tab[:, L.TAB_META_OFF] == ne)


def test_header_is_plain_c(tmp_path):
    """include/idealgan.h is the boundary a C (cgo, JNI, Julia ccall ...) binding compiles against: it must be valid C99 on its own."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "use.c"
    src.write_text('#include "idealgan.h"\n'
                   "int probe(void) { ig_peer *p = 0; ig_ctx *c = 0; (void)p; (void)c; return IG_VERSION + IG_E_ARG + (int)sizeof(size_t); }\n")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                   check=True, capture_output=True, text=True)


@pytest.mark.gpu
def test_c_program_round_trip_without_python_or_torch(tmp_path):
    """tests/c_abi/roundtrip.c: the library driven from plain C (gcc, cudaMalloc, raw pointers): tables -> forward model ->
    LS solve gives the water / fat maps back to 1e-5, and bad arguments come back as IG_E_* codes."""
    import subprocess
    exe = tmp_path / "roundtrip"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "roundtrip.c"), "-L", libdir, "-lidealgan", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                    "-o", str(exe)], check=True, capture_output=True)
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and "C ABI round trip ok" in out.stdout, out.stdout + out.stderr
