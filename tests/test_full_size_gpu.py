"""BASELINE.json's configurations at their FULL sizes, through size-independent properties (the oracle would need minutes
to hours at these sizes): C2 64 x 384 x 384 x 6, C3 256 x 192 x 192 x 6 with per-sample echo times, C4 bipolar mag/phase
at nb = 3 and nb = 64, C5's decoder on a 256-slice shard streamed in chunks.  Inputs are drawn on the device."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from idealgan import _lib as L
from idealgan import dist as igdist
from idealgan import ops, synth

pytestmark = pytest.mark.gpu


def _gen(seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


def _mask(H, W):
    return torch.from_numpy(synth.disc_mask(H, W).astype(np.float32)).cuda()[None, None, :, :, None]


def _wfpm(nb, H, W, g):
    maps = torch.empty((nb, 3, H, W, 2), device="cuda")
    maps[:, :2] = torch.rand((nb, 2, H, W, 2), device="cuda", generator=g) - 0.5
    maps[:, 2, :, :, 0] = 2 * torch.rand((nb, H, W), device="cuda", generator=g) - 1
    maps[:, 2, :, :, 1] = torch.rand((nb, H, W), device="cuda", generator=g)
    return (maps * _mask(H, W)).contiguous()


def _dot(a, b):
    return (a.double() * b.double()).sum().item()


def test_c2_full_batch_fused_objective_equals_composition():
    nb, H, W, ne = 64, 384, 384, 6
    g = _gen(1234)
    maps = _wfpm(nb, H, W, g)
    tab = ops.gen_tables(torch.from_numpy(synth.te_orig(nb, ne)).cuda(), 1.5)
    S = ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne)
    pm = maps[:, 2:3].contiguous()
    rho, _ = ops.get_rho_fwd(S, pm, tab)
    assert_close(rho.cpu().numpy(), maps[:, :2].cpu().numpy(), 3e-6, "encode -> solve round trip")
    del rho
    noisy = (S + 0.02 * torch.randn(S.shape, device="cuda", generator=g) * (S != 0)).contiguous()
    del S
    pm2 = (pm * 0.95).contiguous()
    loss, gl, _, _ = ops.a2a_loss(noisy, pm2, tab)
    _, shat = ops.a2a_fwd(noisy, pm2, tab, want_rho=False)
    resid = torch.where(noisy != 0, shat - noisy, torch.zeros_like(shat))
    del shat
    ref = (resid.double() ** 2).mean().item()
    assert abs(loss.item() - ref) <= 2e-6 * ref
    up = (2.0 / noisy.numel()) * resid
    del resid
    _, gp = ops.a2a_bwd(noisy, pm2, tab, None, up.contiguous(), need_acqs=False)
    assert_close(gl.cpu().numpy(), gp.cpu().numpy(), 2e-5, "fused vs unfused gradient")
    # background voxels: zero gradient, exactly
    bg = (noisy[:, 0, :, :, 0] == 0) & (noisy[:, 0, :, :, 1] == 0)
    assert gl[:, 0][bg].abs().max().item() == 0.0


def test_c3_per_sample_echo_times_round_trip_and_adjoint():
    nb, H, W, ne = 256, 192, 192, 6
    g = _gen(3)
    rng = np.random.default_rng(3)
    maps = _wfpm(nb, H, W, g)
    te = torch.from_numpy(synth.te_random(nb, ne, rng)).cuda()             # a different echo train for every sample
    assert len({tuple(r) for r in te[:, :, 0].cpu().numpy().round(7).tolist()}) == nb
    tab = ops.gen_tables(te, 1.5)
    S = ops.ideal_fwd(L.MODEL_WFPM, maps, tab, ne)
    pm = maps[:, 2:3].contiguous()
    rho, _ = ops.get_rho_fwd(S, pm, tab)
    assert_close(rho.cpu().numpy(), maps[:, :2].cpu().numpy(), 3e-6, "round trip")
    # the solve is linear in the acquisitions: <J v, u> == <v, J^T u> with J^T from the adjoint kernel
    v = torch.randn(S.shape, device="cuda", generator=g)
    u = torch.randn(rho.shape, device="cuda", generator=g)
    Jv, _ = ops.get_rho_fwd(v, pm, tab)
    JTu, g_pm = ops.get_rho_bwd(v, pm, tab, u, None)
    lhs, rhs = _dot(Jv, u), _dot(v, JTu)
    # fp32 outputs, products added in fp64: rounding errors of random sign add up like sqrt(N)
    assert abs(lhs - rhs) <= 2e-6 * max(abs(lhs), abs(rhs)) + 1e-5 * float(np.sqrt(v.numel())), (lhs, rhs)
    # d<rho, u>/d pm by central differences along one random direction (fp32 solve, fp64 accumulation of the dot products)
    d = torch.randn(pm.shape, device="cuda", generator=g)
    eps = 1e-3
    fp = _dot(ops.get_rho_fwd(v, (pm + eps * d).contiguous(), tab)[0], u)
    fm = _dot(ops.get_rho_fwd(v, (pm - eps * d).contiguous(), tab)[0], u)
    fd, an = (fp - fm) / (2 * eps), _dot(g_pm, d)
    assert abs(fd - an) <= 2e-3 * max(abs(an), 1.0), (fd, an)


@pytest.mark.parametrize("nb", [3, 64])
def test_c4_bipolar_mag_phase_objective(nb):
    H, W, ne = 384, 384, 6
    g = _gen(4)
    rng = np.random.default_rng(4)
    mp = torch.empty((nb, 2, H, W, 4), device="cuda")
    mp[:, 0, :, :, 0:2] = 0.7 * torch.rand((nb, H, W, 2), device="cuda", generator=g)
    mp[:, 0, :, :, 2] = torch.rand((nb, H, W), device="cuda", generator=g)
    mp[:, 0, :, :, 3] = 0
    mp[:, 1, :, :, 0:2] = 0.5 * torch.rand((nb, H, W, 2), device="cuda", generator=g) - 0.25
    mp[:, 1, :, :, 2] = 2 * torch.rand((nb, H, W), device="cuda", generator=g) - 1
    mp[:, 1, :, :, 3] = 0.12 * torch.rand((nb, H, W), device="cuda", generator=g) - 0.06
    mp = (mp * _mask(H, W)).contiguous()
    te = torch.from_numpy(synth.te_random(nb, ne, rng, te_ini_d=0.4e-3, d_te_min=0.9e-3, d_te_d=0.3e-3)).cuda()     # bipolar spacing
    tab = ops.gen_tables(te, 1.5)
    S = ops.ideal_fwd(L.MODEL_MAGPHA, mp, tab, ne)
    assert torch.isfinite(S).all()
    # odd/even echo phase: removing the bipolar channel changes the echoes with opposite signs of the phase increment
    mp0 = mp.clone()
    mp0[:, 1, :, :, 3] = 0
    S0 = ops.ideal_fwd(L.MODEL_MAGPHA, mp0, tab, ne)
    z, z0 = torch.view_as_complex(S.contiguous()), torch.view_as_complex(S0.contiguous())
    ang = torch.angle(z * z0.conj())
    tissue = z0.abs() > 0.05
    b = (4 * np.pi) * mp[:, 1, :, :, 3]
    for e in range(ne):
        want = (b if e % 2 else -b)[tissue[:, e]]
        err = torch.remainder(ang[:, e][tissue[:, e]] - want + np.pi, 2 * np.pi) - np.pi
        assert err.abs().max().item() < 2e-5, e
    noisy = (S + 0.02 * torch.randn(S.shape, device="cuda", generator=g) * (S != 0)).contiguous()
    mp2 = mp.clone()
    mp2[:, 0, :, :, :3] *= 0.97
    loss, gm, shat = ops.ideal_loss(L.MODEL_MAGPHA, mp2, noisy, tab, want_shat=True)
    resid = torch.where(noisy != 0, shat - noisy, torch.zeros_like(shat))
    ref = (resid.double() ** 2).mean().item()
    assert abs(loss.item() - ref) <= 2e-6 * ref
    gref = ops.ideal_bwd(L.MODEL_MAGPHA, mp2, tab, ne, ((2.0 / noisy.numel()) * resid).contiguous())
    assert_close(gm.cpu().numpy(), gref.cpu().numpy(), 2e-5, "fused vs forward + mask + MSE + adjoint")


def test_c5_decoder_shard_streamed_in_chunks():
    nb, H, W, ne = 256, 384, 384, 6            # one rank's shard of the 16 384 slices is 2 048; 256 keeps the host buffers at 1.8 GB
    rng = np.random.default_rng(5)
    one = synth.ffpd_maps(8, H, W, rng)
    maps_h = torch.from_numpy(np.tile(one, (nb // 8, 1, 1, 1, 1))).pin_memory()
    maps_h[:, 1, :, :, 0] *= torch.linspace(0.5, 1.0, nb)[:, None, None]          # every slice differs
    te = synth.te_orig(nb, ne)
    out = igdist.synthesize_to_host(L.MODEL_FFPD, maps_h, te, chunk_nb=48)         # ragged last chunk
    assert tuple(out.shape) == (nb, ne, H, W, 2)
    tab = ops.gen_tables(torch.from_numpy(te).cuda(), 1.5)
    for lo in (0, 96, 240):
        direct = ops.ideal_fwd(L.MODEL_FFPD, maps_h[lo:lo + 16].cuda(), tab[lo:lo + 16].contiguous(), ne)
        assert torch.equal(out[lo:lo + 16], direct.cpu())
    # proton density scales the signal linearly; zero maps give zero signal
    assert out[:, :, 0, 0].abs().max().item() == 0.0
    half = maps_h[:16].clone()
    half[:, 1, :, :, 0] *= 0.5
    s_half = ops.ideal_fwd(L.MODEL_FFPD, half.cuda(), tab[:16].contiguous(), ne)
    assert_close(s_half.cpu().numpy(), 0.5 * out[:16].numpy(), 1e-6, "linear in PD")
