// Device helpers of the uncertainty-aware objective shared by the plain persistent kernel (ig_uq.cu) and the
// TMA-pipelined one (ig_solve.cu).  See ig_uq.cu for the math and the reference lines.
#pragma once
#include "ig_common.cuh"

namespace ig {

constexpr float kVarFloor = 1e-5f;      // tf2gan/loss.py:135
constexpr float kLn2 = 0.6931471805599453f;

// 1 - e^{-x}, x >= 0, without the cancellation the reference's fp32 `1 - exp(-x)` suffers at x ~ 1e-3 (its own error
// there is ~1e-4 relative); also returns e^{-x}
__device__ __forceinline__ float one_minus_exp_neg(float x, float &e) {
    e = fast_ex2(-x * kLog2e);
    if (x < 0.25f) {
        // x - x^2/2 + x^3/6 - ... (7 terms: relative error < 1e-7 below 0.25)
        float s = fmaf(x, -1.0f / 5040.0f, 1.0f / 720.0f);
        s = fmaf(x, -s, 1.0f / 120.0f);
        s = fmaf(x, -s, 1.0f / 24.0f);
        s = fmaf(x, -s, 1.0f / 6.0f);
        s = fmaf(x, -s, 0.5f);
        s = fmaf(x, -s, 1.0f);
        return x * s;
    }
    return 1.0f - e;
}

// SFU approximations without the denormal-range wrappers of rsqrtf / __logf / __fdividef (arguments here are >= 1e-5)
__device__ __forceinline__ float lg2_ftz(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// per-echo uncertainty terms of one voxel: variance, its floor gate, 1/std, and the accumulation of the moment gradients
struct UqAcc {
    float g_sphi, g_mu, g_sr, loss;
};
__device__ __forceinline__ float uq_echo(float te, float a2, float msd, float s_phi, float mu, float s_r, bool rem, UqAcc &acc) {
    const float k = kTwoPi * te, k2 = k * k;
    float ephi;
    const float vphi = one_minus_exp_neg(k2 * s_phi, ephi);
    const float er = rem ? 0.f : fast_ex2(-te * mu * kLog2e) * te * te;
    const float var = fmaf(er, s_r, vphi) * a2;
    const bool gate = var >= kVarFloor;
    const float varc = gate ? var : kVarFloor;
    const float inv_std = rsqrt_ftz(varc);
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(varc));
    acc.loss += fmaf(msd, inv_std, lg * kLn2);
    const float gv = gate ? inv_std * inv_std * fmaf(-0.5f * msd, inv_std, 1.0f) : 0.f;
    const float ga = gv * a2;
    acc.g_sphi = fmaf(ga * k2, ephi, acc.g_sphi);
    acc.g_mu = fmaf(-ga * te, er * s_r, acc.g_mu);
    acc.g_sr = fmaf(ga, er, acc.g_sr);
    return inv_std;
}

// the same for the two packed voxels of a lane pair: FP32 work as f32x2 instructions, transcendentals per lane on the SFU
struct UqAcc2 {
    pk g_sphi, g_mu, g_sr, loss;
};
__device__ __forceinline__ pk uq_echo(float te, pk a2, pk msd, pk s_phi, pk mu, pk s_r, bool rem, UqAcc2 &acc) {
    const float k = kTwoPi * te, k2 = k * k;
    const pk x = vmul(k2, s_phi);
    const pk ephi = fast_ex2(vmul(-kLog2e, x));
    // 1 - e^{-x}: series below 0.25 (see one_minus_exp_neg), 1 - e above
    pk s = vfma(x, splat<pk>(-1.0f / 5040.0f), splat<pk>(1.0f / 720.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 120.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 24.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 6.0f));
    s = vfma(vneg(x), s, splat<pk>(0.5f));
    s = vfma(vneg(x), s, splat<pk>(1.0f));
    const pk series = vmul(x, s), direct = vsub(splat<pk>(1.0f), ephi);
    const pk vphi = mk(x.d.x < 0.25f ? series.d.x : direct.d.x, x.d.y < 0.25f ? series.d.y : direct.d.y);
    pk er = splat<pk>(0.f);
    if (!rem) er = vmul(te * te, fast_ex2(vmul(-te * kLog2e, mu)));
    const pk var = vmul(vfma(er, s_r, vphi), a2);
    const bool g0 = var.d.x >= kVarFloor, g1 = var.d.y >= kVarFloor;
    const pk varc = mk(g0 ? var.d.x : kVarFloor, g1 ? var.d.y : kVarFloor);
    const pk inv_std = mk(rsqrt_ftz(varc.d.x), rsqrt_ftz(varc.d.y));
    float l0, l1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(varc.d.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(varc.d.y));
    acc.loss = vadd(acc.loss, vfma(msd, inv_std, vmul(kLn2, mk(l0, l1))));
    pk gv = vmul(vmul(inv_std, inv_std), vfma(vmul(-0.5f, msd), inv_std, splat<pk>(1.0f)));
    gv = mk(g0 ? gv.d.x : 0.f, g1 ? gv.d.y : 0.f);
    const pk ga = vmul(gv, a2);
    acc.g_sphi = vfma(vmul(k2, ga), ephi, acc.g_sphi);
    acc.g_mu = vfma(vmul(-te, ga), vmul(er, s_r), acc.g_mu);
    acc.g_sr = vfma(ga, er, acc.g_sr);
    return inv_std;
}

// scalar path with the per-component mask (train-IDEAL-unsup.py:218) and the general adjoint of acq_to_acq
template <int NE>
__device__ __forceinline__ void uq_slow_voxel(const SampleTab<NE> &T, const float *acq_b, int ne, int nv, int v, float phi_t, float r2, float s_phi,
                                              float mu, float s_r, bool rem, float r2_sc, UqAcc &acc, float &gphi, float &gr2, cx<float> &rw,
                                              cx<float> &rf) {
    cx<float> y[NE];
    Mod<float> m[NE];
    rw = czero<float>();
    rf = czero<float>();
#pragma unroll 1
    for (int e = 0; e < ne; ++e) {
        m[e] = modulator(T, e, phi_t, r2, 0.f);
        const float2 s = reinterpret_cast<const float2 *>(acq_b + static_cast<size_t>(e) * nv * 2)[v];
        y[e] = demod(m[e], cx<float>{s.x, s.y});
        cmac(rw, T.r[e].pw_re, T.r[e].pw_im, y[e]);
        cmac(rf, T.r[e].pf_re, T.r[e].pf_im, y[e]);
    }
    cx<float> gw = czero<float>(), gf = czero<float>(), X = czero<float>();
#pragma unroll 1
    for (int e = 0; e < ne; ++e) {
        const float2 s = reinterpret_cast<const float2 *>(acq_b + static_cast<size_t>(e) * nv * 2)[v];
        const cx<float> yhat = caffine(rw, T.r[e].c_re, T.r[e].c_im, rf);
        const cx<float> sh = remod(m[e], yhat);
        const cx<float> E{mask_sub(sh.re, s.x), mask_sub(sh.im, s.y)};
        const float msd = E.re * E.re + E.im * E.im;
        const float inv_std = uq_echo(T.r[e].te, yhat.re * yhat.re + yhat.im * yhat.im, msd, s_phi, mu, s_r, rem, acc);
        const cx<float> vv = demod_fwd(m[e], cx<float>{inv_std * E.re, inv_std * E.im});
        gw.re += vv.re;
        gw.im += vv.im;
        cmac(gf, T.r[e].c_re, -T.r[e].c_im, vv);
        const cx<float> q = cmulc(vv, yhat);
        X.re = fmaf(-T.r[e].te, q.re, X.re);
        X.im = fmaf(-T.r[e].te, q.im, X.im);
    }
#pragma unroll 1
    for (int e = 0; e < ne; ++e) {
        cx<float> gy = czero<float>();
        cmac(gy, T.r[e].pw_re, -T.r[e].pw_im, gw);
        cmac(gy, T.r[e].pf_re, -T.r[e].pf_im, gf);
        const cx<float> q = cmulc(gy, y[e]);
        X.re = fmaf(T.r[e].te, q.re, X.re);
        X.im = fmaf(T.r[e].te, q.im, X.im);
    }
    gphi = kTwoPi * kFmSc * X.im;      // caller applies 2 / N
    gr2 = r2_sc * X.re;
}

// ------------------------------------------------------------------------------------------------
// Rician objective (VarMeanSquaredErrorR2, tf2gan/loss.py:143-162): exponentially scaled Bessel terms from a table.
// log i0e(z) and om(z) = 1 - I1(z) / I0(z) are smooth on a logarithmic scale: the float's exponent and top two mantissa
// bits pick one of four polynomial pieces per octave, the remaining mantissa bits are the local variable t in [0, 1).
// 136 rows of 12 floats (degree 4 for log i0e, degree 5 for om) generated by tools/gen_bessel_coeffs.py against
// scipy.special; float32 evaluation error 1.9e-7 / 1.8e-7 relative.  No branch on z, so no divergence between lanes
// (the earlier power-series / asymptotic-fit pair cost ~100 scalar instructions per element in mixed warps, this ~25).
// ------------------------------------------------------------------------------------------------
constexpr int kBesselRows = 136, kBesselRowBase = 468;          // rows for 2^-10 <= z < 2^24; (bits >> 21) - 4 (127 - 10)
__device__ __align__(16) const float kBesselTab[kBesselRows * 12] = {
#include "ig_bessel_tab.inc"
};

// cooperative copy of the table into shared memory (16-byte rows of a float4 array); caller synchronises
__device__ __forceinline__ void stage_bessel_table(float4 *smem_tab) {
    const float4 *src = reinterpret_cast<const float4 *>(kBesselTab);
    for (int i = threadIdx.x; i < kBesselRows * 3; i += blockDim.x) smem_tab[i] = src[i];
}

// z >= 0 -> log(i0e(z)) and om = 1 - I1(z) / I0(z)
__device__ __forceinline__ void bessel_terms(const float4 *__restrict__ tab, float z, float &log_i0e, float &om) {
    const float zc = fminf(fmaxf(z, 0.0009765625f), 16777215.0f);
    const unsigned bits = __float_as_uint(zc);
    const float4 *row = tab + ((bits >> 21) - kBesselRowBase) * 3;
    const float t = __uint_as_float(((bits << 2) & 0x007fffffu) | 0x3f800000u) - 1.0f;
    const float4 a = row[0], b = row[1], c = row[2];
    const float p = fmaf(fmaf(fmaf(fmaf(b.x, t, a.w), t, a.z), t, a.y), t, a.x);
    const float q = fmaf(fmaf(fmaf(fmaf(fmaf(c.z, t, c.y), t, c.x), t, b.w), t, b.z), t, b.y);
    const bool tiny = z < 0.0009765625f;                  // incl. z = 0 (masked / background): I0 = 1, I1 = 0
    log_i0e = tiny ? z * fmaf(0.25f, z, -1.0f) : p;       // -z + z^2 / 4 (next term z^4 / 64)
    om = tiny ? fmaf(-0.5f, z, 1.0f) : q;                 // 1 - z / 2   (next term z^3 / 16)
}

// One (echo, voxel) term of the Rician objective and its derivatives.  y = |A_e| observed, nu = |S_hat_e| (0 where masked),
// var = V_e |yhat_e|^2.  With s2 = max(var, 1e-5), z = y nu / s2:
//   -loglik = -[y > 1e-5] log y + log s2 + (y - nu)^2 / (2 s2) - log i0e(z)        ((y^2 + nu^2) / (2 s2) - z, without the cancellation)
//   d/d nu  = ((nu - y) + y om) / s2,    d/d s2 = [var >= 1e-5] (1 - ((y - nu)^2 / 2 + y nu om) / s2) / s2,    om = 1 - I1/I0 (z)
// Returns d/d nu (0 where masked); accumulates the loss and the moment gradients like uq_echo.  The two logarithms are one:
// log s2 - [y > 1e-5] log y = log(s2 / y') with y' = y or 1.
__device__ __forceinline__ float rician_echo(const float4 *__restrict__ btab, float te, float a2, float y, float nu_unmasked, bool keep, float s_phi,
                                             float mu, float s_r, bool rem, UqAcc &acc) {
    const float k = kTwoPi * te, k2 = k * k;
    float ephi;
    const float vphi = one_minus_exp_neg(k2 * s_phi, ephi);
    const float er = rem ? 0.f : fast_ex2(-te * mu * kLog2e) * te * te;
    const float var = fmaf(er, s_r, vphi) * a2;
    const bool gate = var >= kVarFloor;
    const float s2 = gate ? var : kVarFloor;
    const float inv = rcp_ftz(s2);
    const float nu = keep ? nu_unmasked : 0.f;
    const float z = y * nu * inv;
    float li0e, om;
    bessel_terms(btab, z, li0e, om);
    const float diff = y - nu;
    const float hd = 0.5f * diff * diff;
    const float ratio = y > 1e-5f ? s2 * rcp_ftz(y) : s2;
    acc.loss += fmaf(hd, inv, fmaf(kLn2, lg2_ftz(ratio), -li0e));
    const float gv = gate ? inv * fmaf(-fmaf(y * nu, om, hd), inv, 1.0f) : 0.f;
    const float ga = gv * a2;
    acc.g_sphi = fmaf(ga * k2, ephi, acc.g_sphi);
    acc.g_mu = fmaf(-ga * te, er * s_r, acc.g_mu);
    acc.g_sr = fmaf(ga, er, acc.g_sr);
    return keep ? fmaf(y, om, -diff) * inv : 0.f;
}

// the same for the two packed voxels of a lane pair (f32x2 arithmetic; SFU operations and the table rows per lane).
// ys = observed magnitude with the mask in its sign bit (negative = masked).
__device__ __forceinline__ pk rician_echo(const float4 *__restrict__ btab, float te, pk a2, pk ys, pk nu_unmasked, pk s_phi, pk mu, pk s_r, bool rem,
                                          UqAcc2 &acc) {
    const float k = kTwoPi * te, k2 = k * k;
    const pk x = vmul(k2, s_phi);
    const pk ephi = fast_ex2(vmul(-kLog2e, x));
    pk s = vfma(x, splat<pk>(-1.0f / 5040.0f), splat<pk>(1.0f / 720.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 120.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 24.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 6.0f));
    s = vfma(vneg(x), s, splat<pk>(0.5f));
    s = vfma(vneg(x), s, splat<pk>(1.0f));
    const pk series = vmul(x, s), direct = vsub(splat<pk>(1.0f), ephi);
    const pk vphi = mk(x.d.x < 0.25f ? series.d.x : direct.d.x, x.d.y < 0.25f ? series.d.y : direct.d.y);
    pk er = splat<pk>(0.f);
    if (!rem) er = vmul(te * te, fast_ex2(vmul(-te * kLog2e, mu)));
    const pk var = vmul(vfma(er, s_r, vphi), a2);
    const bool g0 = var.d.x >= kVarFloor, g1 = var.d.y >= kVarFloor;
    const pk s2 = mk(g0 ? var.d.x : kVarFloor, g1 ? var.d.y : kVarFloor);
    const pk inv = mk(rcp_ftz(s2.d.x), rcp_ftz(s2.d.y));
    const bool k0 = !signbit(ys.d.x), k1 = !signbit(ys.d.y);
    const pk y = mk(fabsf(ys.d.x), fabsf(ys.d.y));
    const pk nu = mk(k0 ? nu_unmasked.d.x : 0.f, k1 ? nu_unmasked.d.y : 0.f);
    const pk ynu = vmul(y, nu), z = vmul(ynu, inv);
    pk li0e, om;
    bessel_terms(btab, z.d.x, li0e.d.x, om.d.x);
    bessel_terms(btab, z.d.y, li0e.d.y, om.d.y);
    const pk diff = vsub(y, nu);
    const pk hd = vmul(vmul(0.5f, diff), diff);
    const pk ratio = mk(y.d.x > 1e-5f ? s2.d.x * rcp_ftz(y.d.x) : s2.d.x, y.d.y > 1e-5f ? s2.d.y * rcp_ftz(y.d.y) : s2.d.y);
    const pk lg = mk(lg2_ftz(ratio.d.x), lg2_ftz(ratio.d.y));
    acc.loss = vadd(acc.loss, vfma(hd, inv, vfma(splat<pk>(kLn2), lg, vneg(li0e))));
    pk gv = vmul(inv, vfma(vneg(vfma(ynu, om, hd)), inv, splat<pk>(1.0f)));
    gv = mk(g0 ? gv.d.x : 0.f, g1 ? gv.d.y : 0.f);
    const pk ga = vmul(gv, a2);
    acc.g_sphi = vfma(vmul(k2, ga), ephi, acc.g_sphi);
    acc.g_mu = vfma(vmul(-te, ga), vmul(er, s_r), acc.g_mu);
    acc.g_sr = vfma(ga, er, acc.g_sr);
    const pk g = vmul(vfma(y, om, vneg(diff)), inv);
    return mk(k0 ? g.d.x : 0.f, k1 ? g.d.y : 0.f);
}

// host: the same objective on the TMA ring of ig_solve.cu (IG_E_UNSUPPORTED when the shape is not covered)
int a2a_uq_loss_ring(const float *acqs, const float *pm, long pm_bstride, const float *phi_var, const float *r2_mean, const float *r2_var,
                     const float *tab, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm, float *g_phi_var, float *g_r2_mean,
                     float *g_r2_var, float *rho, float *loss, void *scratch, cudaStream_t st);

}  // namespace ig
