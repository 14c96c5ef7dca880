// Least-squares water/fat solve (get_rho), project-and-resynthesise (acq_to_acq), their adjoints, and the
// fused config-2 physics objective.
//
// Replaces get_rho / acq_to_acq of the reference (/root/reference/wflib/IDEAL_model.py:527-624, 142-200)
// and TF autodiff through them (train-IDEAL-unsup.py:214-218,236,255; train-IDEAL-TEaug.py:304).
// Per voxel, with y_e = Wm_e S_e the demodulated echoes (Wm_e = exp(+te_e R - i(2 pi te_e phi + s_e beta))):
//     rho = M^+ y          yhat = M rho          S_hat_e = Wp_e yhat_e         (Wp_e = 1 / Wm_e)
// The pseudo-inverse M^+ depends only on the sample's echo times, so it is a shared-memory table and the
// "solve" is a 2 x ne complex contraction held in registers.
#include <cuda.h>
#include <stdlib.h>

#include "ig_uq.cuh"

namespace ig {

// The fused objectives form the phase in radians and hand it to sin.approx / cos.approx without the exact turn reduction: 4 fewer FP32
// lane-operations per voxel-echo on a kernel bound by the FMA pipe (measured 0.1192 -> 0.1131 ms masked, 0.1395 -> 0.1297 ms
// unmasked at 64 x 384 x 384 x 6; error against the fp64 oracle unchanged, profiles/history_r02.md).  -DIG_PHASE_RAD=0 restores it.
#ifndef IG_PHASE_RAD
#define IG_PHASE_RAD 1
#endif
constexpr bool kPhaseRad = IG_PHASE_RAD != 0;

struct SolveParams {
    const float *acqs, *pm, *bip, *tab;
    const float *g_rho, *g_demod, *g_shat;
    float *rho, *demod, *shat;
    float *g_acqs, *g_pm, *g_bip;
    const float *phi_var, *r2_mean, *r2_var;       // uncertainty-aware objective: (nb, nv) moment maps (r2_* NULL = rem_R2)
    float *g_phi_var, *g_r2_mean, *g_r2_var;
    float *loss;
    float *pdff, *r2s;                             // get_rho epilogue: PDFF (nb, nv) and R2* [1/s] (nb, nv) maps, optional
    int pdff_mode;
    void *scratch;
    long pm_bstride, bip_bstride;
    int nb, ne, nv, flags;
    int tile_stride;      // persistent kernels: visit a sample's tiles in the order (j * tile_stride) % tiles_per_sample
    float r2_sc, inv_n;
    PeerPub peer;         // a2a objective only: scalar exchange with the other ranks in the finishing block (off when boxes == nullptr)
};

// echo loads / stores for both acquisition layouts
template <typename V, bool FLAT> __device__ __forceinline__ cx<V> ld_echo(const float *acq_b, int e, int ne, int nv, int v0) {
    if constexpr (!FLAT) {
        return ld_cx(acq_b + static_cast<size_t>(e) * nv * 2, v0, V{});
    } else {
        cx<V> z;
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            const float2 t = __ldcs(reinterpret_cast<const float2 *>(acq_b + (static_cast<size_t>(v0) + l) * 2 * ne) + e);
            lane_set(z.re, l, t.x);
            lane_set(z.im, l, t.y);
        }
        return z;
    }
}
template <typename V, bool FLAT> __device__ __forceinline__ void st_echo(float *acq_b, int e, int ne, int nv, int v0, const cx<V> &z) {
    if constexpr (!FLAT) {
        st_cx(acq_b + static_cast<size_t>(e) * nv * 2, v0, z);
    } else {
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l)
            __stcs(reinterpret_cast<float2 *>(acq_b + (static_cast<size_t>(v0) + l) * 2 * ne) + e, make_float2(lane_get(z.re, l), lane_get(z.im, l)));
    }
}

// (phi map, R2 map) of a voxel; the flat layout stores (R2*, phi) (IDEAL_model.py:559-560)
template <typename V, bool FLAT> __device__ __forceinline__ void ld_pm(const float *pm_b, int v0, V &phi_t, V &r2) {
    const cx<V> m = ld_cx(pm_b, v0, V{});
    if constexpr (FLAT) { phi_t = m.im; r2 = m.re; } else { phi_t = m.re; r2 = m.im; }
}

// half-angle helpers for the phase-constrained solve: theta = 0.5 arg(z) -> (cos theta, sin theta)
__device__ __forceinline__ void half_angle(float zr, float zi, float &ct, float &st) {
    const float th = 0.5f * atan2f(zi, zr);
    sincosf(th, &st, &ct);
}

// =================================================================================================
// get_rho forward
// =================================================================================================
template <int NE, typename V, bool FLAT> __global__ void __launch_bounds__(kThreads) get_rho_fwd_kernel(const SolveParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v0 >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
    cx<V> S[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e)
        if (e < ne) S[e] = ld_echo<V, FLAT>(p.acqs + acq_b, e, ne, nv, v0);
    V phi_t, r2, bturn = splat<V>(0.f);
    ld_pm<V, FLAT>(p.pm + b * p.pm_bstride, v0, phi_t, r2);
    if (p.bip) bturn = vmul(0.5f, ld_cx(p.bip + b * p.bip_bstride, v0, V{}).re);
    cx<V> rw = czero<V>(), rf = czero<V>();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const Mod<V> m = modulator(T, e, phi_t, r2, bturn);
            const cx<V> y = demod(m, S[e]);
            if (p.demod) st_cx(p.demod + acq_b + static_cast<size_t>(e) * nv * 2, v0, y);
            cmac(rw, T.r[e].pw_re, T.r[e].pw_im, y);
            cmac(rf, T.r[e].pf_re, T.r[e].pf_im, y);
        }
    }
    if (p.flags & IG_F_PHASE_CONSTRAINT) {
        // theta = 0.5 arg(rho_W^2 + rho_F^2) (H^+ = Re(M^+ M)^+ is the identity to rounding, :64-68,584-592);
        // rho_s <- Re(rho_s e^{-i theta}) e^{i theta}
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            const float wr = lane_get(rw.re, l), wi = lane_get(rw.im, l), fr = lane_get(rf.re, l), fi = lane_get(rf.im, l);
            float ct, st;
            half_angle(wr * wr - wi * wi + fr * fr - fi * fi, 2.f * (wr * wi + fr * fi), ct, st);
            const float mw = wr * ct + wi * st, mf = fr * ct + fi * st;
            lane_set(rw.re, l, mw * ct); lane_set(rw.im, l, mw * st);
            lane_set(rf.re, l, mf * ct); lane_set(rf.im, l, mf * st);
        }
    }
    if (p.pdff || p.r2s) {
        // PDFF / R2* maps straight from the registers (ROI-analysis.py:301-306,344-354; gen_LDM_dataset.py:217-218,226-227):
        // mode 0 |F| / |W + F|, 1 |F| / (|W| + |F|), 2 magnitude-discriminated; 0/0 -> 0.  The ratio is scale-free, so rho_sc drops out.
        V pd = splat<V>(0.f);
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            const float wr = lane_get(rw.re, l), wi = lane_get(rw.im, l), fr = lane_get(rf.re, l), fi = lane_get(rf.im, l);
            const float wa = sqrtf(wr * wr + wi * wi), fa = sqrtf(fr * fr + fi * fi);
            float r;
            if (p.pdff_mode == 1) {
                r = fa / (wa + fa);
            } else {
                const float wf = sqrtf((wr + fr) * (wr + fr) + (wi + fi) * (wi + fi));
                r = (p.pdff_mode == 0 || fa >= wa) ? fa / wf : 1.0f - wa / wf;
            }
            lane_set(pd, l, (isnan(r) || isinf(r)) ? 0.f : r);
        }
        if (p.pdff) st_real(p.pdff + static_cast<size_t>(b) * nv, v0, pd);
        if (p.r2s) st_real(p.r2s + static_cast<size_t>(b) * nv, v0, vmul(p.r2_sc, r2));
    }
    const float inv = 1.0f / kRhoSc;
    rw = cx<V>{vmul(inv, rw.re), vmul(inv, rw.im)};
    rf = cx<V>{vmul(inv, rf.re), vmul(inv, rf.im)};
    if constexpr (!FLAT) {
        float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
        st_cx(rho_b, v0, rw);
        st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, rf);
    } else {
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l)
            __stcs(reinterpret_cast<float4 *>(p.rho + (static_cast<size_t>(b) * nv + v0 + l) * 4),
                   make_float4(lane_get(rw.re, l), lane_get(rw.im, l), lane_get(rf.re, l), lane_get(rf.im, l)));
    }
}

// =================================================================================================
// get_rho backward.  With gy_e = sum_s conj(M^+[s,e]) g_rho_s / rho_sc + g_demod_e:
//   dL/dS_e = conj(Wm_e) gy_e ;  X = sum_e te_e conj(gy_e) y_e ;  B = sum_e s_e Im(conj(gy_e) y_e)
//   dL/dphi~ = 2 pi fm_sc Im X ;  dL/dR~ = r2_sc Re X ;  dL/db~ = pi B
// =================================================================================================
template <int NE, typename V, bool FLAT> __global__ void __launch_bounds__(kThreads) get_rho_bwd_kernel(const SolveParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v0 >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
    cx<V> S[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e)
        if (e < ne) S[e] = ld_echo<V, FLAT>(p.acqs + acq_b, e, ne, nv, v0);
    V phi_t, r2, bturn = splat<V>(0.f);
    ld_pm<V, FLAT>(p.pm + b * p.pm_bstride, v0, phi_t, r2);
    if (p.bip) bturn = vmul(0.5f, ld_cx(p.bip + b * p.bip_bstride, v0, V{}).re);
    cx<V> gw = czero<V>(), gf = czero<V>();
    if (p.g_rho) {
        const float inv = 1.0f / kRhoSc;
        if constexpr (!FLAT) {
            const float *g_b = p.g_rho + static_cast<size_t>(b) * 2 * nv * 2;
            gw = ld_cx(g_b, v0, V{});
            gf = ld_cx(g_b + static_cast<size_t>(nv) * 2, v0, V{});
        } else {
#pragma unroll
            for (int l = 0; l < lanes<V>::n; ++l) {
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(p.g_rho + (static_cast<size_t>(b) * nv + v0 + l) * 4));
                lane_set(gw.re, l, t.x); lane_set(gw.im, l, t.y); lane_set(gf.re, l, t.z); lane_set(gf.im, l, t.w);
            }
        }
        gw = cx<V>{vmul(inv, gw.re), vmul(inv, gw.im)};
        gf = cx<V>{vmul(inv, gf.re), vmul(inv, gf.im)};
    }
    if (p.flags & IG_F_PHASE_CONSTRAINT) {
        // Adjoint of rho_s <- Re(rho_s e^{-i theta}) e^{i theta}, theta = 0.5 arg(z), z = rho_W^2 + rho_F^2 (IDEAL_model.py:584-592): the
        // upstream on the constrained estimate is carried back onto the unconstrained one, then the plain adjoint below applies.
        //   a_s = Re(conj(G_s) u), b_s = -Im(conj(G_s) u), m_s = Re(rho_s conj(u)), n_s = Im(rho_s conj(u)),  u = e^{i theta}
        //   G_s <- a_s u + (sum_s a_s n_s + b_s m_s) / |z|^2 . i z conj(rho_s)          (0 for the second term where z = 0)
        cx<V> rw = czero<V>(), rf = czero<V>();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const cx<V> y = demod(modulator(T, e, phi_t, r2, bturn), S[e]);
                cmac(rw, T.r[e].pw_re, T.r[e].pw_im, y);
                cmac(rf, T.r[e].pf_re, T.r[e].pf_im, y);
            }
        }
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            const float wr = lane_get(rw.re, l), wi = lane_get(rw.im, l), fr = lane_get(rf.re, l), fi = lane_get(rf.im, l);
            const float gwr = lane_get(gw.re, l), gwi = lane_get(gw.im, l), gfr = lane_get(gf.re, l), gfi = lane_get(gf.im, l);
            const float zr = wr * wr - wi * wi + fr * fr - fi * fi, zi = 2.f * (wr * wi + fr * fi);
            float ct, st;
            half_angle(zr, zi, ct, st);
            const float aw = gwr * ct + gwi * st, bw = gwi * ct - gwr * st, mw = wr * ct + wi * st, nw = wi * ct - wr * st;
            const float af = gfr * ct + gfi * st, bf = gfi * ct - gfr * st, mf = fr * ct + fi * st, nf = fi * ct - fr * st;
            const float z2 = zr * zr + zi * zi;
            const float k = z2 > 0.f ? (aw * nw + bw * mw + af * nf + bf * mf) / z2 : 0.f;
            lane_set(gw.re, l, aw * ct - k * (zi * wr - zr * wi)); lane_set(gw.im, l, aw * st + k * (zr * wr + zi * wi));
            lane_set(gf.re, l, af * ct - k * (zi * fr - zr * fi)); lane_set(gf.im, l, af * st + k * (zr * fr + zi * fi));
        }
    }
    cx<V> X = czero<V>();
    V B = splat<V>(0.f);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const Mod<V> m = modulator(T, e, phi_t, r2, bturn);
            const cx<V> y = demod(m, S[e]);
            cx<V> gy = czero<V>();
            cmac(gy, T.r[e].pw_re, -T.r[e].pw_im, gw);
            cmac(gy, T.r[e].pf_re, -T.r[e].pf_im, gf);
            if (p.g_demod) {
                const cx<V> gd = ld_cx(p.g_demod + acq_b + static_cast<size_t>(e) * nv * 2, v0, V{});
                gy.re = vadd(gy.re, gd.re);
                gy.im = vadd(gy.im, gd.im);
            }
            if (p.g_acqs) st_echo<V, FLAT>(p.g_acqs + acq_b, e, ne, nv, v0, remod_inv(m, gy));
            const cx<V> q = cmulc(gy, y);
            X.re = vfma(T.r[e].te, q.re, X.re);
            X.im = vfma(T.r[e].te, q.im, X.im);
            B = vfma(T.r[e].sgn, q.im, B);
        }
    }
    const V gphi = vmul(kTwoPi * kFmSc, X.im), gr2 = vmul(p.r2_sc, X.re);
    float *gpm_b = p.g_pm + static_cast<size_t>(b) * nv * 2;
    st_cx(gpm_b, v0, FLAT ? cx<V>{gr2, gphi} : cx<V>{gphi, gr2});
    if (p.g_bip) st_cx(p.g_bip + static_cast<size_t>(b) * nv * 2, v0, cx<V>{vmul(0.5f * kTwoPi, B), splat<V>(0.f)});
}

// =================================================================================================
// acq_to_acq forward: rho_hat / rho_sc and S_hat (or |S_hat|)
// =================================================================================================
template <int NE, typename V> __global__ void __launch_bounds__(kThreads) a2a_fwd_kernel(const SolveParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v0 >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
    cx<V> S[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e)
        if (e < ne) S[e] = ld_cx(p.acqs + acq_b + static_cast<size_t>(e) * nv * 2, v0, V{});
    V phi_t, r2;
    ld_pm<V, false>(p.pm + b * p.pm_bstride, v0, phi_t, r2);
    const V zero = splat<V>(0.f);
    Mod<V> m[NE];
    cx<V> rw = czero<V>(), rf = czero<V>();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            m[e] = modulator(T, e, phi_t, r2, zero);
            const cx<V> y = demod(m[e], S[e]);
            cmac(rw, T.r[e].pw_re, T.r[e].pw_im, y);
            cmac(rf, T.r[e].pf_re, T.r[e].pf_im, y);
        }
    }
    if (p.rho) {
        const float inv = 1.0f / kRhoSc;
        float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
        st_cx(rho_b, v0, cx<V>{vmul(inv, rw.re), vmul(inv, rw.im)});
        st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, cx<V>{vmul(inv, rf.re), vmul(inv, rf.im)});
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const cx<V> sh = remod(m[e], caffine(rw, T.r[e].c_re, T.r[e].c_im, rf));
            if (p.flags & IG_F_ONLY_MAG) {
                st_real(p.shat + static_cast<size_t>(b) * ne * nv + static_cast<size_t>(e) * nv, v0, vsqrt(vfma(sh.re, sh.re, vmul(sh.im, sh.im))));
            } else {
                st_cx(p.shat + acq_b + static_cast<size_t>(e) * nv * 2, v0, sh);
            }
        }
    }
}

// =================================================================================================
// acq_to_acq backward.  Upstream G_e on S_hat (or g_e on |S_hat|: G = g S_hat / |S_hat|) and Gamma_s on rho/rho_sc.
//   v_e = conj(Wp_e) G_e ; g_rho = Gamma / rho_sc + M^H v ; gy = (M^+)^H g_rho ; dL/dS_e = conj(Wm_e) gy_e
//   X = sum_e te_e (conj(gy_e) y_e - conj(v_e) yhat_e) ; dL/dphi~ = 2 pi fm_sc Im X ; dL/dR~ = r2_sc Re X
// Both sums collapse onto per-voxel accumulators, so nothing per echo stays live and the packed lanes fit:
//   sum_e te_e conj(gy_e) y_e    = conj(g_w) t_w + conj(g_f) t_f        (t = M^+ (te . y), as in the fused objective)
//   sum_e te_e conj(v_e) yhat_e  = conj(a_w) rho_w + conj(a_f) rho_f    (a_w = sum te_e v_e, a_f = sum te_e conj(c_e) v_e)
// One loop over the echoes forms y_e and v_e from the same modulator; only dL/dS (optional) needs a second loop,
// which recomputes the modulator instead of keeping six of them in registers.
// =================================================================================================
template <int NE, typename V> __global__ void __launch_bounds__(kThreads) a2a_bwd_kernel(const SolveParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v0 >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
    const size_t plane = static_cast<size_t>(nv) * 2;
    const V zero = splat<V>(0.f);
    const bool only_mag = p.flags & IG_F_ONLY_MAG;
    V phi_t, r2;
    ld_pm<V, false>(p.pm + b * p.pm_bstride, v0, phi_t, r2);
    cx<V> rw = czero<V>(), rf = czero<V>(), tw = czero<V>(), tf = czero<V>();
    cx<V> gw = czero<V>(), gf = czero<V>(), aw = czero<V>(), af = czero<V>();
    [[maybe_unused]] V dec[NE];          // decay per echo, kept for the magnitude upstream only
    const bool direct = p.g_shat && !only_mag;
    // every echo and upstream load is issued before the math: 2 ne 16-byte loads in flight per thread
    cx<V> S[NE], G[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            S[e] = ld_cx(p.acqs + acq_b + e * plane, v0, V{});
            if (direct) G[e] = ld_cx(p.g_shat + acq_b + e * plane, v0, V{});
        }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const EchoRec R = T.r[e];
            const Mod<V> m = modulator_rec<V, false>(R, phi_t, r2, zero);
            const cx<V> y = demod(m, S[e]);
            cmac(rw, R.pw_re, R.pw_im, y);
            cmac(rf, R.pf_re, R.pf_im, y);
            cmac(tw, R.tpw_re, R.tpw_im, y);
            cmac(tf, R.tpf_re, R.tpf_im, y);
            dec[e] = m.d;
            if (direct) {
                const cx<V> v = demod_fwd(m, G[e]);
                gw.re = vadd(gw.re, v.re);
                gw.im = vadd(gw.im, v.im);
                cmac(gf, R.c_re, -R.c_im, v);
                aw.re = vfma(R.te, v.re, aw.re);
                aw.im = vfma(R.te, v.im, aw.im);
                cmac(af, R.te * R.c_re, -R.te * R.c_im, v);
            }
        }
    }
    if (p.g_shat && only_mag) {
        // |S_hat| = d |yhat| ; G = g S_hat / |S_hat| ; v = conj(Wp) G = g d yhat / |yhat|  (0 where |yhat| = 0)
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const EchoRec R = T.r[e];
                const cx<V> yhat = caffine(rw, R.c_re, R.c_im, rf);
                const V g = ld_real(p.g_shat + static_cast<size_t>(b) * ne * nv + static_cast<size_t>(e) * nv, v0, V{});
                V sc = zero;
#pragma unroll
                for (int l = 0; l < lanes<V>::n; ++l) {
                    const float a2 = lane_get(yhat.re, l) * lane_get(yhat.re, l) + lane_get(yhat.im, l) * lane_get(yhat.im, l);
                    lane_set(sc, l, a2 > 0.f ? lane_get(g, l) * lane_get(dec[e], l) * rsqrtf(a2) : 0.f);
                }
                const cx<V> v = cscale(sc, yhat);
                gw.re = vadd(gw.re, v.re);
                gw.im = vadd(gw.im, v.im);
                cmac(gf, R.c_re, -R.c_im, v);
                aw.re = vfma(R.te, v.re, aw.re);
                aw.im = vfma(R.te, v.im, aw.im);
                cmac(af, R.te * R.c_re, -R.te * R.c_im, v);
            }
        }
    }
    if (p.g_rho) {
        const float inv = 1.0f / kRhoSc;
        const float *g_b = p.g_rho + static_cast<size_t>(b) * 2 * nv * 2;
        const cx<V> a = ld_cx(g_b, v0, V{}), c = ld_cx(g_b + plane, v0, V{});
        gw.re = vfma(inv, a.re, gw.re); gw.im = vfma(inv, a.im, gw.im);
        gf.re = vfma(inv, c.re, gf.re); gf.im = vfma(inv, c.im, gf.im);
    }
    // X = conj(g_w) t_w + conj(g_f) t_f - conj(a_w) rho_w - conj(a_f) rho_f
    cx<V> X = cmulc(gw, tw);
    const cx<V> x1 = cmulc(gf, tf), x2 = cmulc(aw, rw), x3 = cmulc(af, rf);
    X.re = vsub(vadd(X.re, x1.re), vadd(x2.re, x3.re));
    X.im = vsub(vadd(X.im, x1.im), vadd(x2.im, x3.im));
    st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, cx<V>{vmul(kTwoPi * kFmSc, X.im), vmul(p.r2_sc, X.re)});
    if (p.g_acqs) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const EchoRec R = T.r[e];
                const Mod<V> m = modulator_rec<V, false>(R, phi_t, r2, zero);
                cx<V> gy = czero<V>();
                cmac(gy, R.pw_re, -R.pw_im, gw);
                cmac(gy, R.pf_re, -R.pf_im, gf);
                st_cx(p.g_acqs + acq_b + e * plane, v0, remod_inv(m, gy));
            }
        }
    }
}

// =================================================================================================
// Fused config-2 objective: loss = inv_n sum |mask(S_hat) - A|^2 and d loss / d (phi~, R~).
//
// Fast path (every voxel of the warp is either fully non-zero or fully zero, the only cases real,
// body-masked data produce): the residual is taken in the demodulated frame, r = yhat - y, so that
//     S_hat - A = Wp r,     loss = sum_e d_e^2 |r_e|^2,     K = sum_e d_e^2 conj(r_e) (te_e yhat_e - h_e),
//     h = M M^+ (te . y),   dL/dphi~ = -fm_sc (4 pi / N) Im K,   dL/dR~ = -r2_sc (2 / N) Re K
// which needs neither Wp nor a second contraction with M^+ (derivation in DESIGN.md).
// Slow path (some component exactly zero while others are not): the mask is applied per real/imag
// component exactly as the reference does (train-IDEAL-unsup.py:218) with the general adjoint.
// =================================================================================================
template <int NE>
__device__ __forceinline__ void a2a_loss_slow_voxel(const SampleTab<NE> &T, const float *acq_b, int ne, int nv, int v, float phi_t, float r2,
                                                 float r2_sc, float &loss, float &gphi, float &gr2) {
    cx<float> y[NE], vv[NE];
    Mod<float> m[NE];
    cx<float> rw = czero<float>(), rf = czero<float>();
#pragma unroll 1
    for (int e = 0; e < ne; ++e) {
        m[e] = modulator(T, e, phi_t, r2, 0.f);
        const float2 s = reinterpret_cast<const float2 *>(acq_b + static_cast<size_t>(e) * nv * 2)[v];
        y[e] = demod(m[e], cx<float>{s.x, s.y});
        cmac(rw, T.r[e].pw_re, T.r[e].pw_im, y[e]);
        cmac(rf, T.r[e].pf_re, T.r[e].pf_im, y[e]);
    }
    cx<float> gw = czero<float>(), gf = czero<float>(), X = czero<float>();
    loss = 0.f;
#pragma unroll 1
    for (int e = 0; e < ne; ++e) {
        const float2 s = reinterpret_cast<const float2 *>(acq_b + static_cast<size_t>(e) * nv * 2)[v];
        const cx<float> yhat = caffine(rw, T.r[e].c_re, T.r[e].c_im, rf);
        const cx<float> sh = remod(m[e], yhat);
        const cx<float> E{mask_sub(sh.re, s.x), mask_sub(sh.im, s.y)};
        loss += E.re * E.re + E.im * E.im;
        vv[e] = demod_fwd(m[e], E);
        gw.re += vv[e].re;
        gw.im += vv[e].im;
        cmac(gf, T.r[e].c_re, -T.r[e].c_im, vv[e]);
        const cx<float> q = cmulc(vv[e], yhat);
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

// raw echoes of the thread's voxels exactly as loaded (one 8- or 16-byte access per echo)
template <typename V> struct RawEcho;
template <> struct RawEcho<float> { float2 v; };
template <> struct RawEcho<pk> { float4 v; };
__device__ __forceinline__ RawEcho<float> ld_raw(const float *plane, int v0, float) {
    RawEcho<float> r; r.v = __ldcs(reinterpret_cast<const float2 *>(plane) + v0); return r;
}
__device__ __forceinline__ RawEcho<pk> ld_raw(const float *plane, int v0, pk) {
    RawEcho<pk> r; r.v = __ldcs(reinterpret_cast<const float4 *>(plane) + (v0 >> 1)); return r;
}
// running min / max of |component| per voxel: a voxel is "ragged" (needs the per-component mask) iff
// min == 0 < max.  FMNMX runs on the ALU pipe, which this kernel otherwise leaves idle.
struct AbsRange { float lo0, hi0, lo1, hi1; };
__device__ __forceinline__ void abs_range(AbsRange &a, const RawEcho<float> &r) {
    a.lo0 = fminf(fminf(a.lo0, fabsf(r.v.x)), fabsf(r.v.y));
    a.hi0 = fmaxf(fmaxf(a.hi0, fabsf(r.v.x)), fabsf(r.v.y));
}
__device__ __forceinline__ void abs_range(AbsRange &a, const RawEcho<pk> &r) {
    a.lo0 = fminf(fminf(a.lo0, fabsf(r.v.x)), fabsf(r.v.y));
    a.hi0 = fmaxf(fmaxf(a.hi0, fabsf(r.v.x)), fabsf(r.v.y));
    a.lo1 = fminf(fminf(a.lo1, fabsf(r.v.z)), fabsf(r.v.w));
    a.hi1 = fmaxf(fmaxf(a.hi1, fabsf(r.v.z)), fabsf(r.v.w));
}
__device__ __forceinline__ bool is_ragged(const AbsRange &a) { return (a.lo0 == 0.f && a.hi0 > 0.f) || (a.lo1 == 0.f && a.hi1 > 0.f); }
__device__ __forceinline__ cx<float> raw_cx(const RawEcho<float> &r) { return cx<float>{r.v.x, r.v.y}; }
__device__ __forceinline__ cx<pk> raw_cx(const RawEcho<pk> &r) { return cx<pk>{mk(r.v.x, r.v.z), mk(r.v.y, r.v.w)}; }
// y = (cd - i sd) * S, written lane by lane with scalar FMAs straight from the loaded (re, im, re, im) tuple:
// this de-interleaves into the packed layout for free instead of paying register moves before FFMA2.
__device__ __forceinline__ cx<float> demod_raw(float cd, float sd, const RawEcho<float> &r) {
    return cx<float>{fmaf(sd, r.v.y, cd * r.v.x), fmaf(-sd, r.v.x, cd * r.v.y)};
}
__device__ __forceinline__ cx<pk> demod_raw(pk cd, pk sd, const RawEcho<pk> &r) {
    cx<pk> y;
    y.re = mk(fmaf(sd.d.x, r.v.y, cd.d.x * r.v.x), fmaf(sd.d.y, r.v.w, cd.d.y * r.v.z));
    y.im = mk(fmaf(-sd.d.x, r.v.x, cd.d.x * r.v.y), fmaf(-sd.d.y, r.v.z, cd.d.y * r.v.w));
    return y;
}

// Persistent grid: each block owns a contiguous range of (sample, tile) work items, restaging the table only
// when it crosses into the next sample, and takes part in the loss reduction once at the very end.
template <int NE, typename V, bool OUTPUTS, int MINB> __global__ void __launch_bounds__(kThreads, MINB) a2a_loss_kernel(const SolveParams p) {
    __shared__ SampleTab<NE> T;
    const int nv = p.nv, ne = p.ne;
    const int tiles_ps = (nv + kThreads * lanes<V>::n - 1) / (kThreads * lanes<V>::n);
    const long total = static_cast<long>(p.nb) * tiles_ps;
    const int tile_end = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
    int cur_b = -1;
    float loss_part = 0.f;
    for (int tile = static_cast<int>(total * blockIdx.x / gridDim.x); tile < tile_end; ++tile) {
    const int b = tile / tiles_ps;
    if (b != cur_b) {
        if (cur_b >= 0) __syncthreads();          // everyone is done reading the previous sample's table
        stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, ne, p.r2_sc);
        cur_b = b;
    }
    const int v0 = ((tile - b * tiles_ps) * kThreads + threadIdx.x) * lanes<V>::n;
    const bool active = v0 < nv;
    const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
    bool ragged = false;      // a voxel with some, but not all, components exactly zero
    RawEcho<V> raw[NE];
    V phi_t = splat<V>(0.f), r2 = splat<V>(0.f);
    if (active) {
        const float *plane = p.acqs + acq_b;
        const size_t plane_stride = static_cast<size_t>(nv) * 2;
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (e < ne) raw[e] = ld_raw(plane + e * plane_stride, v0, V{});
        ld_pm<V, false>(p.pm + b * p.pm_bstride, v0, phi_t, r2);
        AbsRange ar{3.0e38f, 0.f, 3.0e38f, 0.f};
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (e < ne) abs_range(ar, raw[e]);
        ragged = is_ragged(ar);
    }
    const bool warp_ragged = __any_sync(0xffffffffu, ragged);
    if (active && !warp_ragged) {
        const V zero = splat<V>(0.f);
        V d2[NE];
        cx<V> y[NE];
        cx<V> rw = czero<V>(), rf = czero<V>(), tw = czero<V>(), tf = czero<V>();
        [[maybe_unused]] Mod<V> mods[OUTPUTS ? NE : 1];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const EchoRec R = T.r[e];
                const Mod<V> m = modulator_rec<V, false, kPhaseRad>(R, phi_t, r2, zero);
                if constexpr (OUTPUTS) mods[e] = m;
                d2[e] = vmul(m.d, m.d);
                y[e] = demod_raw(vmul(m.c, m.dinv), vmul(m.s, m.dinv), raw[e]);
                cmac(rw, R.pw_re, R.pw_im, y[e]);
                cmac(rf, R.pf_re, R.pf_im, y[e]);
                cmac(tw, R.tpw_re, R.tpw_im, y[e]);
                cmac(tf, R.tpf_re, R.tpf_im, y[e]);
            }
        }
        V lsum = zero;
        cx<V> K = czero<V>();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const EchoRec R = T.r[e];
                const cx<V> yhat = caffine(rw, R.c_re, R.c_im, rf);
                const cx<V> h = caffine(tw, R.c_re, R.c_im, tf);
                const cx<V> r{vsub(yhat.re, y[e].re), vsub(yhat.im, y[e].im)};
                const cx<V> w{vmul(d2[e], r.re), vmul(d2[e], r.im)};                      // d^2 r
                lsum = vfma(w.re, r.re, lsum);
                lsum = vfma(w.im, r.im, lsum);
                const cx<V> g{vfma(R.te, yhat.re, vneg(h.re)), vfma(R.te, yhat.im, vneg(h.im))};
                K.re = vfma(w.re, g.re, K.re);                                            // K += conj(w) g
                K.re = vfma(w.im, g.im, K.re);
                K.im = vfma(w.re, g.im, K.im);
                K.im = vfma(vneg(w.im), g.re, K.im);
                if constexpr (OUTPUTS) {
                    if (p.shat) st_cx(p.shat + acq_b + static_cast<size_t>(e) * nv * 2, v0, remod(mods[e], yhat));
                }
            }
        }
        loss_part += hsum(lsum);
        const cx<V> g{vmul(-2.0f * kTwoPi * kFmSc * p.inv_n, K.im), vmul(-2.0f * p.r2_sc * p.inv_n, K.re)};
        st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, g);
        if constexpr (OUTPUTS) {
            if (p.rho) {
                const float inv = 1.0f / kRhoSc;
                float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
                st_cx(rho_b, v0, cx<V>{vmul(inv, rw.re), vmul(inv, rw.im)});
                st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, cx<V>{vmul(inv, rf.re), vmul(inv, rf.im)});
            }
        }
    } else if (active) {
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            float ls, gphi, gr2;
            a2a_loss_slow_voxel<NE>(T, p.acqs + acq_b, ne, nv, v0 + l, lane_get(phi_t, l), lane_get(r2, l), p.r2_sc, ls, gphi, gr2);
            loss_part += ls;
            reinterpret_cast<float2 *>(p.g_pm + static_cast<size_t>(b) * nv * 2)[v0 + l] = make_float2(2.0f * p.inv_n * gphi, 2.0f * p.inv_n * gr2);
        }
        if constexpr (OUTPUTS) {
            // materialised outputs do not depend on the mask: recompute them with the plain forward formulas
            const V zero = splat<V>(0.f);
            cx<V> rw = czero<V>(), rf = czero<V>();
            Mod<V> mods[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (e < ne) {
                    mods[e] = modulator(T, e, phi_t, r2, zero);
                    const cx<V> ye = demod(mods[e], raw_cx(raw[e]));
                    cmac(rw, T.r[e].pw_re, T.r[e].pw_im, ye);
                    cmac(rf, T.r[e].pf_re, T.r[e].pf_im, ye);
                }
            }
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (e < ne && p.shat) st_cx(p.shat + acq_b + static_cast<size_t>(e) * nv * 2, v0, remod(mods[e], caffine(rw, T.r[e].c_re, T.r[e].c_im, rf)));
            if (p.rho) {
                const float inv = 1.0f / kRhoSc;
                float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
                st_cx(rho_b, v0, cx<V>{vmul(inv, rw.re), vmul(inv, rw.im)});
                st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, cx<V>{vmul(inv, rf.re), vmul(inv, rf.im)});
            }
        }
    }
    }   // tile loop
    block_loss_reduce(loss_part, p.scratch, p.loss, p.inv_n, p.peer);
}

// =================================================================================================
// TMA-pipelined ring kernel of the fused objectives (packed path): the headline kernel and its two siblings.
//   UQ  = false, OUT = false : config-2 objective, loss + d/dPM only (64 B/voxel)            -- bench.py's kernel
//   UQ  = false, OUT = true  : the same with rho_hat / S_hat materialised (128 B/voxel)
//   UQ  = true               : the uncertainty-aware objective of ig_uq.cu (88 B/voxel)
// Template parameters: NE echo bucket, MINB blocks per SM, STAGES ring depth, EXACT (ne == NE: no per-echo predicates),
// CH chunks (of 64 voxels) per tile, MODE 1 = y in registers / 0 = y parked in the stage, NCW consumer warps,
// TMAP = tensor-map loads / bulk copies.
//
// Warp-specialised persistent blocks: NCW consumer warps + 1 producer warp.  The producer's lane 0 claims tiles
// from a global counter and streams each one (ne echo planes + the PM row + the sample's echo records) into a
// ring of shared-memory stages with TMA -- one 3-D tensor-map box for all echo planes, one for the PM row --
// completing on a "full" mbarrier.  Consumer warps draw 64-voxel chunks of the ring from a shared counter,
// read their 16 bytes per plane with conflict-free LDS.128 exactly once, keep y and d^2 in registers between
// the two passes (MODE 1; MODE 0 parks y in the stage for NE > 8) and release the stage through an "empty"
// mbarrier.  Loads therefore cost the consumers no address arithmetic and no long-scoreboard stalls, and up
// to STAGES tiles per block are in flight to HBM.
// =================================================================================================
constexpr int kChunkVox = 64;                          // voxels per consumer-warp chunk (two per lane)
template <int NE, int STAGES, int CH, bool UQ = false> struct TmaCfg {
    static constexpr int tile_vox = CH * kChunkVox;                               // voxels per tile = per ring stage
    static constexpr int plane_bytes = tile_vox * 8;                              // one complex plane of a tile
    static constexpr int tab_bytes = NE * IG_REC_FLOATS * 4;                      // the sample's echo records ride along with the tile
    static constexpr int uq_off = (NE + 1) * plane_bytes + ((tab_bytes + 127) / 128) * 128;   // UQ: three real planes (phi_var, r2_mean, r2_var)
    static constexpr int real_bytes = tile_vox * 4;
    static constexpr int stage_bytes = uq_off + (UQ ? 3 * real_bytes : 0);
    static constexpr int smem_bytes = STAGES * stage_bytes;
};

// Tiles are handed out dynamically (atomic counter in the scratch header) because background tiles are ~15x
// cheaper than tissue tiles; the producer publishes the tile index of each stage next to its data.
template <int NE, int MINB, int STAGES, bool EXACT, int CH, int MODE, int NCW, bool TMAP, bool UQ, bool OUT>
__global__ void __launch_bounds__(NCW * 32 + 32, MINB)
a2a_loss_tma_kernel(const SolveParams p, const __grid_constant__ CUtensorMap map_acq, const __grid_constant__ CUtensorMap map_pm) {
    extern __shared__ __align__(128) unsigned char stage_mem[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int2 stage_tile[STAGES];             // (sample, first voxel) of the tile in each stage; sample < 0 = end
    __shared__ int chunk_ctr;                       // consumer warps draw 64-voxel chunks of the ring from here
    static_assert(!UQ || MODE == 1, "the uncertainty-aware objective lives in the register-resident path");
    static_assert(!OUT || (MODE == 1 && !UQ), "materialised outputs: register-resident MSE path only (the UQ path writes rho itself)");
    using Cfg = TmaCfg<NE, STAGES, CH, UQ>;
    constexpr int kTileVox = Cfg::tile_vox, kPlaneBytes = Cfg::plane_bytes;
    const int nv = p.nv, ne = EXACT ? NE : p.ne;
    [[maybe_unused]] const bool rem = p.r2_mean == nullptr;
    const int tiles_ps = (nv + kTileVox - 1) / kTileVox;
    const int total = p.nb * tiles_ps;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CH);
        }
        chunk_ctr = 0;
        mbar_fence_init();
    }
    // Programmatic dependent launch: everything above overlaps the tail of the kernel before this one in the stream
    // (normally ig_gen_tables, which releases its dependents at once); nothing below may run before that kernel's
    // writes are visible.  A no-op when the launch carried no such dependency.
    grid_dependency_wait();
    __syncthreads();
    float loss_part = 0.f;
    constexpr int kConsumers = NCW * 32;
    if (threadIdx.x >= kConsumers) {
        // ---------------- producer warp ----------------
        // Lane 0 owns the work counter, the barriers and the copies.  With tensor maps (TMAP) a tile is three TMA
        // instructions: one 3-D box {256 floats, tile rows, ne planes} for all echoes, one for the PM row, one bulk copy for the
        // echo records.  Without (nv not a multiple of 128) it is ne + 2 bulk copies, ~130 cycles of issue each.
        const int lane = threadIdx.x & 31;
        unsigned *next_tile = reinterpret_cast<unsigned *>(p.scratch) + 1;
        const size_t plane_stride = static_cast<size_t>(nv) * 2;
        // the claim for the next tile (an L2 atomic) is in flight while this one is waited for and issued
        // The counter is bumped with a FLOAT atomic add (exact up to 2^24 work items; zero bits = 0.0f, so the scratch header
        // is reused as is): ptxas warp-aggregates every integer atom.add/inc, and the shuffle it puts right behind the
        // atomic would make the producer sit out the full L2 round trip (~1400 cycles) for each tile.
        auto claim_async = [&]() -> float {
            float k;
            asm volatile("atom.global.add.f32 %0, [%1], 0f3F800000;" : "=f"(k) : "l"(next_tile) : "memory");
            return k;
        };
        float k_nxt = 0.f;
        if (lane == 0) k_nxt = claim_async();
        int ends_left = (NCW + CH - 1) / CH;
        bool released = false;
        for (int it = 0;; ++it) {
            const int s = it % STAGES;
            const int k = static_cast<int>(__shfl_sync(0xffffffffu, k_nxt, 0));
            const bool end = k >= total;
            if (lane == 0 && !end) k_nxt = claim_async();
            // Past the middle of the work, whatever is launched behind this kernel with PDL (the next step's table, the next objective) may be
            // scheduled as SM resources allow: a small table kernel runs next to our blocks, the blocks of a following objective take our
            // blocks' place as they exit and run their prologue under our tail.  Every such kernel waits for our completion before it touches
            // global memory (grid_dependency_wait here / in ig_ring.cuh; ig_gen_tables_ahead by contract).
            if (!released && 2 * k >= total) {
                released = true;
                if (lane == 0) grid_launch_dependents();
            }
            // Tiles of a sample are visited in a strided order so that cheap background tiles (load-bound) and tissue
            // tiles (math-bound) are in flight together chip-wide instead of in alternating phases.
            const int b = end ? -1 : k / tiles_ps;
            const int j = k - b * tiles_ps;
            const int vs = end ? 0 : static_cast<int>((static_cast<unsigned>(j) * static_cast<unsigned>(p.tile_stride)) % static_cast<unsigned>(tiles_ps)) * kTileVox;   // < 2^32: see coprime_stride
            const int nvox = (nv - vs < kTileVox) ? nv - vs : kTileVox;
            const uint32_t bytes = TMAP ? static_cast<uint32_t>(kPlaneBytes) : static_cast<uint32_t>(nvox) * 8u;   // a TMA box counts in full (out-of-range rows arrive as zeros)
            unsigned char *stage = stage_mem + s * Cfg::stage_bytes;
            if (lane == 0) {
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                stage_tile[s] = make_int2(b, vs);
                // end marker: completes the phase without data
                if (end) mbar_arrive(&full_bar[s]);
                else mbar_expect_tx(&full_bar[s], bytes * static_cast<uint32_t>(ne + 1) + Cfg::tab_bytes +
                                                      (UQ ? static_cast<uint32_t>(nvox) * 4u * (rem ? 1u : 3u) : 0u));
            }
            __syncwarp();
            if (end) {
                // every consumer warp draws exactly one chunk past the data: publish enough marker stages to cover all of them
                if (--ends_left == 0) break;
                continue;
            }
            if (lane == 0) {
                if constexpr (TMAP) {
                    tma_load_3d(stage, &map_acq, 0, vs >> 7, b * ne, &full_bar[s]);
                    tma_load_3d(stage + ne * kPlaneBytes, &map_pm, 0, vs >> 7, b, &full_bar[s]);
                } else {
                    const float *src = p.acqs + (static_cast<size_t>(b) * ne * nv + vs) * 2;
                    for (int e = 0; e < ne; ++e) bulk_g2s(stage + e * kPlaneBytes, src + e * plane_stride, bytes, &full_bar[s]);
                    bulk_g2s(stage + ne * kPlaneBytes, p.pm + b * p.pm_bstride + static_cast<size_t>(vs) * 2, bytes, &full_bar[s]);
                }
                bulk_g2s(stage + (NE + 1) * kPlaneBytes, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, Cfg::tab_bytes, &full_bar[s]);
                if constexpr (UQ) {
                    const size_t off = static_cast<size_t>(b) * nv + vs;
                    const uint32_t rbytes = static_cast<uint32_t>(nvox) * 4u;
                    bulk_g2s(stage + Cfg::uq_off, p.phi_var + off, rbytes, &full_bar[s]);
                    if (!rem) {
                        bulk_g2s(stage + Cfg::uq_off + Cfg::real_bytes, p.r2_mean + off, rbytes, &full_bar[s]);
                        bulk_g2s(stage + Cfg::uq_off + 2 * Cfg::real_bytes, p.r2_var + off, rbytes, &full_bar[s]);
                    }
                }
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        // Warps are decoupled: each draws the next 64-voxel chunk (1/8 of a tile) from a shared counter, so a warp
        // that lands on background (skipped in ~40 instructions) immediately moves on instead of idling until the
        // tissue warps of its tile finish.  A stage returns to the producer when its 8 chunks have been released.
        constexpr int kChunks = CH;
        const float r2_sc = p.r2_sc;
        const int lane = threadIdx.x & 31;
        [[maybe_unused]] const bool want_shat = OUT && p.shat != nullptr;       // OUT: materialise S_hat / rho_hat (MODE 1 only)
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(&chunk_ctr, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            const int it = g / kChunks, slot = (g % kChunks) * 32 + lane;
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            const int2 where = stage_tile[s];
            if (where.x < 0) break;
            const int b = where.x, vs = where.y;
            const int v0 = vs + slot * 2;
            const bool active = v0 < nv;
            unsigned char *stage = stage_mem + s * Cfg::stage_bytes;
            float4 *sraw = reinterpret_cast<float4 *>(stage) + slot;
            const SampleTab<NE> &T = *reinterpret_cast<const SampleTab<NE> *>(stage + (NE + 1) * kPlaneBytes);   // kdec = -te log2(e), unscaled
            constexpr int kPlaneF4 = kPlaneBytes / 16;
            const pk zero = splat<pk>(0.f);
            if constexpr (MODE != 0) {
                // Register-resident variant: every echo is read from the stage ONCE.  Echo 0 pre-filters background chunks;
                // the ragged test rides along with pass 1 on the ALU pipe; y stays in registers for pass 2.
                float4 raw0 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) raw0 = sraw[0];
                const bool nz0 = raw0.x != 0.f || raw0.y != 0.f || raw0.z != 0.f || raw0.w != 0.f;
                bool slow = false;
                if (!__any_sync(0xffffffffu, nz0)) {
                    // echo 0 is zero across the chunk: background unless a later echo is not (then the chunk is ragged)
                    AbsRange ar{3.0e38f, 0.f, 3.0e38f, 0.f};
                    if (active) {
#pragma unroll
                        for (int e = 1; e < NE; ++e) {
                            if (EXACT || e < ne) {
                                RawEcho<pk> raw;
                                raw.v = sraw[e * kPlaneF4];
                                abs_range(ar, raw);
                            }
                        }
                    }
                    if (!__any_sync(0xffffffffu, active && (ar.hi0 > 0.f || ar.hi1 > 0.f))) {
                        if (active) {
                            st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, czero<pk>());
                            if (want_shat) {
#pragma unroll
                                for (int e = 0; e < NE; ++e)
                                    if (EXACT || e < ne) st_cx(p.shat + (static_cast<size_t>(b) * ne + e) * nv * 2, v0, czero<pk>());
                            }
                            if (OUT && !UQ && p.rho) {
                                float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
                                st_cx(rho_b, v0, czero<pk>());
                                st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, czero<pk>());
                            }
                            if constexpr (UQ) {
                                // rho = 0 -> var = 0 -> the floor: every (echo, component) contributes log sqrt(1e-5), no gradient
                                loss_part += static_cast<float>(ne) * 2.0f * -11.512925464970229f;
                                const size_t o = static_cast<size_t>(b) * nv;
                                st_real(p.g_phi_var + o, v0, zero);
                                if (p.g_r2_mean) st_real(p.g_r2_mean + o, v0, zero);
                                if (p.g_r2_var) st_real(p.g_r2_var + o, v0, zero);
                                if (p.rho) {
                                    float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
                                    st_cx(rho_b, v0, czero<pk>());
                                    st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, czero<pk>());
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[s]);
                        continue;
                    }
                    slow = true;
                }
                pk phi_t = zero, r2s = zero;                       // r2s = R2* in 1/s
                [[maybe_unused]] pk s_phi = zero, mu = zero, s_r = zero;   // UQ: variances / mean in physical units
                if (active) {
                    const float4 m4 = sraw[ne * kPlaneF4];
                    phi_t = mk(m4.x, m4.z);
                    r2s = vmul(r2_sc, mk(m4.y, m4.w));
                    if constexpr (UQ) {
                        const float2 *real = reinterpret_cast<const float2 *>(stage + Cfg::uq_off) + slot;
                        pk t; t.d = real[0];
                        s_phi = vmul(kFmSc * kFmSc, t);
                        if (!rem) {
                            t.d = real[Cfg::real_bytes / 8];
                            mu = vmul(r2_sc, t);
                            t.d = real[2 * Cfg::real_bytes / 8];
                            s_r = vmul(r2_sc * r2_sc, t);
                        }
                    }
                }
                cx<pk> y[NE];
                [[maybe_unused]] pk d2[MODE == 1 ? NE : 1];
                [[maybe_unused]] UqAcc2 acc2{zero, zero, zero, zero};
                cx<pk> rw = czero<pk>(), rf = czero<pk>(), tw = czero<pk>(), tf = czero<pk>();
                if (!slow) {
                    AbsRange ar{3.0e38f, 0.f, 3.0e38f, 0.f};
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (EXACT || e < ne) {
                            const EchoRec R = T.r[e];
                            const Mod<pk> m = modulator_rec<pk, false, kPhaseRad>(R, phi_t, r2s, zero);
                            if constexpr (MODE == 1) d2[e] = vmul(m.d, m.d);
                            RawEcho<pk> raw;
                            raw.v = e == 0 ? raw0 : sraw[e * kPlaneF4];
                            abs_range(ar, raw);
                            y[e] = demod_raw(vmul(m.c, m.dinv), vmul(m.s, m.dinv), raw);
                            if (want_shat) {
                                // Wp_e = d (c + i s) goes into this thread's own 16 bytes of the stage (the raw echo is consumed)
                                const pk cd = vmul(m.c, m.d), sd = vmul(m.s, m.d);
                                sraw[e * kPlaneF4] = make_float4(cd.d.x, cd.d.y, sd.d.x, sd.d.y);
                            }
                            cmac(rw, R.pw_re, R.pw_im, y[e]);
                            cmac(rf, R.pf_re, R.pf_im, y[e]);
                            cmac(tw, R.tpw_re, R.tpw_im, y[e]);
                            cmac(tf, R.tpf_re, R.tpf_im, y[e]);
                        }
                    }
                    slow = __any_sync(0xffffffffu, active && is_ragged(ar));
                }
                if (!slow) {
                    pk lsum = zero;
                    cx<pk> K = czero<pk>();
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (EXACT || e < ne) {
                            const EchoRec R = T.r[e];
                            pk dd;
                            if constexpr (MODE == 1) dd = d2[e]; else dd = fast_ex2(vmul(2.0f * R.kdec, r2s));
                            const cx<pk> yhat = caffine(rw, R.c_re, R.c_im, rf);
                            const cx<pk> h = caffine(tw, R.c_re, R.c_im, tf);
                            const cx<pk> r{vsub(yhat.re, y[e].re), vsub(yhat.im, y[e].im)};
                            if (want_shat && active) {
                                const float4 wp = sraw[e * kPlaneF4];
                                st_cx(p.shat + (static_cast<size_t>(b) * ne + e) * nv * 2, v0, cmulv(cx<pk>{mk(wp.x, wp.y), mk(wp.z, wp.w)}, yhat));
                            }
                            cx<pk> w;
                            if constexpr (UQ) {
                                // residual weighted by 1 / std_e; the variance terms run lane by lane (rsqrt, lg2, ex2 on the SFU)
                                const pk a2 = vfma(yhat.re, yhat.re, vmul(yhat.im, yhat.im));
                                const pk msd = vmul(dd, vfma(r.re, r.re, vmul(r.im, r.im)));
                                const pk inv_std = uq_echo(R.te, a2, msd, s_phi, mu, s_r, rem, acc2);
                                const pk wgt = vmul(inv_std, dd);
                                w = cx<pk>{vmul(wgt, r.re), vmul(wgt, r.im)};
                            } else {
                                w = cx<pk>{vmul(dd, r.re), vmul(dd, r.im)};
                                lsum = vfma(w.re, r.re, lsum);
                                lsum = vfma(w.im, r.im, lsum);
                            }
                            const cx<pk> g{vfma(R.te, yhat.re, vneg(h.re)), vfma(R.te, yhat.im, vneg(h.im))};
                            K.re = vfma(w.re, g.re, K.re);
                            K.re = vfma(w.im, g.im, K.re);
                            K.im = vfma(w.re, g.im, K.im);
                            K.im = vfma(vneg(w.im), g.re, K.im);
                        }
                    }
                    if (active) {
                        if constexpr (!UQ) loss_part += hsum(lsum);
                        st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, cx<pk>{vmul(-2.0f * kTwoPi * kFmSc * p.inv_n, K.im), vmul(-2.0f * r2_sc * p.inv_n, K.re)});
                    }
                } else if (active) {
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        float ls = 0.f, gphi, gr2;
                        if constexpr (UQ) {
                            cx<float> w1, f1;
                            UqAcc a1{0.f, 0.f, 0.f, 0.f};
                            uq_slow_voxel<NE>(T, p.acqs + static_cast<size_t>(b) * ne * nv * 2, ne, nv, v0 + l, lane_get(phi_t, l), lane_get(r2s, l),
                                              lane_get(s_phi, l), lane_get(mu, l), lane_get(s_r, l), rem, r2_sc, a1, gphi, gr2, w1, f1);
                            lane_set(acc2.g_sphi, l, a1.g_sphi); lane_set(acc2.g_mu, l, a1.g_mu); lane_set(acc2.g_sr, l, a1.g_sr);
                            lane_set(acc2.loss, l, a1.loss);
                            lane_set(rw.re, l, w1.re); lane_set(rw.im, l, w1.im);
                            lane_set(rf.re, l, f1.re); lane_set(rf.im, l, f1.im);
                        } else {
                            a2a_loss_slow_voxel<NE>(T, p.acqs + static_cast<size_t>(b) * ne * nv * 2, ne, nv, v0 + l, lane_get(phi_t, l),
                                                    lane_get(r2s, l), r2_sc, ls, gphi, gr2);
                        }
                        loss_part += ls;
                        reinterpret_cast<float2 *>(p.g_pm + static_cast<size_t>(b) * nv * 2)[v0 + l] = make_float2(2.0f * p.inv_n * gphi, 2.0f * p.inv_n * gr2);
                    }
                }
                if constexpr (!UQ && OUT) {
                    if (active && (p.rho || (slow && want_shat))) {
                        if (slow) {
                            // ragged chunk: the materialised outputs do not depend on the mask; recompute them with the plain formulas
                            rw = czero<pk>(); rf = czero<pk>();
#pragma unroll
                            for (int e = 0; e < NE; ++e) {
                                if (EXACT || e < ne) {
                                    const EchoRec R = T.r[e];
                                    const Mod<pk> m = modulator_rec<pk, false, kPhaseRad>(R, phi_t, r2s, zero);
                                    const float4 q = __ldcs(reinterpret_cast<const float4 *>(p.acqs + (static_cast<size_t>(b) * ne + e) * nv * 2) + (v0 >> 1));
                                    const cx<pk> ye = demod(m, cx<pk>{mk(q.x, q.z), mk(q.y, q.w)});
                                    cmac(rw, R.pw_re, R.pw_im, ye);
                                    cmac(rf, R.pf_re, R.pf_im, ye);
                                }
                            }
                            if (want_shat) {
#pragma unroll
                                for (int e = 0; e < NE; ++e) {
                                    if (EXACT || e < ne) {
                                        const EchoRec R = T.r[e];
                                        const Mod<pk> m = modulator_rec<pk, false, kPhaseRad>(R, phi_t, r2s, zero);
                                        st_cx(p.shat + (static_cast<size_t>(b) * ne + e) * nv * 2, v0, remod(m, caffine(rw, R.c_re, R.c_im, rf)));
                                    }
                                }
                            }
                        }
                        if (p.rho) {
                            const float inv = 1.0f / kRhoSc;
                            float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
                            st_cx(rho_b, v0, cx<pk>{vmul(inv, rw.re), vmul(inv, rw.im)});
                            st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, cx<pk>{vmul(inv, rf.re), vmul(inv, rf.im)});
                        }
                    }
                }
                if constexpr (UQ) {
                    if (active) {
                        loss_part += hsum(acc2.loss);
                        const size_t o = static_cast<size_t>(b) * nv;
                        st_real(p.g_phi_var + o, v0, vmul(kFmSc * kFmSc * p.inv_n, acc2.g_sphi));
                        if (p.g_r2_mean) st_real(p.g_r2_mean + o, v0, rem ? zero : vmul(r2_sc * p.inv_n, acc2.g_mu));
                        if (p.g_r2_var) st_real(p.g_r2_var + o, v0, rem ? zero : vmul(r2_sc * r2_sc * p.inv_n, acc2.g_sr));
                        if (p.rho) {
                            const float inv = 1.0f / kRhoSc;
                            float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
                            st_cx(rho_b, v0, cx<pk>{vmul(inv, rw.re), vmul(inv, rw.im)});
                            st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, cx<pk>{vmul(inv, rf.re), vmul(inv, rf.im)});
                        }
                    }
                }
            } else {
                AbsRange ar{3.0e38f, 0.f, 3.0e38f, 0.f};
                if (active) {
    #pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (EXACT || e < ne) {
                            RawEcho<pk> raw;
                            raw.v = sraw[e * kPlaneF4];
                            abs_range(ar, raw);
                        }
                    }
                }
                // background: every component of every voxel of the warp is exactly zero -> loss 0, gradient 0
                if (!__any_sync(0xffffffffu, active && (ar.hi0 > 0.f || ar.hi1 > 0.f))) {
                    if (active) st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, czero<pk>());
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                    continue;
                }
                const bool warp_ragged = __any_sync(0xffffffffu, active && is_ragged(ar));
                pk phi_t = zero, r2s = zero;                       // r2s = R2* in 1/s
                if (active) {
                    const float4 m4 = sraw[ne * kPlaneF4];
                    phi_t = mk(m4.x, m4.z);
                    r2s = vmul(r2_sc, mk(m4.y, m4.w));
                }
                pk d2[NE];
                cx<pk> rw = czero<pk>(), rf = czero<pk>(), tw = czero<pk>(), tf = czero<pk>();
                const bool fast = active && !warp_ragged;
                if (fast) {
    #pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (EXACT || e < ne) {
                            const EchoRec R = T.r[e];
                            const Mod<pk> m = modulator_rec<pk, false, kPhaseRad>(R, phi_t, r2s, zero);
                            d2[e] = vmul(m.d, m.d);
                            RawEcho<pk> raw;
                            raw.v = sraw[e * kPlaneF4];              // second read of the stage: cheaper than 24 live registers
                            const cx<pk> y = demod_raw(vmul(m.c, m.dinv), vmul(m.s, m.dinv), raw);
                            // park y in this thread's own 16 bytes of the stage (the raw echo is no longer needed): 24 fewer
                            // live registers across pass 2 than keeping it, for 6 STS + 6 LDS
                            sraw[e * kPlaneF4] = make_float4(y.re.d.x, y.re.d.y, y.im.d.x, y.im.d.y);
                            cmac(rw, R.pw_re, R.pw_im, y);
                            cmac(rf, R.pf_re, R.pf_im, y);
                            cmac(tw, R.tpw_re, R.tpw_im, y);
                            cmac(tf, R.tpf_re, R.tpf_im, y);
                        }
                    }
                }
                if (fast) {
                    pk lsum = zero;
                    cx<pk> K = czero<pk>();
    #pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        if (EXACT || e < ne) {
                            const EchoRec R = T.r[e];
                            const float4 yv = sraw[e * kPlaneF4];
                            const cx<pk> y{mk(yv.x, yv.y), mk(yv.z, yv.w)};
                            const cx<pk> yhat = caffine(rw, R.c_re, R.c_im, rf);
                            const cx<pk> h = caffine(tw, R.c_re, R.c_im, tf);
                            const cx<pk> r{vsub(yhat.re, y.re), vsub(yhat.im, y.im)};
                            const cx<pk> w{vmul(d2[e], r.re), vmul(d2[e], r.im)};
                            lsum = vfma(w.re, r.re, lsum);
                            lsum = vfma(w.im, r.im, lsum);
                            const cx<pk> g{vfma(R.te, yhat.re, vneg(h.re)), vfma(R.te, yhat.im, vneg(h.im))};
                            K.re = vfma(w.re, g.re, K.re);
                            K.re = vfma(w.im, g.im, K.re);
                            K.im = vfma(w.re, g.im, K.im);
                            K.im = vfma(vneg(w.im), g.re, K.im);
                        }
                    }
                    loss_part += hsum(lsum);
                    st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, cx<pk>{vmul(-2.0f * kTwoPi * kFmSc * p.inv_n, K.im), vmul(-2.0f * r2_sc * p.inv_n, K.re)});
                } else if (active) {
    #pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        float ls, gphi, gr2;
                        // the table in the stage carries the unscaled decay constant: hand the slow path R2* in 1/s
                        a2a_loss_slow_voxel<NE>(T, p.acqs + static_cast<size_t>(b) * ne * nv * 2, ne, nv, v0 + l, lane_get(phi_t, l), lane_get(r2s, l),
                                                r2_sc, ls, gphi, gr2);
                        loss_part += ls;
                        reinterpret_cast<float2 *>(p.g_pm + static_cast<size_t>(b) * nv * 2)[v0 + l] = make_float2(2.0f * p.inv_n * gphi, 2.0f * p.inv_n * gr2);
                    }
                }
            }
            // consumers that parked y / Wp in the stage wrote it through the generic proxy; the producer refills the same bytes
            // through the async proxy (TMA), and the PTX memory model orders the two only across a proxy fence
            if constexpr (MODE == 0 || OUT) fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);                        // this warp no longer touches the stage
        }
    }
    block_loss_reduce(loss_part, p.scratch, p.loss, p.inv_n, p.peer);
}

// -------------------------------------------------------------------------------------------------
// launch helpers
// -------------------------------------------------------------------------------------------------
static int check_common(const char *fn, int nb, int ne, int nv, int min_ne) {
    IG_REQUIRE(nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "%s: nb=%d (1..65535), nv=%d", fn, nb, nv);
    IG_REQUIRE(ne >= min_ne && ne <= IG_MAX_NE, IG_E_NE, "%s: ne=%d outside [%d, %d]", fn, ne, min_ne, IG_MAX_NE);
    return 0;
}

template <typename K1, typename K2> static int launch_pair(bool packed, const SolveParams &p, cudaStream_t st, K1 kp, K2 ks) {
    if (packed) {
        kp<<<grid_for(p.nb, p.nv, 2), kThreads, 0, st>>>(p);
    } else {
        ks<<<grid_for(p.nb, p.nv, 1), kThreads, 0, st>>>(p);
    }
    IG_CUDA(cudaGetLastError());
    return 0;
}

// persistent launch: one resident wave (SMs x occupancy), capped by the number of tiles
template <typename K> static int persistent_grid(K kernel, int nb, int nv, int vpt, int *grid) {
    int dev = 0, sms = 0, occ = 0;
    IG_CUDA(cudaGetDevice(&dev));
    IG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    IG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0));
    const long tiles = static_cast<long>(nb) * ((nv + kThreads * vpt - 1) / (kThreads * vpt));
    long g = static_cast<long>(sms) * (occ > 0 ? occ : 1);
    *grid = static_cast<int>(g < tiles ? g : tiles);
    return 0;
}
template <typename K1, typename K2> static int launch_persistent(bool packed, const SolveParams &p, cudaStream_t st, K1 kp, K2 ks) {
    int grid = 1;
    if (packed) {
        if (int rc = persistent_grid(kp, p.nb, p.nv, 2, &grid)) return rc;
        kp<<<grid, kThreads, 0, st>>>(p);
    } else {
        if (int rc = persistent_grid(ks, p.nb, p.nv, 1, &grid)) return rc;
        ks<<<grid, kThreads, 0, st>>>(p);
    }
    IG_CUDA(cudaGetLastError());
    return 0;
}

// a stride near 0.618 * n that is coprime with n: consecutive work items land on distant tiles of the slice
static int coprime_stride(int n) {
    if (n <= 2 || n > 46340) return 1;      // j * stride must fit 32 bits
    int s = static_cast<int>(n * 0.6180339887) | 1;
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    while (gcd(s, n) != 1) s += 2;
    return s % n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// `planes` complex planes of nv voxels, `plane_stride` floats apart, seen as {256 floats, nv / 128 rows, planes}; box = {256, tile_rows, box_planes}
static bool plane_tensor_map(CUtensorMap *m, const float *base, int nv, long planes, long plane_stride, int tile_rows, int box_planes) {
    const EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {256, static_cast<cuuint64_t>(nv / 128), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {1024, static_cast<cuuint64_t>(plane_stride) * 4};
    const cuuint32_t box[3] = {256, static_cast<cuuint32_t>(tile_rows), static_cast<cuuint32_t>(box_planes)};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename K> static int launch_tma(SolveParams p, cudaStream_t st, K kernel, int smem, int tile_vox, int threads, const CUtensorMap &ma, const CUtensorMap &mp) {
    p.tile_stride = coprime_stride((p.nv + tile_vox - 1) / tile_vox);
    int dev = 0, sms = 0, occ = 0;
    IG_CUDA(cudaGetDevice(&dev));
    IG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    IG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    const long tiles = static_cast<long>(p.nb) * ((p.nv + tile_vox - 1) / tile_vox);
    long g = static_cast<long>(sms) * (occ > 0 ? occ : 1);
    if (g > tiles) g = tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(g));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = static_cast<size_t>(smem);
    cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    IG_CUDA(cudaLaunchKernelEx(&cfg, kernel, p, ma, mp));
    return 0;
}

static bool all_aligned(std::initializer_list<const void *> ps) {
    for (const void *q : ps)
        if (q && !aligned16(q)) return false;
    return true;
}

// Ring configuration of the fused objectives: two blocks per SM, 8 consumer warps + 1 producer warp each, 512-voxel tiles,
// as many ring stages as fit the 227 KB of shared memory.  Measured alternatives (profiles/history_r01.md): smaller tiles, one
// 17-warp block per SM, 10-12 consumer warps per block and batched tile claims are all slower.
static long ring_tiles(int nb, int nv) { return static_cast<long>(nb) * ((nv + 8 * kChunkVox - 1) / (8 * kChunkVox)); }

template <int NE, bool UQ, bool OUT = false> static int launch_a2a_ring(const SolveParams &p, cudaStream_t st) {
    constexpr int kBudget = 216 * 1024;
    const int nb = p.nb, ne = p.ne, nv = p.nv;
    auto go = [&](auto st_c, auto minb_c, auto mode_c, auto ncw_c) {
        constexpr int S = decltype(st_c)::value, MB = decltype(minb_c)::value, MD = decltype(mode_c)::value;
        constexpr int C = 8, W = decltype(ncw_c)::value;
        using Cfg = TmaCfg<NE, S, C, UQ>;
        // one instantiation per echo count up to 12 (the callers dispatch on ne; ring kernels with a run-time count lose the lifting of the later
        // echoes' stage reads over the earlier echoes' math: profiles/history_r02.md section 11); the 16-echo bucket keeps a run-time count, and
        // so does the 12-echo instantiation (serves ne = 12 only; its compile-time-count build spills 76 B and measured 0.2380 against 0.2278 ms)
        constexpr bool kExact = NE <= 11;
        CUtensorMap ma{}, mp{};
        // tensor maps need whole 128-voxel rows; otherwise the tile is fetched plane by plane with bulk copies
        const bool tmap = nv % 128 == 0 && plane_tensor_map(&ma, p.acqs, nv, static_cast<long>(nb) * ne, static_cast<long>(nv) * 2, C / 2, ne) &&
                          plane_tensor_map(&mp, p.pm, nv, nb, p.pm_bstride, C / 2, 1);
        if (kExact && ne != NE) {
            set_error("a2a ring: ne=%d dispatched to the %d-echo instantiation", ne, NE);
            return static_cast<int>(IG_E_NE);
        }
        if (tmap) return launch_tma(p, st, a2a_loss_tma_kernel<NE, MB, S, kExact, C, MD, W, true, UQ, OUT && MD == 1>, Cfg::smem_bytes, Cfg::tile_vox, W * 32 + 32, ma, mp);
        return launch_tma(p, st, a2a_loss_tma_kernel<NE, MB, S, kExact, C, MD, W, false, UQ, OUT && MD == 1>, Cfg::smem_bytes, Cfg::tile_vox, W * 32 + 32, ma, mp);
    };
    using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    // up to 8 echoes y (and d^2) stay in registers between the two passes (95 registers at NE = 6, no spills);
    // beyond that y is parked in the thread's own 16 bytes of the stage
    // consumer warps per block: 8 for the C2 objective (95 registers, no spills); the uncertainty-aware one spills at 96 registers and runs
    // 4 % faster on 7 warps + the producer = 8 warps per block = 128 registers per thread (same-call A/B 0.1889 -> 0.1808 ms; C2: 0.1131 -> 0.1194)
    // (one block of 15 consumer warps on a six-stage ring, which wins for the Rician objective, measured 0.1931 ms masked / 0.2101 unmasked here)
    // the variant that materialises rho_hat / S_hat spills 96-320 B at 96 registers as well: 7 warps measured 0.1951 -> 0.1911 ms at 6 echoes
    // (95 -> 97 % of HBM; unmasked 92 -> 97 %), 0.2504 -> 0.2420 ms at 8
    using W8 = std::integral_constant<int, (UQ || OUT) ? 7 : 8>;
    if constexpr (NE <= 8) {
        if constexpr (2 * TmaCfg<NE, 3, 8, UQ>::smem_bytes <= kBudget) return go(I3{}, I2{}, I1{}, W8{});
        else return go(I2{}, I2{}, I1{}, W8{});
    } else if constexpr (!UQ) {
        if constexpr (2 * TmaCfg<NE, 3, 8>::smem_bytes <= kBudget) return go(I3{}, I2{}, I0{}, W8{});
        else if constexpr (2 * TmaCfg<NE, 2, 8>::smem_bytes <= kBudget) return go(I2{}, I2{}, I0{}, W8{});
        else return go(I3{}, I1{}, I0{}, W8{});
    } else {
        set_error("uncertainty-aware ring kernel: more than 8 echoes");
        return IG_E_UNSUPPORTED;
    }
}

// Entry used by ig_uq.cu: the uncertainty-aware objective on the TMA ring.  IG_E_UNSUPPORTED = shape not covered (the caller
// then runs its plain persistent kernel): needs <= 8 echoes and 128-voxel rows (16-byte bulk copies of the real planes).
int a2a_uq_loss_ring(const float *acqs, const float *pm, long pm_bstride, const float *phi_var, const float *r2_mean, const float *r2_var,
                     const float *tab, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm, float *g_phi_var, float *g_r2_mean,
                     float *g_r2_var, float *rho, float *loss, void *scratch, cudaStream_t st) {
    if (ne > 8 || nv % 128 != 0 || pm_bstride % 4 != 0 || ring_tiles(nb, nv) >= (1L << 24) ||
        !all_aligned({acqs, pm, phi_var, r2_mean, r2_var, g_pm, g_phi_var, g_r2_mean, g_r2_var, rho}))
        return IG_E_UNSUPPORTED;
    SolveParams p{};
    p.acqs = acqs; p.pm = pm; p.pm_bstride = pm_bstride; p.tab = tab; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.inv_n = inv_n;
    p.g_pm = g_pm; p.rho = rho; p.loss = loss; p.scratch = scratch;
    p.phi_var = phi_var; p.r2_mean = r2_mean; p.r2_var = r2_var; p.g_phi_var = g_phi_var; p.g_r2_mean = g_r2_mean; p.g_r2_var = g_r2_var;
    return dispatch_exact_ne<2, 8>(ne, [&](auto ne_c) { return launch_a2a_ring<decltype(ne_c)::value, true>(p, st); });
}

}  // namespace ig

using namespace ig;

extern "C" int ig_get_rho_fwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *bip_d, long bip_bstride,
                              const float *tab_d, int nb, int ne, int nv, float r2_sc, int flags, float *rho_d, float *demod_d, void *stream) {
    IG_REQUIRE(acqs_d && pm_d && tab_d && rho_d, IG_E_ARG, "ig_get_rho_fwd: null pointer");
    if (int rc = check_common("ig_get_rho_fwd", nb, ne, nv, 2)) return rc;
    const bool flat = flags & IG_F_FLAT;
    IG_REQUIRE(!(flat && (bip_d || demod_d)), IG_E_UNSUPPORTED, "ig_get_rho_fwd: flat layout has no bipolar / demod output");
    IG_REQUIRE(!flat || aligned16(rho_d), IG_E_ALIGN, "ig_get_rho_fwd: flat rho output must be 16-byte aligned");
    SolveParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.bip = bip_d; p.bip_bstride = bip_bstride; p.tab = tab_d;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.flags = flags; p.rho = rho_d; p.demod = demod_d;
    const bool packed = nv % 2 == 0 && pm_bstride % 4 == 0 && bip_bstride % 4 == 0 && all_aligned({acqs_d, pm_d, bip_d, rho_d, demod_d});
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        if (flat) return launch_pair(packed, p, st, get_rho_fwd_kernel<NE, pk, true>, get_rho_fwd_kernel<NE, float, true>);
        return launch_pair(packed, p, st, get_rho_fwd_kernel<NE, pk, false>, get_rho_fwd_kernel<NE, float, false>);
    });
}

extern "C" int ig_get_rho_maps(const float *acqs_d, const float *pm_d, long pm_bstride, const float *bip_d, long bip_bstride, const float *tab_d,
                               int nb, int ne, int nv, float r2_sc, int flags, int pdff_mode, float *rho_d, float *pdff_d, float *r2s_d, void *stream) {
    IG_REQUIRE(acqs_d && pm_d && tab_d && rho_d && (pdff_d || r2s_d), IG_E_ARG, "ig_get_rho_maps: null pointer");
    IG_REQUIRE(pdff_mode >= 0 && pdff_mode <= 2, IG_E_ARG, "ig_get_rho_maps: pdff_mode %d", pdff_mode);
    if (int rc = check_common("ig_get_rho_maps", nb, ne, nv, 2)) return rc;
    const bool flat = flags & IG_F_FLAT;
    IG_REQUIRE(!(flat && bip_d), IG_E_UNSUPPORTED, "ig_get_rho_maps: flat layout has no bipolar term");
    IG_REQUIRE(!flat || aligned16(rho_d), IG_E_ALIGN, "ig_get_rho_maps: flat rho output must be 16-byte aligned");
    SolveParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.bip = bip_d; p.bip_bstride = bip_bstride; p.tab = tab_d;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.flags = flags; p.rho = rho_d; p.pdff = pdff_d; p.r2s = r2s_d; p.pdff_mode = pdff_mode;
    bool packed = nv % 2 == 0 && pm_bstride % 4 == 0 && bip_bstride % 4 == 0 && all_aligned({acqs_d, pm_d, bip_d, rho_d});
    packed = packed && (!pdff_d || (reinterpret_cast<uintptr_t>(pdff_d) & 7u) == 0) && (!r2s_d || (reinterpret_cast<uintptr_t>(r2s_d) & 7u) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        if (flat) return launch_pair(packed, p, st, get_rho_fwd_kernel<NE, pk, true>, get_rho_fwd_kernel<NE, float, true>);
        return launch_pair(packed, p, st, get_rho_fwd_kernel<NE, pk, false>, get_rho_fwd_kernel<NE, float, false>);
    });
}

extern "C" int ig_get_rho_bwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *bip_d, long bip_bstride,
                              const float *tab_d, int nb, int ne, int nv, float r2_sc, int flags, const float *g_rho_d,
                              const float *g_demod_d, float *g_acqs_d, float *g_pm_d, float *g_bip_d, void *stream) {
    IG_REQUIRE(acqs_d && pm_d && tab_d && g_pm_d, IG_E_ARG, "ig_get_rho_bwd: null pointer");
    if (int rc = check_common("ig_get_rho_bwd", nb, ne, nv, 2)) return rc;
    const bool flat = flags & IG_F_FLAT;
    IG_REQUIRE(!(flat && (bip_d || g_demod_d || g_bip_d)), IG_E_UNSUPPORTED, "ig_get_rho_bwd: flat layout has no bipolar / demod terms");
    IG_REQUIRE(!flat || !g_rho_d || aligned16(g_rho_d), IG_E_ALIGN, "ig_get_rho_bwd: flat g_rho must be 16-byte aligned");
    SolveParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.bip = bip_d; p.bip_bstride = bip_bstride; p.tab = tab_d;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.flags = flags; p.g_rho = g_rho_d; p.g_demod = g_demod_d;
    p.g_acqs = g_acqs_d; p.g_pm = g_pm_d; p.g_bip = g_bip_d;
    // beyond 8 echoes two voxels per thread need 94-106 registers (16 warps per SM): one voxel per thread (57 registers, 32 warps) measured
    // 0.301 against 0.332 ms at 9 echoes and 0.361 against 0.388 ms at 12 (64 x 384 x 384, dPM + dS; profiles/history_r02.md section 11)
    const bool packed = nv % 2 == 0 && pm_bstride % 4 == 0 && bip_bstride % 4 == 0 && (ne <= 8 || flat) &&
                        all_aligned({acqs_d, pm_d, bip_d, g_rho_d, g_demod_d, g_acqs_d, g_pm_d, g_bip_d});
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        if (flat) return launch_pair(packed, p, st, get_rho_bwd_kernel<NE, pk, true>, get_rho_bwd_kernel<NE, float, true>);
        return launch_pair(packed, p, st, get_rho_bwd_kernel<NE, pk, false>, get_rho_bwd_kernel<NE, float, false>);
    });
}

extern "C" int ig_a2a_fwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
                          float r2_sc, int flags, float *rho_d, float *shat_d, void *stream) {
    IG_REQUIRE(acqs_d && pm_d && tab_d && shat_d, IG_E_ARG, "ig_a2a_fwd: null pointer");
    if (int rc = check_common("ig_a2a_fwd", nb, ne, nv, 2)) return rc;
    SolveParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.tab = tab_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    p.flags = flags; p.rho = rho_d; p.shat = shat_d;
    const bool packed = nv % 2 == 0 && pm_bstride % 4 == 0 && all_aligned({acqs_d, pm_d, rho_d, shat_d});
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        return launch_pair(packed, p, st, a2a_fwd_kernel<NE, pk>, a2a_fwd_kernel<NE, float>);
    });
}

extern "C" int ig_a2a_bwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
                          float r2_sc, int flags, const float *g_rho_d, const float *g_shat_d, float *g_acqs_d, float *g_pm_d, void *stream) {
    IG_REQUIRE(acqs_d && pm_d && tab_d && g_pm_d, IG_E_ARG, "ig_a2a_bwd: null pointer");
    if (int rc = check_common("ig_a2a_bwd", nb, ne, nv, 2)) return rc;
    SolveParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.tab = tab_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    p.flags = flags; p.g_rho = g_rho_d; p.g_shat = g_shat_d; p.g_acqs = g_acqs_d; p.g_pm = g_pm_d;
    const bool packed = nv % 2 == 0 && pm_bstride % 4 == 0 && all_aligned({acqs_d, pm_d, g_rho_d, g_shat_d, g_acqs_d, g_pm_d});
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (packed && !(flags & IG_F_ONLY_MAG)) {        // TMA ring (128-voxel rows, <= 8 echoes): ig_ring_ops.cu
        const int rc = a2a_bwd_ring(acqs_d, pm_d, pm_bstride, tab_d, nb, ne, nv, r2_sc, g_rho_d, g_shat_d, g_acqs_d, g_pm_d, st);
        if (rc != IG_E_UNSUPPORTED) return rc;
    }
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        return launch_pair(packed, p, st, a2a_bwd_kernel<NE, pk>, a2a_bwd_kernel<NE, float>);
    });
}

static int a2a_loss_impl(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv, float r2_sc, float inv_n,
                         float *g_pm_d, float *rho_d, float *shat_d, float *loss_d, void *scratch_d, size_t scratch_bytes, void *stream,
                         const PeerPub &peer) {
    IG_REQUIRE(acqs_d && pm_d && tab_d && g_pm_d && loss_d && scratch_d, IG_E_ARG, "ig_a2a_loss: null pointer");
    if (int rc = check_common("ig_a2a_loss", nb, ne, nv, 2)) return rc;
    IG_REQUIRE(scratch_bytes >= ig_loss_scratch_bytes(nb, nv), IG_E_SCRATCH, "ig_a2a_loss: scratch %zu < %zu bytes", scratch_bytes,
               ig_loss_scratch_bytes(nb, nv));
    SolveParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.tab = tab_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    p.inv_n = inv_n; p.g_pm = g_pm_d; p.rho = rho_d; p.shat = shat_d; p.loss = loss_d; p.scratch = scratch_d; p.peer = peer;
    const bool packed = nv % 2 == 0 && pm_bstride % 4 == 0 && all_aligned({acqs_d, pm_d, g_pm_d, rho_d, shat_d});
    const bool outputs = rho_d || shat_d;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // The ring (its work counter is a float atomic, exact up to 2^24 tiles: larger batches take the plain persistent kernel); materialised
    // rho_hat / S_hat ride on its register-resident path, which holds up to 8 echoes.  One instantiation per echo count up to 12.
    if (packed && ring_tiles(nb, nv) < (1L << 24) && (!outputs || ne <= 8)) {
        if (ne > 12) return launch_a2a_ring<16, false>(p, st);
        return dispatch_exact_ne<2, 12>(ne, [&](auto ne_c) {
            constexpr int NE = decltype(ne_c)::value;
            if constexpr (NE <= 8) {
                if (outputs) return launch_a2a_ring<NE, false, true>(p, st);
            }
            return launch_a2a_ring<NE, false>(p, st);
        });
    }
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        if (outputs) return launch_persistent(packed, p, st, a2a_loss_kernel<NE, pk, true, 2>, a2a_loss_kernel<NE, float, true, 2>);
        return launch_persistent(packed, p, st, a2a_loss_kernel<NE, pk, false, 2>, a2a_loss_kernel<NE, float, false, 2>);
    });
}

extern "C" int ig_a2a_loss(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
                           float r2_sc, float inv_n, float *g_pm_d, float *rho_d, float *shat_d, float *loss_d, void *scratch_d,
                           size_t scratch_bytes, void *stream) {
    return a2a_loss_impl(acqs_d, pm_d, pm_bstride, tab_d, nb, ne, nv, r2_sc, inv_n, g_pm_d, rho_d, shat_d, loss_d, scratch_d, scratch_bytes, stream,
                         PeerPub{});
}

extern "C" int ig_a2a_loss_peer(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
                                float r2_sc, float inv_n, float *g_pm_d, float *rho_d, float *shat_d, float *loss_d, void *scratch_d,
                                size_t scratch_bytes, ig_peer *peer, unsigned step, int lag, float *loss_prev_d, void *stream) {
    IG_REQUIRE(peer, IG_E_ARG, "ig_a2a_loss_peer: null peer context");
    PeerPub pub{};
    if (int rc = peer_pub(peer, step, lag, loss_prev_d, &pub)) return rc;
    return a2a_loss_impl(acqs_d, pm_d, pm_bstride, tab_d, nb, ne, nv, r2_sc, inv_n, g_pm_d, rho_d, shat_d, loss_d, scratch_d, scratch_bytes, stream, pub);
}
