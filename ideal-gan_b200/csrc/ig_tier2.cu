// Second-tier operators of the reference's physics library: the magnitude-only fit with its
// closed-form 2x2 eigen-decomposition, the signal-domain and parameter-domain uncertainty propagation, and PDFF
// extraction.  All are one thread per voxel, every echo in registers, per-sample tables in shared memory; none of
// them materialises the (nb, ne, nv) intermediates -- or, for PDFF_uncertainty, the (nv, nb, ne, ne) diagonal
// matrices -- that the TensorFlow op chains of /root/reference/wflib/IDEAL_model.py:100-138, 314-401, 628-767 do.
#include "ig_common.cuh"

namespace ig {

constexpr float kEigEps = 1e-12f;        // IDEAL_model.py:107

struct Eig {
    float x, y, ratio;
};

// SFU square root / division (1-2 ulp): the IEEE forms are 6-10 instructions plus a slow path each, and the eigen-decomposition
// takes 3 roots and up to 6 quotients per voxel
__device__ __forceinline__ float fsqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }

// principal eigenpair of [[a, b/2], [b/2, c]]  (IDEAL_model.py:100-138)
__device__ __forceinline__ Eig eig_fwd(float a, float b, float c) {
    const float hd = (a - c) * 0.5f, hb = b * 0.5f;
    const float delta = fsqrt(hd * hd + hb * hb + kEigEps);
    const float mean = (a + c) * 0.5f;
    const float lmax = mean + delta, lmin = mean - delta;
    const float lmax_p = fmaxf(lmax, 0.f), lmin_p = fmaxf(lmin, 0.f);
    const float vx = hb, vy = lmax - a;
    const float norm = fsqrt(vx * vx + vy * vy + kEigEps);
    const float scale = fsqrt(lmax_p);
    Eig r;
    r.x = scale * fdiv(vx, norm);
    r.y = scale * fdiv(vy, norm);
    r.ratio = lmax_p > 0.f ? fdiv(lmin_p, lmax_p) : 0.f;
    return r;
}

// adjoint of eig_fwd: upstream (gx, gy, gr) -> (ga, gb, gc)
__device__ __forceinline__ void eig_bwd(float a, float b, float c, float gx, float gy, float gr, float &ga, float &gb, float &gc) {
    const float hd = (a - c) * 0.5f, hb = b * 0.5f;
    const float delta = fsqrt(hd * hd + hb * hb + kEigEps);
    const float mean = (a + c) * 0.5f;
    const float lmax = mean + delta, lmin = mean - delta;
    const float lmax_p = fmaxf(lmax, 0.f), lmin_p = fmaxf(lmin, 0.f);
    const float vx = hb, vy = lmax - a;
    const float norm = fsqrt(vx * vx + vy * vy + kEigEps);
    const float inv_norm = fdiv(1.0f, norm);
    const float scale = fsqrt(lmax_p);
    const float vxn = vx * inv_norm, vyn = vy * inv_norm;
    const float g_scale = gx * vxn + gy * vyn;
    const float g_vxn = gx * scale, g_vyn = gy * scale;
    const float dot = (g_vxn * vx + g_vyn * vy) * inv_norm * inv_norm * inv_norm;
    float g_vx = g_vxn * inv_norm - dot * vx;
    const float g_vy = g_vyn * inv_norm - dot * vy;
    float g_lmax = g_vy, g_lmin = 0.f;
    if (lmax > 0.f) {
        g_lmax += fdiv(g_scale * 0.5f, scale);
        g_lmax += fdiv(-gr * lmin_p, lmax_p * lmax_p);
        if (lmin > 0.f) g_lmin = fdiv(gr, lmax_p);
    }
    const float g_mean = g_lmax + g_lmin, g_delta = g_lmax - g_lmin;
    const float inv_delta = fdiv(1.0f, delta);
    const float g_hd = g_delta * hd * inv_delta;
    g_vx += g_delta * hb * inv_delta;
    ga = 0.5f * g_mean + 0.5f * g_hd - g_vy;
    gc = 0.5f * g_mean - 0.5f * g_hd;
    gb = 0.5f * g_vx;
}

__global__ void eigenvals_kernel(const float *__restrict__ X, long n, float *__restrict__ xy, float *__restrict__ ratio) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Eig e = eig_fwd(X[3 * i], X[3 * i + 1], X[3 * i + 2]);
    xy[2 * i] = e.x;
    xy[2 * i + 1] = e.y;
    ratio[i] = e.ratio;
}

__global__ void eigenvals_bwd_kernel(const float *__restrict__ X, long n, const float *__restrict__ g_xy, const float *__restrict__ g_ratio,
                                     float *__restrict__ gX) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ga, gb, gc;
    eig_bwd(X[3 * i], X[3 * i + 1], X[3 * i + 2], g_xy ? g_xy[2 * i] : 0.f, g_xy ? g_xy[2 * i + 1] : 0.f, g_ratio ? g_ratio[i] : 0.f, ga, gb, gc);
    gX[3 * i] = ga;
    gX[3 * i + 1] = gb;
    gX[3 * i + 2] = gc;
}

// -------------------------------------------------------------------------------------------------------------
// CSE_mag: y_e = (e^{te R} |S_e|)^2 ; abc = A^+ y ; fit = A abc ; S_hat_e = e^{-te R} sqrt(fit_e) [fit > 1e-6]
// -------------------------------------------------------------------------------------------------------------
struct CseParams {
    const float *mag, *r2, *r2nu, *tab;                 // (nb, ne, nv), (nb, nv), optional (nb, nv)
    float *rho, *fit, *demod, *ls, *unc;                // forward outputs (all optional)
    const float *g_rho, *g_fit, *g_demod, *g_ls, *g_unc;   // backward upstreams (all optional)
    float *g_mag, *g_r2, *g_r2nu;
    int nb, ne, nv;
    float r2_sc;
};

template <int NE> struct MagTab {
    float te[NE], a1[NE], a2[NE], p0[NE], p1[NE], p2[NE];      // A_e = (1, a1, a2) ; A^+[:, e] = (p0, p1, p2)
};
template <int NE> __device__ __forceinline__ void stage_mag_table(MagTab<NE> &t, const float *tab_b, int ne) {
    for (int e = threadIdx.x; e < NE; e += blockDim.x) {
        const bool live = e < ne;
        const float cr = live ? tab_b[e * IG_REC_FLOATS + IG_REC_C_RE] : 0.f, ci = live ? tab_b[e * IG_REC_FLOATS + IG_REC_C_IM] : 0.f;
        t.te[e] = live ? tab_b[e * IG_REC_FLOATS + IG_REC_TE] : 0.f;
        t.a1[e] = cr;
        t.a2[e] = cr * cr + ci * ci;
        t.p0[e] = live ? tab_b[IG_TAB_AP_OFF + 0 * IG_MAX_NE + e] : 0.f;
        t.p1[e] = live ? tab_b[IG_TAB_AP_OFF + 1 * IG_MAX_NE + e] : 0.f;
        t.p2[e] = live ? tab_b[IG_TAB_AP_OFF + 2 * IG_MAX_NE + e] : 0.f;
    }
    __syncthreads();
}

// Forward pass with two neighbouring voxels per thread (8-byte streaming loads / stores, f32x2 arithmetic for the fit): the
// one-voxel kernel below moved 100 bytes per voxel in 4-byte transactions and was bound by instruction issue (78 % of the slots,
// a third of them load / store instructions) at 67 % of the HBM rate.  Same arithmetic; the eigen-decomposition runs per lane.
__device__ __forceinline__ pk sqrt_approx(pk x) {
    pk r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r.d.x) : "f"(x.d.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r.d.y) : "f"(x.d.y));
    return r;
}
template <int NE> __global__ void __launch_bounds__(kThreads) cse_mag_fwd2_kernel(const CseParams p) {
    __shared__ MagTab<NE> T;
    const int b = blockIdx.y;
    stage_mag_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne);
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (v >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t eb = static_cast<size_t>(b) * ne * nv, vb = static_cast<size_t>(b) * nv;
    const pk R = vmul(p.r2_sc * kLog2e, ld_real(p.r2 + vb, v, pk{}));          // R2* in units of log2(e): the growth factor is one EX2
    pk S[NE], y[NE], wm[NE];
    pk a = splat<pk>(0.f), bb = a, c = a;
#pragma unroll
    for (int e = 0; e < NE; ++e)
        if (e < ne) S[e] = ld_real(p.mag + eb + static_cast<size_t>(e) * nv, v, pk{});
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            wm[e] = fast_ex2(vmul(T.te[e], R));
            const pk t = vmul(wm[e], S[e]);
            y[e] = vmul(t, t);
            a = vfma(T.p0[e], y[e], a);
            bb = vfma(T.p1[e], y[e], bb);
            c = vfma(T.p2[e], y[e], c);
        }
    }
    const float inv_rho = 1.0f / kRhoSc, inv_rho2 = 1.0f / (kRhoSc * kRhoSc);
    const Eig e0 = eig_fwd(a.d.x, bb.d.x, c.d.x), e1 = eig_fwd(a.d.y, bb.d.y, c.d.y);
    if (p.rho) {
        st_real(p.rho + (static_cast<size_t>(b) * 2 + 0) * nv, v, mk(e0.x * inv_rho, e1.x * inv_rho));
        st_real(p.rho + (static_cast<size_t>(b) * 2 + 1) * nv, v, mk(e0.y * inv_rho, e1.y * inv_rho));
    }
    if (p.ls) {
        st_real(p.ls + (static_cast<size_t>(b) * 3 + 0) * nv, v, vmul(inv_rho2, a));
        st_real(p.ls + (static_cast<size_t>(b) * 3 + 1) * nv, v, vmul(inv_rho2, bb));
        st_real(p.ls + (static_cast<size_t>(b) * 3 + 2) * nv, v, vmul(inv_rho2, c));
    }
    if (p.unc) st_real(p.unc + vb, v, mk(e0.ratio, e1.ratio));
    pk Rnu = splat<pk>(0.f);
    if (p.r2nu) Rnu = vmul(p.r2_sc * kLog2e, ld_real(p.r2nu + vb, v, pk{}));
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const size_t o = eb + static_cast<size_t>(e) * nv;
            if (p.fit) {
                const pk f = vfma(T.a2[e], c, vfma(T.a1[e], bb, a));
                const pk sh = vmul(sqrt_approx(f), fast_ex2(vneg(vmul(T.te[e], R))));                 // sqrt(fit) e^{-te R}
                st_real(p.fit + o, v, mk(f.d.x > 1e-6f ? sh.d.x : 0.f, f.d.y > 1e-6f ? sh.d.y : 0.f));
            }
            if (p.demod) {
                if (p.r2nu) {
                    const pk t = vmul(fast_ex2(vmul(T.te[e], Rnu)), S[e]);
                    st_real(p.demod + o, v, vmul(t, t));
                } else {
                    st_real(p.demod + o, v, y[e]);
                }
            }
        }
    }
}

// Adjoint with two voxels per thread and every load issued before the math: the one-voxel kernel below fetched each upstream
// echo inside the `fit > 1e-6` branch, one dependent 4-byte load after the other (0.535 ms = 35 % of the HBM rate for 128 B/voxel).
template <int NE> __global__ void __launch_bounds__(kThreads, NE <= 8 ? 3 : 2) cse_mag_bwd2_kernel(const CseParams p) {
    __shared__ MagTab<NE> T;
    const int b = blockIdx.y;
    stage_mag_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne);
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (v >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t eb = static_cast<size_t>(b) * ne * nv, vb = static_cast<size_t>(b) * nv;
    const pk zero = splat<pk>(0.f);
    const pk R = vmul(p.r2_sc * kLog2e, ld_real(p.r2 + vb, v, pk{}));
    pk S[NE], Gf[NE], Gd[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const size_t o = eb + static_cast<size_t>(e) * nv;
            S[e] = ld_real(p.mag + o, v, pk{});
            Gf[e] = p.g_fit ? ld_real(p.g_fit + o, v, pk{}) : zero;
            Gd[e] = p.g_demod ? ld_real(p.g_demod + o, v, pk{}) : zero;
        }
    }
    const float inv_rho = 1.0f / kRhoSc, inv_rho2 = 1.0f / (kRhoSc * kRhoSc);
    pk gx = zero, gy0 = zero, gr = zero, gl0 = zero, gl1 = zero, gl2 = zero;
    if (p.g_rho) {
        gx = vmul(inv_rho, ld_real(p.g_rho + (static_cast<size_t>(b) * 2 + 0) * nv, v, pk{}));
        gy0 = vmul(inv_rho, ld_real(p.g_rho + (static_cast<size_t>(b) * 2 + 1) * nv, v, pk{}));
    }
    if (p.g_unc) gr = ld_real(p.g_unc + vb, v, pk{});
    if (p.g_ls) {
        gl0 = ld_real(p.g_ls + (static_cast<size_t>(b) * 3 + 0) * nv, v, pk{});
        gl1 = ld_real(p.g_ls + (static_cast<size_t>(b) * 3 + 1) * nv, v, pk{});
        gl2 = ld_real(p.g_ls + (static_cast<size_t>(b) * 3 + 2) * nv, v, pk{});
    }
    pk Rnu = zero;
    if (p.r2nu) Rnu = vmul(p.r2_sc * kLog2e, ld_real(p.r2nu + vb, v, pk{}));
    pk wm[NE];      // y_e = (wm_e S_e)^2 is re-formed where needed: holding it costs 2 NE registers and a block per SM
    pk a = zero, bb = zero, c = zero;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            wm[e] = fast_ex2(vmul(T.te[e], R));
            const pk t = vmul(wm[e], S[e]);
            const pk y = vmul(t, t);
            a = vfma(T.p0[e], y, a);
            bb = vfma(T.p1[e], y, bb);
            c = vfma(T.p2[e], y, c);
        }
    }
    pk ga = zero, gb = zero, gc = zero;
    if (p.g_rho || p.g_unc) {
        float a0, b0, c0, a1, b1, c1;
        eig_bwd(a.d.x, bb.d.x, c.d.x, gx.d.x, gy0.d.x, gr.d.x, a0, b0, c0);
        eig_bwd(a.d.y, bb.d.y, c.d.y, gx.d.y, gy0.d.y, gr.d.y, a1, b1, c1);
        ga = mk(a0, a1); gb = mk(b0, b1); gc = mk(c0, c1);
    }
    ga = vfma(inv_rho2, gl0, ga);
    gb = vfma(inv_rho2, gl1, gb);
    gc = vfma(inv_rho2, gl2, gc);
    pk gR = zero, gRnu = zero;
    if (p.g_fit) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                // S_hat = sqrt(f) / wm where f > 1e-6 (the reference's where(fit > 1e-6, sqrt(fit), 0): zero gradient elsewhere)
                const pk f = vfma(T.a2[e], c, vfma(T.a1[e], bb, a));
                const pk fs = mk(f.d.x > 1e-6f ? f.d.x : 1.0f, f.d.y > 1e-6f ? f.d.y : 1.0f);
                const pk sq = sqrt_approx(fs);
                const pk den = vmul(sq, wm[e]);
                const pk inv = mk(__fdividef(1.0f, den.d.x), __fdividef(1.0f, den.d.y));
                pk gfi = vmul(vmul(0.5f, Gf[e]), inv);                              // d / d fit
                gfi = mk(f.d.x > 1e-6f ? gfi.d.x : 0.f, f.d.y > 1e-6f ? gfi.d.y : 0.f);
                ga = vadd(ga, gfi);
                gb = vfma(T.a1[e], gfi, gb);
                gc = vfma(T.a2[e], gfi, gc);
                gR = vfma(vmul(-2.0f * T.te[e], gfi), fs, gR);                       // -te gf S_hat = -te gf f / (sq wm) = -2 te gfi f
            }
        }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            pk gy = vfma(T.p2[e], gc, vfma(T.p1[e], gb, vmul(T.p0[e], ga)));
            pk gS = zero;
            if (p.g_demod) {
                if (p.r2nu) {
                    const pk wn = fast_ex2(vmul(T.te[e], Rnu));
                    const pk w2s = vmul(vmul(wn, wn), S[e]);
                    gS = vmul(vmul(2.0f, Gd[e]), w2s);
                    gRnu = vfma(vmul(2.0f * T.te[e], Gd[e]), vmul(w2s, S[e]), gRnu);
                } else {
                    gy = vadd(gy, Gd[e]);
                }
            }
            const pk t = vmul(wm[e], S[e]);
            gS = vfma(vmul(vmul(2.0f, gy), wm[e]), t, gS);
            gR = vfma(vmul(2.0f * T.te[e], gy), vmul(t, t), gR);
            if (p.g_mag) st_real(p.g_mag + eb + static_cast<size_t>(e) * nv, v, gS);
        }
    }
    if (p.g_r2) st_real(p.g_r2 + vb, v, vmul(p.r2_sc, gR));
    if (p.g_r2nu) st_real(p.g_r2nu + vb, v, vmul(p.r2_sc, gRnu));
}

template <int NE, bool BWD> __global__ void __launch_bounds__(kThreads) cse_mag_kernel(const CseParams p) {
    __shared__ MagTab<NE> T;
    const int b = blockIdx.y;
    stage_mag_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne);
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t eb = static_cast<size_t>(b) * ne * nv, vb = static_cast<size_t>(b) * nv;
    const float R = p.r2[vb + v] * p.r2_sc;
    float S[NE], y[NE], wm[NE];
    float a = 0.f, bb = 0.f, c = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e)
        if (e < ne) S[e] = __ldcs(p.mag + eb + static_cast<size_t>(e) * nv + v);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            wm[e] = __expf(T.te[e] * R);
            const float t = wm[e] * S[e];
            y[e] = t * t;
            a = fmaf(T.p0[e], y[e], a);
            bb = fmaf(T.p1[e], y[e], bb);
            c = fmaf(T.p2[e], y[e], c);
        }
    }
    const float inv_rho = 1.0f / kRhoSc, inv_rho2 = 1.0f / (kRhoSc * kRhoSc);
    if constexpr (!BWD) {
        const Eig eg = eig_fwd(a, bb, c);
        if (p.rho) {
            p.rho[(static_cast<size_t>(b) * 2 + 0) * nv + v] = eg.x * inv_rho;
            p.rho[(static_cast<size_t>(b) * 2 + 1) * nv + v] = eg.y * inv_rho;
        }
        if (p.ls) {
            p.ls[(static_cast<size_t>(b) * 3 + 0) * nv + v] = a * inv_rho2;
            p.ls[(static_cast<size_t>(b) * 3 + 1) * nv + v] = bb * inv_rho2;
            p.ls[(static_cast<size_t>(b) * 3 + 2) * nv + v] = c * inv_rho2;
        }
        if (p.unc) p.unc[vb + v] = eg.ratio;
        const float Rnu = p.r2nu ? p.r2nu[vb + v] * p.r2_sc : 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const size_t o = eb + static_cast<size_t>(e) * nv + v;
                if (p.fit) {
                    const float f = fmaf(T.a2[e], c, fmaf(T.a1[e], bb, a));
                    p.fit[o] = f > 1e-6f ? __fdividef(sqrtf(f), wm[e]) : 0.f;
                }
                if (p.demod) {
                    if (p.r2nu) {
                        const float t = __expf(T.te[e] * Rnu) * S[e];
                        p.demod[o] = t * t;
                    } else {
                        p.demod[o] = y[e];
                    }
                }
            }
        }
    } else {
        float ga = 0.f, gb = 0.f, gc = 0.f;
        {
            const float gx = p.g_rho ? p.g_rho[(static_cast<size_t>(b) * 2 + 0) * nv + v] * inv_rho : 0.f;
            const float gy = p.g_rho ? p.g_rho[(static_cast<size_t>(b) * 2 + 1) * nv + v] * inv_rho : 0.f;
            const float gr = p.g_unc ? p.g_unc[vb + v] : 0.f;
            if (p.g_rho || p.g_unc) eig_bwd(a, bb, c, gx, gy, gr, ga, gb, gc);
        }
        if (p.g_ls) {
            ga += p.g_ls[(static_cast<size_t>(b) * 3 + 0) * nv + v] * inv_rho2;
            gb += p.g_ls[(static_cast<size_t>(b) * 3 + 1) * nv + v] * inv_rho2;
            gc += p.g_ls[(static_cast<size_t>(b) * 3 + 2) * nv + v] * inv_rho2;
        }
        float gR = 0.f, gRnu = 0.f;
        if (p.g_fit) {
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (e < ne) {
                    const float f = fmaf(T.a2[e], c, fmaf(T.a1[e], bb, a));
                    if (f > 1e-6f) {                   // the reference's where(fit > 1e-6, sqrt(fit), 0): zero gradient elsewhere
                        const float gf = p.g_fit[eb + static_cast<size_t>(e) * nv + v];
                        const float sq = sqrtf(f), shat = sq / wm[e];
                        const float gfit = gf * 0.5f / (sq * wm[e]);
                        ga += gfit;
                        gb = fmaf(gfit, T.a1[e], gb);
                        gc = fmaf(gfit, T.a2[e], gc);
                        gR = fmaf(-T.te[e] * gf, shat, gR);
                    }
                }
            }
        }
        const float Rnu = p.r2nu ? p.r2nu[vb + v] * p.r2_sc : 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const size_t o = eb + static_cast<size_t>(e) * nv + v;
                float gy = fmaf(T.p2[e], gc, fmaf(T.p1[e], gb, T.p0[e] * ga));
                float gS = 0.f;
                if (p.g_demod) {
                    const float gd = p.g_demod[o];
                    if (p.r2nu) {
                        const float wn = __expf(T.te[e] * Rnu);
                        gS = gd * 2.f * wn * wn * S[e];
                        gRnu = fmaf(gd * 2.f * T.te[e], wn * wn * S[e] * S[e], gRnu);
                    } else {
                        gy += gd;
                    }
                }
                gS = fmaf(gy * 2.f * wm[e], wm[e] * S[e], gS);
                gR = fmaf(gy * 2.f * T.te[e], y[e], gR);
                if (p.g_mag) p.g_mag[o] = gS;
            }
        }
        if (p.g_r2) p.g_r2[vb + v] = gR * p.r2_sc;
        if (p.g_r2nu) p.g_r2nu[vb + v] = gRnu * p.r2_sc;
    }
}

// -------------------------------------------------------------------------------------------------------------
// acq_uncertainty: Var_e = V_e |rho_W + c_e rho_F|^2,  V_e = 1 - e^{-(2 pi te)^2 s_phi} + e^{-te mu_R} te^2 s_R
// -------------------------------------------------------------------------------------------------------------
struct UncParams {
    const float *rho, *phi_var, *r2_mean, *r2_var, *tab;   // (nb,2,nv,2), (nb,nv) x3 (r2_* NULL = rem_R2)
    float *out;                                            // (nb, ne, nv, ch)
    const float *g_out;
    float *g_phi_var, *g_r2_mean, *g_r2_var;
    int nb, ne, nv, ch;
    float r2_sc;
};

// V = pk: two neighbouring voxels per thread (16-byte loads of rho, 8-byte loads of the moment maps, 16- or 8-byte stores per echo);
// V = float: odd voxel counts / unaligned planes.  (One voxel per thread with 4- and 8-byte accesses measured 88 % / 74 % of the HBM rate.)
template <int NE, bool BWD, typename V, int CH> __global__ void __launch_bounds__(kThreads) acq_unc_kernel(const UncParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, 1.0f);
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    constexpr int ch = CH;                      // compile-time: a run-time channel count predicated every upstream load and serialised them
    const size_t vb = static_cast<size_t>(b) * nv;
    const V zero = splat<V>(0.f);
    const float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
    const cx<V> w2 = ld_cx(rho_b, v, V{}), f2 = ld_cx(rho_b + static_cast<size_t>(nv) * 2, v, V{});
    const cx<V> rw{vmul(kRhoSc, w2.re), vmul(kRhoSc, w2.im)}, rf{vmul(kRhoSc, f2.re), vmul(kRhoSc, f2.im)};
    const V s_phi = vmul(kFmSc * kFmSc, ld_real(p.phi_var + vb, v, V{}));
    const bool r2 = p.r2_mean != nullptr;
    const V mu = r2 ? vmul(p.r2_sc, ld_real(p.r2_mean + vb, v, V{})) : zero;
    const V s_r = r2 ? vmul(p.r2_sc * p.r2_sc, ld_real(p.r2_var + vb, v, V{})) : zero;
    V g_sphi = zero, g_mu = zero, g_sr = zero;
    [[maybe_unused]] V G[NE];
    if constexpr (BWD) {
        // every upstream load in flight before the first is consumed
        [[maybe_unused]] cx<V> G2[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const float *g_e = p.g_out + (static_cast<size_t>(b) * ne + e) * nv * ch;
                if constexpr (ch == 2) G2[e] = ld_cx(g_e, v, V{});
                else G[e] = ld_real(g_e, v, V{});
            }
        }
        if constexpr (ch == 2) {
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (e < ne) G[e] = vadd(G2[e].re, G2[e].im);
        }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const float te = T.r[e].te, k = kTwoPi * te, k2 = k * k;
            const V ephi = fast_ex2(vmul(-k2 * kLog2e, s_phi));
            const V er = r2 ? vmul(te * te, fast_ex2(vmul(-te * kLog2e, mu))) : zero;
            const cx<V> m = caffine(rw, T.r[e].c_re, T.r[e].c_im, rf);
            const V a2 = vfma(m.re, m.re, vmul(m.im, m.im));
            if constexpr (!BWD) {
                const V var = vmul(vfma(er, s_r, vsub(splat<V>(1.0f), ephi)), a2);        // 1 - e^{-x} literally, as the reference forms it
                float *o_e = p.out + (static_cast<size_t>(b) * ne + e) * nv * ch;
                if constexpr (ch == 2) st_cx(o_e, v, cx<V>{var, var});
                else st_real(o_e, v, var);
            } else {
                const V g = vmul(G[e], a2);
                g_sphi = vfma(vmul(k2, g), ephi, g_sphi);
                g_mu = vfma(vmul(-te, g), vmul(er, s_r), g_mu);
                g_sr = vfma(g, er, g_sr);
            }
        }
    }
    if constexpr (BWD) {
        st_real(p.g_phi_var + vb, v, vmul(kFmSc * kFmSc, g_sphi));
        if (p.g_r2_mean) st_real(p.g_r2_mean + vb, v, vmul(p.r2_sc, g_mu));
        if (p.g_r2_var) st_real(p.g_r2_var + vb, v, vmul(p.r2_sc * p.r2_sc, g_sr));
    }
}

// -------------------------------------------------------------------------------------------------------------
// PDFF_uncertainty (IDEAL_model.py:628-706): weighted LS with Sigma_e = V_e (|Wp (P0 Wm)|_e^2 + |S_e|^2)
// -------------------------------------------------------------------------------------------------------------
struct PdffUncParams {
    const float *acqs, *phi_mean, *phi_var, *r2_mean, *r2_var, *tab;   // r2_* NULL = rem_R2
    float *rho, *cov;                                                  // (nb,2,nv,2), (nb,4,nv)
    int nb, ne, nv;
    float r2_sc;
};

__device__ __forceinline__ float vrcp_or_zero(float x) { return x != 0.f ? __fdividef(1.0f, x) : 0.f; }
__device__ __forceinline__ pk vrcp_or_zero(pk x) { return mk(vrcp_or_zero(x.d.x), vrcp_or_zero(x.d.y)); }
__device__ __forceinline__ float vdiv1(float x) { return 1.0f / x; }
__device__ __forceinline__ pk vdiv1(pk x) { return mk(1.0f / x.d.x, 1.0f / x.d.y); }
__device__ __forceinline__ float vabs(float x) { return fabsf(x); }
__device__ __forceinline__ pk vabs(pk x) { return mk(fabsf(x.d.x), fabsf(x.d.y)); }

// V = float: one voxel per thread (the instantiation that is launched); V = pk: two neighbouring voxels on packed f32x2 lanes
template <int NE, typename V> __global__ void __launch_bounds__(kThreads) pdff_unc_kernel(const PdffUncParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v >= p.nv) return;
    const int nv = p.nv, ne = p.ne;
    const size_t pb = static_cast<size_t>(b) * nv;
    const V zero = splat<V>(0.f);
    const bool r2 = p.r2_mean != nullptr;
    const V phi_t = ld_real(p.phi_mean + pb, v, V{});
    const V r2map = r2 ? ld_real(p.r2_mean + pb, v, V{}) : zero;
    const V s_phi = vmul(kFmSc * kFmSc, ld_real(p.phi_var + pb, v, V{}));
    const V s_r = r2 ? vmul(p.r2_sc * p.r2_sc, ld_real(p.r2_var + pb, v, V{})) : zero;
    // every echo load is in flight before the (long) modulator / projector arithmetic starts
    cx<V> S[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e)
        if (e < ne) S[e] = ld_cx(p.acqs + (static_cast<size_t>(b) * ne + e) * nv * 2, v, V{});
    Mod<V> m[NE];
    cx<V> q_w = czero<V>(), q_f = czero<V>();          // M^+ Wm
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            m[e] = modulator(T, e, phi_t, r2map, zero);
            const cx<V> wm{vmul(m[e].dinv, m[e].c), vneg(vmul(m[e].dinv, m[e].s))};
            cmac(q_w, T.r[e].pw_re, T.r[e].pw_im, wm);
            cmac(q_f, T.r[e].pf_re, T.r[e].pf_im, wm);
        }
    }
    // normal equations of the weighted fit: G = M^H W M (Hermitian 2x2), rhs = M^H W y
    V g00 = zero, g11 = zero;
    cx<V> g01 = czero<V>(), r0 = czero<V>(), r1 = czero<V>();
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const float te = T.r[e].te, k = kTwoPi * te;
            V Vv = one_minus_exp_neg_fast(vmul(k * k, s_phi));
            if (r2) Vv = vfma(vmul(te * te, m[e].dinv), s_r, Vv);                      // e^{te mu} is the demodulator's own growth factor
            const cx<V> wm{vmul(m[e].dinv, m[e].c), vneg(vmul(m[e].dinv, m[e].s))};
            const cx<V> pw = caffine(q_w, T.r[e].c_re, T.r[e].c_im, q_f);              // (M M^+ Wm)_e
            const cx<V> res{vsub(wm.re, pw.re), vsub(wm.im, pw.im)};                    // (P0 Wm)_e
            // |Wp_e (P0 Wm)_e|^2 = d_e^2 |(P0 Wm)_e|^2: the phasor has unit modulus
            const V g2 = vmul(vmul(m[e].d, m[e].d), vfma(res.re, res.re, vmul(res.im, res.im)));
            const V s2 = vfma(S[e].re, S[e].re, vmul(S[e].im, S[e].im));
            const V w = vrcp_or_zero(vmul(Vv, vadd(g2, s2)));
            const cx<V> y = demod(m[e], S[e]);
            const float cr = T.r[e].c_re, ci = T.r[e].c_im;
            g00 = vadd(g00, w);
            g01.re = vfma(cr, w, g01.re);
            g01.im = vfma(ci, w, g01.im);
            g11 = vfma(cr * cr + ci * ci, w, g11);
            r0.re = vfma(w, y.re, r0.re);
            r0.im = vfma(w, y.im, r0.im);
            r1.re = vfma(w, vfma(ci, y.im, vmul(cr, y.re)), r1.re);                     // conj(c) y
            r1.im = vfma(w, vfma(-ci, y.re, vmul(cr, y.im)), r1.im);
        }
    }
    // C = G^-1 = 1/det [[g11, -g01], [-conj(g01), g00]]
    const V det = vsub(vmul(g00, g11), vfma(g01.re, g01.re, vmul(g01.im, g01.im)));
    const V id = vdiv1(det);
    const V c00 = vmul(g11, id), c11 = vmul(g00, id);
    const cx<V> c01{vneg(vmul(g01.re, id)), vneg(vmul(g01.im, id))};
    const cx<V> rho_w{vfma(c00, r0.re, vsub(vmul(c01.re, r1.re), vmul(c01.im, r1.im))), vfma(c00, r0.im, vfma(c01.re, r1.im, vmul(c01.im, r1.re)))};
    const cx<V> rho_f{vfma(c11, r1.re, vfma(c01.re, r0.re, vmul(c01.im, r0.im))), vfma(c11, r1.im, vsub(vmul(c01.re, r0.im), vmul(c01.im, r0.re)))};
    const float inv = 1.0f / kRhoSc, inv2 = 1.0f / (kRhoSc * kRhoSc);
    float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
    st_cx(rho_b, v, cx<V>{vmul(inv, rho_w.re), vmul(inv, rho_w.im)});
    st_cx(rho_b + static_cast<size_t>(nv) * 2, v, cx<V>{vmul(inv, rho_f.re), vmul(inv, rho_f.im)});
    const V a01 = vmul(inv2, vsqrt(vfma(c01.re, c01.re, vmul(c01.im, c01.im))));
    float *cov_b = p.cov + static_cast<size_t>(b) * 4 * nv;
    st_real(cov_b, v, vmul(inv2, vabs(c00)));
    st_real(cov_b + nv, v, a01);
    st_real(cov_b + 2 * static_cast<size_t>(nv), v, a01);
    st_real(cov_b + 3 * static_cast<size_t>(nv), v, vmul(inv2, vabs(c11)));
}

// PDFF extraction (ROI-analysis.py:301-306,344-354; gen_LDM_dataset.py:217-218): mode 0 |F|/|W+F|, 1 |F|/(|W|+|F|),
// 2 magnitude-discriminated; NaN (0/0) -> 0
// Two voxels per thread where the shape allows (16-byte streaming loads of both species, one 8-byte store): the one-voxel version
// moved 20 bytes per thread and reached 54 % of the HBM rate.
template <typename V> __global__ void __launch_bounds__(kThreads) pdff_extract_kernel(const float *__restrict__ rho, int nb, int nv, int mode, float *__restrict__ out) {
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n, b = blockIdx.y;
    if (v0 >= nv) return;
    const float *rho_b = rho + static_cast<size_t>(b) * 2 * nv * 2;
    const cx<V> w = ld_cx(rho_b, v0, V{}), f = ld_cx(rho_b + static_cast<size_t>(nv) * 2, v0, V{});
    V res;
#pragma unroll
    for (int l = 0; l < lanes<V>::n; ++l) {
        const float wx = lane_get(w.re, l), wy = lane_get(w.im, l), fx = lane_get(f.re, l), fy = lane_get(f.im, l);
        const float wa = sqrtf(wx * wx + wy * wy), fa = sqrtf(fx * fx + fy * fy);
        float r;
        if (mode == 1) {
            r = fa / (wa + fa);
        } else {
            const float wf = sqrtf((wx + fx) * (wx + fx) + (wy + fy) * (wy + fy));
            r = (mode == 0 || fa >= wa) ? fa / wf : 1.0f - wa / wf;
        }
        lane_set(res, l, (isnan(r) || isinf(r)) ? 0.f : r);
    }
    st_real(out + static_cast<size_t>(b) * nv, v0, res);
}

static int check_nv(const char *fn, int nb, int ne, int nv, int min_ne) {
    IG_REQUIRE(nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "%s: nb=%d (1..65535), nv=%d", fn, nb, nv);
    IG_REQUIRE(ne >= min_ne && ne <= IG_MAX_NE, IG_E_NE, "%s: ne=%d outside [%d, %d]", fn, ne, min_ne, IG_MAX_NE);
    return 0;
}

}  // namespace ig

using namespace ig;

extern "C" int ig_eigenvals(const float *x_d, long n, float *xy_d, float *ratio_d, void *stream) {
    IG_REQUIRE(x_d && xy_d && ratio_d && n > 0, IG_E_ARG, "ig_eigenvals: null pointer or n <= 0");
    eigenvals_kernel<<<static_cast<unsigned>((n + kThreads - 1) / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x_d, n, xy_d, ratio_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_eigenvals_bwd(const float *x_d, long n, const float *g_xy_d, const float *g_ratio_d, float *gx_d, void *stream) {
    IG_REQUIRE(x_d && gx_d && n > 0, IG_E_ARG, "ig_eigenvals_bwd: null pointer or n <= 0");
    eigenvals_bwd_kernel<<<static_cast<unsigned>((n + kThreads - 1) / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x_d, n, g_xy_d, g_ratio_d, gx_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_cse_mag_fwd(const float *mag_d, const float *r2_d, const float *r2nu_d, const float *tab_d, int nb, int ne, int nv, float r2_sc,
                              float *rho_d, float *fit_d, float *demod_d, float *ls_d, float *unc_d, void *stream) {
    IG_REQUIRE(mag_d && r2_d && tab_d, IG_E_ARG, "ig_cse_mag_fwd: null pointer");
    if (int rc = check_nv("ig_cse_mag_fwd", nb, ne, nv, 3)) return rc;
    CseParams p{};
    p.mag = mag_d; p.r2 = r2_d; p.r2nu = r2nu_d; p.tab = tab_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    p.rho = rho_d; p.fit = fit_d; p.demod = demod_d; p.ls = ls_d; p.unc = unc_d;
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        auto al8 = [](const void *q) { return !q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
        if (nv % 2 == 0 && al8(mag_d) && al8(r2_d) && al8(r2nu_d) && al8(rho_d) && al8(fit_d) && al8(demod_d) && al8(ls_d) && al8(unc_d))
            cse_mag_fwd2_kernel<NE><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        else
            cse_mag_kernel<NE, false><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int ig_cse_mag_bwd(const float *mag_d, const float *r2_d, const float *r2nu_d, const float *tab_d, int nb, int ne, int nv, float r2_sc,
                              const float *g_rho_d, const float *g_fit_d, const float *g_demod_d, const float *g_ls_d, const float *g_unc_d,
                              float *g_mag_d, float *g_r2_d, float *g_r2nu_d, void *stream) {
    IG_REQUIRE(mag_d && r2_d && tab_d && (g_mag_d || g_r2_d || g_r2nu_d), IG_E_ARG, "ig_cse_mag_bwd: null pointer");
    if (int rc = check_nv("ig_cse_mag_bwd", nb, ne, nv, 3)) return rc;
    CseParams p{};
    p.mag = mag_d; p.r2 = r2_d; p.r2nu = r2nu_d; p.tab = tab_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    p.g_rho = g_rho_d; p.g_fit = g_fit_d; p.g_demod = g_demod_d; p.g_ls = g_ls_d; p.g_unc = g_unc_d;
    p.g_mag = g_mag_d; p.g_r2 = g_r2_d; p.g_r2nu = g_r2nu_d;
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        auto al8 = [](const void *q) { return !q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
        if (nv % 2 == 0 && al8(mag_d) && al8(r2_d) && al8(r2nu_d) && al8(g_rho_d) && al8(g_fit_d) && al8(g_demod_d) && al8(g_ls_d) && al8(g_unc_d) &&
            al8(g_mag_d) && al8(g_r2_d) && al8(g_r2nu_d))
            cse_mag_bwd2_kernel<NE><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        else
            cse_mag_kernel<NE, true><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int ig_acq_unc_fwd(const float *rho_d, const float *phi_var_d, const float *r2_mean_d, const float *r2_var_d, const float *tab_d,
                              int nb, int ne, int nv, float r2_sc, int only_mag, float *out_d, void *stream) {
    IG_REQUIRE(rho_d && phi_var_d && tab_d && out_d && (!r2_mean_d == !r2_var_d), IG_E_ARG, "ig_acq_unc_fwd: null pointer");
    if (int rc = check_nv("ig_acq_unc_fwd", nb, ne, nv, 1)) return rc;
    UncParams p{};
    p.rho = rho_d; p.phi_var = phi_var_d; p.r2_mean = r2_mean_d; p.r2_var = r2_var_d; p.tab = tab_d; p.out = out_d;
    p.nb = nb; p.ne = ne; p.nv = nv; p.ch = only_mag ? 1 : 2; p.r2_sc = r2_sc;
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        auto al8 = [](const void *q) { return !q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
        if (nv % 2 == 0 && aligned16(rho_d) && al8(phi_var_d) && al8(r2_mean_d) && al8(r2_var_d) && (only_mag ? al8(out_d) : aligned16(out_d)))
            only_mag ? acq_unc_kernel<NE, false, pk, 1><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p)
                     : acq_unc_kernel<NE, false, pk, 2><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        else
            only_mag ? acq_unc_kernel<NE, false, float, 1><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p)
                     : acq_unc_kernel<NE, false, float, 2><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int ig_acq_unc_bwd(const float *rho_d, const float *phi_var_d, const float *r2_mean_d, const float *r2_var_d, const float *tab_d,
                              int nb, int ne, int nv, float r2_sc, int only_mag, const float *g_out_d, float *g_phi_var_d, float *g_r2_mean_d,
                              float *g_r2_var_d, void *stream) {
    IG_REQUIRE(rho_d && phi_var_d && tab_d && g_out_d && g_phi_var_d && (!r2_mean_d == !r2_var_d), IG_E_ARG, "ig_acq_unc_bwd: null pointer");
    if (int rc = check_nv("ig_acq_unc_bwd", nb, ne, nv, 1)) return rc;
    UncParams p{};
    p.rho = rho_d; p.phi_var = phi_var_d; p.r2_mean = r2_mean_d; p.r2_var = r2_var_d; p.tab = tab_d; p.g_out = g_out_d;
    p.g_phi_var = g_phi_var_d; p.g_r2_mean = g_r2_mean_d; p.g_r2_var = g_r2_var_d;
    p.nb = nb; p.ne = ne; p.nv = nv; p.ch = only_mag ? 1 : 2; p.r2_sc = r2_sc;
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        auto al8 = [](const void *q) { return !q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
        if (nv % 2 == 0 && aligned16(rho_d) && al8(phi_var_d) && al8(r2_mean_d) && al8(r2_var_d) && (only_mag ? al8(g_out_d) : aligned16(g_out_d)) &&
            al8(g_phi_var_d) && al8(g_r2_mean_d) && al8(g_r2_var_d))
            only_mag ? acq_unc_kernel<NE, true, pk, 1><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p)
                     : acq_unc_kernel<NE, true, pk, 2><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        else
            only_mag ? acq_unc_kernel<NE, true, float, 1><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p)
                     : acq_unc_kernel<NE, true, float, 2><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int ig_pdff_unc(const float *acqs_d, const float *phi_mean_d, const float *phi_var_d, const float *r2_mean_d, const float *r2_var_d,
                           const float *tab_d, int nb, int ne, int nv, float r2_sc, float *rho_d, float *cov_d, void *stream) {
    IG_REQUIRE(acqs_d && phi_mean_d && phi_var_d && tab_d && rho_d && cov_d && (!r2_mean_d == !r2_var_d), IG_E_ARG, "ig_pdff_unc: null pointer");
    if (int rc = check_nv("ig_pdff_unc", nb, ne, nv, 2)) return rc;
    PdffUncParams p{};
    p.acqs = acqs_d; p.phi_mean = phi_mean_d; p.phi_var = phi_var_d; p.r2_mean = r2_mean_d; p.r2_var = r2_var_d; p.tab = tab_d;
    p.rho = rho_d; p.cov = cov_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    {
        // 128-voxel rows, <= 12 echoes: the packed fit on the generic TMA ring (ig_ring_ops.cu, PdffUncOp)
        const int rc = pdff_unc_ring(acqs_d, phi_mean_d, phi_var_d, r2_mean_d, r2_var_d, tab_d, nb, ne, nv, r2_sc, rho_d, cov_d, static_cast<cudaStream_t>(stream));
        if (rc != IG_E_UNSUPPORTED) return rc;
    }
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        // measured at 64 x 384 x 384 x 6: one voxel per thread 0.272 ms (58 registers, issue slots 79 % busy); the packed instantiation
        // (V = pk) 0.279 ms: fewer instructions, but 112 registers leave 16 warps per SM and the kernel waits on its own dependency chains
        pdff_unc_kernel<NE, float><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int ig_pdff_extract(const float *rho_d, int nb, int nv, int mode, float *out_d, void *stream) {
    IG_REQUIRE(rho_d && out_d && nb > 0 && nv > 0 && nb <= 65535 && mode >= 0 && mode <= 2, IG_E_ARG, "ig_pdff_extract: bad arguments");
    if (nv % 2 == 0 && aligned16(rho_d) && (reinterpret_cast<uintptr_t>(out_d) & 7u) == 0)
        pdff_extract_kernel<pk><<<grid_for(nb, nv, 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(rho_d, nb, nv, mode, out_d);
    else
        pdff_extract_kernel<float><<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(rho_d, nb, nv, mode, out_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}
