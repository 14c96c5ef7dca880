// Fused uncertainty-aware physics objective of the published AI-DEAL model.
//
// Replaces, in one pass over the echoes, the op chain of train-IDEAL-unsup.py:214-231:
//     A2B_WF, A2B2A = wf.acq_to_acq(A, PM)                       (wflib/IDEAL_model.py:142-200)
//     A2B2A     = where(A != 0, A2B2A, 0)
//     A2B2A_var = wf.acq_uncertainty(stop_gradient(A2B_WF), FM, R2)   (wflib/IDEAL_model.py:710-767)
//     loss      = VarMeanSquaredError()(A, concat([A2B2A, A2B2A_var], -1))   (tf2gan/loss.py:130-140)
// and TF autodiff through all of it.  Per voxel, with y = Wm A, rho = M^+ y, yhat = M rho, r = yhat - y:
//     var_e  = V_e |yhat_e|^2,   V_e = 1 - e^{-(2 pi te_e)^2 s_phi} + e^{-te_e mu_R} te_e^2 s_R        (rho is a constant here)
//     loss   = (1/N) sum_e [ d_e^2 |r_e|^2 / std_e + 2 log std_e ],   std_e = sqrt(max(var_e, 1e-5))   (sigma, not sigma^2: loss.py:139)
// The gradient w.r.t. (phi~, R~) is the MSE objective's with the residual weighted by 1/std_e (see ig_solve.cu); the
// gradients w.r.t. the moment maps follow from d loss / d var_e = [var_e >= 1e-5] (1 - msd_e / (2 std_e)) / var_e.
// Voxels with some, but not all, components exactly zero take the reference's per-component mask on a scalar path.
#include "ig_uq.cuh"

namespace ig {

struct UqParams {
    const float *acqs, *pm, *phi_var, *r2_mean, *r2_var, *tab;
    float *g_pm, *g_phi_var, *g_r2_mean, *g_r2_var, *rho, *loss;
    void *scratch;
    long pm_bstride;
    int nb, ne, nv;
    float r2_sc, inv_n;
};

// a voxel is "ragged" iff some but not all of its 2 ne components are exactly zero
__device__ __forceinline__ void zero_range(float &lo, float &hi, float re, float im) {
    lo = fminf(fminf(lo, fabsf(re)), fabsf(im));
    hi = fmaxf(fmaxf(hi, fabsf(re)), fabsf(im));
}

template <int NE, typename V> __global__ void __launch_bounds__(kThreads, 2) a2a_uq_loss_kernel(const UqParams p) {
    __shared__ SampleTab<NE> T;
    const int nv = p.nv, ne = p.ne;
    constexpr int L = lanes<V>::n;
    const int tiles_ps = (nv + kThreads * L - 1) / (kThreads * L);
    const long total = static_cast<long>(p.nb) * tiles_ps;
    const int tile_end = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
    const bool rem = p.r2_mean == nullptr;
    const float fm2 = kFmSc * kFmSc, r22 = p.r2_sc * p.r2_sc;
    int cur_b = -1;
    float loss_part = 0.f;
    for (int tile = static_cast<int>(total * blockIdx.x / gridDim.x); tile < tile_end; ++tile) {
        const int b = tile / tiles_ps;
        if (b != cur_b) {
            if (cur_b >= 0) __syncthreads();
            stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, ne, p.r2_sc);
            cur_b = b;
        }
        const int v0 = ((tile - b * tiles_ps) * kThreads + threadIdx.x) * L;
        const bool active = v0 < nv;
        const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
        const size_t plane = static_cast<size_t>(nv) * 2;
        const size_t vb = static_cast<size_t>(b) * nv + v0;
        bool ragged = false;
        cx<V> S[NE];
        V phi_t = splat<V>(0.f), r2 = splat<V>(0.f), pv = splat<V>(0.f), rm = splat<V>(0.f), rv = splat<V>(0.f);
        if (active) {
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (e < ne) S[e] = ld_cx(p.acqs + acq_b + e * plane, v0, V{});
            const cx<V> m = ld_cx(p.pm + b * p.pm_bstride, v0, V{});
            phi_t = m.re;
            r2 = m.im;
            pv = ld_real(p.phi_var + static_cast<size_t>(b) * nv, v0, V{});
            if (!rem) {
                rm = ld_real(p.r2_mean + static_cast<size_t>(b) * nv, v0, V{});
                rv = ld_real(p.r2_var + static_cast<size_t>(b) * nv, v0, V{});
            }
#pragma unroll
            for (int l = 0; l < L; ++l) {
                float lo = 3.0e38f, hi = 0.f;
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (e < ne) zero_range(lo, hi, lane_get(S[e].re, l), lane_get(S[e].im, l));
                ragged = ragged || (lo == 0.f && hi > 0.f);
            }
        }
        const bool warp_ragged = __any_sync(0xffffffffu, ragged);
        if (!active) continue;
        UqAcc acc[L];
        cx<V> rw = czero<V>(), rf = czero<V>();
        V gphi, gr2;
        if (!warp_ragged) {
            const V zero = splat<V>(0.f);
            V d2[NE];
            cx<V> y[NE];
            cx<V> tw = czero<V>(), tf = czero<V>();
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (e < ne) {
                    const EchoRec R = T.r[e];
                    const Mod<V> m = modulator_rec<V, false, true>(R, phi_t, r2, zero);
                    d2[e] = vmul(m.d, m.d);
                    y[e] = demod(m, S[e]);
                    cmac(rw, R.pw_re, R.pw_im, y[e]);
                    cmac(rf, R.pf_re, R.pf_im, y[e]);
                    cmac(tw, R.tpw_re, R.tpw_im, y[e]);
                    cmac(tf, R.tpf_re, R.tpf_im, y[e]);
                }
            }
#pragma unroll
            for (int l = 0; l < L; ++l) acc[l] = UqAcc{0.f, 0.f, 0.f, 0.f};
            cx<V> K = czero<V>();
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (e < ne) {
                    const EchoRec R = T.r[e];
                    const cx<V> yhat = caffine(rw, R.c_re, R.c_im, rf);
                    const cx<V> h = caffine(tw, R.c_re, R.c_im, tf);
                    const cx<V> r{vsub(yhat.re, y[e].re), vsub(yhat.im, y[e].im)};
                    const V a2 = vfma(yhat.re, yhat.re, vmul(yhat.im, yhat.im));
                    const V msd = vmul(d2[e], vfma(r.re, r.re, vmul(r.im, r.im)));
                    V inv_std = zero;
#pragma unroll
                    for (int l = 0; l < L; ++l)
                        lane_set(inv_std, l, uq_echo(R.te, lane_get(a2, l), lane_get(msd, l), lane_get(pv, l) * fm2, lane_get(rm, l) * p.r2_sc,
                                                     lane_get(rv, l) * r22, rem, acc[l]));
                    const V wgt = vmul(inv_std, d2[e]);
                    const cx<V> w{vmul(wgt, r.re), vmul(wgt, r.im)};
                    const cx<V> g{vfma(R.te, yhat.re, vneg(h.re)), vfma(R.te, yhat.im, vneg(h.im))};
                    K.re = vfma(w.re, g.re, K.re);
                    K.re = vfma(w.im, g.im, K.re);
                    K.im = vfma(w.re, g.im, K.im);
                    K.im = vfma(vneg(w.im), g.re, K.im);
                }
            }
            gphi = vmul(-2.0f * kTwoPi * kFmSc * p.inv_n, K.im);
            gr2 = vmul(-2.0f * p.r2_sc * p.inv_n, K.re);
        } else {
#pragma unroll
            for (int l = 0; l < L; ++l) {
                acc[l] = UqAcc{0.f, 0.f, 0.f, 0.f};
                float gp, gr;
                cx<float> w1, f1;
                uq_slow_voxel<NE>(T, p.acqs + acq_b, ne, nv, v0 + l, lane_get(phi_t, l), lane_get(r2, l), lane_get(pv, l) * fm2,
                                  lane_get(rm, l) * p.r2_sc, lane_get(rv, l) * r22, rem, p.r2_sc, acc[l], gp, gr, w1, f1);
                lane_set(gphi, l, 2.0f * p.inv_n * gp);
                lane_set(gr2, l, 2.0f * p.inv_n * gr);
                lane_set(rw.re, l, w1.re); lane_set(rw.im, l, w1.im);
                lane_set(rf.re, l, f1.re); lane_set(rf.im, l, f1.im);
            }
        }
        st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, cx<V>{gphi, gr2});
        V o_pv, o_rm, o_rv;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            loss_part += acc[l].loss;
            lane_set(o_pv, l, acc[l].g_sphi * fm2 * p.inv_n);
            lane_set(o_rm, l, acc[l].g_mu * p.r2_sc * p.inv_n);
            lane_set(o_rv, l, acc[l].g_sr * r22 * p.inv_n);
        }
        st_real(p.g_phi_var + static_cast<size_t>(b) * nv, v0, o_pv);
        if (p.g_r2_mean) st_real(p.g_r2_mean + static_cast<size_t>(b) * nv, v0, rem ? splat<V>(0.f) : o_rm);
        if (p.g_r2_var) st_real(p.g_r2_var + static_cast<size_t>(b) * nv, v0, rem ? splat<V>(0.f) : o_rv);
        if (p.rho) {
            const float inv = 1.0f / kRhoSc;
            float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
            st_cx(rho_b, v0, cx<V>{vmul(inv, rw.re), vmul(inv, rw.im)});
            st_cx(rho_b + plane, v0, cx<V>{vmul(inv, rf.re), vmul(inv, rf.im)});
        }
        (void)vb;
    }
    block_loss_reduce(loss_part, p.scratch, p.loss, p.inv_n);
}

// =================================================================================================
// Rician (magnitude) objective of the R2* stage, train-IDEAL-unsup.py:267-292:
//     A2B_WF, |S_hat| = acq_to_acq(A, PM, only_mag=True);  |S_hat| <- where(A[..., :1] != 0, |S_hat|, 0)     (real channel decides, :281)
//     var = acq_uncertainty(stop_gradient(A2B_WF), FM, R2, only_mag=True);  loss = VarMeanSquaredErrorR2()(|A|, [|S_hat|, var])
// The upstream on |S_hat_e| = d_e |yhat_e| enters the adjoint of acq_to_acq as v_e = g_e d_e yhat_e / |yhat_e| (ig_solve.cu,
// a2a_bwd_kernel); the per-voxel accumulators are the same.  No ragged path: the mask only gates g_e.
// =================================================================================================
template <int NE, typename V> __global__ void __launch_bounds__(kThreads, 2) a2a_rician_loss_kernel(const UqParams p) {
    __shared__ SampleTab<NE> T;
    __shared__ float4 btab[kBesselRows * 3];
    stage_bessel_table(btab);            // visible after the __syncthreads of the first stage_table
    // up to 8 echoes: the decay and the observed magnitude of each echo wait for pass 2 in the thread's own shared-memory slots
    // instead of 4 registers per echo (24 at ne = 6, where the kernel sat at the 128-register cap with spills)
    constexpr bool PARK = lanes<V>::n == 2 && NE <= 8;
    __shared__ float4 park[PARK ? NE * kThreads : 1];
    const int nv = p.nv, ne = p.ne;
    constexpr int L = lanes<V>::n;
    const int tiles_ps = (nv + kThreads * L - 1) / (kThreads * L);
    const long total = static_cast<long>(p.nb) * tiles_ps;
    const int tile_end = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
    const bool rem = p.r2_mean == nullptr;
    const float fm2 = kFmSc * kFmSc, r22 = p.r2_sc * p.r2_sc;
    int cur_b = -1;
    float loss_part = 0.f;
    for (int tile = static_cast<int>(total * blockIdx.x / gridDim.x); tile < tile_end; ++tile) {
        const int b = tile / tiles_ps;
        if (b != cur_b) {
            if (cur_b >= 0) __syncthreads();
            stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, ne, p.r2_sc);
            cur_b = b;
        }
        const int v0 = ((tile - b * tiles_ps) * kThreads + threadIdx.x) * L;
        if (v0 >= nv) continue;
        const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
        const size_t plane = static_cast<size_t>(nv) * 2;
        const V zero = splat<V>(0.f);
        cx<V> S[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (e < ne) S[e] = ld_cx(p.acqs + acq_b + e * plane, v0, V{});
        const cx<V> m0 = ld_cx(p.pm + b * p.pm_bstride, v0, V{});
        const V phi_t = m0.re, r2 = m0.im;
        const V pv = ld_real(p.phi_var + static_cast<size_t>(b) * nv, v0, V{});
        V rm = zero, rv = zero;
        if (!rem) {
            rm = ld_real(p.r2_mean + static_cast<size_t>(b) * nv, v0, V{});
            rv = ld_real(p.r2_var + static_cast<size_t>(b) * nv, v0, V{});
        }
        cx<V> rw = czero<V>(), rf = czero<V>(), tw = czero<V>(), tf = czero<V>();
        V dec[NE], ys[NE];            // decay; observed magnitude with the mask in its sign bit (negative = masked)
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const EchoRec R = T.r[e];
                const Mod<V> m = modulator_rec<V, false, true>(R, phi_t, r2, zero);
                const cx<V> y = demod(m, S[e]);
                cmac(rw, R.pw_re, R.pw_im, y);
                cmac(rf, R.pf_re, R.pf_im, y);
                cmac(tw, R.tpw_re, R.tpw_im, y);
                cmac(tf, R.tpf_re, R.tpf_im, y);
                dec[e] = m.d;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const float re = lane_get(S[e].re, l), im = lane_get(S[e].im, l);
                    float mag;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(re, re, im * im)));
                    lane_set(ys[e], l, copysignf(mag, re != 0.f ? 1.0f : -1.0f));
                }
                if constexpr (PARK) park[e * kThreads + threadIdx.x] = make_float4(lane_get(dec[e], 0), lane_get(dec[e], 1), lane_get(ys[e], 0), lane_get(ys[e], 1));
            }
        }
        UqAcc acc[L];
#pragma unroll
        for (int l = 0; l < L; ++l) acc[l] = UqAcc{0.f, 0.f, 0.f, 0.f};
        UqAcc2 acc2{splat<pk>(0.f), splat<pk>(0.f), splat<pk>(0.f), splat<pk>(0.f)};
        cx<V> gw = czero<V>(), gf = czero<V>(), aw = czero<V>(), af = czero<V>();
        [[maybe_unused]] V sp2, mu2, sr2;
        if constexpr (L == 2) {
            sp2 = vmul(fm2, pv);
            mu2 = vmul(p.r2_sc, rm);
            sr2 = vmul(r22, rv);
        }
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < ne) {
                const EchoRec R = T.r[e];
                const cx<V> yhat = caffine(rw, R.c_re, R.c_im, rf);
                const V a2 = vfma(yhat.re, yhat.re, vmul(yhat.im, yhat.im));
                V sc = zero;       // g_nu d / |yhat|
                if constexpr (L == 2) {
                    const pk ra = mk(a2.d.x > 1e-30f ? rsqrt_ftz(a2.d.x) : 0.f, a2.d.y > 1e-30f ? rsqrt_ftz(a2.d.y) : 0.f);
                    pk de = dec[e], yo = ys[e];
                    if constexpr (PARK) {
                        const float4 q = park[e * kThreads + threadIdx.x];
                        de = mk(q.x, q.y);
                        yo = mk(q.z, q.w);
                    }
                    const pk dra = vmul(de, ra);
                    const pk g = rician_echo(btab, R.te, a2, yo, vmul(dra, a2), sp2, mu2, sr2, rem, acc2);
                    sc = vmul(g, dra);
                } else {
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        const float a2l = lane_get(a2, l), d = lane_get(dec[e], l), ysl = lane_get(ys[e], l);
                        const float ra = a2l > 1e-30f ? rsqrt_ftz(a2l) : 0.f;
                        const float g = rician_echo(btab, R.te, a2l, fabsf(ysl), d * a2l * ra, !signbit(ysl), lane_get(pv, l) * fm2,
                                                    lane_get(rm, l) * p.r2_sc, lane_get(rv, l) * r22, rem, acc[l]);
                        lane_set(sc, l, g * d * ra);
                    }
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
        if constexpr (L == 2) {
            acc[0] = UqAcc{acc2.g_sphi.d.x, acc2.g_mu.d.x, acc2.g_sr.d.x, acc2.loss.d.x};
            acc[1] = UqAcc{acc2.g_sphi.d.y, acc2.g_mu.d.y, acc2.g_sr.d.y, acc2.loss.d.y};
        }
        cx<V> X = cmulc(gw, tw);
        const cx<V> x1 = cmulc(gf, tf), x2 = cmulc(aw, rw), x3 = cmulc(af, rf);
        X.re = vsub(vadd(X.re, x1.re), vadd(x2.re, x3.re));
        X.im = vsub(vadd(X.im, x1.im), vadd(x2.im, x3.im));
        st_cx(p.g_pm + static_cast<size_t>(b) * nv * 2, v0, cx<V>{vmul(kTwoPi * kFmSc * p.inv_n, X.im), vmul(p.r2_sc * p.inv_n, X.re)});
        V o_pv, o_rm, o_rv;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            loss_part += acc[l].loss;
            lane_set(o_pv, l, acc[l].g_sphi * fm2 * p.inv_n);
            lane_set(o_rm, l, acc[l].g_mu * p.r2_sc * p.inv_n);
            lane_set(o_rv, l, acc[l].g_sr * r22 * p.inv_n);
        }
        st_real(p.g_phi_var + static_cast<size_t>(b) * nv, v0, o_pv);
        if (p.g_r2_mean) st_real(p.g_r2_mean + static_cast<size_t>(b) * nv, v0, rem ? zero : o_rm);
        if (p.g_r2_var) st_real(p.g_r2_var + static_cast<size_t>(b) * nv, v0, rem ? zero : o_rv);
        if (p.rho) {
            const float inv = 1.0f / kRhoSc;
            float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
            st_cx(rho_b, v0, cx<V>{vmul(inv, rw.re), vmul(inv, rw.im)});
            st_cx(rho_b + plane, v0, cx<V>{vmul(inv, rf.re), vmul(inv, rf.im)});
        }
    }
    block_loss_reduce(loss_part, p.scratch, p.loss, p.inv_n);
}

template <typename K> static int uq_grid(K kernel, int nb, int nv, int vpt, int *grid) {
    int dev = 0, sms = 0, occ = 0;
    IG_CUDA(cudaGetDevice(&dev));
    IG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    IG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0));
    const long tiles = static_cast<long>(nb) * ((nv + kThreads * vpt - 1) / (kThreads * vpt));
    const long g = static_cast<long>(sms) * (occ > 0 ? occ : 1);
    *grid = static_cast<int>(g < tiles ? g : tiles);
    return 0;
}

}  // namespace ig

using namespace ig;

static int uq_common_checks(const char *fn, const void *acqs_d, const void *pm_d, const void *phi_var_d, const void *r2_mean_d, const void *r2_var_d,
                            const void *tab_d, const void *g_pm_d, const void *g_phi_var_d, const void *loss_d, const void *scratch_d, int nb,
                            int ne, int nv, size_t scratch_bytes) {
    IG_REQUIRE(acqs_d && pm_d && phi_var_d && tab_d && g_pm_d && g_phi_var_d && loss_d && scratch_d && (!r2_mean_d == !r2_var_d), IG_E_ARG,
               "%s: null pointer (r2_mean and r2_var go together)", fn);
    IG_REQUIRE(nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "%s: nb=%d (1..65535), nv=%d", fn, nb, nv);
    IG_REQUIRE(ne >= 2 && ne <= IG_MAX_NE, IG_E_NE, "%s: ne=%d outside [2, %d]", fn, ne, IG_MAX_NE);
    IG_REQUIRE(scratch_bytes >= ig_loss_scratch_bytes(nb, nv), IG_E_SCRATCH, "%s: scratch %zu < %zu bytes", fn, scratch_bytes,
               ig_loss_scratch_bytes(nb, nv));
    return 0;
}

extern "C" int ig_a2a_rician_loss(const float *acqs_d, const float *pm_d, long pm_bstride, const float *phi_var_d, const float *r2_mean_d,
                                  const float *r2_var_d, const float *tab_d, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm_d,
                                  float *g_phi_var_d, float *g_r2_mean_d, float *g_r2_var_d, float *rho_d, float *loss_d, void *scratch_d,
                                  size_t scratch_bytes, void *stream) {
    if (int rc = uq_common_checks("ig_a2a_rician_loss", acqs_d, pm_d, phi_var_d, r2_mean_d, r2_var_d, tab_d, g_pm_d, g_phi_var_d, loss_d, scratch_d,
                                  nb, ne, nv, scratch_bytes))
        return rc;
    UqParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.phi_var = phi_var_d; p.r2_mean = r2_mean_d; p.r2_var = r2_var_d; p.tab = tab_d;
    p.g_pm = g_pm_d; p.g_phi_var = g_phi_var_d; p.g_r2_mean = g_r2_mean_d; p.g_r2_var = g_r2_var_d; p.rho = rho_d; p.loss = loss_d;
    p.scratch = scratch_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.inv_n = inv_n;
    bool packed = nv % 2 == 0 && pm_bstride % 4 == 0;
    for (const void *q : {static_cast<const void *>(acqs_d), static_cast<const void *>(pm_d), static_cast<const void *>(g_pm_d),
                          static_cast<const void *>(rho_d)})
        packed = packed && (!q || aligned16(q));
    for (const void *q : {static_cast<const void *>(phi_var_d), static_cast<const void *>(r2_mean_d), static_cast<const void *>(r2_var_d),
                          static_cast<const void *>(g_phi_var_d), static_cast<const void *>(g_r2_mean_d), static_cast<const void *>(g_r2_var_d)})
        packed = packed && (!q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (packed) {                                    // TMA ring (128-voxel rows, <= 12 echoes): ig_ring_ops.cu
        const int rc = a2a_rician_loss_ring(acqs_d, pm_d, pm_bstride, phi_var_d, r2_mean_d, r2_var_d, tab_d, nb, ne, nv, r2_sc, inv_n, g_pm_d,
                                            g_phi_var_d, g_r2_mean_d, g_r2_var_d, rho_d, loss_d, scratch_d, st);
        if (rc != IG_E_UNSUPPORTED) return rc;
    }
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        int grid = 1;
        if (packed) {
            if (int rc = uq_grid(a2a_rician_loss_kernel<NE, pk>, nb, nv, 2, &grid)) return rc;
            a2a_rician_loss_kernel<NE, pk><<<grid, kThreads, 0, st>>>(p);
        } else {
            if (int rc = uq_grid(a2a_rician_loss_kernel<NE, float>, nb, nv, 1, &grid)) return rc;
            a2a_rician_loss_kernel<NE, float><<<grid, kThreads, 0, st>>>(p);
        }
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

extern "C" int ig_a2a_uq_loss(const float *acqs_d, const float *pm_d, long pm_bstride, const float *phi_var_d, const float *r2_mean_d,
                              const float *r2_var_d, const float *tab_d, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm_d,
                              float *g_phi_var_d, float *g_r2_mean_d, float *g_r2_var_d, float *rho_d, float *loss_d, void *scratch_d,
                              size_t scratch_bytes, void *stream) {
    IG_REQUIRE(acqs_d && pm_d && phi_var_d && tab_d && g_pm_d && g_phi_var_d && loss_d && scratch_d && (!r2_mean_d == !r2_var_d), IG_E_ARG,
               "ig_a2a_uq_loss: null pointer (r2_mean and r2_var go together)");
    IG_REQUIRE(nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "ig_a2a_uq_loss: nb=%d (1..65535), nv=%d", nb, nv);
    IG_REQUIRE(ne >= 2 && ne <= IG_MAX_NE, IG_E_NE, "ig_a2a_uq_loss: ne=%d outside [2, %d]", ne, IG_MAX_NE);
    IG_REQUIRE(scratch_bytes >= ig_loss_scratch_bytes(nb, nv), IG_E_SCRATCH, "ig_a2a_uq_loss: scratch %zu < %zu bytes", scratch_bytes,
               ig_loss_scratch_bytes(nb, nv));
    UqParams p{};
    p.acqs = acqs_d; p.pm = pm_d; p.pm_bstride = pm_bstride; p.phi_var = phi_var_d; p.r2_mean = r2_mean_d; p.r2_var = r2_var_d; p.tab = tab_d;
    p.g_pm = g_pm_d; p.g_phi_var = g_phi_var_d; p.g_r2_mean = g_r2_mean_d; p.g_r2_var = g_r2_var_d; p.rho = rho_d; p.loss = loss_d;
    p.scratch = scratch_d; p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.inv_n = inv_n;
    bool packed = nv % 2 == 0 && pm_bstride % 4 == 0;
    for (const void *q : {static_cast<const void *>(acqs_d), static_cast<const void *>(pm_d), static_cast<const void *>(g_pm_d),
                          static_cast<const void *>(rho_d)})
        packed = packed && (!q || aligned16(q));
    for (const void *q : {static_cast<const void *>(phi_var_d), static_cast<const void *>(r2_mean_d), static_cast<const void *>(r2_var_d),
                          static_cast<const void *>(g_phi_var_d), static_cast<const void *>(g_r2_mean_d), static_cast<const void *>(g_r2_var_d)})
        packed = packed && (!q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        const int rc = a2a_uq_loss_ring(acqs_d, pm_d, pm_bstride, phi_var_d, r2_mean_d, r2_var_d, tab_d, nb, ne, nv, r2_sc, inv_n, g_pm_d,
                                        g_phi_var_d, g_r2_mean_d, g_r2_var_d, rho_d, loss_d, scratch_d, st);
        if (rc != IG_E_UNSUPPORTED) return rc;
    }
    return dispatch_ne(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        int grid = 1;
        if (packed) {
            if (int rc = uq_grid(a2a_uq_loss_kernel<NE, pk>, nb, nv, 2, &grid)) return rc;
            a2a_uq_loss_kernel<NE, pk><<<grid, kThreads, 0, st>>>(p);
        } else {
            if (int rc = uq_grid(a2a_uq_loss_kernel<NE, float>, nb, nv, 1, &grid)) return rc;
            a2a_uq_loss_kernel<NE, float><<<grid, kThreads, 0, st>>>(p);
        }
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}
