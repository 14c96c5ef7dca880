// Per-voxel state of the three forward-model parameterisations (IDEAL_model / IDEAL_mag / IDEAL_mag_phase,
// /root/reference/wflib/IDEAL_model.py:220-299,404-509) and the write-out of their map gradients: shared by the plain kernels
// of ig_forward.cu and the TMA-ring objective of ig_ring_ops.cu.
#pragma once
#include "ig_common.cuh"

namespace ig {

// decoded per-voxel model state
template <typename V> struct Voxel {
    cx<V> rhoW, rhoF;   // already multiplied by rho_sc
    V phi_t;            // phi / fm_sc map value
    V r2raw, r2;        // R2* map value before / after the relu gate
    V bturn;            // bipolar phase in turns
    // model-specific leftovers needed by the adjoint
    cx<V> uW, uF;       // unit phasors of the species phases (FFPD: uW = common phasor)
    V ff, pd;           // FFPD; MAGPHA: the raw |W|, |F| map values
};

template <typename V> __device__ __forceinline__ cx<V> ld_row(const float *maps_b, int row, int nv, int v0) {
    return ld_cx(maps_b + static_cast<size_t>(row) * nv * 2, v0, V{});
}
template <typename V> __device__ __forceinline__ void st_row(float *g_b, int row, int nv, int v0, const cx<V> &z) {
    st_cx(g_b + static_cast<size_t>(row) * nv * 2, v0, z);
}

template <typename V, int MODEL> __device__ __forceinline__ Voxel<V> decode(const float *maps_b, int rows_or_ch, int nv, int v0, int flags) {
    Voxel<V> x;
    const V zero = splat<V>(0.f);
    x.bturn = zero;
    x.ff = zero;
    x.pd = zero;
    x.uW = cx<V>{zero, zero};
    x.uF = cx<V>{zero, zero};
    if constexpr (MODEL == IG_MODEL_WFPM) {
        const cx<V> m0 = ld_row<V>(maps_b, 0, nv, v0), m1 = ld_row<V>(maps_b, 1, nv, v0), m2 = ld_row<V>(maps_b, 2, nv, v0);
        x.rhoW = cx<V>{vmul(kRhoSc, m0.re), vmul(kRhoSc, m0.im)};
        x.rhoF = cx<V>{vmul(kRhoSc, m1.re), vmul(kRhoSc, m1.im)};
        x.phi_t = m2.re;
        x.r2raw = m2.im;
        x.r2 = (flags & IG_F_NO_RELU) ? m2.im : vrelu(m2.im);
        if (rows_or_ch > 3) x.bturn = vmul(0.5f, ld_row<V>(maps_b, rows_or_ch - 1, nv, v0).re);   // pi * b / (2 pi)
    } else if constexpr (MODEL == IG_MODEL_FFPD) {
        const cx<V> m0 = ld_row<V>(maps_b, 0, nv, v0), m1 = ld_row<V>(maps_b, 1, nv, v0), m2 = ld_row<V>(maps_b, 2, nv, v0);
        x.ff = m0.re;
        x.pd = m1.re;
        x.r2raw = x.r2 = m1.im;
        x.phi_t = m2.im;
        unit_phasor(vmul(2.0f, m2.re), x.uW.re, x.uW.im);                                        // 4 pi p / (2 pi)
        const V a = vmul(kRhoSc, x.pd);
        const V aw = vfma(vneg(a), x.ff, a), af = vmul(a, x.ff);
        x.rhoW = cscale(aw, x.uW);
        x.rhoF = cscale(af, x.uW);
    } else {
        V magW, magF, pW, pF;
        if (lanes<V>::n == 2 && rows_or_ch == 3) {
            // the pair's six floats of a row are 24 contiguous bytes (8-byte aligned: v0 is even): three 8-byte loads per row instead of six 4-byte ones
            const float2 *r0 = reinterpret_cast<const float2 *>(maps_b + static_cast<size_t>(v0) * 3);
            const float2 *r1 = reinterpret_cast<const float2 *>(maps_b + (static_cast<size_t>(nv) + v0) * 3);
            const float2 a = __ldcs(r0), b = __ldcs(r0 + 1), c = __ldcs(r0 + 2), d = __ldcs(r1), e = __ldcs(r1 + 1), f = __ldcs(r1 + 2);
            lane_set(magW, 0, a.x); lane_set(magF, 0, a.y); lane_set(x.r2, 0, b.x); lane_set(magW, 1, b.y); lane_set(magF, 1, c.x); lane_set(x.r2, 1, c.y);
            lane_set(pW, 0, d.x); lane_set(pF, 0, d.y); lane_set(x.phi_t, 0, e.x); lane_set(pW, 1, e.y); lane_set(pF, 1, f.x); lane_set(x.phi_t, 1, f.y);
        } else {
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            const float *r0 = maps_b + (static_cast<size_t>(v0) + l) * rows_or_ch;
            const float *r1 = r0 + static_cast<size_t>(nv) * rows_or_ch;
            float a0, a1, a2, b0, b1, b2, b3 = 0.f;
            if (rows_or_ch == 4) {
                const float4 p = __ldcs(reinterpret_cast<const float4 *>(r0)), q = __ldcs(reinterpret_cast<const float4 *>(r1));
                a0 = p.x; a1 = p.y; a2 = p.z; b0 = q.x; b1 = q.y; b2 = q.z; b3 = q.w;
            } else {
                a0 = __ldcs(r0); a1 = __ldcs(r0 + 1); a2 = __ldcs(r0 + 2);
                b0 = __ldcs(r1); b1 = __ldcs(r1 + 1); b2 = __ldcs(r1 + 2);
            }
            lane_set(magW, l, a0); lane_set(magF, l, a1); lane_set(x.r2, l, a2);
            lane_set(pW, l, b0); lane_set(pF, l, b1); lane_set(x.phi_t, l, b2); lane_set(x.bturn, l, 2.0f * b3);
        }
        }
        x.r2raw = x.r2;
        x.ff = magW;                                     // raw (signed) magnitudes: the PDFF image of ig_ideal_decode divides these
        x.pd = magF;
        unit_phasor(vmul(2.0f, pW), x.uW.re, x.uW.im);
        unit_phasor(vmul(2.0f, pF), x.uF.re, x.uF.im);
        x.rhoW = cscale(vmul(kRhoSc, magW), x.uW);
        x.rhoF = cscale(vmul(kRhoSc, magF), x.uF);
    }
    return x;
}

// adjoint accumulators over echoes, all in the demodulated frame (g_e = conj(w_e) G_e, q_e = conj(g_e) yhat_e)
template <typename V> struct Adj {
    cx<V> sg;    // sum_e g_e                    = conj(a_W)
    cx<V> sgc;   // sum_e conj(c_e) g_e          = conj(a_F)
    cx<V> tq;    // sum_e te_e q_e
    cx<V> q;     // sum_e q_e
    V bq;        // sum_e s_e Im q_e
};

template <typename V, int MODEL>
__device__ __forceinline__ void write_grads(float *g_b, int rows_or_ch, int nv, int v0, int flags, const Voxel<V> &x, const Adj<V> &a,
                                            float r2_sc, float scale) {
    // d/d(phi map) = fm_sc sum_e Re(conj(G) 2 pi i te S) = -2 pi fm_sc Im(tq);  d/d(R2 map) = -r2_sc Re(tq)
    const V gphi = vmul(-kTwoPi * kFmSc * scale, a.tq.im);
    V gr2 = vmul(-r2_sc * scale, a.tq.re);
    const V zero = splat<V>(0.f);
    if constexpr (MODEL == IG_MODEL_WFPM) {
        if (!(flags & IG_F_NO_RELU)) gr2 = vgate(x.r2raw, gr2);
        st_row<V>(g_b, 0, nv, v0, cx<V>{vmul(kRhoSc * scale, a.sg.re), vmul(kRhoSc * scale, a.sg.im)});
        st_row<V>(g_b, 1, nv, v0, cx<V>{vmul(kRhoSc * scale, a.sgc.re), vmul(kRhoSc * scale, a.sgc.im)});
        st_row<V>(g_b, 2, nv, v0, cx<V>{gphi, gr2});
        for (int r = 3; r < rows_or_ch - 1; ++r) st_row<V>(g_b, r, nv, v0, cx<V>{zero, zero});      // rows the model never reads (IDEAL_model.py:246 takes the LAST row)
        if (rows_or_ch > 3) st_row<V>(g_b, rows_or_ch - 1, nv, v0, cx<V>{vmul(-0.5f * kTwoPi * scale, a.bq), zero});
    } else if constexpr (MODEL == IG_MODEL_FFPD) {
        // Re(u0 conj(z)) = u.re z.re + u.im z.im with z = sg / sgc
        const V uw = vfma(x.uW.im, a.sg.im, vmul(x.uW.re, a.sg.re));
        const V uf = vfma(x.uW.im, a.sgc.im, vmul(x.uW.re, a.sgc.re));
        const V dff = vmul(vmul(kRhoSc * scale, x.pd), vsub(uf, uw));
        const V mix = vfma(x.ff, vsub(uf, uw), uw);                                   // (1 - ff) uw + ff uf
        const V dpd = vmul(kRhoSc * scale, mix);
        const V dpha = vmul(-2.0f * kTwoPi * scale, a.q.im);                          // 4 pi Re(i T)
        st_row<V>(g_b, 0, nv, v0, cx<V>{dff, zero});
        st_row<V>(g_b, 1, nv, v0, cx<V>{dpd, gr2});
        st_row<V>(g_b, 2, nv, v0, cx<V>{dpha, gphi});
    } else {
        const V dmw = vmul(kRhoSc * scale, vfma(x.uW.im, a.sg.im, vmul(x.uW.re, a.sg.re)));
        const V dmf = vmul(kRhoSc * scale, vfma(x.uF.im, a.sgc.im, vmul(x.uF.re, a.sgc.re)));
        // 4 pi Re(i rho conj(sg)) = -4 pi (rho.im sg.re - rho.re sg.im)
        const V dpw = vmul(-2.0f * kTwoPi * scale, vfma(vneg(x.rhoW.re), a.sg.im, vmul(x.rhoW.im, a.sg.re)));
        const V dpf = vmul(-2.0f * kTwoPi * scale, vfma(vneg(x.rhoF.re), a.sgc.im, vmul(x.rhoF.im, a.sgc.re)));
        const V dbip = vmul(-2.0f * kTwoPi * scale, a.bq);
        if (lanes<V>::n == 2 && rows_or_ch == 3) {
            float2 *r0 = reinterpret_cast<float2 *>(g_b + static_cast<size_t>(v0) * 3);
            float2 *r1 = reinterpret_cast<float2 *>(g_b + (static_cast<size_t>(nv) + v0) * 3);
            __stcs(r0, make_float2(lane_get(dmw, 0), lane_get(dmf, 0)));
            __stcs(r0 + 1, make_float2(lane_get(gr2, 0), lane_get(dmw, 1)));
            __stcs(r0 + 2, make_float2(lane_get(dmf, 1), lane_get(gr2, 1)));
            __stcs(r1, make_float2(lane_get(dpw, 0), lane_get(dpf, 0)));
            __stcs(r1 + 1, make_float2(lane_get(gphi, 0), lane_get(dpw, 1)));
            __stcs(r1 + 2, make_float2(lane_get(dpf, 1), lane_get(gphi, 1)));
        } else {
#pragma unroll
        for (int l = 0; l < lanes<V>::n; ++l) {
            float *r0 = g_b + (static_cast<size_t>(v0) + l) * rows_or_ch;
            float *r1 = r0 + static_cast<size_t>(nv) * rows_or_ch;
            if (rows_or_ch == 4) {
                __stcs(reinterpret_cast<float4 *>(r0), make_float4(lane_get(dmw, l), lane_get(dmf, l), lane_get(gr2, l), 0.f));
                __stcs(reinterpret_cast<float4 *>(r1), make_float4(lane_get(dpw, l), lane_get(dpf, l), lane_get(gphi, l), lane_get(dbip, l)));
            } else {
                __stcs(r0, lane_get(dmw, l)); __stcs(r0 + 1, lane_get(dmf, l)); __stcs(r0 + 2, lane_get(gr2, l));
                __stcs(r1, lane_get(dpw, l)); __stcs(r1 + 1, lane_get(dpf, l)); __stcs(r1 + 2, lane_get(gphi, l));
            }
        }
        }
    }
}

}  // namespace ig
