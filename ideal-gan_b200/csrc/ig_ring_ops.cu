// Operators on the generic TMA ring (ig_ring.cuh): the acq_to_acq adjoint an unmodified training script reaches through autodiff
// (train-IDEAL-unsup.py:216-218,255), the fused bipolar mag/phase objective of train-IDEAL-single.py:154-157,175, and the Rician
// objective of the R2* stage (train-IDEAL-unsup.py:267-292).  Same arithmetic as their plain kernels in ig_solve.cu / ig_forward.cu /
// ig_uq.cu, which remain the path for shapes the ring does not cover (nv % 128 != 0, unaligned bases, more than 8 echoes); what
// changes is where the operands come from: a shared-memory stage filled by TMA instead of 2 ne + 1 global loads per thread.
// Each entry returns IG_E_UNSUPPORTED (without touching the error string) when the caller should use its plain kernel.
#include "ig_model.cuh"
#include "ig_ring.cuh"
#include "ig_uq.cuh"

namespace ig {

constexpr int kRingSmemBudget = 216 * 1024;
constexpr int kPlaneF4 = kRingTileVox * 8 / 16;          // float4 per complex plane of a tile

// (cw - i sw) S for the thread's two voxels, S as loaded (re0, im0, re1, im1): demodulation with the scale folded into (cw, sw)
__device__ __forceinline__ cx<pk> conj_rot(pk cw, pk sw, const float4 &r) {
    cx<pk> y;
    y.re = mk(fmaf(sw.d.x, r.y, cw.d.x * r.x), fmaf(sw.d.y, r.w, cw.d.y * r.z));
    y.im = mk(fmaf(-sw.d.x, r.x, cw.d.x * r.y), fmaf(-sw.d.y, r.z, cw.d.y * r.w));
    return y;
}
__device__ __forceinline__ bool all_zero(const float4 &r) { return r.x == 0.f && r.y == 0.f && r.z == 0.f && r.w == 0.f; }


// =================================================================================================
// acq_to_acq adjoint (math: ig_solve.cu, a2a_bwd_kernel)
// =================================================================================================
struct A2aBwdParams {
    const float *acqs, *pm, *tab, *g_rho, *g_shat;
    float *g_acqs, *g_pm, *loss;
    void *scratch;
    long pm_bstride;
    int nb, ne, nv, tile_stride;
    float r2_sc, inv_n;
};

template <int NE, bool EXACT, bool DS> struct A2aBwdOp {
    using Params = A2aBwdParams;
    struct Shared {};
    static constexpr int kNE = NE, kMaps = 3;
    static constexpr bool kExact = EXACT, kDynamic = false, kLoss = false, kWritesStage = false;
    static constexpr int fpv(int) { return 2; }
    static constexpr int planes_max(int m) { return m < 2 ? NE : 1; }
    static constexpr int kStageBytes = (2 * NE + 1) * kRingTileVox * 8 + ((NE * 64 + 127) / 128) * 128;
    // dPM only: two blocks per SM with three (or two) stages each where they fit, one block with as many stages as fit otherwise (9..12 echoes:
    // two).  With dS as well (ne more output planes per voxel) one block per SM with up to four stages is as fast or faster at every echo
    // count (same-call A/B at 4 / 6 / 8 echoes: 0.1765 -> 0.1642, 0.2626 -> 0.2604, 0.3291 -> 0.3283 ms).
    static constexpr bool kTwoBlocks = !DS && 2 * 2 * kStageBytes <= kRingSmemBudget;
    static constexpr int kOneBlockStages = 4 * kStageBytes <= kRingSmemBudget ? 4 : (3 * kStageBytes <= kRingSmemBudget ? 3 : 2);
    static constexpr int kStages = kTwoBlocks ? (2 * 3 * kStageBytes > kRingSmemBudget ? 2 : 3) : kOneBlockStages;
    static constexpr int kMinBlocks = kTwoBlocks ? 2 : 1;
    static_assert(kStages * kStageBytes <= kRingSmemBudget, "the ring needs at least two stages in shared memory");
    __host__ __device__ static int planes(int m, int ne, const Params &p) { return m == 0 ? ne : (m == 1 ? (p.g_shat ? ne : 0) : 1); }
    __device__ static void prologue(Shared &) {}

    __device__ static __forceinline__ void chunk(const Params &p, Shared &, unsigned char *stage, const SampleTab<NE> &T, int slot, int b, int v0,
                                                 bool active, int ne, float &) {
        using Lay = RingLayout<A2aBwdOp>;
        const float4 *sS = reinterpret_cast<const float4 *>(stage + Lay::off(0)) + slot;
        const float4 *sG = reinterpret_cast<const float4 *>(stage + Lay::off(1)) + slot;
        if constexpr (!DS) {
            // background chunk (every measured component zero): y = 0, so rho = t = 0 and d/dPM vanishes whatever the upstream is
            bool nz = false;
            if (active) {
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (EXACT || e < ne) nz = nz || !all_zero(sS[e * kPlaneF4]);
            }
            if (!__any_sync(0xffffffffu, nz)) {
                if (active) st_cx(p.g_pm + static_cast<size_t>(b) * p.nv * 2, v0, czero<pk>());
                return;
            }
        }
        if (!active) return;
        const float4 m4 = reinterpret_cast<const float4 *>(stage + Lay::off(2))[slot];
        const pk zero = splat<pk>(0.f);
        const pk phi_t = mk(m4.x, m4.z), r2s = vmul(p.r2_sc, mk(m4.y, m4.w));
        const bool has_g = p.g_shat != nullptr;
        const int nv = p.nv;
        cx<pk> rw = czero<pk>(), rf = czero<pk>(), tw = czero<pk>(), tf = czero<pk>();
        cx<pk> gw = czero<pk>(), gf = czero<pk>(), aw = czero<pk>(), af = czero<pk>();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                const Mod<pk> m = modulator_rec<pk, false, true>(R, phi_t, r2s, zero);
                const cx<pk> y = conj_rot(vmul(m.c, m.dinv), vmul(m.s, m.dinv), sS[e * kPlaneF4]);
                cmac(rw, R.pw_re, R.pw_im, y);
                cmac(rf, R.pf_re, R.pf_im, y);
                cmac(tw, R.tpw_re, R.tpw_im, y);
                cmac(tf, R.tpf_re, R.tpf_im, y);
                if (has_g) {
                    const cx<pk> v = conj_rot(vmul(m.c, m.d), vmul(m.s, m.d), sG[e * kPlaneF4]);      // conj(Wp) G
                    gw.re = vadd(gw.re, v.re);
                    gw.im = vadd(gw.im, v.im);
                    cmac(gf, R.c_re, -R.c_im, v);
                    aw.re = vfma(R.te, v.re, aw.re);
                    aw.im = vfma(R.te, v.im, aw.im);
                    cmac(af, R.te * R.c_re, -R.te * R.c_im, v);
                }
            }
        }
        const size_t plane = static_cast<size_t>(nv) * 2;
        if (p.g_rho) {
            const float inv = 1.0f / kRhoSc;
            const float *g_b = p.g_rho + static_cast<size_t>(b) * 2 * plane;
            const cx<pk> a = ld_cx(g_b, v0, pk{}), c = ld_cx(g_b + plane, v0, pk{});
            gw.re = vfma(inv, a.re, gw.re); gw.im = vfma(inv, a.im, gw.im);
            gf.re = vfma(inv, c.re, gf.re); gf.im = vfma(inv, c.im, gf.im);
        }
        cx<pk> X = cmulc(gw, tw);
        const cx<pk> x1 = cmulc(gf, tf), x2 = cmulc(aw, rw), x3 = cmulc(af, rf);
        X.re = vsub(vadd(X.re, x1.re), vadd(x2.re, x3.re));
        X.im = vsub(vadd(X.im, x1.im), vadd(x2.im, x3.im));
        st_cx(p.g_pm + static_cast<size_t>(b) * plane, v0, cx<pk>{vmul(kTwoPi * kFmSc, X.im), vmul(p.r2_sc, X.re)});
        if constexpr (DS) {
            float *ga_b = p.g_acqs + static_cast<size_t>(b) * ne * plane;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (EXACT || e < ne) {
                    const EchoRec R = T.r[e];
                    const Mod<pk> m = modulator_rec<pk, false, true>(R, phi_t, r2s, zero);
                    cx<pk> gy = czero<pk>();
                    cmac(gy, R.pw_re, -R.pw_im, gw);
                    cmac(gy, R.pf_re, -R.pf_im, gf);
                    st_cx(ga_b + e * plane, v0, remod_inv(m, gy));
                }
            }
        }
    }
};

template <class Op> static int launch_bwd(const A2aBwdParams &p, cudaStream_t st) {
    RingMaps maps{};
    const long acq_planes = static_cast<long>(p.nb) * p.ne, plane = static_cast<long>(p.nv) * 2;
    if (!ring_tensor_map(&maps.m[0], p.acqs, p.nv, 2, acq_planes, plane, p.ne)) return IG_E_UNSUPPORTED;
    if (p.g_shat && !ring_tensor_map(&maps.m[1], p.g_shat, p.nv, 2, acq_planes, plane, p.ne)) return IG_E_UNSUPPORTED;
    if (!ring_tensor_map(&maps.m[2], p.pm, p.nv, 2, p.nb, p.pm_bstride, 1)) return IG_E_UNSUPPORTED;
    return ring_launch<Op>(p, maps, st);
}

int a2a_bwd_ring(const float *acqs, const float *pm, long pm_bstride, const float *tab, int nb, int ne, int nv, float r2_sc, const float *g_rho,
                 const float *g_shat, float *g_acqs, float *g_pm, cudaStream_t st) {
    if (ne > 12 || nv % 128 != 0 || !aligned16(g_pm) || (g_acqs && !aligned16(g_acqs)) || (g_rho && !aligned16(g_rho))) return IG_E_UNSUPPORTED;
    A2aBwdParams p{};
    p.acqs = acqs; p.pm = pm; p.pm_bstride = pm_bstride; p.tab = tab; p.g_rho = g_rho; p.g_shat = g_shat; p.g_acqs = g_acqs; p.g_pm = g_pm;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    // 2..8 echoes: two blocks per SM; 9..12 (train-IDEAL-TEaug.py:614-618): 76-100 KB stages, one block per SM, two stages
    return dispatch_exact_ne<2, 12>(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        return g_acqs ? launch_bwd<A2aBwdOp<NE, true, true>>(p, st) : launch_bwd<A2aBwdOp<NE, true, false>>(p, st);
    });
}

// =================================================================================================
// fused forward -> mask -> MSE -> adjoint of the mag/phase model with 4-channel rows (math: ig_forward.cu, ideal_voxels<MODE_LOSS>)
// =================================================================================================
struct MagphaLossParams {
    const float *maps, *acqs, *tab;
    float *gmaps, *shat, *loss;
    void *scratch;
    int nb, ne, nv, tile_stride;
    float r2_sc, inv_n;
};

template <int NE, bool EXACT> struct MagphaLossOp {
    using Params = MagphaLossParams;
    struct Shared {};
    static constexpr int kNE = NE, kMaps = 2;
    static constexpr bool kExact = EXACT, kDynamic = true, kLoss = true, kWritesStage = false;
    static constexpr int fpv(int m) { return m == 0 ? 4 : 2; }
    static constexpr int planes_max(int m) { return m == 0 ? 2 : NE; }
    static constexpr int kStageBytes = 2 * kRingTileVox * 16 + NE * kRingTileVox * 8 + ((NE * 64 + 127) / 128) * 128;
    static constexpr int kStages = 2 * 3 * kStageBytes <= kRingSmemBudget ? 3 : 2, kMinBlocks = 2;
    __host__ __device__ static int planes(int m, int ne, const Params &) { return m == 0 ? 2 : ne; }
    __device__ static void prologue(Shared &) {}

    __device__ static __forceinline__ void chunk(const Params &p, Shared &, unsigned char *stage, const SampleTab<NE> &T, int slot, int b, int v0,
                                                 bool active, int ne, float &loss_part) {
        using Lay = RingLayout<MagphaLossOp>;
        const float4 *sA = reinterpret_cast<const float4 *>(stage + Lay::off(1)) + slot;
        const int nv = p.nv;
        const size_t map_elems = static_cast<size_t>(2) * nv * 4;
        float *g_b = p.gmaps + b * map_elems;
        // background chunk (every measured component zero): the mask removes every residual -> loss 0, gradient 0
        // (BWD: a chunk whose upstream is zero everywhere has a zero gradient as well)
        bool nz = false;
        if (active) {
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (EXACT || e < ne) nz = nz || !all_zero(sA[e * kPlaneF4]);
        }
        if (!__any_sync(0xffffffffu, nz) && !p.shat) {
            if (active) {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 *r0 = reinterpret_cast<float4 *>(g_b) + v0, *r1 = r0 + nv;
                __stcs(r0, z); __stcs(r0 + 1, z); __stcs(r1, z); __stcs(r1 + 1, z);
            }
            return;
        }
        if (!active) return;
        // decode the two voxels' rows: row 0 = (|W|, |F|, R2*, -), row 1 = (pW, pF, phi, bip) / (4 pi | fm_sc)
        const float4 *sM0 = reinterpret_cast<const float4 *>(stage + Lay::off(0)) + slot * 2;
        const float4 *sM1 = sM0 + kRingTileVox;
        const float4 a0 = sM0[0], a1 = sM0[1], b0 = sM1[0], b1 = sM1[1];
        Voxel<pk> x;
        const pk zero = splat<pk>(0.f);
        x.ff = zero; x.pd = zero;
        x.r2raw = x.r2 = mk(a0.z, a1.z);
        x.phi_t = mk(b0.z, b1.z);
        x.bturn = mk(2.0f * b0.w, 2.0f * b1.w);
        unit_phasor(mk(2.0f * b0.x, 2.0f * b1.x), x.uW.re, x.uW.im);
        unit_phasor(mk(2.0f * b0.y, 2.0f * b1.y), x.uF.re, x.uF.im);
        x.rhoW = cscale(vmul(kRhoSc, mk(a0.x, a1.x)), x.uW);
        x.rhoF = cscale(vmul(kRhoSc, mk(a0.y, a1.y)), x.uF);
        const pk r2s = vmul(p.r2_sc, x.r2);                               // the stage carries the unscaled decay constant
        Adj<pk> a;
        a.sg = czero<pk>(); a.sgc = czero<pk>(); a.tq = czero<pk>(); a.q = czero<pk>(); a.bq = zero;
        pk lsum = zero;
        const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                pk c, s;
                unit_phasor(vfma(R.sgn, x.bturn, vmul(R.kphi, x.phi_t)), c, s);
                const pk d = fast_ex2(vmul(R.kdec, r2s));
                const cx<pk> w{vmul(d, c), vmul(d, s)};
                const cx<pk> yhat = caffine(x.rhoW, R.c_re, R.c_im, x.rhoF);
                const cx<pk> shat = cmulv(w, yhat);
                const float4 A = sA[e * kPlaneF4];
                const cx<pk> G{mask_sub(shat.re, mk(A.x, A.z)), mask_sub(shat.im, mk(A.y, A.w))};
                lsum = vfma(G.re, G.re, lsum);
                lsum = vfma(G.im, G.im, lsum);
                if (p.shat) st_cx(p.shat + acq_b + static_cast<size_t>(e) * nv * 2, v0, shat);
                const cx<pk> g = cmulc(w, G);
                a.sg.re = vadd(a.sg.re, g.re);
                a.sg.im = vadd(a.sg.im, g.im);
                cmac(a.sgc, R.c_re, -R.c_im, g);
                const cx<pk> q = cmulc(g, yhat);
                a.tq.re = vfma(R.te, q.re, a.tq.re);
                a.tq.im = vfma(R.te, q.im, a.tq.im);
                a.bq = vfma(R.sgn, q.im, a.bq);
            }
        }
        loss_part += hsum(lsum);
        write_grads<pk, IG_MODEL_MAGPHA>(g_b, 4, nv, v0, 0, x, a, p.r2_sc, 2.0f * p.inv_n);
    }
};

int magpha_loss_ring(const float *maps, const float *acqs, const float *tab, int nb, int ne, int nv, float r2_sc, float inv_n, float *gmaps, float *shat,
                     float *loss, void *scratch, cudaStream_t st) {
    if (ne > 8 || nv % 128 != 0 || !aligned16(gmaps) || (shat && !aligned16(shat)) || static_cast<long>(nb) * (nv / kRingTileVox + 1) >= (1L << 24))
        return IG_E_UNSUPPORTED;
    MagphaLossParams p{};
    p.maps = maps; p.acqs = acqs; p.tab = tab; p.gmaps = gmaps; p.shat = shat; p.loss = loss; p.scratch = scratch;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.inv_n = inv_n;
    RingMaps m{};
    if (!ring_tensor_map(&m.m[0], maps, nv, 4, static_cast<long>(nb) * 2, static_cast<long>(nv) * 4, 2)) return IG_E_UNSUPPORTED;
    if (!ring_tensor_map(&m.m[1], acqs, nv, 2, static_cast<long>(nb) * ne, static_cast<long>(nv) * 2, ne)) return IG_E_UNSUPPORTED;
    return dispatch_exact_ne<1, 8>(ne, [&](auto ne_c) { return ring_launch<MagphaLossOp<decltype(ne_c)::value, true>>(p, m, st); });
}

// =================================================================================================
// the same fused objective for the complex-row parameterisations (WF-PM maps with 3 or 4 rows, ff/pd/phase maps): the rows are
// planes of (Re, Im) pairs like the echoes, so one tensor map with `rows` planes delivers them (math: ideal_voxels<MODE_LOSS>)
// =================================================================================================
struct RowLossParams {
    const float *maps, *acqs, *tab;
    float *gmaps, *shat, *loss;
    void *scratch;
    int nb, ne, nv, tile_stride, rows, flags;
    float r2_sc, inv_n;
};

// BWD: the adjoint alone (IDEAL_Layer's autodiff, train-IDEAL-TEaug.py:192 with up to 12 echoes): `acqs` carries the upstream gradient, which
// takes the place of the masked residual; no loss, no work counter (tiles are dealt round-robin), gradients unscaled.
template <int MODEL, int NE, bool EXACT, bool BWD = false> struct RowLossOp {
    using Params = RowLossParams;
    struct Shared {};
    static constexpr int kNE = NE, kMaps = 2;
    static constexpr bool kExact = EXACT, kDynamic = !BWD, kLoss = !BWD, kWritesStage = false;
    static constexpr int fpv(int) { return 2; }
    static constexpr int planes_max(int m) { return m == 0 ? 4 : NE; }
    static constexpr int kStageBytes = (4 + NE) * kRingTileVox * 8 + ((NE * 64 + 127) / 128) * 128;
    // two blocks per SM where two stages each fit (<= 8 echoes), one block with three 65 KB stages for 9..12 echoes (as A2aBwdOp)
    static constexpr bool kTwoBlocks = 2 * 2 * kStageBytes <= kRingSmemBudget;
    static constexpr int kStages = kTwoBlocks ? (2 * 3 * kStageBytes <= kRingSmemBudget ? 3 : 2) : (3 * kStageBytes <= kRingSmemBudget ? 3 : 2);
    static constexpr int kMinBlocks = kTwoBlocks ? 2 : 1;
    static_assert(kStages * kStageBytes <= kRingSmemBudget, "the ring needs at least two stages in shared memory");
    __host__ __device__ static int planes(int m, int ne, const Params &p) { return m == 0 ? p.rows : ne; }
    __device__ static void prologue(Shared &) {}

    __device__ static __forceinline__ void chunk(const Params &p, Shared &, unsigned char *stage, const SampleTab<NE> &T, int slot, int b, int v0,
                                                 bool active, int ne, float &loss_part) {
        using Lay = RingLayout<RowLossOp>;
        const float4 *sA = reinterpret_cast<const float4 *>(stage + Lay::off(1)) + slot;
        const float4 *sM = reinterpret_cast<const float4 *>(stage + Lay::off(0)) + slot;
        const int nv = p.nv, rows = p.rows;
        float *g_b = p.gmaps + static_cast<size_t>(b) * rows * nv * 2;
        const pk zero = splat<pk>(0.f);
        // background chunk (every measured component zero): the mask removes every residual -> loss 0, gradient 0
        // (BWD: a chunk whose upstream is zero everywhere has a zero gradient as well)
        bool nz = false;
        if (active) {
#pragma unroll
            for (int e = 0; e < NE; ++e)
                if (EXACT || e < ne) nz = nz || !all_zero(sA[e * kPlaneF4]);
        }
        if (!__any_sync(0xffffffffu, nz) && !p.shat) {
            if (active)
                for (int r = 0; r < rows; ++r) st_row<pk>(g_b, r, nv, v0, czero<pk>());
            return;
        }
        if (!active) return;
        auto row = [&](int r) { const float4 q = sM[r * kPlaneF4]; return cx<pk>{mk(q.x, q.z), mk(q.y, q.w)}; };
        Voxel<pk> x;
        x.bturn = zero; x.ff = zero; x.pd = zero; x.uW = czero<pk>(); x.uF = czero<pk>();
        const cx<pk> m0 = row(0), m1 = row(1), m2 = row(2);
        if constexpr (MODEL == IG_MODEL_WFPM) {
            x.rhoW = cx<pk>{vmul(kRhoSc, m0.re), vmul(kRhoSc, m0.im)};
            x.rhoF = cx<pk>{vmul(kRhoSc, m1.re), vmul(kRhoSc, m1.im)};
            x.phi_t = m2.re;
            x.r2raw = m2.im;
            x.r2 = (p.flags & IG_F_NO_RELU) ? m2.im : vrelu(m2.im);
            if (rows > 3) x.bturn = vmul(0.5f, row(rows - 1).re);
        } else {
            x.ff = m0.re;
            x.pd = m1.re;
            x.r2raw = x.r2 = m1.im;
            x.phi_t = m2.im;
            unit_phasor(vmul(2.0f, m2.re), x.uW.re, x.uW.im);
            const pk aa = vmul(kRhoSc, x.pd);
            x.rhoW = cscale(vfma(vneg(aa), x.ff, aa), x.uW);
            x.rhoF = cscale(vmul(aa, x.ff), x.uW);
        }
        const pk r2s = vmul(p.r2_sc, x.r2);                               // the stage carries the unscaled decay constant
        Adj<pk> a;
        a.sg = czero<pk>(); a.sgc = czero<pk>(); a.tq = czero<pk>(); a.q = czero<pk>(); a.bq = zero;
        pk lsum = zero;
        const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                pk c, s;
                unit_phasor(vfma(R.sgn, x.bturn, vmul(R.kphi, x.phi_t)), c, s);
                const pk d = fast_ex2(vmul(R.kdec, r2s));
                const cx<pk> w{vmul(d, c), vmul(d, s)};
                const cx<pk> yhat = caffine(x.rhoW, R.c_re, R.c_im, x.rhoF);
                const cx<pk> shat = cmulv(w, yhat);
                const float4 A = sA[e * kPlaneF4];
                cx<pk> G;
                if constexpr (BWD) {
                    G = cx<pk>{mk(A.x, A.z), mk(A.y, A.w)};
                } else {
                    G = cx<pk>{mask_sub(shat.re, mk(A.x, A.z)), mask_sub(shat.im, mk(A.y, A.w))};
                    lsum = vfma(G.re, G.re, lsum);
                    lsum = vfma(G.im, G.im, lsum);
                    if (p.shat) st_cx(p.shat + acq_b + static_cast<size_t>(e) * nv * 2, v0, shat);
                }
                const cx<pk> g = cmulc(w, G);
                a.sg.re = vadd(a.sg.re, g.re);
                a.sg.im = vadd(a.sg.im, g.im);
                cmac(a.sgc, R.c_re, -R.c_im, g);
                const cx<pk> q = cmulc(g, yhat);
                a.tq.re = vfma(R.te, q.re, a.tq.re);
                a.tq.im = vfma(R.te, q.im, a.tq.im);
                if constexpr (MODEL == IG_MODEL_FFPD) { a.q.re = vadd(a.q.re, q.re); a.q.im = vadd(a.q.im, q.im); }
                else a.bq = vfma(R.sgn, q.im, a.bq);
            }
        }
        if constexpr (!BWD) loss_part += hsum(lsum);
        write_grads<pk, MODEL>(g_b, rows, nv, v0, p.flags, x, a, p.r2_sc, BWD ? 1.0f : 2.0f * p.inv_n);
    }
};

template <bool BWD>
static int row_ring(int model, const float *maps, int rows, const float *acqs_or_gout, const float *tab, int nb, int ne, int nv, float r2_sc, int flags,
                    float inv_n, float *gmaps, float *shat, float *loss, void *scratch, cudaStream_t st) {
    if (ne > 12 || rows > 4 || nv % 128 != 0 || !aligned16(gmaps) || (shat && !aligned16(shat)) || static_cast<long>(nb) * (nv / kRingTileVox + 1) >= (1L << 24))
        return IG_E_UNSUPPORTED;
    RowLossParams p{};
    p.maps = maps; p.acqs = acqs_or_gout; p.tab = tab; p.gmaps = gmaps; p.shat = shat; p.loss = loss; p.scratch = scratch;
    p.nb = nb; p.ne = ne; p.nv = nv; p.rows = rows; p.flags = flags; p.r2_sc = r2_sc; p.inv_n = inv_n;
    RingMaps m{};
    if (!ring_tensor_map(&m.m[0], maps, nv, 2, static_cast<long>(nb) * rows, static_cast<long>(nv) * 2, rows)) return IG_E_UNSUPPORTED;
    if (!ring_tensor_map(&m.m[1], acqs_or_gout, nv, 2, static_cast<long>(nb) * ne, static_cast<long>(nv) * 2, ne)) return IG_E_UNSUPPORTED;
    // 1..8 echoes: two blocks per SM; 9..12: one block per SM, three 53-65 KB stages.  The adjoint alone runs here beyond 8 echoes only.
    return dispatch_exact_ne<(BWD ? 9 : 1), 12>(ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        if (model == IG_MODEL_WFPM) return ring_launch<RowLossOp<IG_MODEL_WFPM, NE, true, BWD>>(p, m, st);
        return ring_launch<RowLossOp<IG_MODEL_FFPD, NE, true, BWD>>(p, m, st);
    });
}

int row_loss_ring(int model, const float *maps, int rows, const float *acqs, const float *tab, int nb, int ne, int nv, float r2_sc, int flags, float inv_n,
                  float *gmaps, float *shat, float *loss, void *scratch, cudaStream_t st) {
    return row_ring<false>(model, maps, rows, acqs, tab, nb, ne, nv, r2_sc, flags, inv_n, gmaps, shat, loss, scratch, st);
}

// the adjoint of the complex-row forward models for 9..12 echoes (the plain kernel holds 12 upstream echoes of two voxels in 110-126 registers
// and reaches 80-84 % of the HBM rate there; below 9 echoes it is at 95-99 % and stays the path)
int row_bwd_ring(int model, const float *maps, int rows, const float *gout, const float *tab, int nb, int ne, int nv, float r2_sc, int flags, float *gmaps,
                 cudaStream_t st) {
    if (ne <= 8 || !aligned16(gout)) return IG_E_UNSUPPORTED;
    return row_ring<true>(model, maps, rows, gout, tab, nb, ne, nv, r2_sc, flags, 1.0f, gmaps, nullptr, nullptr, nullptr, st);
}

// =================================================================================================
// Rician objective of the R2* stage (math: ig_uq.cu, a2a_rician_loss_kernel, packed lanes)
// =================================================================================================
struct RicianParams {
    const float *acqs, *pm, *phi_var, *r2_mean, *r2_var, *tab;
    float *g_pm, *g_phi_var, *g_r2_mean, *g_r2_var, *rho, *loss;
    void *scratch;
    long pm_bstride;
    int nb, ne, nv, tile_stride;
    float r2_sc, inv_n;
};

template <int NE, bool EXACT> struct RicianOp {
    using Params = RicianParams;
    struct Shared {
        float4 btab[kBesselRows * 3];
    };
    static constexpr int kNE = NE, kMaps = 5;
    // the decay and the observed magnitude of each echo wait for pass 2 in the thread's own 16 bytes of that echo's plane
    static constexpr bool kExact = EXACT, kDynamic = true, kLoss = true, kWritesStage = true;
    static constexpr int fpv(int m) { return m < 2 ? 2 : 1; }
    static constexpr int planes_max(int m) { return m == 0 ? NE : 1; }
    static constexpr int kStageBytes = (NE + 1) * kRingTileVox * 8 + 3 * kRingTileVox * 4 + ((NE * 64 + 127) / 128) * 128;
    static constexpr int kStatic = kBesselRows * 48 + 256;
    // One block of 15 consumer warps + the producer per SM (16 warps = 4 per scheduler = 128 registers per thread, no spills) on one ring of
    // six stages.  Same-call A/B at 64 x 384 x 384 x 6, masked / unmasked: two blocks of 8 + 1 warps at 96 registers (88 B of spills)
    // 0.2728 / 0.3200 ms, two blocks of 7 + 1 at 128 registers 0.2605 / 0.3014, this 0.2452 / 0.2933.
    static constexpr int kFit = (kRingSmemBudget - kStatic) / kStageBytes;      // 6 stages up to 6 echoes, 4 at 7..9, 3 at 10..12
    static constexpr int kStages = kFit > 6 ? 6 : kFit, kMinBlocks = 1, kConsumerWarps = 15;
    static_assert(kStages >= 2, "the ring needs at least two stages in shared memory");
    __host__ __device__ static int planes(int m, int ne, const Params &p) { return m == 0 ? ne : ((m >= 3 && !p.r2_mean) ? 0 : 1); }
    __device__ static void prologue(Shared &sh) { stage_bessel_table(sh.btab); }

    __device__ static __forceinline__ void chunk(const Params &p, Shared &sh, unsigned char *stage, const SampleTab<NE> &T, int slot, int b, int v0,
                                                 bool active, int ne, float &loss_part) {
        using Lay = RingLayout<RicianOp>;
        float4 *sS = reinterpret_cast<float4 *>(stage + Lay::off(0)) + slot;
        {
            // background chunk (every measured component zero): rho = 0 -> var = 0 -> the floor; y = nu = 0: each (echo, voxel) term is
            // log(1e-5), every gradient vanishes
            bool nz = false;
            if (active) {
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (EXACT || e < ne) nz = nz || !all_zero(sS[e * kPlaneF4]);
            }
            if (!__any_sync(0xffffffffu, nz)) {
                if (active) {
                    const pk z2 = splat<pk>(0.f);
                    const size_t pl = static_cast<size_t>(p.nv) * 2, o = static_cast<size_t>(b) * p.nv;
                    loss_part += static_cast<float>(ne) * 2.0f * -11.512925464970229f;
                    st_cx(p.g_pm + static_cast<size_t>(b) * pl, v0, czero<pk>());
                    st_real(p.g_phi_var + o, v0, z2);
                    if (p.g_r2_mean) st_real(p.g_r2_mean + o, v0, z2);
                    if (p.g_r2_var) st_real(p.g_r2_var + o, v0, z2);
                    if (p.rho) {
                        float *rho_b = p.rho + static_cast<size_t>(b) * 2 * pl;
                        st_cx(rho_b, v0, czero<pk>());
                        st_cx(rho_b + pl, v0, czero<pk>());
                    }
                }
                return;
            }
        }
        if (!active) return;
        const float4 m4 = reinterpret_cast<const float4 *>(stage + Lay::off(1))[slot];
        const bool rem = p.r2_mean == nullptr;
        const pk zero = splat<pk>(0.f);
        const int nv = p.nv;
        const float fm2 = kFmSc * kFmSc, r22 = p.r2_sc * p.r2_sc;
        const pk phi_t = mk(m4.x, m4.z), r2s = vmul(p.r2_sc, mk(m4.y, m4.w));
        pk t;
        t.d = reinterpret_cast<const float2 *>(stage + Lay::off(2))[slot];
        const pk sp2 = vmul(fm2, t);
        pk mu2 = zero, sr2 = zero;
        if (!rem) {
            t.d = reinterpret_cast<const float2 *>(stage + Lay::off(3))[slot];
            mu2 = vmul(p.r2_sc, t);
            t.d = reinterpret_cast<const float2 *>(stage + Lay::off(4))[slot];
            sr2 = vmul(r22, t);
        }
        cx<pk> rw = czero<pk>(), rf = czero<pk>(), tw = czero<pk>(), tf = czero<pk>();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                const Mod<pk> m = modulator_rec<pk, false, true>(R, phi_t, r2s, zero);
                const float4 S = sS[e * kPlaneF4];
                const cx<pk> y = conj_rot(vmul(m.c, m.dinv), vmul(m.s, m.dinv), S);
                cmac(rw, R.pw_re, R.pw_im, y);
                cmac(rf, R.pf_re, R.pf_im, y);
                cmac(tw, R.tpw_re, R.tpw_im, y);
                cmac(tf, R.tpf_re, R.tpf_im, y);
                float m0, m1;                  // observed magnitude with the mask (real channel == 0, :281) in its sign bit
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m0) : "f"(fmaf(S.x, S.x, S.y * S.y)));
                asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m1) : "f"(fmaf(S.z, S.z, S.w * S.w)));
                sS[e * kPlaneF4] = make_float4(m.d.d.x, m.d.d.y, copysignf(m0, S.x != 0.f ? 1.0f : -1.0f), copysignf(m1, S.z != 0.f ? 1.0f : -1.0f));
            }
        }
        UqAcc2 acc2{zero, zero, zero, zero};
        cx<pk> gw = czero<pk>(), gf = czero<pk>(), aw = czero<pk>(), af = czero<pk>();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                const cx<pk> yhat = caffine(rw, R.c_re, R.c_im, rf);
                const pk a2 = vfma(yhat.re, yhat.re, vmul(yhat.im, yhat.im));
                const pk ra = mk(a2.d.x > 1e-30f ? rsqrt_ftz(a2.d.x) : 0.f, a2.d.y > 1e-30f ? rsqrt_ftz(a2.d.y) : 0.f);
                const float4 q = sS[e * kPlaneF4];
                const pk dra = vmul(mk(q.x, q.y), ra);
                const pk g = rician_echo(sh.btab, R.te, a2, mk(q.z, q.w), vmul(dra, a2), sp2, mu2, sr2, rem, acc2);
                const cx<pk> v = cscale(vmul(g, dra), yhat);
                gw.re = vadd(gw.re, v.re);
                gw.im = vadd(gw.im, v.im);
                cmac(gf, R.c_re, -R.c_im, v);
                aw.re = vfma(R.te, v.re, aw.re);
                aw.im = vfma(R.te, v.im, aw.im);
                cmac(af, R.te * R.c_re, -R.te * R.c_im, v);
            }
        }
        cx<pk> X = cmulc(gw, tw);
        const cx<pk> x1 = cmulc(gf, tf), x2 = cmulc(aw, rw), x3 = cmulc(af, rf);
        X.re = vsub(vadd(X.re, x1.re), vadd(x2.re, x3.re));
        X.im = vsub(vadd(X.im, x1.im), vadd(x2.im, x3.im));
        const size_t plane = static_cast<size_t>(nv) * 2, o = static_cast<size_t>(b) * nv;
        st_cx(p.g_pm + static_cast<size_t>(b) * plane, v0, cx<pk>{vmul(kTwoPi * kFmSc * p.inv_n, X.im), vmul(p.r2_sc * p.inv_n, X.re)});
        loss_part += hsum(acc2.loss);
        st_real(p.g_phi_var + o, v0, vmul(fm2 * p.inv_n, acc2.g_sphi));
        if (p.g_r2_mean) st_real(p.g_r2_mean + o, v0, rem ? zero : vmul(p.r2_sc * p.inv_n, acc2.g_mu));
        if (p.g_r2_var) st_real(p.g_r2_var + o, v0, rem ? zero : vmul(r22 * p.inv_n, acc2.g_sr));
        if (p.rho) {
            const float inv = 1.0f / kRhoSc;
            float *rho_b = p.rho + static_cast<size_t>(b) * 2 * plane;
            st_cx(rho_b, v0, cx<pk>{vmul(inv, rw.re), vmul(inv, rw.im)});
            st_cx(rho_b + plane, v0, cx<pk>{vmul(inv, rf.re), vmul(inv, rf.im)});
        }
    }
};

int a2a_rician_loss_ring(const float *acqs, const float *pm, long pm_bstride, const float *phi_var, const float *r2_mean, const float *r2_var,
                         const float *tab, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm, float *g_phi_var, float *g_r2_mean,
                         float *g_r2_var, float *rho, float *loss, void *scratch, cudaStream_t st) {
    auto al8 = [](const void *q) { return !q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
    if (ne > 12 || nv % 128 != 0 || !aligned16(g_pm) || (rho && !aligned16(rho)) || !al8(g_phi_var) || !al8(g_r2_mean) || !al8(g_r2_var) ||
        static_cast<long>(nb) * (nv / kRingTileVox + 1) >= (1L << 24))
        return IG_E_UNSUPPORTED;
    RicianParams p{};
    p.acqs = acqs; p.pm = pm; p.pm_bstride = pm_bstride; p.phi_var = phi_var; p.r2_mean = r2_mean; p.r2_var = r2_var; p.tab = tab;
    p.g_pm = g_pm; p.g_phi_var = g_phi_var; p.g_r2_mean = g_r2_mean; p.g_r2_var = g_r2_var; p.rho = rho; p.loss = loss; p.scratch = scratch;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc; p.inv_n = inv_n;
    RingMaps m{};
    if (!ring_tensor_map(&m.m[0], acqs, nv, 2, static_cast<long>(nb) * ne, static_cast<long>(nv) * 2, ne)) return IG_E_UNSUPPORTED;
    if (!ring_tensor_map(&m.m[1], pm, nv, 2, nb, pm_bstride, 1)) return IG_E_UNSUPPORTED;
    if (!ring_tensor_map(&m.m[2], phi_var, nv, 1, nb, nv, 1)) return IG_E_UNSUPPORTED;
    if (r2_mean && (!ring_tensor_map(&m.m[3], r2_mean, nv, 1, nb, nv, 1) || !ring_tensor_map(&m.m[4], r2_var, nv, 1, nb, nv, 1))) return IG_E_UNSUPPORTED;
    return dispatch_exact_ne<2, 12>(ne, [&](auto ne_c) { return ring_launch<RicianOp<decltype(ne_c)::value, true>>(p, m, st); });
}

// =================================================================================================
// PDFF_uncertainty: per-voxel weighted least squares with an echo-wise noise model (math: ig_tier2.cu, pdff_unc_kernel;
// IDEAL_model.py:628-706).  The plain kernel keeps six modulators in registers on scalar lanes (1080 instructions per voxel, issue-bound at
// 51 % of the HBM rate); its packed instantiation needs 112 registers and lost to its own dependency chains at 16 warps per SM.  Here the
// echoes come from the stage and the demodulated echoes go back into it, the demodulator of each echo is the only per-echo state held
// in registers, and the two voxels of a lane share every instruction of the fit on f32x2 lanes.
// =================================================================================================
struct PdffUncRingParams {
    const float *acqs, *phi_mean, *phi_var, *r2_mean, *r2_var, *tab;
    float *rho, *cov, *loss;
    void *scratch;
    int nb, ne, nv, tile_stride;
    float r2_sc, inv_n;
};

template <int NE, bool EXACT> struct PdffUncOp {
    using Params = PdffUncRingParams;
    struct Shared {};
    static constexpr int kNE = NE, kMaps = 5;
    static constexpr bool kExact = EXACT, kDynamic = false, kLoss = false, kWritesStage = true;       // y_e is parked in the echo planes
    static constexpr int fpv(int m) { return m == 0 ? 2 : 1; }
    static constexpr int planes_max(int m) { return m == 0 ? NE : 1; }
    static constexpr int kStageBytes = NE * kRingTileVox * 8 + 4 * kRingTileVox * 4 + ((NE * 64 + 127) / 128) * 128;
    // one block of 15 consumer warps + the producer per SM at 128 registers, one ring of six (four) stages: same-call A/B against two blocks of
    // 8 + 1 warps at 96 registers 0.1842 -> 0.1683 ms (75 -> 82 % of HBM); seven warps per block at 128 registers measured 0.1829
    static constexpr int kFit = kRingSmemBudget / kStageBytes;
    static constexpr int kStages = kFit > 6 ? 6 : kFit, kMinBlocks = 1, kConsumerWarps = 15;
    static_assert(kStages >= 2, "the ring needs at least two stages in shared memory");
    __host__ __device__ static int planes(int m, int ne, const Params &p) { return m == 0 ? ne : ((m >= 3 && !p.r2_mean) ? 0 : 1); }
    __device__ static void prologue(Shared &) {}

    __device__ static __forceinline__ void chunk(const Params &p, Shared &, unsigned char *stage, const SampleTab<NE> &T, int slot, int b, int v0,
                                                 bool active, int ne, float &) {
        using Lay = RingLayout<PdffUncOp>;
        if (!active) return;
        const bool r2 = p.r2_mean != nullptr;
        const pk zero = splat<pk>(0.f);
        pk phi_t, s_phi, r2s = zero, s_r = zero;
        phi_t.d = reinterpret_cast<const float2 *>(stage + Lay::off(1))[slot];
        s_phi.d = reinterpret_cast<const float2 *>(stage + Lay::off(2))[slot];
        s_phi = vmul(kFmSc * kFmSc, s_phi);
        if (r2) {
            r2s.d = reinterpret_cast<const float2 *>(stage + Lay::off(3))[slot];
            r2s = vmul(p.r2_sc, r2s);                                      // the stage carries the unscaled decay constant
            s_r.d = reinterpret_cast<const float2 *>(stage + Lay::off(4))[slot];
            s_r = vmul(p.r2_sc * p.r2_sc, s_r);
        }
        // Pass 1: the demodulator Wm_e = e^{te R} conj(u_e) is formed ONCE per echo and kept (two register pairs); the demodulated echo
        // y_e = Wm_e S_e replaces the raw echo in the thread's own 16 bytes of the stage.  Everything pass 2 needs follows from those two:
        // |S_e|^2 = d_e^2 |y_e|^2 and d_e = |Wm_e|^-1, so no modulator is formed twice and no echo is read twice.
        cx<pk> q_w = czero<pk>(), q_f = czero<pk>();                        // M^+ Wm
        cx<pk> wm[NE];
        float4 *sY = reinterpret_cast<float4 *>(stage + Lay::off(0)) + slot;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                pk c, sn;
                unit_phasor(vmul(R.kphi, phi_t), c, sn);
                const pk dinv = fast_ex2(vneg(vmul(R.kdec, r2s)));
                wm[e] = cx<pk>{vmul(dinv, c), vneg(vmul(dinv, sn))};
                const cx<pk> y = conj_rot(wm[e].re, vneg(wm[e].im), sY[e * kPlaneF4]);      // (re - i(-im)) ... = Wm S
                sY[e * kPlaneF4] = make_float4(y.re.d.x, y.re.d.y, y.im.d.x, y.im.d.y);
                cmac(q_w, R.pw_re, R.pw_im, wm[e]);
                cmac(q_f, R.pf_re, R.pf_im, wm[e]);
            }
        }
        // Pass 2: normal equations of the weighted fit, G = M^H W M (Hermitian 2x2), rhs = M^H W y, with
        // 1 / w_e = V_e (d_e^2 |(P0 Wm)_e|^2 + |S_e|^2) = (vphi_e d_e^2 + te_e^2 s_r d_e) (|(P0 Wm)_e|^2 + |y_e|^2)       (V_e = vphi_e + te_e^2 s_r / d_e)
        pk g00 = zero, g11 = zero;
        cx<pk> g01 = czero<pk>(), r0 = czero<pk>(), r1 = czero<pk>();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (EXACT || e < ne) {
                const EchoRec R = T.r[e];
                const float k = kTwoPi * R.te;
                const pk n2 = vfma(wm[e].re, wm[e].re, vmul(wm[e].im, wm[e].im));     // |Wm|^2 = 1 / d^2
                const pk d = mk(rsqrt_ftz(n2.d.x), rsqrt_ftz(n2.d.y));
                const pk d2 = vmul(d, d);
                const pk vphi = one_minus_exp_neg_fast(vmul(k * k, s_phi));
                pk vd = vmul(vphi, d2);
                if (r2) vd = vfma(vmul(R.te * R.te, s_r), d, vd);
                const cx<pk> pw = caffine(q_w, R.c_re, R.c_im, q_f);                  // (M M^+ Wm)_e
                const cx<pk> res{vsub(wm[e].re, pw.re), vsub(wm[e].im, pw.im)};        // (P0 Wm)_e
                const float4 q = sY[e * kPlaneF4];
                const cx<pk> y{mk(q.x, q.y), mk(q.z, q.w)};
                const pk sum = vadd(vfma(res.re, res.re, vmul(res.im, res.im)), vfma(y.re, y.re, vmul(y.im, y.im)));
                const pk den = vmul(vd, sum);
                const pk w = mk(den.d.x != 0.f ? rcp_ftz(den.d.x) : 0.f, den.d.y != 0.f ? rcp_ftz(den.d.y) : 0.f);
                const float cr = R.c_re, ci = R.c_im;
                g00 = vadd(g00, w);
                g01.re = vfma(cr, w, g01.re);
                g01.im = vfma(ci, w, g01.im);
                g11 = vfma(cr * cr + ci * ci, w, g11);
                r0.re = vfma(w, y.re, r0.re);
                r0.im = vfma(w, y.im, r0.im);
                r1.re = vfma(w, vfma(ci, y.im, vmul(cr, y.re)), r1.re);                 // conj(c) y
                r1.im = vfma(w, vfma(-ci, y.re, vmul(cr, y.im)), r1.im);
            }
        }
        // C = G^-1 = 1/det [[g11, -g01], [-conj(g01), g00]]
        const pk det = vsub(vmul(g00, g11), vfma(g01.re, g01.re, vmul(g01.im, g01.im)));
        const pk id = mk(__fdividef(1.0f, det.d.x), __fdividef(1.0f, det.d.y));      // SFU reciprocal (2 ulp); the weights above carry more rounding than that
        const pk c00 = vmul(g11, id), c11 = vmul(g00, id);
        const cx<pk> c01{vneg(vmul(g01.re, id)), vneg(vmul(g01.im, id))};
        const cx<pk> rho_w{vfma(c00, r0.re, vsub(vmul(c01.re, r1.re), vmul(c01.im, r1.im))), vfma(c00, r0.im, vfma(c01.re, r1.im, vmul(c01.im, r1.re)))};
        const cx<pk> rho_f{vfma(c11, r1.re, vfma(c01.re, r0.re, vmul(c01.im, r0.im))), vfma(c11, r1.im, vsub(vmul(c01.re, r0.im), vmul(c01.im, r0.re)))};
        const float inv = 1.0f / kRhoSc, inv2 = 1.0f / (kRhoSc * kRhoSc);
        const int nv = p.nv;
        float *rho_b = p.rho + static_cast<size_t>(b) * 2 * nv * 2;
        st_cx(rho_b, v0, cx<pk>{vmul(inv, rho_w.re), vmul(inv, rho_w.im)});
        st_cx(rho_b + static_cast<size_t>(nv) * 2, v0, cx<pk>{vmul(inv, rho_f.re), vmul(inv, rho_f.im)});
        const pk m01 = vfma(c01.re, c01.re, vmul(c01.im, c01.im));
        float s0, s1;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(m01.d.x));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(m01.d.y));
        const pk a01 = vmul(inv2, mk(s0, s1));
        float *cov_b = p.cov + static_cast<size_t>(b) * 4 * nv;
        st_real(cov_b, v0, vmul(inv2, mk(fabsf(c00.d.x), fabsf(c00.d.y))));
        st_real(cov_b + nv, v0, a01);
        st_real(cov_b + 2 * static_cast<size_t>(nv), v0, a01);
        st_real(cov_b + 3 * static_cast<size_t>(nv), v0, vmul(inv2, mk(fabsf(c11.d.x), fabsf(c11.d.y))));
    }
};

int pdff_unc_ring(const float *acqs, const float *phi_mean, const float *phi_var, const float *r2_mean, const float *r2_var, const float *tab, int nb,
                  int ne, int nv, float r2_sc, float *rho, float *cov, cudaStream_t st) {
    auto al8 = [](const void *q) { return !q || (reinterpret_cast<uintptr_t>(q) & 7u) == 0; };
    if (ne > 12 || nv % 128 != 0 || !aligned16(rho) || !al8(cov) || static_cast<long>(nb) * (nv / kRingTileVox + 1) >= (1L << 24)) return IG_E_UNSUPPORTED;
    PdffUncRingParams p{};
    p.acqs = acqs; p.phi_mean = phi_mean; p.phi_var = phi_var; p.r2_mean = r2_mean; p.r2_var = r2_var; p.tab = tab; p.rho = rho; p.cov = cov;
    p.nb = nb; p.ne = ne; p.nv = nv; p.r2_sc = r2_sc;
    RingMaps m{};
    if (!ring_tensor_map(&m.m[0], acqs, nv, 2, static_cast<long>(nb) * ne, static_cast<long>(nv) * 2, ne)) return IG_E_UNSUPPORTED;
    if (!ring_tensor_map(&m.m[1], phi_mean, nv, 1, nb, nv, 1) || !ring_tensor_map(&m.m[2], phi_var, nv, 1, nb, nv, 1)) return IG_E_UNSUPPORTED;
    if (r2_mean && (!ring_tensor_map(&m.m[3], r2_mean, nv, 1, nb, nv, 1) || !ring_tensor_map(&m.m[4], r2_var, nv, 1, nb, nv, 1))) return IG_E_UNSUPPORTED;
    return dispatch_exact_ne<2, 12>(ne, [&](auto ne_c) { return ring_launch<PdffUncOp<decltype(ne_c)::value, true>>(p, m, st); });
}

}  // namespace ig
