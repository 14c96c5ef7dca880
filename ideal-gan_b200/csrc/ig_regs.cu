// Script-level per-voxel reductions that sit around the physics operators:
//
//  * ig_mag_regs  -- the regularisers train-IDEAL-mag.py:288-289,308-316 adds to the generator objective from the
//    outputs of CSE_mag: anisotropic total variation of the demodulated echoes and of the R2* map
//    (tf.image.total_variation), the negative-coefficient penalty LS_NZ and the discriminant penalty LS_cond of the
//    quadratic fit (a, b, c).  One pass: four sums and the gradient of their weighted sum, instead of ~25 TF ops
//    with (nb, ne, H, W) temporaries each and their autodiff mirror images.
//  * ig_roi_maps  -- the map assembly and PDFF-variance propagation of ROI-analysis.py:301-322.
//
// Both are read-once / write-once streams: a thread owns VEC consecutive pixels of a row; the row neighbours of
// the total variation come from L1/L2 (the rows above and below are read by neighbouring blocks at the same time).
#include "ig_common.cuh"

namespace ig {

constexpr int kRegSums = 4;      // Ad_TV, LS_NZ, LS_cond, R2_TV  (WF_NZ is identically zero, see ig_mag_regs)

struct RegParams {
    const float *ls, *demod, *r2;
    float *g_ls, *g_demod, *g_r2;
    int nb, ne, H, W;
    float w_ad_tv, w_ls_nz, w_ls_cond, w_r2_tv;
};

// d|x|/dx with TF's sign(0) = 0: sign bit of x OR'ed onto [x != 0] (one compare, one select, one logic op)
__device__ __forceinline__ float sgn(float x) {
    return __uint_as_float((__float_as_uint(x) & 0x80000000u) | __float_as_uint(x != 0.f ? 1.0f : 0.0f));
}

// total variation of one plane at the VEC pixels starting at (row, col): forward differences are summed where they
// exist (tf.image.total_variation: |x[1:,:] - x[:-1,:]| + |x[:,1:] - x[:,:-1]|), the gradient collects the four
// differences a pixel takes part in.
template <int VEC>
__device__ __forceinline__ float tv_plane(const float *__restrict__ x, float *__restrict__ g, float w, int row, int col, int H, int W) {
    const size_t o = static_cast<size_t>(row) * W + col;
    float c[VEC], u[VEC], d[VEC];
    if constexpr (VEC == 4) {
        const float4 c4 = __ldg(reinterpret_cast<const float4 *>(x + o));
        const float4 u4 = row > 0 ? __ldg(reinterpret_cast<const float4 *>(x + o - W)) : c4;
        const float4 d4 = row < H - 1 ? __ldg(reinterpret_cast<const float4 *>(x + o + W)) : c4;
        c[0] = c4.x; c[1] = c4.y; c[2] = c4.z; c[3] = c4.w;
        u[0] = u4.x; u[1] = u4.y; u[2] = u4.z; u[3] = u4.w;
        d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
    } else {
        c[0] = __ldg(x + o);
        u[0] = row > 0 ? __ldg(x + o - W) : c[0];
        d[0] = row < H - 1 ? __ldg(x + o + W) : c[0];
    }
    const float l = col > 0 ? __ldg(x + o - 1) : c[0];
    const float r = col + VEC < W ? __ldg(x + o + VEC) : c[VEC - 1];
    float sum = 0.f, gr[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const float left = k == 0 ? l : c[k - 1], right = k == VEC - 1 ? r : c[k + 1];
        const float dr = right - c[k], dd = d[k] - c[k];            // zero on the last column / row (neighbour := centre)
        sum += fabsf(dr) + fabsf(dd);
        gr[k] = w * ((sgn(c[k] - left) - sgn(dr)) + (sgn(c[k] - u[k]) - sgn(dd)));
    }
    if (g) {
        if constexpr (VEC == 4) __stcs(reinterpret_cast<float4 *>(g + o), make_float4(gr[0], gr[1], gr[2], gr[3]));
        else __stcs(g + o, gr[0]);
    }
    return sum;
}

template <int VEC> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldcs(p);
    }
}
template <int VEC> __device__ __forceinline__ void st_vec(float *p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    else __stcs(p, v[0]);
}

// scratch: [0] ticket, then kRegSums floats per block.  The last block adds them in block order in fp64 (bit-reproducible).
template <int VEC>
__global__ void __launch_bounds__(kThreads) mag_regs_kernel(RegParams p, void *scratch, float *sums_out) {
    const int nv = p.H * p.W, b = blockIdx.y;
    const int i = (blockIdx.x * kThreads + threadIdx.x) * VEC;
    float acc[kRegSums] = {0.f, 0.f, 0.f, 0.f};
    if (i < nv) {
        const int row = i / p.W, col = i - row * p.W;
        if (p.demod) {
#pragma unroll 1
            for (int e = 0; e < p.ne; ++e) {
                const size_t plane = (static_cast<size_t>(b) * p.ne + e) * nv;
                acc[0] += tv_plane<VEC>(p.demod + plane, p.g_demod ? p.g_demod + plane : nullptr, p.w_ad_tv, row, col, p.H, p.W);
            }
        }
        if (p.r2) {
            const size_t plane = static_cast<size_t>(b) * nv;
            acc[3] = tv_plane<VEC>(p.r2 + plane, p.g_r2 ? p.g_r2 + plane : nullptr, p.w_r2_tv, row, col, p.H, p.W);
        }
        if (p.ls) {
            const size_t base = static_cast<size_t>(b) * 3 * nv + i;
            float a[VEC], bb[VEC], c[VEC], ga[VEC], gb[VEC], gc[VEC];
            ld_vec<VEC>(p.ls + base, a);
            ld_vec<VEC>(p.ls + base + nv, bb);
            ld_vec<VEC>(p.ls + base + 2 * static_cast<size_t>(nv), c);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                // LS_NZ: `A2B_ls[..., ::2]` strides the LAST axis, which has one element, so all three rows take part (:310)
                const float na = fminf(a[k], 0.f), nb_ = fminf(bb[k], 0.f), nc = fminf(c[k], 0.f);
                acc[1] += na * na + nb_ * nb_ + nc * nc;
                // LS_cond: positive discriminant b^2 - 4ac of the fitted quadratic (:313-314)
                const float q = bb[k] * bb[k] - 4.0f * (a[k] * c[k]);
                const float qp = q > 0.f ? q : 0.f;
                acc[2] += qp * qp;
                const float dq = 2.0f * p.w_ls_cond * qp;
                ga[k] = 2.0f * p.w_ls_nz * na - 4.0f * dq * c[k];
                gb[k] = 2.0f * p.w_ls_nz * nb_ + 2.0f * dq * bb[k];
                gc[k] = 2.0f * p.w_ls_nz * nc - 4.0f * dq * a[k];
            }
            if (p.g_ls) {
                st_vec<VEC>(p.g_ls + base, ga);
                st_vec<VEC>(p.g_ls + base + nv, gb);
                st_vec<VEC>(p.g_ls + base + 2 * static_cast<size_t>(nv), gc);
            }
        }
    }
    // block reduction of the four sums
    __shared__ float warp_part[kThreads / 32][kRegSums];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < kRegSums; ++s) {
        float v = acc[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) warp_part[warp][s] = v;
    }
    __syncthreads();
    unsigned *ticket = reinterpret_cast<unsigned *>(scratch);
    float *partials = reinterpret_cast<float *>(reinterpret_cast<char *>(scratch) + kScratchHeader);
    const unsigned nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < kRegSums) {
        float s = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) s += warp_part[w][threadIdx.x];
        partials[static_cast<size_t>(bid) * kRegSums + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double dpart[kThreads / 32][kRegSums];
    double dacc[kRegSums] = {0.0, 0.0, 0.0, 0.0};
    for (unsigned k = threadIdx.x; k < nblocks; k += kThreads) {
        const float4 t = __ldcg(reinterpret_cast<const float4 *>(partials) + k);
        dacc[0] += t.x; dacc[1] += t.y; dacc[2] += t.z; dacc[3] += t.w;
    }
#pragma unroll
    for (int s = 0; s < kRegSums; ++s) {
        double v = dacc[s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) dpart[warp][s] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[kRegSums] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < kThreads / 32; ++w)
            for (int k = 0; k < kRegSums; ++k) s[k] += dpart[w][k];
        sums_out[0] = static_cast<float>(s[0]);      // Ad_TV
        sums_out[1] = static_cast<float>(s[1]);      // LS_NZ
        sums_out[2] = 0.f;                           // WF_NZ: `x[..., :1] < x[..., -1:]` compares an element with itself (:311)
        sums_out[3] = static_cast<float>(s[2]);      // LS_cond
        sums_out[4] = static_cast<float>(s[3]);      // R2_TV
        ticket[0] = 0u;
    }
}

// ROI-analysis.py:301-322.  maps (nb,3,nv,2) = W, F, (phi, R2*); var (nb,5,nv,2) = rows |C_WW|, |C_WF|, |C_FW|, |C_FF|
// (second channel zero-padded by the caller, :244) and (var phi, var R2*).  out (nb, nv, nch): |W|, |F|, |W+F|, R2*
// (map units), and for mode >= 1 the propagated PDFF variance (mode 2 = the magnitude model's: the W-F row itself).
// IEEE divisions on purpose: background voxels give the same 0/0 -> NaN as the reference.
__global__ void __launch_bounds__(kThreads) roi_maps_kernel(const float *__restrict__ maps, const float *__restrict__ var, int nb, int nv, int mode,
                                                            float *__restrict__ out) {
    const int v = blockIdx.x * kThreads + threadIdx.x, b = blockIdx.y;
    const int nch = mode ? 5 : 4;
    __shared__ float tile[kThreads * 5];
    const int n_here = min(kThreads, nv - static_cast<int>(blockIdx.x) * kThreads);
    if (v < nv) {
        const float2 *m2 = reinterpret_cast<const float2 *>(maps) + static_cast<size_t>(b) * 3 * nv + v;
        const float2 w = __ldcs(m2), f = __ldcs(m2 + nv), pm = __ldcs(m2 + 2 * static_cast<size_t>(nv));
        const float wa = sqrtf(w.x * w.x + w.y * w.y), fa = sqrtf(f.x * f.x + f.y * f.y);
        const float sx = w.x + f.x, sy = w.y + f.y;
        const float wf = sqrtf(sx * sx + sy * sy);
        float *t = tile + threadIdx.x * nch;
        t[0] = wa; t[1] = fa; t[2] = wf; t[3] = pm.y;
        if (mode) {
            const float2 *v2 = reinterpret_cast<const float2 *>(var) + static_cast<size_t>(b) * 5 * nv + v;
            const float2 cww = __ldcs(v2), cwf = __ldcs(v2 + nv), cff = __ldcs(v2 + 3 * static_cast<size_t>(nv));
            const float w_var = hypotf(cww.x, cww.y), wf_var = hypotf(cwf.x, cwf.y), f_var = hypotf(cff.x, cff.y);
            float pv;
            if (mode == 2) {
                pv = wf_var;
            } else {
                const float wa2 = wa * wa;
                pv = w_var / wa2;
                pv = pv - 2.0f * wf_var / (wa * wf);
                pv = pv + (w_var + f_var + 2.0f * wf_var) / wa;
                pv = pv * (wa2 / (wf * wf));
            }
            t[4] = pv;
        }
    }
    __syncthreads();
    float *dst = out + (static_cast<size_t>(b) * nv + static_cast<size_t>(blockIdx.x) * kThreads) * nch;
    for (int k = threadIdx.x; k < n_here * nch; k += kThreads) __stcs(dst + k, tile[k]);
}

}  // namespace ig

using namespace ig;

extern "C" size_t ig_mag_regs_scratch_bytes(int nb, int H, int W) {
    if (nb <= 0 || H <= 0 || W <= 0) return 0;
    const size_t blocks = (static_cast<size_t>(H) * W + kThreads - 1) / kThreads;      // the one-pixel-per-thread shape is the widest
    return kScratchHeader + sizeof(float) * kRegSums * blocks * static_cast<size_t>(nb);
}

extern "C" int ig_mag_regs(const float *ls_d, const float *demod_d, const float *r2_d, int nb, int ne, int H, int W, float w_ad_tv, float w_ls_nz,
                           float w_ls_cond, float w_r2_tv, float *sums_d, float *g_ls_d, float *g_demod_d, float *g_r2_d, void *scratch_d,
                           size_t scratch_bytes, void *stream) {
    IG_REQUIRE(sums_d && scratch_d && (ls_d || demod_d || r2_d), IG_E_ARG, "ig_mag_regs: null pointer");
    IG_REQUIRE(nb > 0 && nb <= 65535 && H > 0 && W > 0 && static_cast<long>(H) * W < (1L << 30), IG_E_ARG, "ig_mag_regs: nb=%d H=%d W=%d", nb, H, W);
    IG_REQUIRE(!demod_d || ne > 0, IG_E_NE, "ig_mag_regs: ne=%d", ne);
    IG_REQUIRE((!g_ls_d || ls_d) && (!g_demod_d || demod_d) && (!g_r2_d || r2_d), IG_E_ARG, "ig_mag_regs: gradient requested for an absent input");
    IG_REQUIRE(scratch_bytes >= ig_mag_regs_scratch_bytes(nb, H, W), IG_E_ARG, "ig_mag_regs: scratch too small (%zu < %zu)", scratch_bytes,
               ig_mag_regs_scratch_bytes(nb, H, W));
    RegParams p{ls_d, demod_d, r2_d, g_ls_d, g_demod_d, g_r2_d, nb, ne, H, W, w_ad_tv, w_ls_nz, w_ls_cond, w_r2_tv};
    const int nv = H * W;
    auto aligned = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = (W % 4 == 0) && aligned(ls_d) && aligned(demod_d) && aligned(r2_d) && aligned(g_ls_d) && aligned(g_demod_d) && aligned(g_r2_d);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vec) mag_regs_kernel<4><<<grid_for(nb, nv, 4), kThreads, 0, s>>>(p, scratch_d, sums_d);
    else mag_regs_kernel<1><<<grid_for(nb, nv, 1), kThreads, 0, s>>>(p, scratch_d, sums_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_roi_maps(const float *maps_d, const float *var_d, int nb, int nv, int mode, float *out_d, void *stream) {
    IG_REQUIRE(maps_d && out_d && nb > 0 && nb <= 65535 && nv > 0 && mode >= 0 && mode <= 2, IG_E_ARG, "ig_roi_maps: bad arguments");
    IG_REQUIRE(mode == 0 || var_d, IG_E_ARG, "ig_roi_maps: mode %d needs the variance maps", mode);
    roi_maps_kernel<<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(maps_d, var_d, nb, nv, mode, out_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}
