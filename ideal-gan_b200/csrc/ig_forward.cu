// Forward signal models, their adjoint, and the fused forward -> mask -> MSE -> backward objective.
//
// Replaces IDEAL_model / IDEAL_mag / IDEAL_mag_phase of the reference
// (/root/reference/wflib/IDEAL_model.py:220-299, 404-453, 456-509) and the TF autodiff through them
// (train-IDEAL-single.py:154-157,175; train-IDEAL-GAN.py:243,288).  The reference materialises
// ~6 complex (nb, ne, nv) intermediates per call; here each voxel is read once, kept in registers for
// all echoes, and written once:
//     S_e = exp(-te_e R) exp(i (2 pi te_e phi + s_e beta)) (rho_W + c_e rho_F),   s_e = (-1)^e
// Algorithmic HBM bytes per voxel (ne = 6): forward 24|32 read + 48 written; backward 48 + 24|32 read
// + 24|32 written; fused loss 48 + 24|32 read + 24|32 written.
#include <stdlib.h>

#include "ig_model.cuh"

namespace ig {

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_LOSS = 2, MODE_DEC = 3 };

struct FwdParams {
    const float *maps;
    const float *tab;
    const float *gout;   // MODE_BWD: upstream (nb, ne, nv, 2)
    const float *acqs;   // MODE_LOSS: measured echoes (nb, ne, nv, 2)
    float *out;          // MODE_FWD: S_hat ; MODE_LOSS: optional S_hat
    float *gmaps;
    float *loss;
    void *scratch;
    float *mag, *pdff, *r2s;   // MODE_DEC: (nb, ne, nv) |S_e|, (nb, nv) PDFF, (nb, nv) R2* map -- each optional, clipped to [0, 1]
    int rows_or_ch, nb, ne, nv, flags;
    float r2_sc, inv_n;
};

// clip_by_value(x, 0, 1) with TensorFlow's NaN propagation (tf.maximum / tf.minimum return NaN for a NaN operand; fmaxf would not)
__device__ __forceinline__ float clip01(float x) {
    float r;
    asm("{ .reg .f32 t; max.NaN.f32 t, %1, 0f00000000; min.NaN.f32 %0, t, 0f3F800000; }" : "=f"(r) : "f"(x));      // FMNMX.NAN x 2
    return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));       // MUFU.SQRT, 1 ulp: the IEEE sqrtf is six instructions and a slow path
    return r;
}
__device__ __forceinline__ pk sqrt_fast(pk x) { return mk(sqrt_fast(x.d.x), sqrt_fast(x.d.y)); }
__device__ __forceinline__ pk clip01(pk x) { return mk(clip01(x.d.x), clip01(x.d.y)); }
__device__ __forceinline__ float vdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ pk vdiv(pk a, pk b) { return mk(__fdiv_rn(a.d.x, b.d.x), __fdiv_rn(a.d.y, b.d.y)); }

// Dataset synthesis consumer (gen_LDM_dataset.py:156-158,217-235): the forward model and, in the same pass over the registers,
// the three images the script writes per slice -- PDFF = |F| / (|W| + |F|), the R2* map and one magnitude image per echo, each
// clipped to [0, 1].  |S_e| = e^{-te R} |rho_W + c_e rho_F|: when the complex signals are not wanted the field-map phasor (two
// MUFU + the phase arithmetic per echo) is never formed.
template <int NE, typename V, int MODEL>
__device__ __forceinline__ void decode_outputs(const FwdParams &p, const SampleTab<NE> &T, const Voxel<V> &x, int b, int v0) {
    const int nv = p.nv, ne = p.ne;
    const bool clip = !(p.flags & IG_F_NO_CLIP);
    if (p.pdff) {
        const V mw = vsqrt(vfma(x.rhoW.re, x.rhoW.re, vmul(x.rhoW.im, x.rhoW.im)));
        const V mf = vsqrt(vfma(x.rhoF.re, x.rhoF.re, vmul(x.rhoF.im, x.rhoF.im)));
        // the script divides the decoder's magnitude channels as they are (Z2B[i,0,:,:,1] / (Z2B[i,0,:,:,0] + Z2B[i,0,:,:,1])); the
        // complex-row models have no such channels and use |F| / (|W| + |F|).  0 / 0 = NaN on background, as in the script.
        const V q = (MODEL == IG_MODEL_MAGPHA) ? vdiv(x.pd, vadd(x.ff, x.pd)) : vdiv(mf, vadd(mw, mf));
        st_real(p.pdff + static_cast<size_t>(b) * nv, v0, clip ? clip01(q) : q);
    }
    if (p.r2s) st_real(p.r2s + static_cast<size_t>(b) * nv, v0, clip ? clip01(x.r2raw) : x.r2raw);
    if (!p.out && !p.mag) return;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            const V d = fast_ex2(vmul(T.r[e].kdec, x.r2));
            const cx<V> yhat = caffine(x.rhoW, T.r[e].c_re, T.r[e].c_im, x.rhoF);
            if (p.out) {
                V c, s;
                unit_phasor(vfma(T.r[e].sgn, x.bturn, vmul(T.r[e].kphi, x.phi_t)), c, s);
                st_cx(p.out + static_cast<size_t>(b) * ne * nv * 2 + static_cast<size_t>(e) * nv * 2, v0, cmulv(cx<V>{vmul(d, c), vmul(d, s)}, yhat));
            }
            if (p.mag) {
                const V m = vmul(d, sqrt_fast(vfma(yhat.re, yhat.re, vmul(yhat.im, yhat.im))));
                st_real(p.mag + (static_cast<size_t>(b) * ne + e) * nv, v0, clip ? clip01(m) : m);
            }
        }
    }
}

// one thread's voxels of sample b starting at v0: decode, all echoes, write-out; returns the thread's loss partial
template <int NE, typename V, int MODEL, int MODE>
__device__ __forceinline__ float ideal_voxels(const FwdParams &p, const SampleTab<NE> &T, int b, int v0, float2 *stile = nullptr, int spitch = 0) {
    const int nv = p.nv, ne = p.ne;
    const size_t map_elems = (MODEL == IG_MODEL_MAGPHA) ? static_cast<size_t>(2) * nv * p.rows_or_ch
                                                        : static_cast<size_t>(p.rows_or_ch) * nv * 2;
    const Voxel<V> x = decode<V, MODEL>(p.maps + b * map_elems, p.rows_or_ch, nv, v0, p.flags);
    const size_t acq_b = static_cast<size_t>(b) * ne * nv * 2;
    if constexpr (MODE == MODE_DEC) {
        decode_outputs<NE, V, MODEL>(p, T, x, b, v0);
        return 0.f;
    }
    Adj<V> a;
    a.sg = czero<V>(); a.sgc = czero<V>(); a.tq = czero<V>(); a.q = czero<V>(); a.bq = splat<V>(0.f);
    V lsum = splat<V>(0.f);
    // issue every upstream / measurement load before the math so each thread has ne loads in flight
    cx<V> in[NE];
    if constexpr (MODE != MODE_FWD) {
        const float *src = (MODE == MODE_BWD ? p.gout : p.acqs) + acq_b;
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (e < ne) in[e] = ld_cx(src + static_cast<size_t>(e) * nv * 2, v0, V{});
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (e < ne) {
            V c, s;
            unit_phasor(vfma(T.r[e].sgn, x.bturn, vmul(T.r[e].kphi, x.phi_t)), c, s);
            const V d = fast_ex2(vmul(T.r[e].kdec, x.r2));
            const cx<V> w{vmul(d, c), vmul(d, s)};
            const cx<V> yhat = caffine(x.rhoW, T.r[e].c_re, T.r[e].c_im, x.rhoF);
            const cx<V> shat = cmulv(w, yhat);
            if constexpr (MODE == MODE_FWD) {
                if (stile) {
                    // channel-interleaved output: staged as [voxel][echo] in shared memory, written out linearly by the block
#pragma unroll
                    for (int l = 0; l < lanes<V>::n; ++l)
                        stile[(threadIdx.x * lanes<V>::n + l) * spitch + e] = make_float2(lane_get(shat.re, l), lane_get(shat.im, l));
                } else {
                    st_cx(p.out + acq_b + static_cast<size_t>(e) * nv * 2, v0, shat);
                }
            } else {
                cx<V> G;
                if constexpr (MODE == MODE_BWD) {
                    G = in[e];
                } else {
                    G = cx<V>{mask_sub(shat.re, in[e].re), mask_sub(shat.im, in[e].im)};
                    lsum = vfma(G.re, G.re, lsum);
                    lsum = vfma(G.im, G.im, lsum);
                    if (p.out) st_cx(p.out + acq_b + static_cast<size_t>(e) * nv * 2, v0, shat);
                }
                const cx<V> g = cmulc(w, G);
                a.sg.re = vadd(a.sg.re, g.re);
                a.sg.im = vadd(a.sg.im, g.im);
                cmac(a.sgc, T.r[e].c_re, -T.r[e].c_im, g);
                const cx<V> q = cmulc(g, yhat);
                a.tq.re = vfma(T.r[e].te, q.re, a.tq.re);
                a.tq.im = vfma(T.r[e].te, q.im, a.tq.im);
                if constexpr (MODEL == IG_MODEL_FFPD) { a.q.re = vadd(a.q.re, q.re); a.q.im = vadd(a.q.im, q.im); }
                if constexpr (MODEL != IG_MODEL_FFPD) a.bq = vfma(T.r[e].sgn, q.im, a.bq);
            }
        }
    }
    if constexpr (MODE != MODE_FWD) {
        const float scale = (MODE == MODE_LOSS) ? 2.0f * p.inv_n : 1.0f;
        write_grads<V, MODEL>(p.gmaps + b * map_elems, p.rows_or_ch, nv, v0, p.flags, x, a, p.r2_sc, scale);
    }
    return hsum(lsum);
}

template <int NE, typename V, int MODEL, int MODE>
__global__ void __launch_bounds__(kThreads) ideal_kernel(const FwdParams p) {
    __shared__ SampleTab<NE> T;
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    if (v0 < p.nv) ideal_voxels<NE, V, MODEL, MODE>(p, T, b, v0);
}

// Forward model with the channel-interleaved output (nb, nv, 2 ne) of data.A_from_MEBCRN (train-sup.py:242-244 runs IDEAL_op and
// then that adapter): the block's tile is transposed through shared memory ([voxel][echo], odd pitch) and leaves as one
// contiguous, fully coalesced run of float2 -- a strided store straight from the registers measured 3 x slower than the
// planar kernel, slower even than planar kernel + adapter.
template <int NE, typename V, int MODEL>
__global__ void __launch_bounds__(kThreads) ideal_fwd_flat_kernel(const FwdParams p) {
    __shared__ SampleTab<NE> T;
    extern __shared__ float2 flat_tile[];
    const int b = blockIdx.y, ne = p.ne, pitch = ne | 1;
    constexpr int kTile = kThreads * lanes<V>::n;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, ne, p.r2_sc);
    const int t0 = blockIdx.x * kTile;
    const int v0 = t0 + threadIdx.x * lanes<V>::n;
    if (v0 < p.nv) ideal_voxels<NE, V, MODEL, MODE_FWD>(p, T, b, v0, flat_tile, pitch);
    __syncthreads();
    const int nvox = min(kTile, p.nv - t0);
    float2 *dst = reinterpret_cast<float2 *>(p.out) + (static_cast<size_t>(b) * p.nv + t0) * ne;
    for (int i = threadIdx.x; i < nvox * ne; i += kThreads) {
        const int v = i / ne, e = i - v * ne;
        __stcs(dst + i, flat_tile[v * pitch + e]);
    }
}

// Fused objective: a persistent grid (one resident wave), each block walking a contiguous range of (sample, tile)
// work items and taking part in the loss reduction ONCE at the end -- with one block per tile the 18 k ticket atomics
// of a 64-slice batch all hit one address and every block sits out its own L2 round trip before it can retire.
template <int NE, typename V, int MODEL>
__global__ void __launch_bounds__(kThreads, MODEL == IG_MODEL_MAGPHA ? 2 : 3) ideal_loss_kernel(const FwdParams p) {   // measured: 80 registers pay off except for mag/phase (spills)
    __shared__ SampleTab<NE> T;
    const int tiles_ps = (p.nv + kThreads * lanes<V>::n - 1) / (kThreads * lanes<V>::n);
    const long total = static_cast<long>(p.nb) * tiles_ps;
    const int tile_end = static_cast<int>(total * (blockIdx.x + 1) / gridDim.x);
    int cur_b = -1;
    float loss_part = 0.f;
    for (int tile = static_cast<int>(total * blockIdx.x / gridDim.x); tile < tile_end; ++tile) {
        const int b = tile / tiles_ps;
        if (b != cur_b) {
            if (cur_b >= 0) __syncthreads();          // everyone is done reading the previous sample's table
            stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
            cur_b = b;
        }
        const int v0 = ((tile - b * tiles_ps) * kThreads + threadIdx.x) * lanes<V>::n;
        if (v0 < p.nv) loss_part += ideal_voxels<NE, V, MODEL, MODE_LOSS>(p, T, b, v0);
    }
    block_loss_reduce(loss_part, p.scratch, p.loss, p.inv_n);
}

// Alternative split of the same objective: one block per tile (the shape that streams at the roofline for the adjoint), each
// block leaving its partial with a plain store, and a one-block kernel adding the partials in index order in fp64.  No atomics,
// no fences, bit-reproducible.
template <int NE, typename V, int MODEL>
__global__ void __launch_bounds__(kThreads) ideal_loss_grid_kernel(const FwdParams p) {
    __shared__ SampleTab<NE> T;
    __shared__ float warp_part[kThreads / 32];
    const int b = blockIdx.y;
    stage_table(T, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, p.ne, p.r2_sc);
    const int v0 = (blockIdx.x * blockDim.x + threadIdx.x) * lanes<V>::n;
    float v = v0 < p.nv ? ideal_voxels<NE, V, MODEL, MODE_LOSS>(p, T, b, v0) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += warp_part[w];
        reinterpret_cast<float *>(reinterpret_cast<char *>(p.scratch) + kScratchHeader)[blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024) loss_finish_kernel(const void *scratch, int n, float scale, float *loss_out) {
    __shared__ double part[32];
    grid_dependency_wait();
    const float *partials = reinterpret_cast<const float *>(reinterpret_cast<const char *>(scratch) + kScratchHeader);
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(partials[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += part[w];
        loss_out[0] = static_cast<float>(s * static_cast<double>(scale));
    }
}

template <typename K> static int resident_grid(K kernel, int nb, int nv, int vpt, int *grid) {
    int dev = 0, sms = 0, occ = 0;
    IG_CUDA(cudaGetDevice(&dev));
    IG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    IG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0));
    const long tiles = static_cast<long>(nb) * ((nv + kThreads * vpt - 1) / (kThreads * vpt));
    const long g = static_cast<long>(sms) * (occ > 0 ? occ : 1);
    *grid = static_cast<int>(g < tiles ? g : tiles);
    return 0;
}

template <int MODEL, int MODE> static int launch_ideal(const FwdParams &p, cudaStream_t st) {
    bool packed = (p.nv % 2 == 0) && aligned16(p.maps);
    if (MODE == MODE_DEC) packed = packed && (!p.out || aligned16(p.out)) && (!p.mag || aligned16(p.mag)) && (!p.pdff || aligned16(p.pdff)) && (!p.r2s || aligned16(p.r2s));
    else packed = packed && (MODE == MODE_FWD ? aligned16(p.out) : aligned16(p.gmaps));
    if (MODE == MODE_BWD) packed = packed && aligned16(p.gout);
    if (MODE == MODE_LOSS) packed = packed && aligned16(p.acqs) && (!p.out || aligned16(p.out));
    if (MODEL == IG_MODEL_MAGPHA && p.rows_or_ch == 4) {     // 4-channel rows are read/written as float4 in both paths
        IG_REQUIRE(aligned16(p.maps) && (MODE == MODE_FWD || aligned16(p.gmaps)), IG_E_ALIGN, "mag/phase maps must be 16-byte aligned");
    }
    return dispatch_ne(p.ne, [&](auto ne_c) {
        constexpr int NE = decltype(ne_c)::value;
        if constexpr (MODE == MODE_BWD && MODEL != IG_MODEL_MAGPHA) {
            if (packed && p.ne > 8) {      // 9..12 echoes, 128-voxel rows: the adjoint on the TMA ring (ig_ring_ops.cu, RowLossOp<.., BWD>)
                const int rc = row_bwd_ring(MODEL, p.maps, p.rows_or_ch, p.gout, p.tab, p.nb, p.ne, p.nv, p.r2_sc, p.flags, p.gmaps, st);
                if (rc != IG_E_UNSUPPORTED) return rc;
            }
        }
        if constexpr (MODE == MODE_LOSS) {
            // measured at 64 x 384 x 384 x 6: the grid + finish pair wins for the complex-row models (0.159 vs 0.165 ms), the
            // persistent kernel for mag/phase, whose 128 registers leave too few warps for the one-tile-per-block shape (0.213 vs 0.251 ms)
            if (MODEL != IG_MODEL_MAGPHA) {
                if (packed) {      // 128-voxel rows, <= 12 echoes, <= 4 map rows: the TMA ring (ig_ring_ops.cu, RowLossOp)
                    const int rc = row_loss_ring(MODEL, p.maps, p.rows_or_ch, p.acqs, p.tab, p.nb, p.ne, p.nv, p.r2_sc, p.flags, p.inv_n, p.gmaps, p.out, p.loss,
                                                 p.scratch, st);
                    if (rc != IG_E_UNSUPPORTED) return rc;
                }
                const dim3 g = grid_for(p.nb, p.nv, packed ? 2 : 1);
                if (packed) ideal_loss_grid_kernel<NE, pk, MODEL><<<g, kThreads, 0, st>>>(p);
                else ideal_loss_grid_kernel<NE, float, MODEL><<<g, kThreads, 0, st>>>(p);
                IG_CUDA(cudaGetLastError());
                loss_finish_kernel<<<1, 1024, 0, st>>>(p.scratch, static_cast<int>(g.x * g.y), p.inv_n, p.loss);
                IG_CUDA(cudaGetLastError());
                return 0;
            }
            if (packed && p.rows_or_ch == 4) {     // bipolar rows on the TMA ring (128-voxel rows, <= 8 echoes): ig_ring_ops.cu
                const int rc = magpha_loss_ring(p.maps, p.acqs, p.tab, p.nb, p.ne, p.nv, p.r2_sc, p.inv_n, p.gmaps, p.out, p.loss, p.scratch, st);
                if (rc != IG_E_UNSUPPORTED) return rc;
            }
            int grid = 1;
            if (packed) {
                if (int rc = resident_grid(ideal_loss_kernel<NE, pk, MODEL>, p.nb, p.nv, 2, &grid)) return rc;
                ideal_loss_kernel<NE, pk, MODEL><<<grid, kThreads, 0, st>>>(p);
            } else {
                if (int rc = resident_grid(ideal_loss_kernel<NE, float, MODEL>, p.nb, p.nv, 1, &grid)) return rc;
                ideal_loss_kernel<NE, float, MODEL><<<grid, kThreads, 0, st>>>(p);
            }
        } else if (MODE == MODE_FWD && (p.flags & IG_F_FLAT)) {
            if constexpr (MODE == MODE_FWD) {
                const size_t smem = static_cast<size_t>(kThreads) * (packed ? 2 : 1) * (p.ne | 1) * sizeof(float2);
                if (packed) {
                    IG_CUDA(cudaFuncSetAttribute(ideal_fwd_flat_kernel<NE, pk, MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                    ideal_fwd_flat_kernel<NE, pk, MODEL><<<grid_for(p.nb, p.nv, 2), kThreads, smem, st>>>(p);
                } else {
                    ideal_fwd_flat_kernel<NE, float, MODEL><<<grid_for(p.nb, p.nv, 1), kThreads, smem, st>>>(p);
                }
            }
        } else if (packed) {
            ideal_kernel<NE, pk, MODEL, MODE><<<grid_for(p.nb, p.nv, 2), kThreads, 0, st>>>(p);
        } else {
            ideal_kernel<NE, float, MODEL, MODE><<<grid_for(p.nb, p.nv, 1), kThreads, 0, st>>>(p);
        }
        IG_CUDA(cudaGetLastError());
        return 0;
    });
}

template <int MODE> static int launch_model(int model, const FwdParams &p, cudaStream_t st) {
    switch (model) {
        case IG_MODEL_WFPM: return launch_ideal<IG_MODEL_WFPM, MODE>(p, st);
        case IG_MODEL_FFPD: return launch_ideal<IG_MODEL_FFPD, MODE>(p, st);
        case IG_MODEL_MAGPHA: return launch_ideal<IG_MODEL_MAGPHA, MODE>(p, st);
    }
    set_error("unknown model %d", model);
    return IG_E_ARG;
}

static int check_model_args(const char *fn, int model, int rows_or_ch, int nb, int ne, int nv) {
    IG_REQUIRE(nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "%s: nb=%d (1..65535), nv=%d", fn, nb, nv);
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "%s: ne=%d outside [1, %d]", fn, ne, IG_MAX_NE);
    const bool ok = (model == IG_MODEL_WFPM && rows_or_ch >= 3) || (model == IG_MODEL_FFPD && rows_or_ch == 3) ||
                    (model == IG_MODEL_MAGPHA && (rows_or_ch == 3 || rows_or_ch == 4));
    IG_REQUIRE(ok, IG_E_ARG, "%s: model %d does not take rows/channels = %d", fn, model, rows_or_ch);
    return 0;
}

}  // namespace ig

using namespace ig;

extern "C" int ig_ideal_fwd(int model, const float *maps_d, int rows_or_ch, const float *tab_d, int nb, int ne, int nv, float r2_sc,
                            int flags, float *out_d, void *stream) {
    IG_REQUIRE(maps_d && tab_d && out_d, IG_E_ARG, "ig_ideal_fwd: null pointer");
    if (int rc = check_model_args("ig_ideal_fwd", model, rows_or_ch, nb, ne, nv)) return rc;
    FwdParams p{};
    p.maps = maps_d; p.tab = tab_d; p.out = out_d; p.rows_or_ch = rows_or_ch; p.nb = nb; p.ne = ne; p.nv = nv; p.flags = flags; p.r2_sc = r2_sc;
    if (flags & IG_F_ONLY_MAG) {
        // |S_hat| only, (nb, ne, nv): the decode kernel without its clip (half the output bytes; the field-map phasor is never formed)
        IG_REQUIRE(!(flags & IG_F_FLAT), IG_E_UNSUPPORTED, "ig_ideal_fwd: IG_F_ONLY_MAG comes with the planar layout");
        p.out = nullptr; p.mag = out_d; p.flags = flags | IG_F_NO_CLIP;
        return launch_model<MODE_DEC>(model, p, static_cast<cudaStream_t>(stream));
    }
    return launch_model<MODE_FWD>(model, p, static_cast<cudaStream_t>(stream));
}

extern "C" int ig_ideal_decode(int model, const float *maps_d, int rows_or_ch, const float *tab_d, int nb, int ne, int nv, float r2_sc,
                               int flags, float *shat_d, float *mag_d, float *pdff_d, float *r2s_d, void *stream) {
    IG_REQUIRE(maps_d && tab_d, IG_E_ARG, "ig_ideal_decode: null pointer");
    IG_REQUIRE(shat_d || mag_d || pdff_d || r2s_d, IG_E_ARG, "ig_ideal_decode: no output requested");
    if (int rc = check_model_args("ig_ideal_decode", model, rows_or_ch, nb, ne, nv)) return rc;
    IG_REQUIRE(!(flags & IG_F_FLAT), IG_E_UNSUPPORTED, "ig_ideal_decode: planar outputs only");
    FwdParams p{};
    p.maps = maps_d; p.tab = tab_d; p.out = shat_d; p.mag = mag_d; p.pdff = pdff_d; p.r2s = r2s_d; p.rows_or_ch = rows_or_ch; p.nb = nb;
    p.ne = ne; p.nv = nv; p.flags = flags; p.r2_sc = r2_sc;
    return launch_model<MODE_DEC>(model, p, static_cast<cudaStream_t>(stream));
}

extern "C" int ig_ideal_bwd(int model, const float *maps_d, int rows_or_ch, const float *tab_d, int nb, int ne, int nv, float r2_sc,
                            int flags, const float *gout_d, float *gmaps_d, void *stream) {
    IG_REQUIRE(maps_d && tab_d && gout_d && gmaps_d, IG_E_ARG, "ig_ideal_bwd: null pointer");
    if (int rc = check_model_args("ig_ideal_bwd", model, rows_or_ch, nb, ne, nv)) return rc;
    IG_REQUIRE(!(flags & IG_F_FLAT), IG_E_UNSUPPORTED, "ig_ideal_bwd: the interleaved layout is a forward-only output option");
    FwdParams p{};
    p.maps = maps_d; p.tab = tab_d; p.gout = gout_d; p.gmaps = gmaps_d; p.rows_or_ch = rows_or_ch; p.nb = nb; p.ne = ne; p.nv = nv;
    p.flags = flags; p.r2_sc = r2_sc;
    return launch_model<MODE_BWD>(model, p, static_cast<cudaStream_t>(stream));
}

extern "C" int ig_ideal_loss(int model, const float *maps_d, int rows_or_ch, const float *acqs_d, const float *tab_d, int nb, int ne,
                             int nv, float r2_sc, int flags, float inv_n, float *gmaps_d, float *shat_d, float *loss_d, void *scratch_d,
                             size_t scratch_bytes, void *stream) {
    IG_REQUIRE(maps_d && tab_d && acqs_d && gmaps_d && loss_d && scratch_d, IG_E_ARG, "ig_ideal_loss: null pointer");
    if (int rc = check_model_args("ig_ideal_loss", model, rows_or_ch, nb, ne, nv)) return rc;
    IG_REQUIRE(!(flags & IG_F_FLAT), IG_E_UNSUPPORTED, "ig_ideal_loss: the interleaved layout is a forward-only output option");
    IG_REQUIRE(scratch_bytes >= ig_loss_scratch_bytes(nb, nv), IG_E_SCRATCH, "ig_ideal_loss: scratch %zu < %zu bytes", scratch_bytes,
               ig_loss_scratch_bytes(nb, nv));
    FwdParams p{};
    p.maps = maps_d; p.tab = tab_d; p.acqs = acqs_d; p.gmaps = gmaps_d; p.out = shat_d; p.loss = loss_d; p.scratch = scratch_d;
    p.rows_or_ch = rows_or_ch; p.nb = nb; p.ne = ne; p.nv = nv; p.flags = flags; p.r2_sc = r2_sc; p.inv_n = inv_n;
    return launch_model<MODE_LOSS>(model, p, static_cast<cudaStream_t>(stream));
}
