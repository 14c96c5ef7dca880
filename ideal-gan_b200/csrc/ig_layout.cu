// Layout adapters between the echo-planar "MEBCRN" tensors of the physics path and the channel-interleaved ("flat")
// tensors the 2-D networks consume.  Replaces data.A_from_MEBCRN / B_from_MEBCRN / B_to_MEBCRN of the reference
// (/root/reference/data.py:262-329; callers train-sup.py:244-245, ROI-analysis.py:187-209): there, transposes, zero
// stacks, reshapes and concats, each one more full pass over the echo tensor; here one read and one write.
//     acquisitions   (nb, ne, nv, 2)  <->  (nb, nv, 2 ne)          a batched (ne x nv) transpose of complex elements
//     maps           (nb, 3, nv, 2)   <->  (nb, nv, 6)             with the (phi, R2*) row swapped to (R2*, phi)
#include "ig_common.cuh"

namespace ig {

constexpr int kTileV = 256;      // voxels per block (one per thread)

// Tiled transpose through shared memory: both global sides move whole, coalesced 256-byte warp requests of float2.
// Shared layout [voxel][plane] with an odd plane pitch: conflict-free for the 8-byte accesses of both phases.
template <bool TO_FLAT> __global__ void __launch_bounds__(kTileV) acq_relayout_kernel(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                                                                      int ne, int nv) {
    extern __shared__ float2 tile[];
    const int pitch = ne | 1;
    const int b = blockIdx.y, v0 = blockIdx.x * kTileV, t = threadIdx.x;
    const int nvox = min(kTileV, nv - v0);
    const size_t planar = static_cast<size_t>(b) * ne * nv + v0;          // + e * nv + voxel
    const size_t flat = (static_cast<size_t>(b) * nv + v0) * ne;          // + voxel * ne + e
    if constexpr (TO_FLAT) {
        if (t < nvox)
            for (int e = 0; e < ne; ++e) tile[t * pitch + e] = __ldcs(src + planar + static_cast<size_t>(e) * nv + t);
        __syncthreads();
        for (int i = t; i < nvox * ne; i += kTileV) {
            const int v = i / ne, e = i - v * ne;
            __stcs(dst + flat + i, tile[v * pitch + e]);
        }
    } else {
        for (int i = t; i < nvox * ne; i += kTileV) {
            const int v = i / ne, e = i - v * ne;
            tile[v * pitch + e] = __ldcs(src + flat + i);
        }
        __syncthreads();
        if (t < nvox)
            for (int e = 0; e < ne; ++e) __stcs(dst + planar + static_cast<size_t>(e) * nv + t, tile[t * pitch + e]);
    }
}

// B_from_MEBCRN (data.py:283-299).  mode 0: (nb,3,nv,2) rows W, F, (phi, R2*) -> (nb,nv,6) = (W_re, W_im, F_re, F_im, R2*, phi).
// mode 1 (mag_and_phase): (nb,2,nv,ch) rows (|W|, |F|, R2*[, x]) and (pW, pF, phi[, x]) -> (nb,nv,4 + 2 (ch - 2)):
// both species take the phase c_pha * pi * B[:,1,...,1] (the reference uses the fat-phase channel for water too), then the
// trailing channels of row 0, then those of row 1.
__global__ void maps_to_flat_kernel(const float *__restrict__ maps, int nv, int mode, int ch, float c_pha, float *__restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (v >= nv) return;
    if (mode == 0) {
        const float2 *m = reinterpret_cast<const float2 *>(maps) + static_cast<size_t>(b) * 3 * nv + v;
        const float2 w = __ldcs(m), f = __ldcs(m + nv), pm = __ldcs(m + 2 * static_cast<size_t>(nv));
        float2 *o = reinterpret_cast<float2 *>(out) + (static_cast<size_t>(b) * nv + v) * 3;
        o[0] = w;
        o[1] = f;
        o[2] = make_float2(pm.y, pm.x);
    } else {
        const float *r0 = maps + (static_cast<size_t>(b) * 2 * nv + v) * ch;
        const float *r1 = r0 + static_cast<size_t>(nv) * ch;
        float s, c;
        sincosf(c_pha * r1[1] * 3.14159265358979323846f, &s, &c);
        const int tail = ch - 2, nout = 4 + 2 * tail;
        float *o = out + (static_cast<size_t>(b) * nv + v) * nout;
        o[0] = r0[0] * c;
        o[1] = r0[0] * s;
        o[2] = r0[1] * c;
        o[3] = r0[1] * s;
        for (int k = 0; k < tail; ++k) {
            o[4 + k] = r0[2 + k];
            o[4 + tail + k] = r1[2 + k];
        }
    }
}

// B_to_MEBCRN (data.py:302-329).  mode 0 'All': (nb,nv,6) -> (nb,3,nv,2); 1 'WF-PM': (nb,nv,4) = (W, F, R2*, phi) -> (nb,3,nv,2)
// with zero imaginary parts; 2 'WF': (nb,nv,2) -> (nb,2,nv,2); 3 'PM': (nb,nv,2) = (R2*, phi) -> (nb,1,nv,2) = (phi, R2*).
__global__ void maps_from_flat_kernel(const float *__restrict__ flat, int nv, int mode, float *__restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (v >= nv) return;
    const size_t vb = static_cast<size_t>(b) * nv + v;
    float2 *o = reinterpret_cast<float2 *>(out);
    if (mode == 0) {
        const float2 *i = reinterpret_cast<const float2 *>(flat) + vb * 3;
        const float2 pm = i[2];
        o[(static_cast<size_t>(b) * 3 + 0) * nv + v] = i[0];
        o[(static_cast<size_t>(b) * 3 + 1) * nv + v] = i[1];
        o[(static_cast<size_t>(b) * 3 + 2) * nv + v] = make_float2(pm.y, pm.x);
    } else if (mode == 1) {
        const float4 i = reinterpret_cast<const float4 *>(flat)[vb];
        o[(static_cast<size_t>(b) * 3 + 0) * nv + v] = make_float2(i.x, 0.f);
        o[(static_cast<size_t>(b) * 3 + 1) * nv + v] = make_float2(i.y, 0.f);
        o[(static_cast<size_t>(b) * 3 + 2) * nv + v] = make_float2(i.w, i.z);
    } else if (mode == 2) {
        const float2 i = reinterpret_cast<const float2 *>(flat)[vb];
        o[(static_cast<size_t>(b) * 2 + 0) * nv + v] = make_float2(i.x, 0.f);
        o[(static_cast<size_t>(b) * 2 + 1) * nv + v] = make_float2(i.y, 0.f);
    } else {
        const float2 i = reinterpret_cast<const float2 *>(flat)[vb];
        o[vb] = make_float2(i.y, i.x);
    }
}

static int relayout(bool to_flat, const float *src, float *dst, int nb, int ne, int nv, cudaStream_t st) {
    const dim3 grid(static_cast<unsigned>((nv + kTileV - 1) / kTileV), static_cast<unsigned>(nb));
    const size_t smem = static_cast<size_t>(kTileV) * (ne | 1) * sizeof(float2);
    if (to_flat)
        acq_relayout_kernel<true><<<grid, kTileV, smem, st>>>(reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), ne, nv);
    else
        acq_relayout_kernel<false><<<grid, kTileV, smem, st>>>(reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), ne, nv);
    IG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ig

using namespace ig;

extern "C" int ig_acq_to_flat(const float *acqs_d, int nb, int ne, int nv, float *flat_d, void *stream) {
    IG_REQUIRE(acqs_d && flat_d && nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "ig_acq_to_flat: bad arguments");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_acq_to_flat: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    return relayout(true, acqs_d, flat_d, nb, ne, nv, static_cast<cudaStream_t>(stream));
}

extern "C" int ig_acq_from_flat(const float *flat_d, int nb, int ne, int nv, float *acqs_d, void *stream) {
    IG_REQUIRE(acqs_d && flat_d && nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "ig_acq_from_flat: bad arguments");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_acq_from_flat: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    return relayout(false, flat_d, acqs_d, nb, ne, nv, static_cast<cudaStream_t>(stream));
}

extern "C" int ig_maps_to_flat(const float *maps_d, int nb, int nv, int mode, int ch, float c_pha, float *flat_d, void *stream) {
    IG_REQUIRE(maps_d && flat_d && nb > 0 && nv > 0 && nb <= 65535, IG_E_ARG, "ig_maps_to_flat: bad arguments");
    IG_REQUIRE(mode == 0 || (mode == 1 && ch >= 3 && ch <= 8), IG_E_ARG, "ig_maps_to_flat: mode %d with %d channels", mode, ch);
    maps_to_flat_kernel<<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(maps_d, nv, mode, ch, c_pha, flat_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_maps_from_flat(const float *flat_d, int nb, int nv, int mode, float *maps_d, void *stream) {
    IG_REQUIRE(maps_d && flat_d && nb > 0 && nv > 0 && nb <= 65535 && mode >= 0 && mode <= 3, IG_E_ARG, "ig_maps_from_flat: bad arguments");
    IG_REQUIRE(mode != 1 || aligned16(flat_d), IG_E_ALIGN, "ig_maps_from_flat: 'WF-PM' input must be 16-byte aligned");
    maps_from_flat_kernel<<<grid_for(nb, nv, 1), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(flat_d, nv, mode, maps_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}
