// Generic TMA ring: the producer / consumer pipeline of the headline kernel (ig_solve.cu, a2a_loss_tma_kernel) as a reusable
// template, so that the other latency-bound operators (acq_to_acq adjoint, the bipolar mag/phase objective, the Rician
// objective) get the same treatment: loads cost the math warps no registers, no address arithmetic and no scoreboard stalls.
//
//   block  = 8 consumer warps (or Op::kConsumerWarps: 7 or 15) + 1 producer warp, persistent (one resident wave of one or two blocks per SM)
//   tile   = 512 voxels = 8 chunks of 64 voxels (a consumer warp's unit: two voxels per lane)
//   stage  = every input plane of one tile + the sample's echo records, delivered by TMA on a "full" mbarrier:
//            one 3-D tensor-map box {<= 256 floats, rows, planes} per input TENSOR (UTMALDG) + one bulk copy (UBLKCP)
//   ring   = Op::kStages stages; consumers draw chunks from a shared counter and release a stage through its "empty" mbarrier;
//            the producer ends the ring with one marker stage per eight consumer warps
//
// An Op supplies the input tensors (floats per voxel and planes per sample of each), the per-chunk math and its Params; outputs
// leave straight from registers with streaming stores.  Shapes need whole 128-voxel rows (nv % 128 == 0) and 16-byte aligned
// bases; everything else stays on the plain kernels.  Ops are instantiated per echo count (dispatch_exact_ne in ig_common.cuh), not per bucket.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "ig_common.cuh"

namespace ig {

constexpr int kRingTileVox = 512, kRingChunkVox = 64, kRingChunks = kRingTileVox / kRingChunkVox;
// Consumer warps per block: 8 unless the Op names its own count (`kConsumerWarps`).  8 + the producer = 9 warps per block, 18 per SM at two
// blocks: the fuller scheduler holds 5 warps, so a thread gets 96 registers.  7 + 1 = 8 warps per block is 4 per scheduler and 128 registers:
// the issue-bound objectives whose 96-register builds spill (Rician, UQ) run 4-6 % faster on seven warps, the HBM-bound operators and the C2
// objective (no spills at 95 registers) do not (profiles/history_r02.md section 12).
constexpr int kRingConsumerWarps = 8;
template <class Op, class = void> struct RingCw { static constexpr int value = kRingConsumerWarps; };
template <class Op> struct RingCw<Op, std::void_t<decltype(Op::kConsumerWarps)>> { static constexpr int value = Op::kConsumerWarps; };
template <class Op> constexpr int ring_threads() { return RingCw<Op>::value * 32 + 32; }
constexpr int kRingMaxMaps = 5;

struct RingMaps {
    CUtensorMap m[kRingMaxMaps];
};

// byte offsets of each tensor's box inside a stage (sized for the NE bucket), then the echo records
template <class Op> struct RingLayout {
    static constexpr int plane_bytes(int m) { return kRingTileVox * Op::fpv(m) * 4; }
    static constexpr int off(int m) {
        int o = 0;
        for (int k = 0; k < m; ++k) o += Op::planes_max(k) * plane_bytes(k);
        return o;
    }
    static constexpr int tab_off = off(Op::kMaps);
    static constexpr int tab_bytes = Op::kNE * IG_REC_FLOATS * 4;
    static constexpr int stage_bytes = tab_off + ((tab_bytes + 127) / 128) * 128;
    static constexpr int smem_bytes = Op::kStages * stage_bytes;
};

// rows of one tile in tensor m's map: the map's inner dimension is min(256, 128 fpv) floats
__host__ __device__ constexpr int ring_inner_floats(int fpv) { return 128 * fpv < 256 ? 128 * fpv : 256; }
__host__ __device__ constexpr int ring_tile_rows(int fpv) { return kRingTileVox * fpv / ring_inner_floats(fpv); }

template <class Op>
__global__ void __launch_bounds__(ring_threads<Op>(), Op::kMinBlocks) ring_kernel(const typename Op::Params p, const __grid_constant__ RingMaps maps) {
    extern __shared__ __align__(128) unsigned char stage_mem[];
    using Lay = RingLayout<Op>;
    constexpr int STAGES = Op::kStages, NE = Op::kNE;
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int2 stage_tile[STAGES];             // (sample, first voxel) of the tile in each stage; sample < 0 = end
    __shared__ int chunk_ctr;
    __shared__ typename Op::Shared shared;
    const int nv = p.nv, ne = Op::kExact ? NE : p.ne;
    const int tiles_ps = (nv + kRingTileVox - 1) / kRingTileVox;
    const int total = p.nb * tiles_ps;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kRingChunks);
        }
        chunk_ctr = 0;
        mbar_fence_init();
    }
    Op::prologue(shared);
    grid_dependency_wait();          // PDL: everything above overlaps the tail of the previous kernel in the stream (ig_gen_tables)
    __syncthreads();
    float loss_part = 0.f;
    constexpr int kCw = RingCw<Op>::value;
    if (threadIdx.x >= kCw * 32) {
        // ---------------- producer warp: lane 0 owns the work counter, the barriers and the copies ----------------
        const int lane = threadIdx.x & 31;
        unsigned *next_tile = reinterpret_cast<unsigned *>(p.scratch) + 1;
        // dynamic claims (float atomic: see ig_solve.cu) where tiles differ in cost, a static round-robin otherwise
        auto claim_async = [&]() -> float {
            float k;
            asm volatile("atom.global.add.f32 %0, [%1], 0f3F800000;" : "=f"(k) : "l"(next_tile) : "memory");
            return k;
        };
        float k_nxt = 0.f;
        if (Op::kDynamic && lane == 0) k_nxt = claim_async();
        uint32_t tx = Lay::tab_bytes;
#pragma unroll
        for (int m = 0; m < Op::kMaps; ++m) tx += static_cast<uint32_t>(Op::planes(m, ne, p)) * Lay::plane_bytes(m);
        bool released = false;
        int ends_left = (kCw + kRingChunks - 1) / kRingChunks;      // every consumer warp draws one chunk past the data: enough marker stages for all of them
        for (int it = 0;; ++it) {
            const int s = it % STAGES;
            int k;
            if constexpr (Op::kDynamic) k = static_cast<int>(__shfl_sync(0xffffffffu, k_nxt, 0));
            else k = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const bool end = k >= total;
            if (Op::kDynamic && lane == 0 && !end) k_nxt = claim_async();
            if (!released && 2 * k >= total) {      // PDL: past the middle of the work a dependent may be scheduled (see a2a_loss_tma_kernel)
                released = true;
                if (lane == 0) grid_launch_dependents();
            }
            const int b = end ? -1 : k / tiles_ps;
            const int j = k - b * tiles_ps;
            const int tile = end ? 0 : static_cast<int>((static_cast<unsigned>(j) * static_cast<unsigned>(p.tile_stride)) % static_cast<unsigned>(tiles_ps));
            if (lane == 0) {
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                stage_tile[s] = make_int2(b, tile * kRingTileVox);
                if (end) mbar_arrive(&full_bar[s]);          // end marker: completes the phase without data
                else mbar_expect_tx(&full_bar[s], tx);
            }
            __syncwarp();
            if (end) {
                if (--ends_left == 0) break;
                continue;
            }
            if (lane == 0) {
                unsigned char *stage = stage_mem + s * Lay::stage_bytes;
#pragma unroll
                for (int m = 0; m < Op::kMaps; ++m) {
                    const int planes = Op::planes(m, ne, p);
                    if (planes > 0)
                        tma_load_3d(stage + Lay::off(m), &maps.m[m], 0, tile * ring_tile_rows(Op::fpv(m)), b * planes, &full_bar[s]);
                }
                bulk_g2s(stage + Lay::tab_off, p.tab + static_cast<size_t>(b) * IG_TAB_FLOATS, Lay::tab_bytes, &full_bar[s]);
            }
        }
    } else {
        // ---------------- consumer warps: decoupled, each draws the next 64-voxel chunk of the ring ----------------
        const int lane = threadIdx.x & 31;
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(&chunk_ctr, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            const int it = g / kRingChunks, slot = (g % kRingChunks) * 32 + lane;
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            const int2 where = stage_tile[s];
            if (where.x < 0) break;
            const int b = where.x, v0 = where.y + slot * 2;
            unsigned char *stage = stage_mem + s * Lay::stage_bytes;
            const SampleTab<NE> &T = *reinterpret_cast<const SampleTab<NE> *>(stage + Lay::tab_off);       // kdec = -te log2(e), unscaled
            Op::chunk(p, shared, stage, T, slot, b, v0, v0 < nv, ne, loss_part);
            if constexpr (Op::kWritesStage) fence_proxy_async();      // generic-proxy writes to the stage before TMA refills it
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
    if constexpr (Op::kLoss) block_loss_reduce(loss_part, p.scratch, p.loss, p.inv_n);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using RingEncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline RingEncodeFn ring_encode_fn() {
    static const RingEncodeFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<RingEncodeFn>(f);
    }();
    return fn;
}

// `planes` planes of nv voxels x fpv floats, `plane_stride` floats apart; box = one tile of `box_planes` planes
inline bool ring_tensor_map(CUtensorMap *m, const float *base, int nv, int fpv, long planes, long plane_stride, int box_planes) {
    const RingEncodeFn enc = ring_encode_fn();
    const int inner = ring_inner_floats(fpv);
    if (!enc || !base || (static_cast<long>(nv) * fpv) % inner != 0 || (plane_stride * 4) % 16 != 0 || !aligned16(base) || box_planes < 1 || box_planes > 256)
        return false;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(static_cast<long>(nv) * fpv / inner), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(inner) * 4, static_cast<cuuint64_t>(plane_stride) * 4};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(inner), static_cast<cuuint32_t>(ring_tile_rows(fpv)), static_cast<cuuint32_t>(box_planes)};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int ring_coprime_stride(int n) {
    if (n <= 2 || n > 46340) return 1;
    int s = static_cast<int>(n * 0.6180339887) | 1;
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    while (gcd(s, n) != 1) s += 2;
    return s % n;
}

template <class Op> int ring_launch(typename Op::Params p, const RingMaps &maps, cudaStream_t st) {
    using Lay = RingLayout<Op>;
    const int tiles_ps = (p.nv + kRingTileVox - 1) / kRingTileVox;
    p.tile_stride = Op::kDynamic ? ring_coprime_stride(tiles_ps) : 1;
    int dev = 0, sms = 0, occ = 0;
    IG_CUDA(cudaGetDevice(&dev));
    IG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    IG_CUDA(cudaFuncSetAttribute(ring_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay::smem_bytes));
    IG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel<Op>, ring_threads<Op>(), Lay::smem_bytes));
    const long tiles = static_cast<long>(p.nb) * tiles_ps;
    long g = static_cast<long>(sms) * (occ > 0 ? occ : 1);
    if (g > tiles) g = tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(g));
    cfg.blockDim = dim3(ring_threads<Op>());
    cfg.dynamicSmemBytes = static_cast<size_t>(Lay::smem_bytes);
    cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    IG_CUDA(cudaLaunchKernelEx(&cfg, ring_kernel<Op>, p, maps));
    return 0;
}

}  // namespace ig
