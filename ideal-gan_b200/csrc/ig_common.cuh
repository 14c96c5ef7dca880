// Shared device/host helpers for libidealgan (sm_100a only).
//
// Design: every kernel is "one thread = two neighbouring voxels", with the two voxels
// packed in the lanes of Blackwell's f32x2 arithmetic (FFMA2/FMUL2/FADD2 via __ffma2_rn & co., new on
// sm_100), so the per-voxel complex algebra issues half the FP32 instructions of a scalar kernel while
// global loads/stores are 16-byte float4 and fully coalesced.  Per-sample constants (echo times, fat
// phasor, pseudo-inverse rows) are staged once per block in shared memory and enter the packed math as
// scalar-broadcast operands.  Transcendentals go to the SFU (MUFU.SIN/COS/EX2) after an exact range
// reduction in turns.  A scalar (one voxel per thread) instantiation of the same templates handles odd
// voxel counts / unaligned planes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <initializer_list>
#include <type_traits>

#include "../../include/idealgan.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libidealgan targets sm_100a (B200) only"
#endif

namespace ig {

constexpr float kFmSc = 300.0f;     // IDEAL_model.py:18
constexpr float kRhoSc = 1.4f;      // IDEAL_model.py:19
constexpr float kTwoPi = 6.283185307179586f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// host-side error plumbing
// ------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
#define IG_CUDA(call)                                                  \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return ::ig::cuda_fail(e__, #call);    \
    } while (0)
#define IG_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            ::ig::set_error(__VA_ARGS__);    \
            return (code);                   \
        }                                    \
    } while (0)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// lane types: float (one voxel per thread) or pk (two voxels per thread, f32x2 packed)
// ------------------------------------------------------------------------------------------------
struct pk {
    float2 d;
};
__device__ __forceinline__ pk mk(float a, float b) { pk r; r.d = make_float2(a, b); return r; }

template <typename V> struct lanes;
template <> struct lanes<float> { static constexpr int n = 1; };
template <> struct lanes<pk> { static constexpr int n = 2; };

__device__ __forceinline__ float bc(float s, float) { return s; }          // broadcast scalar to lane type
__device__ __forceinline__ pk bc(float s, pk) { return mk(s, s); }
template <typename V> __device__ __forceinline__ V splat(float s) { return bc(s, V{}); }

__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vneg(float a) { return -a; }

#ifdef IG_NO_F32X2
// experiment build (make SUFFIX=_s EXTRA=-DIG_NO_F32X2=1): the two lanes as scalar FFMA / FMUL / FADD.  An FFMA2 holds the issue port and the
// FMA pipe for two cycles (tools/ubench_f32x2.cu), so packing buys no arithmetic rate; this build measures what it costs or saves otherwise.
__device__ __forceinline__ pk vfma(pk a, pk b, pk c) { return mk(fmaf(a.d.x, b.d.x, c.d.x), fmaf(a.d.y, b.d.y, c.d.y)); }
__device__ __forceinline__ pk vmul(pk a, pk b) { return mk(a.d.x * b.d.x, a.d.y * b.d.y); }
__device__ __forceinline__ pk vadd(pk a, pk b) { return mk(a.d.x + b.d.x, a.d.y + b.d.y); }
__device__ __forceinline__ pk vneg(pk a) { return mk(-a.d.x, -a.d.y); }
__device__ __forceinline__ pk vsub(pk a, pk b) { return mk(a.d.x - b.d.x, a.d.y - b.d.y); }
#else
__device__ __forceinline__ pk vfma(pk a, pk b, pk c) { pk r; r.d = __ffma2_rn(a.d, b.d, c.d); return r; }
__device__ __forceinline__ pk vmul(pk a, pk b) { pk r; r.d = __fmul2_rn(a.d, b.d); return r; }
__device__ __forceinline__ pk vadd(pk a, pk b) { pk r; r.d = __fadd2_rn(a.d, b.d); return r; }
__device__ __forceinline__ pk vneg(pk a) { return mk(-a.d.x, -a.d.y); }
__device__ __forceinline__ pk vsub(pk a, pk b) { return vfma(mk(-1.0f, -1.0f), b, a); }
#endif
// scalar (block-uniform table coefficient) x packed: the compiler folds make_float2(s, s) into FFMA2's
// scalar-broadcast operand form (R.F32), so no register pair is materialised.
__device__ __forceinline__ pk vfma(float a, pk b, pk c) { return vfma(mk(a, a), b, c); }
__device__ __forceinline__ pk vmul(float a, pk b) { return vmul(mk(a, a), b); }

template <typename V> struct cx {
    V re, im;
};
template <typename V> __device__ __forceinline__ cx<V> czero() { return cx<V>{splat<V>(0.f), splat<V>(0.f)}; }
// acc += (ar + i ai) * z        (table coefficient times per-voxel complex)
template <typename V> __device__ __forceinline__ void cmac(cx<V> &acc, float ar, float ai, const cx<V> &z) {
    acc.re = vfma(ar, z.re, acc.re);
    acc.re = vfma(-ai, z.im, acc.re);
    acc.im = vfma(ar, z.im, acc.im);
    acc.im = vfma(ai, z.re, acc.im);
}
// a + (cr + i ci) * b
template <typename V> __device__ __forceinline__ cx<V> caffine(const cx<V> &a, float cr, float ci, const cx<V> &b) {
    cx<V> r = a;
    cmac(r, cr, ci, b);
    return r;
}
template <typename V> __device__ __forceinline__ cx<V> cmulv(const cx<V> &a, const cx<V> &b) {
    cx<V> r;
    r.re = vfma(vneg(a.im), b.im, vmul(a.re, b.re));
    r.im = vfma(a.im, b.re, vmul(a.re, b.im));
    return r;
}
// conj(a) * b
template <typename V> __device__ __forceinline__ cx<V> cmulc(const cx<V> &a, const cx<V> &b) {
    cx<V> r;
    r.re = vfma(a.im, b.im, vmul(a.re, b.re));
    r.im = vfma(vneg(a.im), b.re, vmul(a.re, b.im));
    return r;
}
template <typename V> __device__ __forceinline__ cx<V> cscale(V s, const cx<V> &a) {
    return cx<V>{vmul(s, a.re), vmul(s, a.im)};
}

// ------------------------------------------------------------------------------------------------
// SFU transcendentals.  Phases are carried in TURNS: tau -> tau - rint(tau) is exact in fp32, after
// which MUFU.SIN/COS see |x| <= pi where their absolute error is ~2^-21.4.  Decay uses MUFU.EX2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float round_turns(float t) {
    // round-to-nearest-integer by the 1.5 * 2^23 trick: two FADDs on the FMA pipe, no conversion pipe
    const float magic = 12582912.0f;
    float n = __fadd_rn(t, magic);
    n = __fsub_rn(n, magic);
    return __fsub_rn(t, n);
}
__device__ __forceinline__ void unit_phasor(float turns, float &c, float &s) {
    float r = round_turns(turns) * kTwoPi;
    c = __cosf(r);
    s = __sinf(r);
}
__device__ __forceinline__ void unit_phasor(pk turns, pk &c, pk &s) {
    // same reduction, packed: 3 FADD2 + 1 FMUL2 for the two voxels, then the SFU per lane
    const float2 magic = make_float2(12582912.0f, 12582912.0f), nmagic = make_float2(-12582912.0f, -12582912.0f);
    float2 n = __fadd2_rn(turns.d, magic);
    n = __fadd2_rn(n, nmagic);
    const float2 r = __fmul2_rn(__ffma2_rn(n, make_float2(-1.f, -1.f), turns.d), make_float2(kTwoPi, kTwoPi));
    c = mk(__cosf(r.x), __cosf(r.y));
    s = mk(__sinf(r.x), __sinf(r.y));
}
// Phase already in radians, no explicit reduction: sin.approx / cos.approx are FMUL by 1/(2 pi) + MUFU.SIN/COS, and the MUFU
// takes its argument in turns and drops the integer part itself, so the only cost of |x| > pi is the fp32 spacing of x
// (|x| <= 25 rad on this path: 2^-20 rad).  Saves the 3 FADD2 + FMUL2 of the exact reduction per echo.
__device__ __forceinline__ void unit_phasor_rad(float rad, float &c, float &s) {
    c = __cosf(rad);
    s = __sinf(rad);
}
__device__ __forceinline__ void unit_phasor_rad(pk rad, pk &c, pk &s) {
    c = mk(__cosf(rad.d.x), __cosf(rad.d.y));
    s = mk(__sinf(rad.d.x), __sinf(rad.d.y));
}
__device__ __forceinline__ pk fast_ex2(pk x) { return mk(fast_ex2(x.d.x), fast_ex2(x.d.y)); }
__device__ __forceinline__ float vrelu(float x) { return fmaxf(x, 0.f); }
__device__ __forceinline__ pk vrelu(pk x) { return mk(fmaxf(x.d.x, 0.f), fmaxf(x.d.y, 0.f)); }
__device__ __forceinline__ float vgate(float x, float g) { return x > 0.f ? g : 0.f; }      // relu'(x) * g
__device__ __forceinline__ pk vgate(pk x, pk g) { return mk(x.d.x > 0.f ? g.d.x : 0.f, x.d.y > 0.f ? g.d.y : 0.f); }
__device__ __forceinline__ float vsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ pk vsqrt(pk x) { return mk(sqrtf(x.d.x), sqrtf(x.d.y)); }
// (a != 0) ? s - a : 0, per lane -- the reference's where(A != 0, S_hat, 0) followed by (S_hat - A)
__device__ __forceinline__ float mask_sub(float s, float a) { return a != 0.f ? s - a : 0.f; }
__device__ __forceinline__ pk mask_sub(pk s, pk a) { return mk(mask_sub(s.d.x, a.d.x), mask_sub(s.d.y, a.d.y)); }
__device__ __forceinline__ float hsum(float x) { return x; }
__device__ __forceinline__ float hsum(pk x) { return x.d.x + x.d.y; }

// ------------------------------------------------------------------------------------------------
// streaming global access.  A complex plane stores (re, im) per voxel; a thread owns lanes<V>::n
// consecutive voxels starting at voxel index v0 (even for pk), i.e. one 8- or 16-byte access.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cx<float> ld_cx(const float *plane, int v0, float) {
    float2 t = __ldcs(reinterpret_cast<const float2 *>(plane) + v0);
    return cx<float>{t.x, t.y};
}
__device__ __forceinline__ cx<pk> ld_cx(const float *plane, int v0, pk) {
    float4 t = __ldcs(reinterpret_cast<const float4 *>(plane) + (v0 >> 1));
    return cx<pk>{mk(t.x, t.z), mk(t.y, t.w)};
}
__device__ __forceinline__ void st_cx(float *plane, int v0, const cx<float> &z) {
    __stcs(reinterpret_cast<float2 *>(plane) + v0, make_float2(z.re, z.im));
}
__device__ __forceinline__ void st_cx(float *plane, int v0, const cx<pk> &z) {
    __stcs(reinterpret_cast<float4 *>(plane) + (v0 >> 1), make_float4(z.re.d.x, z.im.d.x, z.re.d.y, z.im.d.y));
}
// one real value per voxel (plane of nv floats)
__device__ __forceinline__ void st_real(float *plane, int v0, float x) { __stcs(plane + v0, x); }
__device__ __forceinline__ void st_real(float *plane, int v0, pk x) {
    __stcs(reinterpret_cast<float2 *>(plane) + (v0 >> 1), x.d);
}
__device__ __forceinline__ float ld_real(const float *plane, int v0, float) { return __ldcs(plane + v0); }
__device__ __forceinline__ pk ld_real(const float *plane, int v0, pk) {
    pk r; r.d = __ldcs(reinterpret_cast<const float2 *>(plane) + (v0 >> 1)); return r;
}
// lane-wise access helpers for the few strided layouts (flat, mag/phase rows)
__device__ __forceinline__ float lane_get(float x, int) { return x; }
__device__ __forceinline__ float lane_get(pk x, int l) { return l ? x.d.y : x.d.x; }
__device__ __forceinline__ void lane_set(float &x, int, float v) { x = v; }
__device__ __forceinline__ void lane_set(pk &x, int l, float v) { if (l) x.d.y = v; else x.d.x = v; }

// 1 - e^{-x} for x >= 0 without the cancellation of the literal form (the reference's fp32 `1 - exp(-x)` is off by ~1e-7 / x): a five-term
// series below 0.1 (truncation 1.4e-8 relative), the direct difference above (SFU error 1.2e-7 / x <= 1.2e-6)
__device__ __forceinline__ float one_minus_exp_neg_fast(float x) {
    const float direct = 1.0f - fast_ex2(-kLog2e * x);
    float s = fmaf(x, -1.0f / 120.0f, 1.0f / 24.0f);
    s = fmaf(-x, s, 1.0f / 6.0f);
    s = fmaf(-x, s, 0.5f);
    s = fmaf(-x, s, 1.0f);
    return x < 0.1f ? x * s : direct;
}
__device__ __forceinline__ pk one_minus_exp_neg_fast(pk x) {
    const pk direct = vsub(splat<pk>(1.0f), fast_ex2(vmul(-kLog2e, x)));
    pk s = vfma(x, splat<pk>(-1.0f / 120.0f), splat<pk>(1.0f / 24.0f));
    s = vfma(vneg(x), s, splat<pk>(1.0f / 6.0f));
    s = vfma(vneg(x), s, splat<pk>(0.5f));
    s = vfma(vneg(x), s, splat<pk>(1.0f));
    const pk ser = vmul(x, s);
    return mk(x.d.x < 0.1f ? ser.d.x : direct.d.x, x.d.y < 0.1f ? ser.d.y : direct.d.y);
}

// ------------------------------------------------------------------------------------------------
// per-sample table staged in shared memory
// ------------------------------------------------------------------------------------------------
struct __align__(16) EchoRec {
    float te;      // seconds
    float kphi;    // te * fm_sc            : turns per unit of the phi/300 map
    float kdec;    // -te * r2_sc * log2(e) : log2 of the decay per unit of the R2*/r2_sc map (global table: -te log2 e)
    float sgn;     // (-1)^e, e = 1..ne     : bipolar odd/even sign
    float c_re, c_im;
    float pw_re, pw_im, pf_re, pf_im;
    float tpw_re, tpw_im, tpf_re, tpf_im;   // te * M^+ rows
    float kphi_rad; // 2 pi te fm_sc         : phase in RADIANS per unit of the phi/300 map (one rounding, from fp64)
    float pad1;
};
static_assert(sizeof(EchoRec) == IG_REC_FLOATS * sizeof(float), "echo record layout");

template <int NE> struct SampleTab {
    EchoRec r[NE];
};

// Straight 16-byte copy of the sample's first NE echo records (records beyond `ne` are zero in the global
// table); the one r2_sc-dependent field is scaled on the way.  Only NE*4 threads take part.
template <int NE> __device__ __forceinline__ void stage_table_nosync(SampleTab<NE> &t, const float *__restrict__ tab_b, float r2_sc) {
    if (threadIdx.x < NE * 4) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(tab_b) + threadIdx.x);
        if ((threadIdx.x & 3) == 0) v.z *= r2_sc;
        reinterpret_cast<float4 *>(t.r)[threadIdx.x] = v;
    }
}
template <int NE> __device__ __forceinline__ void stage_table(SampleTab<NE> &t, const float *__restrict__ tab_b, int ne, float r2_sc) {
    stage_table_nosync(t, tab_b, r2_sc);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// per-echo modulation shared by the solve-side kernels
// ------------------------------------------------------------------------------------------------
// demodulator for one echo: Wm = dinv * conj(u), Wp = d * u
template <typename V> struct Mod {
    V c, s, d, dinv;
};
template <int NE, typename V> __device__ __forceinline__ Mod<V> modulator(const SampleTab<NE> &T, int e, V phi_t, V r2, V bturn) {
    Mod<V> m;
    unit_phasor(vfma(T.r[e].sgn, bturn, vmul(T.r[e].kphi, phi_t)), m.c, m.s);
    const V lg = vmul(T.r[e].kdec, r2);
    m.d = fast_ex2(lg);
    m.dinv = fast_ex2(vneg(lg));
    return m;
}
template <typename V, bool BIP = true, bool RAD = false> __device__ __forceinline__ Mod<V> modulator_rec(const EchoRec &R, V phi_t, V r2, V bturn) {
    Mod<V> m;
    if constexpr (BIP) unit_phasor(vfma(R.sgn, bturn, vmul(R.kphi, phi_t)), m.c, m.s);
    else if constexpr (RAD) unit_phasor_rad(vmul(R.kphi_rad, phi_t), m.c, m.s);
    else unit_phasor(vmul(R.kphi, phi_t), m.c, m.s);
    const V lg = vmul(R.kdec, r2);
    m.d = fast_ex2(lg);
    m.dinv = fast_ex2(vneg(lg));
    return m;
}
template <typename V> __device__ __forceinline__ cx<V> demod(const Mod<V> &m, const cx<V> &S) {       // Wm S
    const cx<V> t{vfma(m.s, S.im, vmul(m.c, S.re)), vfma(vneg(m.s), S.re, vmul(m.c, S.im))};
    return cscale(m.dinv, t);
}
template <typename V> __device__ __forceinline__ cx<V> remod(const Mod<V> &m, const cx<V> &y) {       // Wp y
    const cx<V> t{vfma(vneg(m.s), y.im, vmul(m.c, y.re)), vfma(m.s, y.re, vmul(m.c, y.im))};
    return cscale(m.d, t);
}
template <typename V> __device__ __forceinline__ cx<V> remod_inv(const Mod<V> &m, const cx<V> &g) {   // conj(Wm) g = dinv u g
    const cx<V> t{vfma(vneg(m.s), g.im, vmul(m.c, g.re)), vfma(m.s, g.re, vmul(m.c, g.im))};
    return cscale(m.dinv, t);
}
template <typename V> __device__ __forceinline__ cx<V> demod_fwd(const Mod<V> &m, const cx<V> &G) {   // conj(Wp) G = d conj(u) G
    const cx<V> t{vfma(m.s, G.im, vmul(m.c, G.re)), vfma(vneg(m.s), G.re, vmul(m.c, G.im))};
    return cscale(m.d, t);
}

// ------------------------------------------------------------------------------------------------
// loss reduction: per-thread partial -> warp shuffle -> block -> one float per block in scratch; the
// last block to finish (ticket counter) adds the per-block partials in a fixed order in fp64 and
// writes the scalar, then re-zeroes the ticket so the scratch can be reused by the next launch.
// scratch layout: [0] unsigned ticket, [4] unsigned dynamic tile counter, [16..] float partials[gridDim.x * gridDim.y]
// ------------------------------------------------------------------------------------------------
constexpr size_t kScratchHeader = 16;

// Scalar exchange between the GPUs of a data-parallel job, fused into the objective's finishing block (ig_peer.cu): every
// rank owns a mailbox of kPeerSlots x world 8-byte words {step + 1, float bits}; the finishing thread of step i stores its
// scalar into slot i % kPeerSlots of EVERY rank's mailbox (its own through a local pointer, the others through IPC-mapped
// peer memory: posted stores over NVLink, nothing waits for them) and then adds up the `world` words of step i - lag in its
// own mailbox in rank order (identical bits on every rank).  No collective kernel, no host call.  boxes == nullptr: off.
// lag = 1 makes every step wait for the slowest rank's previous step; lag = 2 leaves a step of slack, so the wait is only
// taken when a rank really falls behind.  A rank publishes step j + 1 having seen its peers' step j - lag, whose readers may
// still be on step j - 2 lag: slots must outnumber 2 lag + 1.
constexpr int kPeerSlots = 8;
constexpr int kPeerMaxLag = 3;
constexpr unsigned long long kPeerTimeoutNs = 2000000000ull;       // a peer that never delivers gives NaN, not a hung GPU
struct PeerPub {
    unsigned long long *const *boxes;      // device array [world] of mailbox base pointers
    float *prev_out;                       // <- global scalar of step - lag (optional)
    int rank, world;
    unsigned step, lag;
};

int peer_pub(const ig_peer *peer, unsigned step, int lag, float *prev_out, PeerPub *out);      // ig_peer.cu: fills the kernel-side view

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// sum over ranks of the words of `step` in `box` (fp64, rank order); NaN after kPeerTimeoutNs
__device__ __forceinline__ float peer_collect(const unsigned long long *box, int world, unsigned step) {
    const volatile unsigned long long *slot = box + static_cast<size_t>(step % kPeerSlots) * world;
    const unsigned long long t0 = global_timer_ns();
    double acc = 0.0;
    for (int r = 0; r < world; ++r) {
        unsigned long long w = slot[r];
        while (static_cast<unsigned>(w >> 32) != step + 1u) {
            if (global_timer_ns() - t0 > kPeerTimeoutNs) return __int_as_float(0x7fc00000);
            __nanosleep(64);
            w = slot[r];
        }
        acc += static_cast<double>(__uint_as_float(static_cast<unsigned>(w)));
    }
    return static_cast<float>(acc);
}

__device__ __forceinline__ void peer_exchange(const PeerPub &peer, float value) {
    const unsigned long long word = (static_cast<unsigned long long>(peer.step + 1u) << 32) | __float_as_uint(value);
    const size_t at = static_cast<size_t>(peer.step % kPeerSlots) * peer.world + peer.rank;
    for (int r = 0; r < peer.world; ++r) *reinterpret_cast<volatile unsigned long long *>(peer.boxes[r] + at) = word;
    if (peer.prev_out && peer.step >= peer.lag) peer.prev_out[0] = peer_collect(peer.boxes[peer.rank], peer.world, peer.step - peer.lag);
}

__device__ __forceinline__ void block_loss_reduce(float v, void *scratch, float *loss_out, float scale, const PeerPub peer = PeerPub{}) {
    __shared__ float warp_part[32];
    __shared__ bool is_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    if (lane == 0) warp_part[warp] = v;
    __syncthreads();
    unsigned *ticket = reinterpret_cast<unsigned *>(scratch);
    float *partials = reinterpret_cast<float *>(reinterpret_cast<char *>(scratch) + kScratchHeader);
    const unsigned nblocks = gridDim.x * gridDim.y;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += warp_part[w];
        partials[bid] = s;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < nblocks; i += blockDim.x) acc += static_cast<double>(__ldcg(partials + i));
    __shared__ double dpart[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) dpart[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += dpart[w];
        const float total = static_cast<float>(s * static_cast<double>(scale));
        loss_out[0] = total;
        ticket[0] = 0u;
        ticket[1] = 0u;      // dynamic tile counter of the persistent kernels
        if (peer.boxes) peer_exchange(peer, total);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copies + mbarrier pipeline (sm_90+/sm_100a PTX; shows up as UBLKCP / SYNCS in SASS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes to shared memory made visible to the async proxy (TMA) before the stage is handed back to the producer
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "IG_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra IG_DONE_%=;\n"
        "bra IG_WAIT_%=;\n"
        "IG_DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// tiled TMA load of a 3-D box described by a tensor map (kernel parameter, __grid_constant__); coordinates innermost first
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// programmatic dependent launch (PDL): wait for the kernel this launch depends on / let the dependent kernel start early
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// barrier among the first `count` threads of the block (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void named_barrier(int count) { asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory"); }

// Ring operators: one instantiation per echo count instead of a bucket with a run-time count: behind `if (e < ne)` the compiler cannot lift the stage reads of
// the later echoes over the earlier echoes' math, and with 8-16 consumer warps per SM that latency shows (measured, 64 x 384 x 384: the
// acq_to_acq adjoint at 7 echoes in the 8-echo bucket 0.226 ms against 0.201 ms AT 8 echoes; the WF-PM adjoint at 9 in the 12-echo bucket
// 0.238 against 0.205 ms at 12).
template <int N, int HI, typename F> inline int dispatch_exact_ne(int ne, F &&f) {
    if constexpr (N > HI) {
        return IG_E_UNSUPPORTED;
    } else {
        if (ne == N) return f(std::integral_constant<int, N>{});
        return dispatch_exact_ne<N + 1, HI>(ne, f);
    }
}

// NE buckets of the plain kernels: they are instantiated for these echo counts; a call with `ne` echoes runs in the
// smallest bucket >= ne with the unused echoes predicated off (their table entries are zero).
template <typename F> inline int dispatch_ne(int ne, F &&f) {
    if (ne <= 4) return f(std::integral_constant<int, 4>{});
    if (ne <= 6) return f(std::integral_constant<int, 6>{});
    if (ne <= 8) return f(std::integral_constant<int, 8>{});
    if (ne <= 12) return f(std::integral_constant<int, 12>{});
    return f(std::integral_constant<int, 16>{});
}

// Operators on the generic TMA ring (ig_ring_ops.cu).  IG_E_UNSUPPORTED = shape / alignment not covered: run the plain kernel.
int a2a_bwd_ring(const float *acqs, const float *pm, long pm_bstride, const float *tab, int nb, int ne, int nv, float r2_sc, const float *g_rho,
                 const float *g_shat, float *g_acqs, float *g_pm, cudaStream_t st);
int magpha_loss_ring(const float *maps, const float *acqs, const float *tab, int nb, int ne, int nv, float r2_sc, float inv_n, float *gmaps, float *shat,
                     float *loss, void *scratch, cudaStream_t st);
int row_loss_ring(int model, const float *maps, int rows, const float *acqs, const float *tab, int nb, int ne, int nv, float r2_sc, int flags, float inv_n,
                  float *gmaps, float *shat, float *loss, void *scratch, cudaStream_t st);
int row_bwd_ring(int model, const float *maps, int rows, const float *gout, const float *tab, int nb, int ne, int nv, float r2_sc, int flags, float *gmaps,
                 cudaStream_t st);
int pdff_unc_ring(const float *acqs, const float *phi_mean, const float *phi_var, const float *r2_mean, const float *r2_var, const float *tab, int nb,
                  int ne, int nv, float r2_sc, float *rho, float *cov, cudaStream_t st);
int a2a_rician_loss_ring(const float *acqs, const float *pm, long pm_bstride, const float *phi_var, const float *r2_mean, const float *r2_var,
                         const float *tab, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm, float *g_phi_var, float *g_r2_mean,
                         float *g_r2_var, float *rho, float *loss, void *scratch, cudaStream_t st);

inline dim3 grid_for(int nb, int nv, int vpt) {
    const int per_block = kThreads * vpt;
    return dim3(static_cast<unsigned>((nv + per_block - 1) / per_block), static_cast<unsigned>(nb), 1);
}

}  // namespace ig
