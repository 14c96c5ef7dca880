// Per-sample tables: the model matrix M (fat phasor column), its pseudo-inverse, and the pseudo-inverse
// of the magnitude design matrix.  Replaces gen_M / gen_A of the reference
// (/root/reference/wflib/IDEAL_model.py:48-97): there, complex64 QR + triangular solve per call inside
// the TF graph; here, closed-form normal equations in fp64 (cond(M) ~ 1.3-1.6), one
// warp per sample on the device, written once to a (nb, IG_TAB_FLOATS) fp32 table of per-echo records (layout in
// include/idealgan.h) that every operator kernel stages in shared memory with 16-byte copies.  The same code runs on the host for CPU callers (ig_gen_tables_host).
#include <math.h>

#include "ig_common.cuh"

namespace ig {

// IDEAL_model.py:10,14 -- ppm shifts * 42.58 Hz/ppm/T, stored by the reference as complex64 (fp32).
__host__ __device__ inline void fat_model(float f_hz_per_t[6], float amp[6]) {
    const double ppm[6] = {-3.80, -3.40, -2.60, -1.94, -0.39, 0.60};
    const double a[6] = {0.087, 0.693, 0.128, 0.004, 0.039, 0.048};
    for (int p = 0; p < 6; ++p) {
        f_hz_per_t[p] = static_cast<float>(ppm[p] * 1e-6 * 42.58e6);
        amp[p] = static_cast<float>(a[p]);
    }
}

// partial fat phasor of one echo over peaks [p0, p1): sum_p alpha_p exp(2 pi i te field f_p)
__host__ __device__ inline void fat_phasor(float te, float field, int p0, int p1, double &re, double &im) {
    float f_p[6], amp[6];
    fat_model(f_p, amp);
    re = 0.0;
    im = 0.0;
    for (int p = p0; p < p1; ++p) {
        // the reference forms this phase in complex64 (:54): fl32(fl32(2 pi te) * fl32(field f_p)); the
        // three fp32 roundings are reproduced so that M agrees with TF's to ~1e-7 (at 3 T the phase reaches
        // ~40 rad and an exact product would differ from the reference by 3e-6), then sin/cos are exact
        const float a32 = 6.2831855f * te;
        const float b32 = field * f_p[p];
        const float p32 = a32 * b32;
        double sn, cs;
        sincos(static_cast<double>(p32), &sn, &cs);
        re += static_cast<double>(amp[p]) * cs;
        im += static_cast<double>(amp[p]) * sn;
    }
}

// the 16 floats of one echo record.  M^H M = [[ne, s], [conj(s), q]],  M^+ = (M^H M)^-1 M^H ; inv_det = 0 leaves M^+ zero
__host__ __device__ inline void echo_record(float te, int e, int ne, double cr, double ci, double s_re, double s_im, double q, double inv_det,
                                            float *rec) {
    rec[IG_REC_TE] = te;
    rec[IG_REC_KPHI] = te * 300.0f;                                      // fm_sc (IDEAL_model.py:18)
    rec[IG_REC_NTE_L2E] = static_cast<float>(-static_cast<double>(te) * 1.4426950408889634);
    rec[IG_REC_SGN] = (e & 1) ? 1.f : -1.f;                              // (-1)^(e+1), echoes counted from 1 (:250-251)
    rec[IG_REC_C_RE] = static_cast<float>(cr);
    rec[IG_REC_C_IM] = static_cast<float>(ci);
    // water row: (q - s conj(c_e)) / det ; fat row: (ne conj(c_e) - conj(s)) / det
    const double sc_re = s_re * cr + s_im * ci;      // s * conj(c)
    const double sc_im = s_im * cr - s_re * ci;
    const double pw_re = (q - sc_re) * inv_det, pw_im = (-sc_im) * inv_det;
    const double pf_re = (ne * cr - s_re) * inv_det, pf_im = (-ne * ci + s_im) * inv_det;
    const double t = inv_det != 0.0 ? static_cast<double>(te) : 0.0;
    rec[IG_REC_PW_RE] = static_cast<float>(pw_re);
    rec[IG_REC_PW_IM] = static_cast<float>(pw_im);
    rec[IG_REC_PF_RE] = static_cast<float>(pf_re);
    rec[IG_REC_PF_IM] = static_cast<float>(pf_im);
    rec[IG_REC_TPW_RE] = static_cast<float>(t * pw_re);
    rec[IG_REC_TPW_IM] = static_cast<float>(t * pw_im);
    rec[IG_REC_TPF_RE] = static_cast<float>(t * pf_re);
    rec[IG_REC_TPF_IM] = static_cast<float>(t * pf_im);
    // Radians per unit of the phi map for the kernels that hand the phase straight to sin.approx / cos.approx.  Those compile to
    // FMUL.RZ x, 0f3E22F983 + MUFU.SIN/COS (the MUFU works in turns): the constant is 4.03e-8 below 1/(2 pi) and the truncating
    // multiply loses another 4.2e-8 on average (measured over |x| <= 22 rad), a SYSTEMATIC -8.3e-8 relative phase error that the
    // 5 % mismatch of a typical estimate amplifies 40-fold in the objective.  It is folded into this constant instead.
    rec[IG_REC_KPHI_RAD] = static_cast<float>(static_cast<double>(te) * 300.0 * 6.283185307179586 * (1.0 + 8.3e-8));
    rec[15] = 0.f;
}

__host__ __device__ inline double pinv_inv_det(int ne, double s_re, double s_im, double q) {
    const double det = static_cast<double>(ne) * q - (s_re * s_re + s_im * s_im);
    return (ne >= 2 && det > 1e-12) ? 1.0 / det : 0.0;
}

// inverse of the symmetric 3 x 3 Gram matrix of A = [1, Re c, |c|^2] (gen_A, :80-90) by Gauss-Jordan with partial pivoting
__host__ __device__ inline bool invert_gram3(double G[3][3], double Ginv[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Ginv[i][j] = (i == j) ? 1.0 : 0.0;
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        for (int r = col + 1; r < 3; ++r)
            if (fabs(G[r][col]) > fabs(G[piv][col])) piv = r;
        if (fabs(G[piv][col]) < 1e-300) return false;
        for (int j = 0; j < 3; ++j) {
            double t = G[col][j]; G[col][j] = G[piv][j]; G[piv][j] = t;
            t = Ginv[col][j]; Ginv[col][j] = Ginv[piv][j]; Ginv[piv][j] = t;
        }
        const double d = 1.0 / G[col][col];
        for (int j = 0; j < 3; ++j) { G[col][j] *= d; Ginv[col][j] *= d; }
        for (int r = 0; r < 3; ++r) {
            if (r == col) continue;
            const double f = G[r][col];
            for (int j = 0; j < 3; ++j) { G[r][j] -= f * G[col][j]; Ginv[r][j] -= f * Ginv[col][j]; }
        }
    }
    return true;
}

// host: one sample, serial
inline void build_sample_table(const float *te, int ne, float field, float *tab) {
    for (int i = 0; i < IG_TAB_FLOATS; ++i) tab[i] = 0.f;
    double cr[IG_MAX_NE], ci[IG_MAX_NE];
    double s_re = 0.0, s_im = 0.0, q = 0.0;
    for (int e = 0; e < ne; ++e) {
        fat_phasor(te[e], field, 0, 6, cr[e], ci[e]);
        s_re += cr[e];
        s_im += ci[e];
        q += cr[e] * cr[e] + ci[e] * ci[e];
    }
    const double inv_det = pinv_inv_det(ne, s_re, s_im, q);
    for (int e = 0; e < ne; ++e) echo_record(te[e], e, ne, cr[e], ci[e], s_re, s_im, q, inv_det, tab + e * IG_REC_FLOATS);
    tab[IG_TAB_META_OFF + 0] = static_cast<float>(ne);
    tab[IG_TAB_META_OFF + 1] = field;
    if (ne >= 3) {
        double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Ginv[3][3];
        for (int e = 0; e < ne; ++e) {
            const double a[3] = {1.0, cr[e], cr[e] * cr[e] + ci[e] * ci[e]};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) G[i][j] += a[i] * a[j];
        }
        if (invert_gram3(G, Ginv)) {
            for (int e = 0; e < ne; ++e) {
                const double a[3] = {1.0, cr[e], cr[e] * cr[e] + ci[e] * ci[e]};
                for (int i = 0; i < 3; ++i)
                    tab[IG_TAB_AP_OFF + i * IG_MAX_NE + e] = static_cast<float>(Ginv[i][0] * a[0] + Ginv[i][1] * a[1] + Ginv[i][2] * a[2]);
            }
        }
    }
}

// sum over the 16 lanes of a half-warp (both halves end up with their own total)
__device__ __forceinline__ double half_warp_sum(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// device: one warp per sample.  Lane l works on echo l & 15; the two half-warps split the six fat peaks (the
// 36 fp64 sincos of a 6-echo sample are what the table costs), sums over echoes are half-warp shuffles, and each
// of the 320 floats of the table is written exactly once.
//
// AHEAD (ig_gen_tables_ahead): the launch carries the programmatic-dependent-launch attribute, so the warps start while the kernel in
// front of them in the stream is still running (one warp per block: it fits next to two resident blocks of a fused objective), build a
// table that kernel does not touch, and only then wait for it: a kernel launched behind the table therefore still sees everything
// before it in the stream complete, and pays nothing for the table.
constexpr int kTabWarps = 4;
template <int WARPS, bool AHEAD>
__global__ void __launch_bounds__(WARPS * 32) gen_tables_kernel(const float *__restrict__ te, int nb, int ne, float field, float *__restrict__ tab) {
    grid_launch_dependents();          // a kernel launched behind this one with PDL may start its prologue now; it still waits for our writes
    const int b = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (b >= nb) {
        if (AHEAD) grid_dependency_wait();
        return;
    }
    const int lane = threadIdx.x & 31, e = lane & 15, half = lane >> 4;
    const bool live = e < ne;
    const float te_e = live ? te[static_cast<size_t>(b) * ne + e] : 0.f;
    double cr = 0.0, ci = 0.0;
    if (live) fat_phasor(te_e, field, half * 3, half * 3 + 3, cr, ci);
    cr += __shfl_xor_sync(0xffffffffu, cr, 16);
    ci += __shfl_xor_sync(0xffffffffu, ci, 16);
    const double m = cr * cr + ci * ci;
    const double s_re = half_warp_sum(cr), s_im = half_warp_sum(ci), q = half_warp_sum(m);
    const double g11 = half_warp_sum(cr * cr), g12 = half_warp_sum(cr * m), g22 = half_warp_sum(m * m);
    float *tab_b = tab + static_cast<size_t>(b) * IG_TAB_FLOATS;
    if (half == 0) {
        float rec[IG_REC_FLOATS];
#pragma unroll
        for (int i = 0; i < IG_REC_FLOATS; ++i) rec[i] = 0.f;
        if (live) echo_record(te_e, e, ne, cr, ci, s_re, s_im, q, pinv_inv_det(ne, s_re, s_im, q), rec);
        float4 *dst = reinterpret_cast<float4 *>(tab_b + e * IG_REC_FLOATS);
#pragma unroll
        for (int i = 0; i < IG_REC_FLOATS / 4; ++i) dst[i] = make_float4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
    } else {
        float ap[3] = {0.f, 0.f, 0.f};
        if (ne >= 3) {
            double G[3][3] = {{static_cast<double>(ne), s_re, q}, {s_re, g11, g12}, {q, g12, g22}}, Ginv[3][3];
            if (invert_gram3(G, Ginv) && live) {
#pragma unroll
                for (int i = 0; i < 3; ++i) ap[i] = static_cast<float>(Ginv[i][0] + Ginv[i][1] * cr + Ginv[i][2] * m);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) tab_b[IG_TAB_AP_OFF + i * IG_MAX_NE + e] = ap[i];
        tab_b[IG_TAB_META_OFF + e] = e == 0 ? static_cast<float>(ne) : (e == 1 ? field : 0.f);
    }
    if (AHEAD) grid_dependency_wait();     // completion of this grid implies completion of the one in front of it
}

}  // namespace ig

extern "C" int ig_gen_tables(const float *te_d, int nb, int ne, float field, float *tab_d, void *stream) {
    IG_REQUIRE(te_d && tab_d && nb > 0, IG_E_ARG, "ig_gen_tables: null pointer or nb <= 0");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_gen_tables: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    ig::gen_tables_kernel<ig::kTabWarps, false><<<(nb + ig::kTabWarps - 1) / ig::kTabWarps, ig::kTabWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(te_d, nb, ne, field, tab_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_gen_tables_ahead(const float *te_d, int nb, int ne, float field, float *tab_d, void *stream) {
    IG_REQUIRE(te_d && tab_d && nb > 0, IG_E_ARG, "ig_gen_tables_ahead: null pointer or nb <= 0");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_gen_tables_ahead: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(nb));
    cfg.blockDim = dim3(32);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    IG_CUDA(cudaLaunchKernelEx(&cfg, ig::gen_tables_kernel<1, true>, te_d, nb, ne, field, tab_d));
    return 0;
}

extern "C" int ig_gen_tables_host(const float *te_h, int nb, int ne, float field, float *tab_h) {
    IG_REQUIRE(te_h && tab_h && nb > 0, IG_E_ARG, "ig_gen_tables_host: null pointer or nb <= 0");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_gen_tables_host: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    for (int b = 0; b < nb; ++b) ig::build_sample_table(te_h + static_cast<size_t>(b) * ne, ne, field, tab_h + static_cast<size_t>(b) * IG_TAB_FLOATS);
    return 0;
}
