// Per-sample tables: the model matrix M (fat phasor column), its pseudo-inverse, and the pseudo-inverse
// of the magnitude design matrix.  Replaces gen_M / gen_A of the reference
// (/root/reference/wflib/IDEAL_model.py:48-97): there, complex64 QR + triangular solve per call inside
// the TF graph; here, closed-form normal equations in fp64 (cond(M) ~ 1.3-1.6), one
// thread per sample, written once to a (nb, IG_TAB_FLOATS) fp32 table of per-echo records (layout in
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

__host__ __device__ inline void build_sample_table(const float *te, int ne, float field, float *tab) {
    for (int i = 0; i < IG_TAB_FLOATS; ++i) tab[i] = 0.f;
    float f_p[6], amp[6];
    fat_model(f_p, amp);
    double cr[IG_MAX_NE], ci[IG_MAX_NE];
    double s_re = 0.0, s_im = 0.0, q = 0.0;
    for (int e = 0; e < ne; ++e) {
        double re = 0.0, im = 0.0;
        for (int p = 0; p < 6; ++p) {
            // the reference forms this phase in complex64 (:54): fl32(fl32(2 pi te) * fl32(field f_p)); the
            // three fp32 roundings are reproduced so that M agrees with TF's to ~1e-7 (at 3 T the phase reaches
            // ~40 rad and an exact product would differ from the reference by 3e-6), then sin/cos are exact
            const float a32 = 6.2831855f * te[e];
            const float b32 = field * f_p[p];
            const float p32 = a32 * b32;
            double sn, cs;
            sincos(static_cast<double>(p32), &sn, &cs);
            re += static_cast<double>(amp[p]) * cs;
            im += static_cast<double>(amp[p]) * sn;
        }
        cr[e] = re;
        ci[e] = im;
        s_re += re;
        s_im += im;
        q += re * re + im * im;
        float *rec = tab + e * IG_REC_FLOATS;
        rec[IG_REC_TE] = te[e];
        rec[IG_REC_KPHI] = te[e] * 300.0f;                                   // fm_sc (IDEAL_model.py:18)
        rec[IG_REC_NTE_L2E] = static_cast<float>(-static_cast<double>(te[e]) * 1.4426950408889634);
        rec[IG_REC_SGN] = (e & 1) ? 1.f : -1.f;                              // (-1)^(e+1), echoes counted from 1 (:250-251)
        rec[IG_REC_C_RE] = static_cast<float>(re);
        rec[IG_REC_C_IM] = static_cast<float>(im);
    }
    tab[IG_TAB_META_OFF + 0] = static_cast<float>(ne);
    tab[IG_TAB_META_OFF + 1] = field;
    // M^H M = [[ne, s], [conj(s), q]],  M^+ = (M^H M)^-1 M^H
    const double det = static_cast<double>(ne) * q - (s_re * s_re + s_im * s_im);
    if (ne >= 2 && det > 1e-12) {
        const double inv = 1.0 / det;
        for (int e = 0; e < ne; ++e) {
            // water row: (q - s conj(c_e)) / det ; fat row: (ne conj(c_e) - conj(s)) / det
            const double sc_re = s_re * cr[e] + s_im * ci[e];      // s * conj(c)
            const double sc_im = s_im * cr[e] - s_re * ci[e];
            float *rec = tab + e * IG_REC_FLOATS;
            const double pw_re = (q - sc_re) * inv, pw_im = (-sc_im) * inv;
            const double pf_re = (ne * cr[e] - s_re) * inv, pf_im = (-ne * ci[e] + s_im) * inv;
            const double t = static_cast<double>(te[e]);
            rec[IG_REC_PW_RE] = static_cast<float>(pw_re);
            rec[IG_REC_PW_IM] = static_cast<float>(pw_im);
            rec[IG_REC_PF_RE] = static_cast<float>(pf_re);
            rec[IG_REC_PF_IM] = static_cast<float>(pf_im);
            rec[IG_REC_TPW_RE] = static_cast<float>(t * pw_re);
            rec[IG_REC_TPW_IM] = static_cast<float>(t * pw_im);
            rec[IG_REC_TPF_RE] = static_cast<float>(t * pf_re);
            rec[IG_REC_TPF_IM] = static_cast<float>(t * pf_im);
        }
    }
    // A = [1, Re c, |c|^2] (gen_A, :80-90); A^+ = (A^T A)^-1 A^T by Gauss-Jordan with partial pivoting
    if (ne >= 3) {
        double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        double Ginv[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int e = 0; e < ne; ++e) {
            const double a[3] = {1.0, cr[e], cr[e] * cr[e] + ci[e] * ci[e]};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) G[i][j] += a[i] * a[j];
        }
        bool ok = true;
        for (int col = 0; col < 3; ++col) {
            int piv = col;
            for (int r = col + 1; r < 3; ++r)
                if (fabs(G[r][col]) > fabs(G[piv][col])) piv = r;
            if (fabs(G[piv][col]) < 1e-300) { ok = false; break; }
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
        if (ok) {
            for (int e = 0; e < ne; ++e) {
                const double a[3] = {1.0, cr[e], cr[e] * cr[e] + ci[e] * ci[e]};
                for (int i = 0; i < 3; ++i) {
                    const double v = Ginv[i][0] * a[0] + Ginv[i][1] * a[1] + Ginv[i][2] * a[2];
                    tab[IG_TAB_AP_OFF + i * IG_MAX_NE + e] = static_cast<float>(v);
                }
            }
        }
    }
}

__global__ void gen_tables_kernel(const float *__restrict__ te, int nb, int ne, float field, float *__restrict__ tab) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    float te_l[IG_MAX_NE];
    for (int e = 0; e < ne; ++e) te_l[e] = te[static_cast<size_t>(b) * ne + e];
    float out[IG_TAB_FLOATS];
    build_sample_table(te_l, ne, field, out);
    for (int i = 0; i < IG_TAB_FLOATS; ++i) tab[static_cast<size_t>(b) * IG_TAB_FLOATS + i] = out[i];
}

}  // namespace ig

extern "C" int ig_gen_tables(const float *te_d, int nb, int ne, float field, float *tab_d, void *stream) {
    IG_REQUIRE(te_d && tab_d && nb > 0, IG_E_ARG, "ig_gen_tables: null pointer or nb <= 0");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_gen_tables: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    const int threads = 64;
    ig::gen_tables_kernel<<<(nb + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(te_d, nb, ne, field, tab_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_gen_tables_host(const float *te_h, int nb, int ne, float field, float *tab_h) {
    IG_REQUIRE(te_h && tab_h && nb > 0, IG_E_ARG, "ig_gen_tables_host: null pointer or nb <= 0");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_gen_tables_host: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    for (int b = 0; b < nb; ++b) ig::build_sample_table(te_h + static_cast<size_t>(b) * ne, ne, field, tab_h + static_cast<size_t>(b) * IG_TAB_FLOATS);
    return 0;
}
