/*
 * idealgan.h -- C ABI of libidealgan.so: B200 (sm_100a) kernels for the IDEAL water/fat physics path.
 *
 * The reference (jpmeneses/IDEAL-GAN) has no FFI: its hot path is TensorFlow op chains inside
 * wflib/IDEAL_model.py.  Every entry point below therefore replaces one *Python function* of that file
 * (cited per declaration as IDEAL_model.py:<lines>) and is what the reference-side binding of
 * INTEGRATION.md (ctypes + DLPack + tf.custom_gradient) calls.
 *
 * Conventions
 *  - All tensors are float32, C-contiguous, in the reference's own layouts (data.py:98-137):
 *      acquisitions "MEBCRN"  (nb, ne, nv, 2)      nv = H*W, last axis (Re, Im)
 *      acquisitions "flat"    (nb, nv, 2*ne)       Re/Im interleaved per echo
 *      maps WF-PM             (nb, rows, nv, 2)    rows >= 3; rows > 3: the LAST row = bipolar phase / pi (IDEAL_model.py:246-247)
 *      maps ff/pd/phase       (nb, 3, nv, 2)
 *      maps mag/phase         (nb, 2, nv, ch)      ch = 3 | 4 (channel 3 of row 1 = bipolar / 4pi)
 *      PM ("param maps")      row 0 of a (nb, rows>=1, nv, 2) tensor: (phi/300, R2* / r2_sc)
 *  - Pointers named *_d are DEVICE pointers of the current CUDA device; `stream` is a cudaStream_t
 *    (NULL = legacy default stream).  Calls are asynchronous on that stream, allocate nothing and keep
 *    no global mutable state; the only host-side state is a thread-local error string.
 *  - Per-sample constants (echo times, fat phasor c_e = M[e,1], pseudo-inverse rows) live in a table
 *    of IG_TAB_FLOATS floats per sample (16-byte aligned) built by ig_gen_tables from the (nb, ne) echo times.
 *  - Return value: 0 ok; <0 invalid argument (IG_E_*); >0 a cudaError_t.  Never throws.
 *  - `ne` <= IG_MAX_NE.  Optional outputs may be NULL.
 */
#ifndef IDEALGAN_H_
#define IDEALGAN_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IG_VERSION 100          /* 0.1.0 */
#define IG_MAX_NE 16
#define IG_REC_FLOATS 16                                   /* one record per echo                     */
#define IG_TAB_AP_OFF (IG_MAX_NE * IG_REC_FLOATS)          /* A^+ rows: 3 x IG_MAX_NE                 */
#define IG_TAB_META_OFF (IG_TAB_AP_OFF + 3 * IG_MAX_NE)    /* [0] = ne, [1] = field                   */
#define IG_TAB_FLOATS (IG_TAB_META_OFF + 16)               /* 320 floats = 1280 B per sample          */

/* Per-sample table = IG_MAX_NE echo records (zero beyond ne) + A^+ + meta.  Fields of an echo record: */
enum {
    IG_REC_TE = 0,       /* echo time [s]                                                            */
    IG_REC_KPHI = 1,     /* te * fm_sc: phase in TURNS per unit of the phi/300 map                    */
    IG_REC_NTE_L2E = 2,  /* -te * log2(e): log2 of the decay per unit of R2* [1/s]                    */
    IG_REC_SGN = 3,      /* (-1)^e, e = 1..ne: bipolar odd/even sign                                  */
    IG_REC_C_RE = 4,     /* fat phasor c_e = sum_p alpha_p exp(2 pi i te field f_p)      (M[e,1])     */
    IG_REC_C_IM = 5,
    IG_REC_PW_RE = 6,    /* pseudo-inverse row for water  M^+[0,e]                                    */
    IG_REC_PW_IM = 7,
    IG_REC_PF_RE = 8,    /* pseudo-inverse row for fat    M^+[1,e]                                    */
    IG_REC_PF_IM = 9,
    IG_REC_TPW_RE = 10,  /* te * M^+[0,e]                                                             */
    IG_REC_TPW_IM = 11,
    IG_REC_TPF_RE = 12,  /* te * M^+[1,e]                                                             */
    IG_REC_TPF_IM = 13,
    IG_REC_KPHI_RAD = 14 /* 2 pi te fm_sc: phase in RADIANS per unit of the phi/300 map                */
};

enum { IG_MODEL_WFPM = 0, IG_MODEL_FFPD = 1, IG_MODEL_MAGPHA = 2 };

/* flags */
enum {
    IG_F_PHASE_CONSTRAINT = 1,  /* get_rho(phase_constraint=True)                                 */
    IG_F_FLAT = 2,              /* flat layout (MEBCRN=False)                                     */
    IG_F_ONLY_MAG = 4,          /* acq_to_acq(only_mag=True) / ig_ideal_fwd: the signal output is |S_hat|, 1 channel */
    IG_F_NO_RELU = 8,           /* IG_MODEL_WFPM without the relu gate on R2* (not used by wflib) */
    IG_F_NO_CLIP = 16           /* ig_ideal_decode: images as computed, without clip_by_value(., 0, 1) */
};

enum { IG_E_ARG = -1, IG_E_NE = -2, IG_E_ALIGN = -3, IG_E_SCRATCH = -4, IG_E_UNSUPPORTED = -5 };

int ig_version(void);
const char *ig_last_error(void);
/* 1 if the library was built for, and the current device is, compute capability 10.x */
int ig_device_ok(void);

/* ---- per-sample tables: gen_M / gen_A (IDEAL_model.py:48-97) --------------------------------- */
/* te_d: (nb, ne) seconds -> tab_d: (nb, IG_TAB_FLOATS).  Arithmetic in fp64, stored fp32. */
int ig_gen_tables(const float *te_d, int nb, int ne, float field, float *tab_d, void *stream);
/* The same table, built AHEAD of its use in a training loop: the launch overlaps the kernel in front of it in `stream` (programmatic
 * dependent launch) instead of waiting for it, and completes only after that kernel has.  The caller promises that te_d was complete
 * before that kernel was launched and that tab_d is neither read nor written by it (e.g. three table buffers in rotation: step i
 * launches the table of batch i + 1, then the objective of batch i).  Work launched behind this call is ordered as usual. */
int ig_gen_tables_ahead(const float *te_d, int nb, int ne, float field, float *tab_d, void *stream);
/* same arithmetic on the host (used by the Python gen_M()/gen_A() wrappers for CPU tensors) */
int ig_gen_tables_host(const float *te_h, int nb, int ne, float field, float *tab_h);

/* ---- loss scratch -------------------------------------------------------------------------- */
/* Bytes of zero-initialised device scratch the *_loss entry points need for (nb, nv).  The kernels
 * leave it zeroed again on completion, so one allocation can be reused launch after launch. */
size_t ig_loss_scratch_bytes(int nb, int nv);

/* ---- forward models: IDEAL_model / IDEAL_mag / IDEAL_mag_phase (IDEAL_model.py:220-299,404-509) */
/* maps_d layout by model (see top); rows_or_ch = rows (WFPM: 3|4, FFPD: 3) or channels (MAGPHA: 3|4).
 * out_d: (nb, ne, nv, 2), or with IG_F_FLAT the channel-interleaved (nb, nv, 2 ne) of data.A_from_MEBCRN (forward only), or with
 * IG_F_ONLY_MAG the magnitudes |S_hat| (nb, ne, nv) (forward only; what gen_LDM_dataset.py:234 reduces the signals to). */
int ig_ideal_fwd(int model, const float *maps_d, int rows_or_ch, const float *tab_d, int nb, int ne, int nv,
                 float r2_sc, int flags, float *out_d, void *stream);
/* The forward model as the dataset-synthesis script consumes it (gen_LDM_dataset.py:156-158,217-235): any of
 *   shat_d (nb, ne, nv, 2) complex signals (the TFRecord payload), mag_d (nb, ne, nv) = clip(|S_e|, 0, 1) (the multi-echo images),
 *   pdff_d (nb, nv) = clip(|F| / (|W| + |F|), 0, 1), r2s_d (nb, nv) = clip(R2* map, 0, 1)           (NULL = not wanted)
 * in ONE pass over the maps; IG_F_NO_CLIP leaves the images unclipped (mag_d is then IDEAL_mag(..)'s only_mag counterpart).
 * 0/0 on background stays NaN, as tf.clip_by_value leaves it. */
int ig_ideal_decode(int model, const float *maps_d, int rows_or_ch, const float *tab_d, int nb, int ne, int nv,
                    float r2_sc, int flags, float *shat_d, float *mag_d, float *pdff_d, float *r2s_d, void *stream);
/* adjoint: gout_d (nb, ne, nv, 2) upstream -> gmaps_d (same shape as maps, every element written) */
int ig_ideal_bwd(int model, const float *maps_d, int rows_or_ch, const float *tab_d, int nb, int ne, int nv,
                 float r2_sc, int flags, const float *gout_d, float *gmaps_d, void *stream);
/* fused forward -> where(A != 0) mask -> MSE -> backward (train-IDEAL-single.py:154-157,175).
 * loss_d[0] = inv_n * sum (A - mask(S_hat))^2 ; gmaps_d = d loss / d maps ; shat_d (optional) = unmasked S_hat.
 * inv_n is 1 / (number of elements of the GLOBAL batch), so shards of a multi-GPU batch sum to the mean. */
int ig_ideal_loss(int model, const float *maps_d, int rows_or_ch, const float *acqs_d, const float *tab_d,
                  int nb, int ne, int nv, float r2_sc, int flags, float inv_n, float *gmaps_d, float *shat_d,
                  float *loss_d, void *scratch_d, size_t scratch_bytes, void *stream);

/* ---- LS water/fat solve: get_rho (IDEAL_model.py:527-624) ------------------------------------- */
/* pm_d points at the (phi, R2*) row of sample 0; consecutive samples are pm_bstride floats apart.
 * bip_d (optional, MEBCRN only) points at the bipolar row of sample 0 (its channel 0 is used), stride
 * bip_bstride.  rho_d: (nb, 2, nv, 2) or flat (nb, nv, 4); demod_d (optional): (nb, ne, nv, 2).
 * With IG_F_FLAT: acqs (nb, nv, 2ne), pm (nb, nv, 2) ordered (R2*, phi) (IDEAL_model.py:559-560). */
int ig_get_rho_fwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *bip_d, long bip_bstride,
                   const float *tab_d, int nb, int ne, int nv, float r2_sc, int flags, float *rho_d,
                   float *demod_d, void *stream);
/* the same solve with the PDFF / R2* map extraction its inference callers run next (ROI-analysis.py:301-306,344-354;
 * gen_LDM_dataset.py:217-218,226-227) as an epilogue on the registers: pdff_d (nb, nv) with pdff_mode 0 |F|/|W+F|, 1 |F|/(|W|+|F|),
 * 2 magnitude-discriminated (0/0 -> 0); r2s_d (nb, nv) = R2* map x r2_sc.  Either map may be NULL, not both. */
int ig_get_rho_maps(const float *acqs_d, const float *pm_d, long pm_bstride, const float *bip_d, long bip_bstride,
                    const float *tab_d, int nb, int ne, int nv, float r2_sc, int flags, int pdff_mode, float *rho_d,
                    float *pdff_d, float *r2s_d, void *stream);
/* adjoint; g_rho_d / g_demod_d upstream (either may be NULL), outputs g_acqs_d (optional), g_pm_d (nb, nv, 2)
 * dense (row layout of pm, without the batch stride), g_bip_d (optional, (nb, nv, 2), channel 1 = 0). */
int ig_get_rho_bwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *bip_d, long bip_bstride,
                   const float *tab_d, int nb, int ne, int nv, float r2_sc, int flags, const float *g_rho_d,
                   const float *g_demod_d, float *g_acqs_d, float *g_pm_d, float *g_bip_d, void *stream);

/* ---- project + resynthesise: acq_to_acq (IDEAL_model.py:142-200; 2-result form of its callers) -- */
/* rho_d (optional): (nb, 2, nv, 2) = rho_hat / rho_sc ; shat_d: (nb, ne, nv, 2) or (nb, ne, nv, 1) with ONLY_MAG */
int ig_a2a_fwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
               float r2_sc, int flags, float *rho_d, float *shat_d, void *stream);
int ig_a2a_bwd(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
               float r2_sc, int flags, const float *g_rho_d, const float *g_shat_d, float *g_acqs_d, float *g_pm_d,
               void *stream);
/* fused config-2 objective (train-IDEAL-unsup.py:214-218,236,255): acq_to_acq -> mask -> MSE -> d/dPM.
 * loss_d[0] = inv_n * sum (A - mask(S_hat))^2 ; g_pm_d (nb, nv, 2) ; rho_d / shat_d optional materialisation. */
int ig_a2a_loss(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
                float r2_sc, float inv_n, float *g_pm_d, float *rho_d, float *shat_d, float *loss_d, void *scratch_d,
                size_t scratch_bytes, void *stream);

/* fused uncertainty-aware objective of the published AI-DEAL model (train-IDEAL-unsup.py:214-231; IDEAL_model.py:142-200,710-767;
 * tf2gan/loss.py:130-140): acq_to_acq -> mask -> acq_uncertainty(stop_gradient(rho)) -> VarMeanSquaredError, with all gradients.
 * phi_var_d / r2_mean_d / r2_var_d: (nb, nv) moment maps in network units (r2_* both NULL = rem_R2).
 * loss_d[0] = inv_n * sum over (e, component) of [ (A - mask(S_hat))^2 / std_e + log std_e ], std_e = sqrt(max(var_e, 1e-5));
 * g_pm_d (nb, nv, 2); g_phi_var_d / g_r2_mean_d / g_r2_var_d (nb, nv) (the r2 outputs may be NULL; zero-filled under rem_R2);
 * rho_d optional (nb, 2, nv, 2) = rho_hat / rho_sc. */
int ig_a2a_uq_loss(const float *acqs_d, const float *pm_d, long pm_bstride, const float *phi_var_d, const float *r2_mean_d,
                   const float *r2_var_d, const float *tab_d, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm_d,
                   float *g_phi_var_d, float *g_r2_mean_d, float *g_r2_var_d, float *rho_d, float *loss_d, void *scratch_d,
                   size_t scratch_bytes, void *stream);

/* Rician (magnitude) objective of the R2* stage (train-IDEAL-unsup.py:267-292; tf2gan/loss.py:143-162): same arguments and outputs;
 * loss_d[0] = inv_n * sum over echoes of -loglik(|A_e| ; nu = where(Re A_e != 0, |S_hat_e|, 0), sigma^2 = max(var_e, 1e-5)),
 * var_e from acq_uncertainty(only_mag=True).  inv_n = 1 / (nb ne nv) of the GLOBAL batch (one element per echo and voxel). */
int ig_a2a_rician_loss(const float *acqs_d, const float *pm_d, long pm_bstride, const float *phi_var_d, const float *r2_mean_d,
                       const float *r2_var_d, const float *tab_d, int nb, int ne, int nv, float r2_sc, float inv_n, float *g_pm_d,
                       float *g_phi_var_d, float *g_r2_mean_d, float *g_r2_var_d, float *rho_d, float *loss_d, void *scratch_d,
                       size_t scratch_bytes, void *stream);

/* ---- second tier: magnitude fit, uncertainty propagation, PDFF (IDEAL_model.py:100-138,314-401,628-767) ---- */
/* eigenvals: x_d (n, 3) = (a, b, c) -> xy_d (n, 2), ratio_d (n); adjoint with optional upstreams */
int ig_eigenvals(const float *x_d, long n, float *xy_d, float *ratio_d, void *stream);
int ig_eigenvals_bwd(const float *x_d, long n, const float *g_xy_d, const float *g_ratio_d, float *gx_d, void *stream);
/* CSE_mag: mag_d (nb, ne, nv) magnitudes, r2_d (nb, nv) R2* / r2_sc, r2nu_d optional second R2* map (R2_prob).
 * Outputs (each optional): rho (nb,2,nv) / rho_sc, fit (nb,ne,nv), demod (nb,ne,nv), ls (nb,3,nv) / rho_sc^2, unc (nb,nv). */
int ig_cse_mag_fwd(const float *mag_d, const float *r2_d, const float *r2nu_d, const float *tab_d, int nb, int ne, int nv,
                   float r2_sc, float *rho_d, float *fit_d, float *demod_d, float *ls_d, float *unc_d, void *stream);
int ig_cse_mag_bwd(const float *mag_d, const float *r2_d, const float *r2nu_d, const float *tab_d, int nb, int ne, int nv,
                   float r2_sc, const float *g_rho_d, const float *g_fit_d, const float *g_demod_d, const float *g_ls_d,
                   const float *g_unc_d, float *g_mag_d, float *g_r2_d, float *g_r2nu_d, void *stream);
/* acq_uncertainty: rho_d (nb,2,nv,2), phi_var_d / r2_mean_d / r2_var_d (nb,nv) in map units (r2_* NULL = rem_R2)
 * -> out_d (nb, ne, nv, only_mag ? 1 : 2); adjoint w.r.t. the three moment maps (rho is a constant for its callers) */
int ig_acq_unc_fwd(const float *rho_d, const float *phi_var_d, const float *r2_mean_d, const float *r2_var_d,
                   const float *tab_d, int nb, int ne, int nv, float r2_sc, int only_mag, float *out_d, void *stream);
int ig_acq_unc_bwd(const float *rho_d, const float *phi_var_d, const float *r2_mean_d, const float *r2_var_d,
                   const float *tab_d, int nb, int ne, int nv, float r2_sc, int only_mag, const float *g_out_d,
                   float *g_phi_var_d, float *g_r2_mean_d, float *g_r2_var_d, void *stream);
/* PDFF_uncertainty: acqs_d (nb,ne,nv,2) + moment maps (nb,nv) -> rho_d (nb,2,nv,2), cov_d (nb,4,nv) = |C| / rho_sc^2 */
int ig_pdff_unc(const float *acqs_d, const float *phi_mean_d, const float *phi_var_d, const float *r2_mean_d,
                const float *r2_var_d, const float *tab_d, int nb, int ne, int nv, float r2_sc, float *rho_d, float *cov_d,
                void *stream);
/* PDFF maps from rho (nb,2,nv,2): mode 0 |F|/|W+F|, 1 |F|/(|W|+|F|), 2 magnitude-discriminated; 0/0 -> 0 */
int ig_pdff_extract(const float *rho_d, int nb, int nv, int mode, float *out_d, void *stream);

/* ---- script-level reductions around the magnitude fit and the ROI analysis ---- */
/* Regularisers of train-IDEAL-mag.py:288-289,308-316 on the CSE_mag outputs, in one pass:
 *   ls_d (nb,3,H*W) fit coefficients (a,b,c), demod_d (nb,ne,H,W) demodulated echoes, r2_d (nb,H,W) R2* map; any may be NULL.
 *   sums_d[5] = Ad_TV, LS_NZ, WF_NZ, LS_cond, R2_TV (unweighted, as the script logs them; WF_NZ is identically 0 in
 *   the reference: it compares A2B_ls[...,:1] with A2B_ls[...,-1:] on a last axis of length 1).
 *   g_* (optional) = gradient of  w_ad_tv*Ad_TV + w_ls_nz*LS_NZ + w_ls_cond*LS_cond + w_r2_tv*R2_TV  w.r.t. the input.
 * scratch_d: ig_mag_regs_scratch_bytes(nb,H,W) bytes, zero-initialised once (the kernel re-arms it). */
size_t ig_mag_regs_scratch_bytes(int nb, int H, int W);
int ig_mag_regs(const float *ls_d, const float *demod_d, const float *r2_d, int nb, int ne, int H, int W, float w_ad_tv,
                float w_ls_nz, float w_ls_cond, float w_r2_tv, float *sums_d, float *g_ls_d, float *g_demod_d, float *g_r2_d,
                void *scratch_d, size_t scratch_bytes, void *stream);
/* Map assembly + PDFF-variance propagation of ROI-analysis.py:301-322.  maps_d (nb,3,nv,2) = W, F, (phi, R2*);
 * var_d (nb,5,nv,2) = |C_WW|, |C_WF|, |C_FW|, |C_FF| (2nd channel 0) and (var phi, var R2*), NULL for mode 0.
 * out_d (nb,nv,nch): |W|, |F|, |W+F|, R2* (map units) [, PDFF variance]; mode 0 nch=4, 1 general, 2 magnitude model. */
int ig_roi_maps(const float *maps_d, const float *var_d, int nb, int nv, int mode, float *out_d, void *stream);

/* ---- layout adapters: data.A_from_MEBCRN / B_from_MEBCRN / B_to_MEBCRN (data.py:262-329) ------------ */
/* acquisitions (nb, ne, nv, 2) <-> channel-interleaved (nb, nv, 2 ne); the second is the adjoint (and inverse) of the first */
int ig_acq_to_flat(const float *acqs_d, int nb, int ne, int nv, float *flat_d, void *stream);
int ig_acq_from_flat(const float *flat_d, int nb, int ne, int nv, float *acqs_d, void *stream);
/* B_from_MEBCRN.  mode 0: maps (nb,3,nv,2) -> (nb,nv,6) = (W_re, W_im, F_re, F_im, R2*, phi).
 * mode 1 (mag_and_phase=True): maps (nb,2,nv,ch) -> (nb,nv,4 + 2 (ch - 2)), phase = c_pha * pi * maps[:,1,:,1]. */
int ig_maps_to_flat(const float *maps_d, int nb, int nv, int mode, int ch, float c_pha, float *flat_d, void *stream);
/* B_to_MEBCRN.  mode 0 'All' (nb,nv,6) -> (nb,3,nv,2); 1 'WF-PM' (nb,nv,4) -> (nb,3,nv,2); 2 'WF' (nb,nv,2) -> (nb,2,nv,2);
 * 3 'PM' (nb,nv,2) -> (nb,1,nv,2). */
int ig_maps_from_flat(const float *flat_d, int nb, int nv, int mode, float *maps_d, void *stream);

/* ---- multi-GPU: scalar-loss exchange over peer memory, fused into the objective's kernel ---- */
/* One process per GPU.  Each rank owns a small device mailbox; the finishing thread of the objective's kernel stores the
 * rank's scalar into every rank's mailbox (own: local pointer, others: CUDA-IPC-mapped peer memory, i.e. posted stores over
 * NVLink) and adds up the scalars of the PREVIOUS step, which have arrived by then: no collective kernel, no host call, no
 * extra launch.  Replaces the `all_reduce` a data-parallel run of train-IDEAL-unsup.py would issue for its logged loss.
 *   ig_peer_create   on the current device; rank in [0, world), world <= 64
 *   ig_peer_handle   -> 64 opaque bytes (a cudaIpcMemHandle_t) to hand to the other ranks (any transport)
 *   ig_peer_connect  handles of ALL ranks in rank order (world x 64 bytes); own entry is ignored
 *   ig_peer_connect_local  same-process alternative (threads / tests): the contexts of all ranks, in rank order
 *   ig_a2a_loss_peer = ig_a2a_loss + publication of step `step`'s scalar; loss_prev_d (optional) <- global loss of step - lag
 *                      (lag 1..3; 1 waits for the slowest rank's previous step every step, 2 leaves a step of slack).
 *                      `step` must increase by 1 per call on every rank (mailbox slots rotate; ranks stay within lag steps).
 *   ig_peer_publish  the same exchange for a scalar already in device memory (any other objective: ig_a2a_uq_loss,
 *                      ig_a2a_rician_loss, ig_ideal_loss): one thread, launched after the loss kernel on the same stream
 *   ig_peer_reduce   global loss of `step` into loss_d (tiny kernel; for the last step, or whenever the scalar is needed at once)
 * A rank that never delivers yields NaN after 2 s instead of a hung GPU. */
typedef struct ig_peer ig_peer;
#define IG_PEER_HANDLE_BYTES 64
int ig_peer_create(int rank, int world, ig_peer **out);
int ig_peer_handle(ig_peer *peer, void *handle_out);
int ig_peer_connect(ig_peer *peer, const void *handles);
int ig_peer_connect_local(ig_peer *const *peers, int world);
int ig_a2a_loss_peer(const float *acqs_d, const float *pm_d, long pm_bstride, const float *tab_d, int nb, int ne, int nv,
                     float r2_sc, float inv_n, float *g_pm_d, float *rho_d, float *shat_d, float *loss_d, void *scratch_d,
                     size_t scratch_bytes, ig_peer *peer, unsigned step, int lag, float *loss_prev_d, void *stream);
int ig_peer_publish(ig_peer *peer, unsigned step, int lag, const float *loss_d, float *loss_prev_d, void *stream);
int ig_peer_reduce(ig_peer *peer, unsigned step, float *loss_d, void *stream);
void ig_peer_destroy(ig_peer *peer);

/* ---- host-buffer pipeline (the call timed as `e2e` by bench.py) -------------------------------- */
/* A context owns device staging buffers and streams on `device`; chunks of `chunk_nb` samples are copied
 * host->device, processed and copied back with copy/compute overlap.  Host buffers should be pinned. */
typedef struct ig_ctx ig_ctx;
int ig_ctx_create(int device, int chunk_nb, int ne, int nv, ig_ctx **out);
void ig_ctx_destroy(ig_ctx *ctx);
/* acqs_h (nb, ne, nv, 2), pm_h (nb, 1, nv, 2), te_h (nb, ne) -> loss_h[0], g_pm_h (nb, 1, nv, 2).  Blocking. */
int ig_a2a_loss_host(ig_ctx *ctx, const float *acqs_h, const float *pm_h, const float *te_h, int nb, float field,
                     float r2_sc, float inv_n, float *loss_h, float *g_pm_h);

/* Config 5 (gen_LDM_dataset.py:140-254): a shard of decoder maps streamed through ig_ideal_decode, host buffers on both sides.
 * The context owns three slots of device staging for chunks of chunk_nb samples (want_shat: also for the complex signals).
 * maps_h (nb, ...) in the model's layout, te_h (nb, ne) -> any of shat_h (nb,ne,nv,2), mag_h (nb,ne,nv), pdff_h (nb,nv),
 * r2s_h (nb,nv) (NULL = not wanted; images clipped to [0,1] unless IG_F_NO_CLIP).  Blocking; host buffers should be pinned. */
typedef struct ig_decode_ctx ig_decode_ctx;
int ig_decode_ctx_create(int device, int model, int rows_or_ch, int chunk_nb, int ne, int nv, int want_shat, ig_decode_ctx **out);
void ig_decode_ctx_destroy(ig_decode_ctx *ctx);
int ig_decode_host(ig_decode_ctx *ctx, const float *maps_h, const float *te_h, int nb, float field, float r2_sc, int flags,
                   float *shat_h, float *mag_h, float *pdff_h, float *r2s_h);

/* Pinned host memory placed on the NUMA node of `device` (sysfs numa_node of its PCI function; plain cudaHostAlloc where the host
 * has one node or hides the topology).  IG_HOST_WRITE_COMBINED: write-combined pages -- faster for the GPU to pull, very slow for
 * the CPU to read back: input staging only.  ig_host_numa_node: the node, or -1. */
enum { IG_HOST_WRITE_COMBINED = 1 };
int ig_host_alloc(size_t bytes, int device, int flags, void **out);
int ig_host_free(void *p);
int ig_host_numa_node(int device);
/* Bare cudaMemcpyAsync loop, the ceiling of any host-buffer pipeline on this host: dir 0 host->device, 1 device->host, 2 both at
 * once (second buffer pair).  seconds_out = wall time of `reps` copies of `bytes` per direction. */
int ig_copy_probe(void *host, void *dev, void *host2, void *dev2, size_t bytes, int reps, int dir, double *seconds_out);

#ifdef __cplusplus
}
#endif
#endif /* IDEALGAN_H_ */
