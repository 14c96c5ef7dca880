// Host-buffer entry point for the config-2 objective: the call bench.py times as `e2e`.
//
// The reference feeds every training step from host numpy arrays through tf.data
// (/root/reference/train-IDEAL-unsup.py:118-119,334-351).  Here a context owns three device "slots"
// (acquisitions, PM, tables, gradient, loss, scratch) each with its own stream; a batch is cut into chunks of
// `chunk_nb` samples and chunk k runs H2D -> tables -> fused kernel -> D2H on stream k % 3, so the copy
// engines (one per direction) and the SMs overlap across chunks.  No batched-memcpy API is used.
#include <new>
#include <vector>

#include "ig_common.cuh"

struct ig_ctx {
    static constexpr int kSlots = 3;
    int device = 0, chunk_nb = 0, ne = 0, nv = 0;
    cudaStream_t st[kSlots] = {};
    float *acq[kSlots] = {}, *pm[kSlots] = {}, *gpm[kSlots] = {}, *te[kSlots] = {}, *tab[kSlots] = {}, *loss[kSlots] = {};
    void *scratch[kSlots] = {};
    size_t scratch_bytes = 0;
    float *loss_h = nullptr;      // pinned, one float per chunk
    int loss_h_cap = 0;
};

extern "C" void ig_ctx_destroy(ig_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int s = 0; s < ig_ctx::kSlots; ++s) {
        if (c->st[s]) cudaStreamSynchronize(c->st[s]);
        cudaFree(c->acq[s]); cudaFree(c->pm[s]); cudaFree(c->gpm[s]); cudaFree(c->te[s]); cudaFree(c->tab[s]); cudaFree(c->loss[s]);
        cudaFree(c->scratch[s]);
        if (c->st[s]) cudaStreamDestroy(c->st[s]);
    }
    if (c->loss_h) cudaFreeHost(c->loss_h);
    delete c;
}

extern "C" int ig_ctx_create(int device, int chunk_nb, int ne, int nv, ig_ctx **out) {
    IG_REQUIRE(out && chunk_nb > 0 && nv > 0, IG_E_ARG, "ig_ctx_create: bad arguments");
    IG_REQUIRE(ne >= 2 && ne <= IG_MAX_NE, IG_E_NE, "ig_ctx_create: ne=%d outside [2, %d]", ne, IG_MAX_NE);
    ig_ctx *c = new (std::nothrow) ig_ctx;
    IG_REQUIRE(c, IG_E_ARG, "ig_ctx_create: out of host memory");
    c->device = device; c->chunk_nb = chunk_nb; c->ne = ne; c->nv = nv;
    c->scratch_bytes = ig_loss_scratch_bytes(chunk_nb, nv);
    const size_t vox = static_cast<size_t>(chunk_nb) * nv;
    cudaError_t e = cudaSetDevice(device);
    for (int s = 0; s < ig_ctx::kSlots && e == cudaSuccess; ++s) {
        if ((e = cudaStreamCreateWithFlags(&c->st[s], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->acq[s], vox * ne * 2 * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->pm[s], vox * 2 * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->gpm[s], vox * 2 * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->te[s], static_cast<size_t>(chunk_nb) * ne * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->tab[s], static_cast<size_t>(chunk_nb) * IG_TAB_FLOATS * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->loss[s], sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->scratch[s], c->scratch_bytes)) != cudaSuccess) break;
        if ((e = cudaMemset(c->scratch[s], 0, c->scratch_bytes)) != cudaSuccess) break;
    }
    if (e == cudaSuccess) {
        c->loss_h_cap = 4096;
        e = cudaHostAlloc(&c->loss_h, sizeof(float) * c->loss_h_cap, cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        ig_ctx_destroy(c);
        return ig::cuda_fail(e, "ig_ctx_create");
    }
    *out = c;
    return 0;
}

extern "C" int ig_a2a_loss_host(ig_ctx *c, const float *acqs_h, const float *pm_h, const float *te_h, int nb, float field, float r2_sc,
                                float inv_n, float *loss_h, float *g_pm_h) {
    IG_REQUIRE(c && acqs_h && pm_h && te_h && loss_h && g_pm_h && nb > 0, IG_E_ARG, "ig_a2a_loss_host: null pointer or nb <= 0");
    IG_CUDA(cudaSetDevice(c->device));
    const int nchunks = (nb + c->chunk_nb - 1) / c->chunk_nb;
    IG_REQUIRE(nchunks <= c->loss_h_cap, IG_E_ARG, "ig_a2a_loss_host: %d chunks exceed the context's %d", nchunks, c->loss_h_cap);
    const size_t nv = c->nv, ne = c->ne;
    for (int k = 0; k < nchunks; ++k) {
        const int s = k % ig_ctx::kSlots;
        const int b0 = k * c->chunk_nb;
        const int cb = (nb - b0 < c->chunk_nb) ? nb - b0 : c->chunk_nb;
        cudaStream_t st = c->st[s];
        IG_CUDA(cudaMemcpyAsync(c->te[s], te_h + static_cast<size_t>(b0) * ne, sizeof(float) * cb * ne, cudaMemcpyHostToDevice, st));
        IG_CUDA(cudaMemcpyAsync(c->acq[s], acqs_h + static_cast<size_t>(b0) * ne * nv * 2, sizeof(float) * cb * ne * nv * 2,
                                cudaMemcpyHostToDevice, st));
        IG_CUDA(cudaMemcpyAsync(c->pm[s], pm_h + static_cast<size_t>(b0) * nv * 2, sizeof(float) * cb * nv * 2, cudaMemcpyHostToDevice, st));
        if (int rc = ig_gen_tables(c->te[s], cb, c->ne, field, c->tab[s], st)) return rc;
        if (int rc = ig_a2a_loss(c->acq[s], c->pm[s], static_cast<long>(nv * 2), c->tab[s], cb, c->ne, c->nv, r2_sc, inv_n, c->gpm[s], nullptr,
                                 nullptr, c->loss[s], c->scratch[s], c->scratch_bytes, st))
            return rc;
        IG_CUDA(cudaMemcpyAsync(g_pm_h + static_cast<size_t>(b0) * nv * 2, c->gpm[s], sizeof(float) * cb * nv * 2, cudaMemcpyDeviceToHost, st));
        IG_CUDA(cudaMemcpyAsync(c->loss_h + k, c->loss[s], sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < ig_ctx::kSlots; ++s) IG_CUDA(cudaStreamSynchronize(c->st[s]));
    double acc = 0.0;
    for (int k = 0; k < nchunks; ++k) acc += static_cast<double>(c->loss_h[k]);
    loss_h[0] = static_cast<float>(acc);
    return 0;
}
