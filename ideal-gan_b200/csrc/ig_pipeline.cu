// Host-buffer entry point for the config-2 objective: the call bench.py times as `e2e`.
//
// The reference feeds every training step from host numpy arrays through tf.data
// (/root/reference/train-IDEAL-unsup.py:118-119,334-351).  Here a context owns three device "slots"
// (acquisitions, PM, tables, gradient, loss, scratch) each with its own stream; a batch is cut into chunks of
// `chunk_nb` samples and chunk k runs H2D -> tables -> fused kernel -> D2H on stream k % 3, so the copy
// engines (one per direction) and the SMs overlap across chunks.  No batched-memcpy API is used.
#include <stdio.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

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
    auto enqueue = [&]() -> int {
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
        return 0;
    };
    const int rc = enqueue();
    // Success or not, nothing may still be reading or writing the caller's buffers when this call returns.
    cudaError_t sync_err = cudaSuccess;
    for (int s = 0; s < ig_ctx::kSlots; ++s) {
        const cudaError_t e = cudaStreamSynchronize(c->st[s]);
        if (e != cudaSuccess && sync_err == cudaSuccess) sync_err = e;
    }
    if (rc != 0 || sync_err != cudaSuccess) {
        // a chunk that never ran its kernel to the end can leave the loss scratch (ticket + partials) armed: re-zero it
        for (int s = 0; s < ig_ctx::kSlots; ++s) cudaMemset(c->scratch[s], 0, c->scratch_bytes);
        cudaDeviceSynchronize();
        return rc != 0 ? rc : ig::cuda_fail(sync_err, "ig_a2a_loss_host");
    }
    double acc = 0.0;
    for (int k = 0; k < nchunks; ++k) acc += static_cast<double>(c->loss_h[k]);
    loss_h[0] = static_cast<float>(acc);
    return 0;
}

// ---- config 5: streamed physics decoding, host maps in / host images (and signals) out -------------------------------------
// gen_LDM_dataset.py:140-254 decodes 16 384 slices; the complex signals alone are 116 GB, so a rank's shard is streamed: three
// slots (stream + device staging for maps, tables and every output), chunk k runs H2D -> ig_gen_tables -> ig_ideal_decode -> D2H
// on slot k % 3.  Staging is allocated once per context: a streamed pass issues no allocation (a first version built on a
// framework allocator paid a synchronising cudaMalloc per chunk and ran at 5 GB/s instead of the link's 55).
struct ig_decode_ctx {
    static constexpr int kSlots = 3;
    int device = 0, model = 0, roc = 0, chunk_nb = 0, ne = 0, nv = 0;
    size_t map_floats = 0;        // per sample
    cudaStream_t st[kSlots] = {};
    float *maps[kSlots] = {}, *te[kSlots] = {}, *tab[kSlots] = {}, *shat[kSlots] = {}, *mag[kSlots] = {}, *pdff[kSlots] = {}, *r2s[kSlots] = {};
};

extern "C" void ig_decode_ctx_destroy(ig_decode_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int s = 0; s < ig_decode_ctx::kSlots; ++s) {
        if (c->st[s]) cudaStreamSynchronize(c->st[s]);
        cudaFree(c->maps[s]); cudaFree(c->te[s]); cudaFree(c->tab[s]); cudaFree(c->shat[s]); cudaFree(c->mag[s]); cudaFree(c->pdff[s]);
        cudaFree(c->r2s[s]);
        if (c->st[s]) cudaStreamDestroy(c->st[s]);
    }
    delete c;
}

extern "C" int ig_decode_ctx_create(int device, int model, int rows_or_ch, int chunk_nb, int ne, int nv, int want_shat, ig_decode_ctx **out) {
    IG_REQUIRE(out && chunk_nb > 0 && nv > 0, IG_E_ARG, "ig_decode_ctx_create: bad arguments");
    IG_REQUIRE(ne >= 1 && ne <= IG_MAX_NE, IG_E_NE, "ig_decode_ctx_create: ne=%d outside [1, %d]", ne, IG_MAX_NE);
    IG_REQUIRE(model >= IG_MODEL_WFPM && model <= IG_MODEL_MAGPHA && rows_or_ch >= 3, IG_E_ARG, "ig_decode_ctx_create: model %d rows/channels %d", model,
               rows_or_ch);
    ig_decode_ctx *c = new (std::nothrow) ig_decode_ctx;
    IG_REQUIRE(c, IG_E_ARG, "ig_decode_ctx_create: out of host memory");
    c->device = device; c->model = model; c->roc = rows_or_ch; c->chunk_nb = chunk_nb; c->ne = ne; c->nv = nv;
    c->map_floats = static_cast<size_t>(model == IG_MODEL_MAGPHA ? 2 : rows_or_ch) * nv * (model == IG_MODEL_MAGPHA ? rows_or_ch : 2);
    const size_t vox = static_cast<size_t>(chunk_nb) * nv;
    cudaError_t e = cudaSetDevice(device);
    for (int s = 0; s < ig_decode_ctx::kSlots && e == cudaSuccess; ++s) {
        if ((e = cudaStreamCreateWithFlags(&c->st[s], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->maps[s], c->map_floats * chunk_nb * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->te[s], static_cast<size_t>(chunk_nb) * ne * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->tab[s], static_cast<size_t>(chunk_nb) * IG_TAB_FLOATS * sizeof(float))) != cudaSuccess) break;
        if (want_shat && (e = cudaMalloc(&c->shat[s], vox * ne * 2 * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->mag[s], vox * ne * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->pdff[s], vox * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMalloc(&c->r2s[s], vox * sizeof(float))) != cudaSuccess) break;
    }
    if (e != cudaSuccess) {
        ig_decode_ctx_destroy(c);
        return ig::cuda_fail(e, "ig_decode_ctx_create");
    }
    *out = c;
    return 0;
}

extern "C" int ig_decode_host(ig_decode_ctx *c, const float *maps_h, const float *te_h, int nb, float field, float r2_sc, int flags, float *shat_h,
                              float *mag_h, float *pdff_h, float *r2s_h) {
    IG_REQUIRE(c && maps_h && te_h && nb > 0, IG_E_ARG, "ig_decode_host: null pointer or nb <= 0");
    IG_REQUIRE(shat_h || mag_h || pdff_h || r2s_h, IG_E_ARG, "ig_decode_host: no output requested");
    IG_REQUIRE(!shat_h || c->shat[0], IG_E_ARG, "ig_decode_host: the context was created without staging for the complex signals");
    IG_CUDA(cudaSetDevice(c->device));
    const size_t nv = c->nv, ne = c->ne;
    auto enqueue = [&]() -> int {
        for (int k = 0, b0 = 0; b0 < nb; ++k, b0 += c->chunk_nb) {
            const int s = k % ig_decode_ctx::kSlots;
            const size_t cb = (nb - b0 < c->chunk_nb) ? nb - b0 : c->chunk_nb;
            cudaStream_t st = c->st[s];
            IG_CUDA(cudaMemcpyAsync(c->te[s], te_h + static_cast<size_t>(b0) * ne, sizeof(float) * cb * ne, cudaMemcpyHostToDevice, st));
            IG_CUDA(cudaMemcpyAsync(c->maps[s], maps_h + static_cast<size_t>(b0) * c->map_floats, sizeof(float) * cb * c->map_floats, cudaMemcpyHostToDevice, st));
            if (int rc = ig_gen_tables(c->te[s], static_cast<int>(cb), c->ne, field, c->tab[s], st)) return rc;
            if (int rc = ig_ideal_decode(c->model, c->maps[s], c->roc, c->tab[s], static_cast<int>(cb), c->ne, c->nv, r2_sc, flags, shat_h ? c->shat[s] : nullptr,
                                         mag_h ? c->mag[s] : nullptr, pdff_h ? c->pdff[s] : nullptr, r2s_h ? c->r2s[s] : nullptr, st))
                return rc;
            if (shat_h) IG_CUDA(cudaMemcpyAsync(shat_h + static_cast<size_t>(b0) * ne * nv * 2, c->shat[s], sizeof(float) * cb * ne * nv * 2, cudaMemcpyDeviceToHost, st));
            if (mag_h) IG_CUDA(cudaMemcpyAsync(mag_h + static_cast<size_t>(b0) * ne * nv, c->mag[s], sizeof(float) * cb * ne * nv, cudaMemcpyDeviceToHost, st));
            if (pdff_h) IG_CUDA(cudaMemcpyAsync(pdff_h + static_cast<size_t>(b0) * nv, c->pdff[s], sizeof(float) * cb * nv, cudaMemcpyDeviceToHost, st));
            if (r2s_h) IG_CUDA(cudaMemcpyAsync(r2s_h + static_cast<size_t>(b0) * nv, c->r2s[s], sizeof(float) * cb * nv, cudaMemcpyDeviceToHost, st));
        }
        return 0;
    };
    const int rc = enqueue();
    cudaError_t sync_err = cudaSuccess;                  // success or not: the caller's buffers are quiet when this returns
    for (int s = 0; s < ig_decode_ctx::kSlots; ++s) {
        const cudaError_t e = cudaStreamSynchronize(c->st[s]);
        if (e != cudaSuccess && sync_err == cudaSuccess) sync_err = e;
    }
    if (rc != 0) return rc;
    if (sync_err != cudaSuccess) return ig::cuda_fail(sync_err, "ig_decode_host");
    return 0;
}

// ---- pinned host buffers next to the GPU ---------------------------------------------------------------------------------
// The e2e leg is bound by host -> device staging; on multi-socket hosts a pinned buffer on the far socket halves it.  The
// buffer is allocated while the calling thread's memory policy is bound to the GPU's NUMA node (sysfs numa_node of its PCI
// function; raw set_mempolicy syscall, no libnuma), then the policy is restored.  A host with one node (or a container that
// hides the topology: numa_node = -1) gets a plain cudaHostAlloc.
namespace {
int device_numa_node(int device) {
    char bus[32] = {};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char *p = bus; *p; ++p)
        if (*p >= 'A' && *p <= 'Z') *p = static_cast<char>(*p - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
constexpr int kMpolDefault = 0, kMpolPreferred = 1;
long set_policy(int mode, const unsigned long *mask, unsigned long maxnode) {
#ifdef SYS_set_mempolicy
    return syscall(SYS_set_mempolicy, mode, mask, maxnode);
#else
    return -1;
#endif
}
}  // namespace

extern "C" int ig_host_numa_node(int device) { return device_numa_node(device); }

extern "C" int ig_host_alloc(size_t bytes, int device, int flags, void **out) {
    IG_REQUIRE(out && bytes > 0, IG_E_ARG, "ig_host_alloc: bad arguments");
    IG_CUDA(cudaSetDevice(device));
    const int node = device_numa_node(device);
    bool bound = false;
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = {};
        mask[node / (8 * sizeof(unsigned long))] = 1ul << (node % (8 * sizeof(unsigned long)));
        bound = set_policy(kMpolPreferred, mask, 1024) == 0;
    }
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, bytes, (flags & IG_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e == cudaSuccess && !(flags & IG_HOST_WRITE_COMBINED)) memset(p, 0, bytes);      // first touch under the policy
    if (bound) set_policy(kMpolDefault, nullptr, 0);
    if (e != cudaSuccess) return ig::cuda_fail(e, "ig_host_alloc");
    *out = p;
    return 0;
}

extern "C" int ig_host_free(void *p) {
    if (p) IG_CUDA(cudaFreeHost(p));
    return 0;
}

// Bare cudaMemcpyAsync loop between a host and a device buffer: the ceiling the host-buffer pipeline can reach on this host
// (bench.py prints it per rank and summed over ranks next to the e2e figure).  dir 0: host -> device, 1: device -> host,
// 2: both at once on two streams.  seconds_out = wall time of `reps` copies of `bytes` (per direction).
extern "C" int ig_copy_probe(void *host, void *dev, void *host2, void *dev2, size_t bytes, int reps, int dir, double *seconds_out) {
    IG_REQUIRE(host && dev && seconds_out && bytes > 0 && reps > 0 && dir >= 0 && dir <= 2, IG_E_ARG, "ig_copy_probe: bad arguments");
    IG_REQUIRE(dir != 2 || (host2 && dev2), IG_E_ARG, "ig_copy_probe: the bidirectional probe needs the second buffer pair");
    cudaStream_t st[2] = {};
    IG_CUDA(cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking));
    IG_CUDA(cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    auto run = [&](int n) {
        for (int i = 0; i < n && e == cudaSuccess; ++i) {
            if (dir == 0 || dir == 2) e = cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st[0]);
            if (e == cudaSuccess && dir == 1) e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st[0]);
            if (e == cudaSuccess && dir == 2) e = cudaMemcpyAsync(host2, dev2, bytes, cudaMemcpyDeviceToHost, st[1]);
        }
    };
    run(2);
    cudaStreamSynchronize(st[0]);
    cudaStreamSynchronize(st[1]);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    run(reps);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st[0]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st[1]);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    cudaStreamDestroy(st[0]); cudaStreamDestroy(st[1]);
    if (e != cudaSuccess) return ig::cuda_fail(e, "ig_copy_probe");
    *seconds_out = static_cast<double>(t1.tv_sec - t0.tv_sec) + 1e-9 * static_cast<double>(t1.tv_nsec - t0.tv_nsec);
    return 0;
}
