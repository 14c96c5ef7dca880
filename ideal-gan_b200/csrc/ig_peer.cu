// Scalar exchange between the ranks of a data-parallel job over peer memory (one process per GPU, NVLink / NVSwitch).
// The device side lives in ig_common.cuh (peer_exchange, called by the finishing thread of the objective's kernel); this
// file owns the mailboxes: allocation, CUDA-IPC export / import, the same-process variant for threads and tests, and the
// stand-alone reduction of one step.
#include <vector>

#include "ig_common.cuh"

struct ig_peer {
    int device = 0, rank = 0, world = 1;
    unsigned long long *box = nullptr;                 // own mailbox: kPeerSlots x world words
    unsigned long long **boxes_d = nullptr;            // device array of all ranks' mailbox pointers
    std::vector<unsigned long long *> boxes_h;         // host copy; entries opened through IPC are closed on destroy
    std::vector<bool> ipc;
    bool connected = false;
};

namespace ig {

int peer_pub(const ig_peer *peer, unsigned step, int lag, float *prev_out, PeerPub *out) {
    IG_REQUIRE(peer->connected, IG_E_ARG, "ig_peer: not connected (ig_peer_connect / ig_peer_connect_local first)");
    IG_REQUIRE(lag >= 1 && lag <= kPeerMaxLag, IG_E_ARG, "ig_peer: lag %d outside [1, %d]", lag, kPeerMaxLag);
    out->boxes = peer->boxes_d;
    out->prev_out = prev_out;
    out->rank = peer->rank;
    out->world = peer->world;
    out->step = step;
    out->lag = static_cast<unsigned>(lag);
    return 0;
}

__global__ void peer_reduce_kernel(const unsigned long long *box, int world, unsigned step, float *out) {
    if (threadIdx.x == 0) out[0] = peer_collect(box, world, step);
}

// stand-alone publication for the objectives whose kernels do not carry the hook (UQ, Rician, forward-model losses): one thread
// after the loss kernel on the same stream
__global__ void peer_publish_kernel(PeerPub peer, const float *value) {
    if (threadIdx.x == 0) peer_exchange(peer, value[0]);
}

static int finish_connect(ig_peer *p) {
    IG_CUDA(cudaMemcpy(p->boxes_d, p->boxes_h.data(), sizeof(unsigned long long *) * p->world, cudaMemcpyHostToDevice));
    p->connected = true;
    return 0;
}

}  // namespace ig

using namespace ig;

extern "C" int ig_peer_create(int rank, int world, ig_peer **out) {
    IG_REQUIRE(out && world >= 1 && world <= 64 && rank >= 0 && rank < world, IG_E_ARG, "ig_peer_create: rank %d of %d", rank, world);
    ig_peer *p = new ig_peer();
    p->rank = rank;
    p->world = world;
    p->boxes_h.assign(world, nullptr);
    p->ipc.assign(world, false);
    cudaError_t e = cudaGetDevice(&p->device);
    const size_t bytes = sizeof(unsigned long long) * kPeerSlots * world;
    if (e == cudaSuccess) e = cudaMalloc(&p->box, bytes);
    if (e == cudaSuccess) e = cudaMemset(p->box, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->boxes_d, sizeof(unsigned long long *) * world);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p->box);
        cudaFree(p->boxes_d);
        delete p;
        return cuda_fail(e, "ig_peer_create");
    }
    p->boxes_h[rank] = p->box;
    if (world == 1) {
        if (int rc = finish_connect(p)) return rc;
    }
    *out = p;
    return 0;
}

extern "C" int ig_peer_handle(ig_peer *p, void *handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == IG_PEER_HANDLE_BYTES, "IPC handle size");
    IG_REQUIRE(p && handle_out, IG_E_ARG, "ig_peer_handle: null pointer");
    cudaIpcMemHandle_t h;
    IG_CUDA(cudaIpcGetMemHandle(&h, p->box));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int ig_peer_connect(ig_peer *p, const void *handles) {
    IG_REQUIRE(p && handles, IG_E_ARG, "ig_peer_connect: null pointer");
    IG_REQUIRE(!p->connected || p->world == 1, IG_E_ARG, "ig_peer_connect: already connected");
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + static_cast<size_t>(r) * IG_PEER_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        IG_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->boxes_h[r] = static_cast<unsigned long long *>(ptr);
        p->ipc[r] = true;
    }
    return finish_connect(p);
}

extern "C" int ig_peer_connect_local(ig_peer *const *peers, int world) {
    IG_REQUIRE(peers && world >= 1, IG_E_ARG, "ig_peer_connect_local: null pointer");
    for (int r = 0; r < world; ++r)
        IG_REQUIRE(peers[r] && peers[r]->world == world && peers[r]->rank == r, IG_E_ARG, "ig_peer_connect_local: context %d is not rank %d of %d", r, r, world);
    int cur = 0;
    IG_CUDA(cudaGetDevice(&cur));
    for (int r = 0; r < world; ++r) {
        ig_peer *p = peers[r];
        IG_CUDA(cudaSetDevice(p->device));
        for (int q = 0; q < world; ++q) {
            p->boxes_h[q] = peers[q]->box;
            if (peers[q]->device != p->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(peers[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaSetDevice(cur);
                    return cuda_fail(e, "cudaDeviceEnablePeerAccess");
                }
                (void)cudaGetLastError();
            }
        }
        if (int rc = finish_connect(p)) {
            cudaSetDevice(cur);
            return rc;
        }
    }
    IG_CUDA(cudaSetDevice(cur));
    return 0;
}

extern "C" int ig_peer_reduce(ig_peer *p, unsigned step, float *loss_d, void *stream) {
    IG_REQUIRE(p && loss_d, IG_E_ARG, "ig_peer_reduce: null pointer");
    IG_REQUIRE(p->connected, IG_E_ARG, "ig_peer_reduce: not connected");
    peer_reduce_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p->box, p->world, step, loss_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ig_peer_publish(ig_peer *p, unsigned step, int lag, const float *loss_d, float *loss_prev_d, void *stream) {
    IG_REQUIRE(p && loss_d, IG_E_ARG, "ig_peer_publish: null pointer");
    PeerPub pub{};
    if (int rc = peer_pub(p, step, lag, loss_prev_d, &pub)) return rc;
    peer_publish_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pub, loss_d);
    IG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" void ig_peer_destroy(ig_peer *p) {
    if (!p) return;
    for (int r = 0; r < p->world; ++r)
        if (p->ipc[r] && p->boxes_h[r]) cudaIpcCloseMemHandle(p->boxes_h[r]);
    cudaFree(p->box);
    cudaFree(p->boxes_d);
    delete p;
}
