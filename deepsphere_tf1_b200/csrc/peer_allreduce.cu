// One-shot all-reduce of the small SyncBN vectors over NVLink peer memory (new: the reference is single-device).
// A data-parallel step needs ~13 all-reduces of at most a few hundred doubles (batch-norm sums forward and backward,
// the Gram vector of the fused input layer, the loss); each costs ~15 us as an NCCL launch inside the step's CUDA graph,
// all of it latency.  Here every rank owns a communication buffer that all peers map through CUDA IPC:
//   slots[2][world][kMaxN] doubles  +  flags[2][world] 64-bit sequence numbers
// and one call is ONE single-CTA kernel: (1) push my vector into my slot of EVERY rank's buffer (remote stores),
// __threadfence_system, (2) publish the call's sequence number in my flag on every rank, (3) wait until all flags of my
// own buffer carry the sequence number, (4) add the slots in rank order (bitwise identical result on every rank).
// The sequence number lives in device memory and is advanced by the kernel itself, so a captured CUDA graph can replay
// the call; slots and flags are double buffered by its parity (a rank cannot be two calls ahead of a peer: it needs the
// peer's flag of the previous call first).  Waits are bounded: a protocol failure traps instead of hanging the GPU.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace dsb200 {

constexpr int kPeerMaxN = 1024;             // doubles per call
constexpr int kPeerMaxWorld = 8;

struct PeerComm {
    double slots[2][kPeerMaxWorld][kPeerMaxN];
    unsigned long long flags[2][kPeerMaxWorld];
    unsigned long long seq;                 // calls issued so far by the owner of this buffer
};

struct PeerPtrs { PeerComm* p[kPeerMaxWorld]; };

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerPtrs peers, int rank, int world, double* data, int n)
{
    PeerComm* mine = peers.p[rank];
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) s_seq = ++mine->seq;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull);
    // (1) push
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = data[i];
        for (int r = 0; r < world; ++r) peers.p[r]->slots[par][rank][i] = v;
    }
    __threadfence_system();
    __syncthreads();
    // (2) publish
    if (threadIdx.x < world) {
        volatile unsigned long long* f = &peers.p[threadIdx.x]->flags[par][rank];
        *f = seq;
    }
    // (3) wait for every rank's contribution to MY buffer
    if (threadIdx.x < world) {
        volatile unsigned long long* f = &mine->flags[par][threadIdx.x];
        unsigned long long spins = 0;
        while (*f < seq) {
            if (++spins > 2000000000ull) {
                printf("dsb200 peer all-reduce: rank %d timed out waiting for rank %d (call %llu)\n", rank, (int)threadIdx.x, seq);
                __trap();
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    // (4) reduce in rank order
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += *(volatile double*)&mine->slots[par][r][i];
        data[i] = acc;
    }
}

}  // namespace dsb200

using namespace dsb200;

extern "C" long long dsb200_peer_comm_bytes(void) { return (long long)sizeof(PeerComm); }

/* allocate this rank's communication buffer (cudaMalloc, zeroed) and export its 64-byte IPC handle */
extern "C" int dsb200_peer_comm_create(void** comm_out, void* handle_out64)
{
    DSB_REQUIRE(comm_out && handle_out64, "peer_comm_create: null pointer");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(PeerComm));
    if (e == cudaSuccess) e = cudaMemset(p, 0, sizeof(PeerComm));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out64, &h, 64);
    *comm_out = p;
    return 0;
}

/* map a peer's buffer from its IPC handle */
extern "C" int dsb200_peer_comm_open(const void* handle64, void** comm_out)
{
    DSB_REQUIRE(handle64 && comm_out, "peer_comm_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
    *comm_out = p;
    return 0;
}

/* data[0..n) (doubles, device) <- sum over the ranks; comms[r] = rank r's buffer as mapped in THIS process */
extern "C" int dsb200_peer_allreduce_f64(void* const* comms, int rank, int world, double* data, int n, void* stream)
{
    DSB_REQUIRE(comms && data && n > 0 && n <= kPeerMaxN, "peer_allreduce: at most 1024 doubles per call");
    DSB_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer_allreduce: bad rank / world");
    PeerPtrs pp;
    for (int r = 0; r < kPeerMaxWorld; ++r) pp.p[r] = r < world ? (PeerComm*)comms[r] : nullptr;
    peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pp, rank, world, data, n);
    DSB_LAUNCH_CHECK();
}
