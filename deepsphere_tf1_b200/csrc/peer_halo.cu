// Halo exchange of the vertex-sharded mode over NVLink peer memory (new: the reference is single-device).  Every sparse product of
// a sharded layer is preceded by one exchange of the 1-ring ghost rows (halo.py); as NCCL send / recv pairs that is ~40 launches of
// pure latency per training step (profiles/r01_sharded_bench.jsonl: two GPUs slower than one).  Here every rank owns a mailbox that
// its peers map through CUDA IPC, and one exchange is ONE kernel:
//   (1) push: every boundary row a neighbour needs is copied straight into that neighbour's mailbox (remote stores), at the place
//       where the row sits in the neighbour's ghost block (ghosts are grouped by owner, ascending vertex id);
//   (2) the last CTA to finish publishes the call's sequence number in its flag on every neighbour;
//   (3) wait for the flags of the ranks this rank receives from, then copy the mailbox into the ghost rows [Ml, Ml + H).
// The sequence number lives in device memory and is advanced by the kernel, so a captured CUDA graph replays the exchange; mailbox
// halves and flags alternate with its parity (a rank cannot be two exchanges ahead of a neighbour: it needs the neighbour's flag of
// the previous one first).  All CTAs are co-resident (grid <= number of SMs), waits are bounded and trap instead of hanging.
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace dsb200 {

constexpr int kHaloMaxWorld = 8;

struct HaloHeader {
    unsigned long long flags[2][kHaloMaxWorld];
    unsigned long long seq;             // exchanges issued so far by the owner
    unsigned int arrived;               // CTAs of the running exchange that finished pushing
    unsigned int pad;
    unsigned long long capacity;        // floats per mailbox half
};

struct HaloPtrs { unsigned char* p[kHaloMaxWorld]; };

__device__ __forceinline__ float* mailbox(unsigned char* comm, int half, unsigned long long cap)
{
    return reinterpret_cast<float*>(comm + 256) + (size_t)half * cap;
}

// meta: [0] = number of sends, [1] = number of receives, then per send (peer, rows, first ghost row at the peer, first entry in idx),
// then per receive (peer)
template <typename VT> struct VecOf { static constexpr int n = sizeof(VT) / sizeof(float); };

// VT = float4 when F % 4 == 0 (16-byte pieces), float otherwise (the single-feature input layer)
template <typename VT>
__global__ void __launch_bounds__(256)
peer_halo_kernel(HaloPtrs peers, int rank, float* rows, int F, int Ml, int H, const int* __restrict__ meta,
                 const int* __restrict__ idx)
{
    unsigned char* mine = peers.p[rank];
    HaloHeader* hdr = reinterpret_cast<HaloHeader*>(mine);
    __shared__ unsigned long long s_seq;
    __shared__ int s_last;
    if (threadIdx.x == 0) s_seq = hdr->seq + 1;          // every CTA reads the same value: it is advanced only by the last one
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int half = (int)(seq & 1ull);
    const int nsend = meta[0], nrecv = meta[1];
    const int F4 = F / VecOf<VT>::n;
    // (1) push
    for (int s = 0; s < nsend; ++s) {
        const int peer = meta[2 + 4 * s], n = meta[3 + 4 * s], off = meta[4 + 4 * s], first = meta[5 + 4 * s];
        HaloHeader* ph = reinterpret_cast<HaloHeader*>(peers.p[peer]);
        VT* dst = reinterpret_cast<VT*>(mailbox(peers.p[peer], half, ph->capacity) + (size_t)off * F);
        const long long total = (long long)n * F4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / F4), c = (int)(i - (long long)r * F4);
            dst[i] = reinterpret_cast<const VT*>(rows + (size_t)__ldg(idx + first + r) * F)[c];
        }
    }
    __threadfence_system();
    __syncthreads();
    // (2) the last CTA publishes
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&hdr->arrived, 1u);
        s_last = prev == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < nsend) {
            const int peer = meta[2 + 4 * threadIdx.x];
            volatile unsigned long long* f = &reinterpret_cast<HaloHeader*>(peers.p[peer])->flags[half][rank];
            *f = seq;
        }
        if (threadIdx.x == 0) { hdr->arrived = 0; hdr->seq = seq; __threadfence(); }
    }
    // (3) wait for the senders, then unpack
    if (threadIdx.x < nrecv) {
        const int peer = meta[2 + 4 * nsend + threadIdx.x];
        volatile unsigned long long* f = &hdr->flags[half][peer];
        unsigned long long spins = 0;
        while (*f < seq) {
            if (++spins > 4000000000ull) {
                printf("dsb200 peer halo exchange: rank %d timed out waiting for rank %d (exchange %llu)\n", rank, peer, seq);
                __trap();
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    const VT* src = reinterpret_cast<const VT*>(mailbox(mine, half, hdr->capacity));
    VT* ghost = reinterpret_cast<VT*>(rows + (size_t)Ml * F);
    const long long total = (long long)H * F4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        ghost[i] = __ldcv(src + i);                       // written by the peers: bypass any stale cached line
}

}  // namespace dsb200

using namespace dsb200;

/* allocate this rank's mailbox (cudaMalloc, zeroed; `capacity` floats per half) and export its 64-byte IPC handle */
extern "C" int dsb200_peer_halo_create(long long capacity, void** comm_out, void* handle_out64)
{
    DSB_REQUIRE(comm_out && handle_out64 && capacity > 0, "peer_halo_create: bad arguments");
    void* p = nullptr;
    const size_t bytes = 256 + 2 * (size_t)capacity * sizeof(float);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) {
        HaloHeader h;
        memset(&h, 0, sizeof(h));
        h.capacity = (unsigned long long)capacity;
        e = cudaMemcpy(p, &h, sizeof(h), cudaMemcpyHostToDevice);
    }
    cudaIpcMemHandle_t ih;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ih, p);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
    static_assert(sizeof(HaloHeader) <= 256, "header");
    memcpy(handle_out64, &ih, 64);
    *comm_out = p;
    return 0;
}

/* rows [Ml + H][F] (device, F % 4 == 0, 16-byte aligned): send the boundary rows listed in (meta, idx) to the neighbours' mailboxes,
 * receive this rank's H ghost rows into rows[Ml ..].  comms[r] = rank r's mailbox as mapped in THIS process (HOST array). */
extern "C" int dsb200_peer_halo_exchange(void* const* comms, int rank, int world, float* rows, int F, int Ml, int H,
                                         const int32_t* meta, const int32_t* idx, int max_rows, void* stream)
{
    DSB_REQUIRE(comms && rows && meta && F > 0 && Ml > 0 && H >= 0, "peer_halo_exchange: bad arguments");
    DSB_REQUIRE(world >= 1 && world <= kHaloMaxWorld && rank >= 0 && rank < world, "peer_halo_exchange: bad rank / world");
    const bool vec = F % 4 == 0 && (((uintptr_t)rows) & 15) == 0;
    HaloPtrs pp;
    for (int r = 0; r < kHaloMaxWorld; ++r) pp.p[r] = r < world ? (unsigned char*)comms[r] : nullptr;
    const long long work = (long long)(max_rows > H ? max_rows : H) * (vec ? F / 4 : F);
    int grid = (int)((work + 255) / 256);
    if (grid > kNumSMs) grid = kNumSMs;                   // all CTAs co-resident: the waits cannot starve the pushes
    if (grid < 1) grid = 1;
    if (vec) peer_halo_kernel<float4><<<grid, 256, 0, (cudaStream_t)stream>>>(pp, rank, rows, F, Ml, H, meta, idx);
    else peer_halo_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(pp, rank, rows, F, Ml, H, meta, idx);
    DSB_LAUNCH_CHECK();
}
