// Data-parallel gradient exchange over NVLink / NVSwitch peer memory: a one-shot all-reduce(SUM) of
// the flat fp64 gradient bucket written by the training step.  Every rank's bucket lives in memory mapped into all ranks of
// the node (symmetric memory; the host side gets the peer pointers from torch.distributed's rendezvous -- plumbing only); the
// kernel synchronises the ranks with flags in the peers' signal pads, then every rank pulls all peers' buckets and adds them
// in rank order 0..W-1 -- a fixed order, identical on every rank, so the replicas stay bit-identical (what NCCL's ring would
// only guarantee per topology).  At 508 KB the exchange is latency-bound: two flag round trips plus one pull, instead of the
// multi-step ring / tree protocol of ncclAllReduce.
#include "lgae_common.cuh"

namespace lgae {

constexpr int PEER_MAX = 16;
constexpr int PEER_BLOCKS = 32;      // co-resident on every GPU of the node (the barrier needs all blocks running)
constexpr int PEER_THREADS = 512;
struct PeerArgs {
    const double* buf[PEER_MAX];   // bucket of every rank (peer-mapped)
    uint32_t* sig[PEER_MAX];       // signal pad of every rank: [phase 0/1][block][sender rank] uint32 flags, zero when idle
    double* out;                   // local result (n)
    int32_t* err;                  // local: set to 1 if a flag wait ran into the spin limit
    int64_t n;
    int rank, world;
};

// Flag hand-shake (the CAS protocol of PyTorch's symmetric-memory barrier): the sender flips the receiver's slot 0 -> 1, the
// receiver flips it back 1 -> 0, so the slots are reusable launch after launch (CUDA-graph replay included) without epochs.
LGAE_DEV bool put_signal(uint32_t* addr) {
    for (long long spin = 0; spin < (1LL << 28); ++spin) {
        unsigned old;
        asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
        if (old == 0u) return true;
    }
    return false;
}
LGAE_DEV bool wait_signal(uint32_t* addr) {
    for (long long spin = 0; spin < (1LL << 28); ++spin) {
        unsigned old;
        asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
        if (old == 1u) return true;
    }
    return false;
}
// All ranks' block `blockIdx.x` meet: thread t < world signals rank t and waits for rank t's signal.
LGAE_DEV void peer_barrier(const PeerArgs& a, int phase) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < a.world && t != a.rank) {
        const size_t slot = ((size_t)phase * PEER_BLOCKS + blockIdx.x) * PEER_MAX;
        bool ok = put_signal(a.sig[t] + slot + a.rank);
        ok = wait_signal(a.sig[a.rank] + slot + t) && ok;
        if (!ok) *a.err = 1;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(const PeerArgs a) {
    // every rank's bucket is final in its own stream order before this kernel starts there; after the hand-shake it is final on
    // all ranks (the release / acquire pair orders the peers' earlier writes before the reads below)
    __threadfence_system();
    peer_barrier(a, 0);
    const int64_t n2 = a.n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * PEER_THREADS + threadIdx.x; i < n2; i += (int64_t)PEER_BLOCKS * PEER_THREADS) {
        double2 v[PEER_MAX];
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < a.world) v[r] = __ldcg(reinterpret_cast<const double2*>(a.buf[r]) + i);   // all pulls in flight, L1 bypassed
        double2 s = v[0];
#pragma unroll
        for (int r = 1; r < PEER_MAX; ++r)
            if (r < a.world) { s.x += v[r].x; s.y += v[r].y; }   // rank order: the same sum on every rank
        reinterpret_cast<double2*>(a.out)[i] = s;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) s += __ldcg(a.buf[r] + a.n - 1);
        a.out[a.n - 1] = s;
    }
    // nobody may overwrite its bucket (the next step's gradient init) before every peer has finished reading it
    peer_barrier(a, 1);
}

}  // namespace lgae

using namespace lgae;

extern "C" {

int64_t lgae_peer_signal_bytes(void) { return (int64_t)2 * PEER_BLOCKS * PEER_MAX * sizeof(uint32_t); }

int lgae_peer_allreduce(const double* const* bufs, uint32_t* const* signals, int32_t rank, int32_t world, int64_t n, double* out,
                        int32_t* err_flag, void* stream) {
    if (!bufs || !signals || !out || !err_flag || world < 1 || world > PEER_MAX || rank < 0 || rank >= world || n < 0) return LGAE_E_BADARG;
    if (n == 0) return LGAE_OK;
    PeerArgs a = {};
    for (int r = 0; r < world; ++r) {
        if (!bufs[r] || !signals[r] || ((uintptr_t)bufs[r] & 15)) return LGAE_E_BADARG;
        a.buf[r] = bufs[r];
        a.sig[r] = signals[r];
    }
    if ((uintptr_t)out & 15) return LGAE_E_BADARG;
    a.out = out; a.err = err_flag; a.n = n; a.rank = rank; a.world = world;
    cudaStream_t st = (cudaStream_t)stream;
    LaunchScope ls_("peer_allreduce", st);
    // launched WITHOUT the programmatic-serialisation attribute: the kernel must not start before the bucket is complete
    peer_allreduce_kernel<<<PEER_BLOCKS, PEER_THREADS, 0, st>>>(a);
    return check_launch("peer_allreduce");
}

}  // extern "C"
