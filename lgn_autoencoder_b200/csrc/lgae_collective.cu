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
constexpr int PEER_BLOCKS = 32;      // slots per phase in the signal pad = the largest grid; all blocks are co-resident on every GPU
constexpr int PEER_THREADS = 1024;
constexpr int PEER_PULL_THREADS = 512;   // the pull kernel keeps 16 double2 in flight per thread
constexpr int PEER_MC_BLOCKS = 8;    // the multicast path moves n / W doubles per rank: few blocks, few flags on the wire
struct PeerArgs {
    const double* buf[PEER_MAX];   // bucket of every rank (peer-mapped)
    uint32_t* sig[PEER_MAX];       // signal pad of every rank: [phase 0/1][block][sender rank] uint32 flags, zero when idle
    double* out;                   // local result (n)
    int32_t* err;                  // local: set to 1 if a flag wait ran into the spin limit
    int64_t n;
    int rank, world;
};

// Flag hand-shake.  Slot [phase][block][sender] of the RECEIVER's signal pad: the sender stores 1 (release, system scope), the
// receiver spins on relaxed loads until it reads 1, stores 0 back and fences (acquire).  A slot is always 0 again before its
// sender's next store: the sender only gets past the phase-1 hand-shake of a launch after the receiver has consumed (and
// reset) the phase-0 flag of that launch, and it only reaches the phase-1 store of the next launch after the receiver has
// entered that launch, i.e. finished the previous one including its phase-1 reset.  So the flags are reusable launch after
// launch (CUDA-graph replay included) without epochs or compare-and-swap round trips over NVLink.
LGAE_DEV void put_signal(uint32_t* addr) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(1u) : "memory");
}
LGAE_DEV bool wait_signal(uint32_t* addr) {
    for (long long spin = 0; spin < (1LL << 30); ++spin) {
        unsigned v;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
        if (v == 1u) {
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(0u) : "memory");
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            return true;
        }
    }
    return false;
}
// All ranks' block `blockIdx.x` meet: thread t < world signals rank t and waits for rank t's signal.
LGAE_DEV void peer_barrier(const PeerArgs& a, int phase) {
    __syncthreads();
    const int t = threadIdx.x;
    if (t < a.world && t != a.rank) {
        const size_t slot = ((size_t)phase * PEER_BLOCKS + blockIdx.x) * PEER_MAX;
        put_signal(a.sig[t] + slot + a.rank);
        if (!wait_signal(a.sig[a.rank] + slot + t)) *a.err = 1;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PEER_PULL_THREADS) peer_allreduce_kernel(const PeerArgs a) {
    const int nblk = gridDim.x;
    // every rank's bucket is final in its own stream order before this kernel starts there; after the hand-shake it is final on
    // all ranks (the release / acquire pair orders the peers' earlier writes before the reads below)
    peer_barrier(a, 0);
    const int64_t n2 = a.n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * PEER_PULL_THREADS + threadIdx.x; i < n2; i += (int64_t)nblk * PEER_PULL_THREADS) {
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

// The same exchange through the NVSwitch's multicast object (NVLS): rank r owns the r-th slice of the bucket, loads it with
// multimem.ld_reduce -- the switch adds the slices of all ranks and returns one value -- and stores the sum with multimem.st,
// which the switch writes into every rank's bucket.  In place; every rank moves n / W doubles each way instead of pulling
// (W - 1) n.  The reduction order is the switch's (fixed for a topology) and every rank receives the same broadcast bits.
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_mc_kernel(const PeerArgs a, double* mc) {
    peer_barrier(a, 0);
    const int64_t per = (a.n + a.world - 1) / a.world;
    const int64_t lo = per * a.rank, hi = lo + per < a.n ? lo + per : a.n;
    for (int64_t i = lo + (int64_t)blockIdx.x * PEER_THREADS + threadIdx.x; i < hi; i += (int64_t)gridDim.x * PEER_THREADS) {
        double v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(mc + i) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i), "d"(v) : "memory");
    }
    // every rank's slice has landed everywhere before anyone reads its bucket (or starts the next step's gradient init)
    peer_barrier(a, 1);
}

}  // namespace lgae

using namespace lgae;

extern "C" {

int64_t lgae_peer_signal_bytes(void) { return (int64_t)2 * PEER_BLOCKS * PEER_MAX * sizeof(uint32_t); }

int lgae_peer_allreduce(const double* const* bufs, uint32_t* const* signals, int32_t rank, int32_t world, int64_t n, double* out,
                        double* multicast, int32_t* err_flag, void* stream) {
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
    if (multicast) {
        if (out != bufs[rank] || ((uintptr_t)multicast & 7)) return LGAE_E_BADARG;   // the multicast path works in place
        peer_allreduce_mc_kernel<<<PEER_MC_BLOCKS, PEER_THREADS, 0, st>>>(a, multicast);
    } else {
        peer_allreduce_kernel<<<PEER_BLOCKS, PEER_PULL_THREADS, 0, st>>>(a);
    }
    return check_launch("peer_allreduce");
}

}  // extern "C"
