// envforge_b200 -- episode statistics off the critical path: the replica fold and its fusion with the cross-GPU sum.
//
// The step / rollout kernels add finished-episode statistics into EF_STATS_REPLICAS copies of the vector per env set
// (ef_common.cuh::cta_epilogue).  Reading them used to be a handful of eager tensor ops per env set; here it is ONE tiny
// kernel over all env sets of a rank (graph-capturable), and -- for more than one GPU -- the same kernel also performs the
// engine's only collective, the sum of the 64-byte vector over the ranks, through NVLink peer memory:
//     fold local replicas -> store the vector into every peer's mailbox slot (P2P stores + system-scope release)
//     -> wait for the peers' slots of this sequence number (acquire loads on LOCAL memory) -> add in rank order.
// The env instances are independent (grid_world.py:495-528, cart_pole.py:107-180 touch only `self`), so this 64-byte
// exchange is all that ever crosses NVLink; it is latency-bound, which is why it is one kernel and not a library call.
#include <cstring>
#include <new>

#include "ef_common.cuh"

namespace ef {

constexpr int kMaxPeers = 16;

struct FoldArgs {
    const double* raw[EF_STATS_MAX_SETS];
    int32_t n_sets;
};

// 128 bytes: one rank's vector for one sequence number
struct alignas(128) PeerSlot {
    double v[EF_STATS_LEN];
    unsigned long long seq;
    unsigned long long pad[7];
};

struct PeerArgs {
    PeerSlot* mailbox[kMaxPeers];  // mailbox[r] = rank r's buffer: [2][world] slots, then the sequence counter
    int32_t rank, world;
};

// thread (replica = tid >> 3, field = tid & 7) walks the env sets; the 32 replica partials of a field are then added
// in replica order by one thread: a fixed order of additions, so the result is bit-reproducible.
__device__ __forceinline__ void fold_to_shared(const FoldArgs& f, double (&s_part)[EF_STATS_REPLICAS][EF_STATS_LEN],
                                               double (&s_out)[EF_STATS_LEN]) {
    const int rep = threadIdx.x >> 3, field = threadIdx.x & 7;
    static_assert(EF_STATS_REPLICAS * EF_STATS_LEN == kThreads, "one thread per (replica, field)");
    double acc = 0.0;
    // eight env sets' loads in flight at a time (their lines are as cold as the sets: one DRAM round trip per batch instead
    // of one per set); the additions keep their order, so the sum keeps its bits
    for (int s0 = 0; s0 < f.n_sets; s0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (s0 + u < f.n_sets) ? f.raw[s0 + u][rep * EF_STATS_STRIDE + field] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (s0 + u < f.n_sets) acc += v[u];
    }
    s_part[rep][field] = acc;
    __syncthreads();
    if (threadIdx.x < EF_STATS_LEN) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < EF_STATS_REPLICAS; ++r) t += s_part[r][threadIdx.x];
        s_out[threadIdx.x] = t;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) stats_fold_kernel(const FoldArgs f, double* __restrict__ out) {
    __shared__ double s_part[EF_STATS_REPLICAS][EF_STATS_LEN];
    __shared__ double s_out[EF_STATS_LEN];
    pdl_launch_dependents();
    pdl_wait();
    fold_to_shared(f, s_part, s_out);
    if (threadIdx.x < EF_STATS_LEN) out[threadIdx.x] = s_out[threadIdx.x];
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kThreads) stats_fold_allreduce_kernel(const FoldArgs f, const PeerArgs p,
                                                                        double* __restrict__ out_local,
                                                                        double* __restrict__ out_global,
                                                                        uint32_t* __restrict__ status) {
    __shared__ double s_part[EF_STATS_REPLICAS][EF_STATS_LEN];
    __shared__ double s_out[EF_STATS_LEN];
    __shared__ double s_in[kMaxPeers][EF_STATS_LEN];
    __shared__ unsigned long long s_seq;
    pdl_launch_dependents();
    pdl_wait();
    PeerSlot* mine = p.mailbox[p.rank];
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(mine + 2 * p.world);
    if (threadIdx.x == 0) s_seq = *counter + 1ull;  // sequence number of this call (device-resident: graph replays advance it)
    if (f.n_sets > 0) {
        fold_to_shared(f, s_part, s_out);           // (its barriers also publish s_seq)
    } else {                                        // already folded by ef_stats_fold: out_local is the INPUT
        if (threadIdx.x < EF_STATS_LEN) s_out[threadIdx.x] = out_local[threadIdx.x];
        __syncthreads();
    }
    const unsigned long long seq = s_seq;
    const int parity = static_cast<int>(seq & 1ull);
    if (threadIdx.x < p.world) {
        // my vector -> slot [parity][my rank] of peer `threadIdx.x` (P2P stores over NVLink; the local slot the same way)
        PeerSlot* dst = p.mailbox[threadIdx.x] + parity * p.world + p.rank;
#pragma unroll
        for (int i = 0; i < EF_STATS_LEN; ++i) reinterpret_cast<volatile double*>(dst->v)[i] = s_out[i];
        __threadfence_system();
        st_release_sys(&dst->seq, seq);
        // peer `threadIdx.x`'s vector <- my local slot [parity][peer]; two sequence numbers share no slot, and a peer
        // can be at most one call ahead (it needs my vector of call s+1 to finish s+1), so two slots suffice
        const PeerSlot* src = mine + parity * p.world + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        bool ok = true;
        while (ld_acquire_sys(&src->seq) != seq) {
            if (global_timer_ns() - t0 > 10000000000ull) {  // a rank that never launches must not hang the GPU
                ok = false;
                break;
            }
            __nanosleep(64);
        }
        if (!ok && status) atomicExch(status, 1u);
#pragma unroll
        for (int i = 0; i < EF_STATS_LEN; ++i) s_in[threadIdx.x][i] = reinterpret_cast<const volatile double*>(src->v)[i];
    }
    __syncthreads();
    if (threadIdx.x < EF_STATS_LEN) {
        double t = 0.0;
        for (int r = 0; r < p.world; ++r) t += s_in[r][threadIdx.x];  // rank order on every rank: identical bits everywhere
        out_global[threadIdx.x] = t;
        if (out_local && f.n_sets > 0) out_local[threadIdx.x] = s_out[threadIdx.x];
    }
    if (threadIdx.x == 0) *counter = seq;
}

static int fill_fold(const double* const* raw_sets, int32_t n_sets, FoldArgs& f) {
    if (!raw_sets || n_sets < 1 || n_sets > EF_STATS_MAX_SETS) return EF_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_sets; ++i) {
        if (!raw_sets[i] || !aligned(raw_sets[i], 8)) return EF_ERR_INVALID_ARGUMENT;
        f.raw[i] = raw_sets[i];
    }
    f.n_sets = n_sets;
    return EF_OK;
}

}  // namespace ef

using namespace ef;

struct ef_peer_comm {
    int32_t rank, world, device;
    PeerSlot* local;                 // cudaMalloc'ed here (IPC-exportable)
    PeerSlot* peers[kMaxPeers];      // peers[rank] == local; the others are cudaIpcOpenMemHandle mappings
    bool connected;
};

extern "C" int ef_stats_fold(const double* const* raw_sets, int32_t n_sets, double* out, void* stream) {
    FoldArgs f{};
    if (int s = fill_fold(raw_sets, n_sets, f)) return s;
    if (!out) return EF_ERR_INVALID_ARGUMENT;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cuda_status(cudaLaunchKernelEx(&cfg, stats_fold_kernel, f, out));
}

extern "C" int ef_peer_comm_create(int32_t rank, int32_t world, ef_peer_comm** comm, uint8_t* handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) <= EF_PEER_HANDLE_BYTES, "IPC handle must fit the wire size");
    if (!comm || !handle_out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return EF_ERR_INVALID_ARGUMENT;
    ef_peer_comm* c = new (std::nothrow) ef_peer_comm();
    if (!c) return EF_ERR_INVALID_ARGUMENT;
    c->rank = rank;
    c->world = world;
    c->connected = false;
    for (int i = 0; i < kMaxPeers; ++i) c->peers[i] = nullptr;
    const size_t bytes = sizeof(PeerSlot) * (2 * static_cast<size_t>(world) + 1);  // slots + the sequence counter
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->local), bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) {
        if (c->local) cudaFree(c->local);
        delete c;
        return cuda_status(e);
    }
    memset(handle_out, 0, EF_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    c->peers[rank] = c->local;
    *comm = c;
    return EF_OK;
}

extern "C" int ef_peer_comm_connect(ef_peer_comm* c, const uint8_t* all_handles) {
    if (!c || !all_handles) return EF_ERR_INVALID_ARGUMENT;
    if (c->connected) return EF_OK;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + static_cast<size_t>(r) * EF_PEER_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_status(e);
        c->peers[r] = static_cast<PeerSlot*>(p);
    }
    c->connected = true;
    return EF_OK;
}

extern "C" int ef_peer_comm_destroy(ef_peer_comm* c) {
    if (!c) return EF_OK;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    return EF_OK;
}

extern "C" int ef_stats_fold_allreduce(ef_peer_comm* c, const double* const* raw_sets, int32_t n_sets, double* out_local,
                                       double* out_global, uint32_t* status, void* stream) {
    FoldArgs f{};
    if (n_sets == 0) {  // the vector was folded earlier (ef_stats_fold on the stepping stream): out_local is the input
        if (!out_local) return EF_ERR_INVALID_ARGUMENT;
    } else if (int s = fill_fold(raw_sets, n_sets, f)) {
        return s;
    }
    if (!c || !out_global || (c->world > 1 && !c->connected)) return EF_ERR_INVALID_ARGUMENT;
    PeerArgs p{};
    p.rank = c->rank;
    p.world = c->world;
    for (int r = 0; r < c->world; ++r) p.mailbox[r] = c->peers[r];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cuda_status(cudaLaunchKernelEx(&cfg, stats_fold_allreduce_kernel, f, p, out_local, out_global, status));
}
