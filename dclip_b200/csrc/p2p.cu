// Push-based exchange over NVLink peer memory for the data-parallel objective: every rank owns
// one "symmetric" buffer (same layout on all ranks, cudaIpc-mapped into every peer process); a rank all-gathers
// by WRITING its slice into the same offsets of every peer's buffer and then bumping a counter in the peer's
// flag block; the consumer side is a one-warp wait kernel on its LOCAL flags.  No NCCL call, no host
// synchronisation, graph-capturable (all state lives in device memory):
//
//   flag block (at `flags_offset` inside the symmetric buffer):
//     arrive[kind][src]   bumped by rank `src` (remote atomic) when its push of that kind has landed here
//     expected[kind]      how many pushes of that kind this rank has already consumed (local)
//     done_ctr            last-block detection of the push kernel
//
// Ordering: a push block stores, __threadfence_system(), then counts itself done; the last block fences again
// and bumps the peers' arrive counters (system-scope atomics).  The waiter spins with volatile system-scope
// loads and fences before it lets the stream continue.  Reuse of the buffers across steps is protected by a
// third kind ("step done"): a rank does not push step k+1 into a peer that has not finished reading step k.
// A peer that never arrives is FATAL, as with NCCL's watchdog: after 120 s of %globaltimer the waiting kernel sets
// the library's pipeline error flag and executes `trap`, so the stream never continues on stale operands and the
// next CUDA call on this context raises (expected[] is not advanced, nothing of the next step is published).
#include "common.cuh"
#include "p2p.h"

namespace dclip {

constexpr int kP2PMaxRanges = 4;

struct PushArgs {
    char* base[kP2PMaxWorld];
    size_t off[kP2PMaxRanges], bytes[kP2PMaxRanges];
    int nranges, world, rank;
    size_t flags_offset;
    int wait_kind, signal_kind;
    int wait_after;       // 1: the kernel also waits until every peer's push of `signal_kind` has landed HERE (push + wait in one launch)
    int* error_flag;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int ld_volatile_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// wait until arrive[kind][r] >= target for every r != rank (lane r watches rank r)
__device__ bool wait_arrivals(const P2PFlags* f, int kind, unsigned int target, int world, int rank, int lane) {
    bool ok = true;
    if (lane < world && lane != rank) {
        const unsigned long long t0 = globaltimer_ns();
        while ((int)(ld_volatile_sys(&f->arrive[kind][lane]) - target) < 0) {
            if (globaltimer_ns() - t0 > kP2PTimeoutNs) { ok = false; break; }
            __nanosleep(64);
        }
    }
    return ok;
}
// a wait that timed out: record it and kill the context (no silent continuation on the previous step's data)
__device__ __forceinline__ void p2p_fatal(int* error_flag) {
    if (error_flag) atomicExch(error_flag, 1);
    __threadfence_system();
    __trap();
}

__global__ void __launch_bounds__(256) k_p2p_push(const PushArgs a) {
    P2PFlags* mine = reinterpret_cast<P2PFlags*>(a.base[a.rank] + a.flags_offset);
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (a.wait_kind >= 0) {
        // every block waits (the grid may not be co-resident with a block that would do it for all)
        if (threadIdx.x < 32) {
            const unsigned int target = ld_volatile_sys(&mine->expected[a.wait_kind]) + 1;
            if (!wait_arrivals(mine, a.wait_kind, target, a.world, a.rank, threadIdx.x)) p2p_fatal(a.error_flag);
        }
        __syncthreads();
        __threadfence_system();
    }
    for (int r = 0; r < a.nranges; ++r) {
        const size_t n16 = a.bytes[r] / 16;
        const uint4* src = reinterpret_cast<const uint4*>(a.base[a.rank] + a.off[r]);
        for (int p = 0; p < a.world; ++p) {
            if (p == a.rank) continue;
            uint4* dst = reinterpret_cast<uint4*>(a.base[p] + a.off[r]);
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&mine->done_ctr, 1u) == gridDim.x - 1) {        // last block: everything of this rank has landed
            __threadfence_system();
            if (a.signal_kind >= 0)
                for (int p = 0; p < a.world; ++p)
                    if (p != a.rank)
                        atomicAdd_system(&reinterpret_cast<P2PFlags*>(a.base[p] + a.flags_offset)->arrive[a.signal_kind][a.rank], 1u);
            if (a.wait_kind >= 0) mine->expected[a.wait_kind] += 1;
            mine->done_ctr = 0;
            __threadfence_system();
            s_ok = 2;                  // this block is the last one
        }
    }
    __syncthreads();
    if (a.wait_after && a.signal_kind >= 0 && s_ok == 2) {
        // the kernel (hence the stream) does not proceed before the peers' pushes of the same kind are here
        if (threadIdx.x < 32) {
            const unsigned int target = ld_volatile_sys(&mine->expected[a.signal_kind]) + 1;
            if (!wait_arrivals(mine, a.signal_kind, target, a.world, a.rank, threadIdx.x)) p2p_fatal(a.error_flag);
            __threadfence_system();
            __syncwarp();
            if (threadIdx.x == 0) mine->expected[a.signal_kind] += 1;
        }
    }
}

__global__ void k_p2p_wait(P2PFlags* mine, int kind, int world, int rank, int* error_flag) {
    const unsigned int target = ld_volatile_sys(&mine->expected[kind]) + 1;
    if (!wait_arrivals(mine, kind, target, world, rank, threadIdx.x)) p2p_fatal(error_flag);
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) mine->expected[kind] += 1;
}

int* tc_error_flag();   // loss_tc.cu

}  // namespace dclip

extern "C" {

size_t dclip_p2p_flags_bytes(void) { return dclip::align_up(sizeof(dclip::P2PFlags), 256); }

int dclip_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
    using namespace dclip;
    DCLIP_REQUIRE(dev_ptr && handle64 && bytes > 0, "p2p_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    void* p = nullptr;
    DCLIP_CUDA_OK(cudaMalloc(&p, bytes));
    DCLIP_CUDA_OK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    DCLIP_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return DCLIP_OK;
}

// "step done" credits start at 1: the first push of a run has nothing to wait for
int dclip_p2p_init_flags(void* base, size_t flags_offset, int credit_kind) {
    using namespace dclip;
    DCLIP_REQUIRE(base && credit_kind >= 0 && credit_kind < kP2PKinds, "p2p_init_flags: bad arguments");
    P2PFlags f;
    memset(&f, 0, sizeof(f));
    for (int r = 0; r < kP2PMaxWorld; ++r) f.arrive[credit_kind][r] = 1;
    DCLIP_CUDA_OK(cudaMemcpy(static_cast<char*>(base) + flags_offset, &f, sizeof(f), cudaMemcpyHostToDevice));
    DCLIP_CUDA_OK(cudaDeviceSynchronize());
    return DCLIP_OK;
}

int dclip_p2p_open(const unsigned char* handle64, void** dev_ptr) {
    using namespace dclip;
    DCLIP_REQUIRE(dev_ptr && handle64, "p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    DCLIP_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DCLIP_OK;
}

int dclip_p2p_close(void* dev_ptr) {
    using namespace dclip;
    DCLIP_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
    return DCLIP_OK;
}

int dclip_p2p_free(void* dev_ptr) {
    using namespace dclip;
    DCLIP_CUDA_OK(cudaFree(dev_ptr));
    return DCLIP_OK;
}

int dclip_p2p_push(void* const* peer_bases, int world, int rank, const size_t* offsets, const size_t* bytes, int nranges,
                   size_t flags_offset, int wait_kind, int signal_kind, int wait_after, void* stream) {
    using namespace dclip;
    DCLIP_REQUIRE(peer_bases && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, "p2p_push: bad world / rank");
    DCLIP_REQUIRE(nranges >= 0 && nranges <= kP2PMaxRanges && wait_kind < kP2PKinds && signal_kind < kP2PKinds, "p2p_push: bad ranges / kinds");
    PushArgs a{};
    size_t total = 0;
    for (int p = 0; p < world; ++p) {
        DCLIP_REQUIRE(peer_bases[p] != nullptr, "p2p_push: null peer base");
        a.base[p] = static_cast<char*>(peer_bases[p]);
    }
    for (int r = 0; r < nranges; ++r) {
        DCLIP_REQUIRE(offsets[r] % 16 == 0 && bytes[r] % 16 == 0, "p2p_push: ranges must be 16-byte aligned");
        a.off[r] = offsets[r]; a.bytes[r] = bytes[r];
        total += bytes[r];
    }
    a.nranges = nranges; a.world = world; a.rank = rank; a.flags_offset = flags_offset;
    a.wait_kind = wait_kind; a.signal_kind = signal_kind; a.wait_after = wait_after; a.error_flag = tc_error_flag();
    int blocks = (int)ceil_div((int64_t)(total / 16), 256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
    k_p2p_push<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    DCLIP_LAUNCH_OK("k_p2p_push");
    return DCLIP_OK;
}

int dclip_p2p_wait(void* local_base, size_t flags_offset, int world, int rank, int kind, void* stream) {
    using namespace dclip;
    DCLIP_REQUIRE(local_base && world >= 1 && world <= kP2PMaxWorld && kind >= 0 && kind < kP2PKinds, "p2p_wait: bad arguments");
    k_p2p_wait<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<P2PFlags*>(static_cast<char*>(local_base) + flags_offset), kind, world,
                                                                 rank, tc_error_flag());
    DCLIP_LAUNCH_OK("k_p2p_wait");
    return DCLIP_OK;
}

}  // extern "C"
