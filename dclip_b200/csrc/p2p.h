// Peer-memory exchange state shared by p2p.cu (stand-alone push / wait kernels) and the loss kernels that do their exchange
// themselves (remote stores + flag from the producing kernel, flag wait in the consuming kernel's prologue).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dclip {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PKinds = 4;
constexpr int kP2PKindOperands = 0, kP2PKindStats = 1, kP2PKindDone = 2;     // dclip_b200/peer.py: KIND_*
constexpr unsigned long long kP2PTimeoutNs = 120000000000ull;

// flag block at `flags_offset` of every rank's symmetric buffer
struct P2PFlags {
    unsigned int arrive[kP2PKinds][kP2PMaxWorld];   // arrive[kind][src]: bumped by rank src (remote atomic) when its data of that kind has landed here
    unsigned int expected[kP2PKinds];               // pushes of that kind this rank has consumed (local)
    unsigned int done_ctr;                          // last-block detection of the kernel that signals
};

// the ranks' symmetric buffers as mapped into THIS process; world == 0: no peer exchange inside the kernels
struct PeerCtx {
    char* base[kP2PMaxWorld];
    int world, rank;
    unsigned long long flags_offset;
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long p2p_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int p2p_ld_volatile(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ P2PFlags* p2p_flags(const PeerCtx& c, int r) { return reinterpret_cast<P2PFlags*>(c.base[r] + c.flags_offset); }
// the address in rank r's buffer of a location inside MY buffer
template <typename T> __device__ __forceinline__ T* p2p_remote(const PeerCtx& c, T* mine, int r) {
    return reinterpret_cast<T*>(c.base[r] + (reinterpret_cast<char*>(mine) - c.base[c.rank]));
}
// one thread: wait until every peer's arrive[kind] has reached target (expected[kind] + 1, read by the caller at kernel start).
// A peer that never arrives is fatal (error flag + trap), as in p2p.cu.
__device__ __forceinline__ void p2p_wait_all(const PeerCtx& c, int kind, unsigned int target, int* error_flag) {
    const P2PFlags* f = p2p_flags(c, c.rank);
    const unsigned long long t0 = p2p_globaltimer_ns();
    for (int r = 0; r < c.world; ++r) {
        if (r == c.rank) continue;
        while ((int)(p2p_ld_volatile(&f->arrive[kind][r]) - target) < 0) {
            if (p2p_globaltimer_ns() - t0 > kP2PTimeoutNs) {
                if (error_flag) atomicExch(error_flag, 1);
                __threadfence_system();
                __trap();
            }
            __nanosleep(32);
        }
    }
    __threadfence_system();
}
// lane r of a warp watches rank r (all lanes return together)
__device__ __forceinline__ void p2p_wait_all_warp(const PeerCtx& c, int kind, unsigned int target, int* error_flag, int lane) {
    const P2PFlags* f = p2p_flags(c, c.rank);
    if (lane < c.world && lane != c.rank) {
        const unsigned long long t0 = p2p_globaltimer_ns();
        while ((int)(p2p_ld_volatile(&f->arrive[kind][lane]) - target) < 0) {
            if (p2p_globaltimer_ns() - t0 > kP2PTimeoutNs) {
                if (error_flag) atomicExch(error_flag, 1);
                __threadfence_system();
                __trap();
            }
            __nanosleep(32);
        }
    }
    __threadfence_system();
    __syncwarp();
}
// one thread, after every store of this rank is fenced (system scope): tell every peer that my data of `kind` has landed
__device__ __forceinline__ void p2p_signal_all(const PeerCtx& c, int kind) {
    __threadfence_system();
    for (int r = 0; r < c.world; ++r)
        if (r != c.rank) atomicAdd_system(&p2p_flags(c, r)->arrive[kind][c.rank], 1u);
}
// last-block detection for grid-wide "all my stores are out": every block calls this after __threadfence_system() + __syncthreads()
// from ONE thread; returns true in the last block (which then signals)
__device__ __forceinline__ bool p2p_last_block(const PeerCtx& c, unsigned int nblocks) {
    P2PFlags* f = p2p_flags(c, c.rank);
    if (atomicAdd(&f->done_ctr, 1u) == nblocks - 1) {
        f->done_ctr = 0;
        __threadfence_system();
        return true;
    }
    return false;
}
#endif

}  // namespace dclip
