// Attention cores of the bidirectional cross-attention on tcgen05 / TMEM, fed by TMA — forward and backward.
//
// Replaces the explicit bmm -> softmax -> bmm of nn.MultiheadAttention's need_weights path behind
// training/patch_text_aggregation.py:33 and :42 (torch/nn/functional.py multi_head_attention_forward) and its
// autograd backward (training/train_contrastive_teacher.py:361) for 16-bit activations:
//     P = softmax_j(scale * Q K^T),  O = P V                                   per (batch b, head h, direction)
//     dV = P^T dO,  dP = dO V^T,  dS = P o (dP - rowsum(dP o P)) * scale,  dQ = dS K,  dK = dS^T Q
// One CTA (128 threads) per (b, h, direction); the sequences are short (Lq, Lk <= 128: CLIP's 77 tokens, tens of YOLO
// regions), so one 128-row MMA tile holds a whole problem and nothing is ever written to HBM but O (dQ, dK, dV).
// No masks: zero-padded rows are attended like any other (the reference has none either).
//
// Data path.  The projections leave Q [B*Lq, d] and K|V [B*Lk, 2d] in HBM as 16-bit row-major matrices; head h is the
// column window [h*hd, h*hd + hd).  TMA loads [rows x 64-column] SWIZZLE_128B boxes straight from those matrices
// (ceil(hd/64) boxes per operand; columns past hd, rows past the sample are loaded but never consumed, rows / columns
// past the tensor are zero-filled).  The same shared-memory bytes serve as K-major operands (contraction over hd:
// S = Q K^T, dP = dO V^T) and as MN-major operands (contraction over the sequence: O = P V, dQ = dS K, dK = dS^T Q,
// dV = P^T dO) - only the descriptors differ.  Scores live in TMEM: thread t of the CTA owns row t (TMEM lane t),
// reads it with tcgen05.ld in 16-column slices, does the softmax in registers and writes P (and dS) as 16-bit
// SWIZZLE_128B tiles back to shared memory for the second-level MMAs.  Rows >= Lq and columns >= Lk of P / dS are
// written as exact zeros: they are the contraction padding of the transposed products.
// The backward also leaves the per-(b, head) column sums of dQ, dK, dV (the in-projection bias gradients) in a
// [B, cols] partial buffer: a 32 x 32 shuffle butterfly per warp, fixed order, no atomics.
#include "attn_tc.h"
#include "common.cuh"
#include "tcgen05.cuh"

namespace dclip {
using namespace tc;

namespace at {
constexpr int kThreads = 128;
constexpr int kQAtom = 128 * 128;          // what an M = 128 MMA reads from an A atom: [128 rows x 64 cols] 16-bit = 16 KB
constexpr uint32_t kSpin = 1u << 22;
constexpr float kLog2e = 1.4426950408889634f;
}  // namespace at

struct AttnTcDirDev {
    CUtensorMap tq, tkv, tdo;
    void* O; void* dQ; void* dKV;
    float* dbq; float* dbkv;
    int Lq, Lk, LkP, LqP;
};
struct AttnTcDev {
    AttnTcDirDev dir[2];
    int B, heads, hd, d, bf16;
    float scale, ds_scale;        // ds_scale: power of two applied to dS before its 16-bit rounding (fp16 range), undone on dQ / dK
    int* error_flag;
};

__device__ __forceinline__ float ex2f(float x) {          // arguments are <= 0: ex2.approx.ftz, no range handling needed
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack16(float a, float b, bool bf) {
    if (bf) {
        const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const uint32_t*>(&v);
    }
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// byte offset of the 16-byte chunk holding columns [8c, 8c + 8) of row r inside a [128 x 64] SWIZZLE_128B atom
__device__ __forceinline__ uint32_t sw128_chunk(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// sum over the warp's 32 rows of 32 per-thread column values: lane l ends with column l (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = upper ? v[i] : v[i + s];
            const float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(at::kThreads) k_attn_fwd_tc(const __grid_constant__ AttnTcDev a) {
    using namespace at;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const AttnTcDirDev& D = a.dir[blockIdx.y];
    const int b = blockIdx.x / a.heads, h = blockIdx.x - b * a.heads;
    const int hd = a.hd, na = (hd + 63) >> 6;
    const int Lq = D.Lq, Lk = D.Lk, LkP = D.LkP, LqP = D.LqP;
    const int kvAtom = LkP * 128, qAtom = LqP * 128;   // bytes of one [rows x 64 columns] SWIZZLE_128B atom
    const int nap = (LkP + 63) >> 6;
    // Q and P are A operands of M = 128 MMAs but only LqP rows are loaded / written: the tensor core reads 16 KB from each
    // atom base, i.e. into the buffers behind it (rows >= LqP of S / O are garbage and never used).  Q and P come first
    // so that those reads stay inside this CTA's shared memory (the host pads the allocation).
    uint8_t* Qs = smem;
    uint8_t* Ps = Qs + na * qAtom;
    uint8_t* Ks = Ps + nap * qAtom;
    uint8_t* Vs = Ks + na * kvAtom;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(Vs + na * kvAtom);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int row = threadIdx.x;                       // my query row == TMEM lane
    const bool bf = a.bf16 != 0;

    if (threadIdx.x == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 128);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;

    const uint64_t hiK = make_smem_desc_hi(0, 1024, kSwizzle128B);
    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(bar_tma, (uint32_t)(na * (qAtom + 2 * kvAtom)));
        for (int t = 0; t < na; ++t) {
            tma_load_2d(Qs + t * qAtom, &D.tq, bar_tma, h * hd + t * 64, b * Lq);
            tma_load_2d(Ks + t * kvAtom, &D.tkv, bar_tma, h * hd + t * 64, b * Lk);
            tma_load_2d(Vs + t * kvAtom, &D.tkv, bar_tma, a.d + h * hd + t * 64, b * Lk);
        }
        ok = mbar_wait(bar_tma, 0, kSpin);
        fence_after_thread_sync();
        // S[128 q x LkP j] = Q K^T, contraction over hd in 16-wide steps (4 per 64-column atom)
        const uint32_t idesc = make_idesc_f16(128, (uint32_t)LkP, bf ? kFmtBF16 : kFmtF16, 0, 0);
        const uint32_t qa = smem_u32(Qs), ka = smem_u32(Ks);
        for (int ks = 0; ks < hd / 16; ++ks)
            umma_f16(tmem, smem_desc(hiK, qa + (ks >> 2) * qAtom + (ks & 3) * 32), smem_desc(hiK, ka + (ks >> 2) * kvAtom + (ks & 3) * 32), idesc,
                     ks != 0);
        umma_commit(bar_mma);
    }
    __syncwarp();
    ok = mbar_wait(bar_mma, 0, kSpin) && ok;
    fence_after_thread_sync();

    // ---- softmax of my row (three passes over TMEM: max, sum, write P); rows >= LqP have no place in the P tile ----
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float c = a.scale * kLog2e;
    float mx = -INFINITY;
    for (int c0 = 0; c0 < LkP; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < Lk) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    const float mc = mx * c;
    float sum = 0.f;
    for (int c0 = 0; c0 < LkP; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < Lk) sum += ex2f(fmaf(__uint_as_float(r[i]), c, -mc));
    }
    const float inv = 1.0f / sum;
    for (int c0 = 0; c0 < LkP; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
        tmem_ld_wait();
        if (row < LqP) {
            float p[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = c0 + i < Lk ? ex2f(fmaf(__uint_as_float(r[i]), c, -mc)) * inv : 0.f;
            uint8_t* atom = Ps + (c0 >> 6) * qAtom;
            const int ch = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(atom + sw128_chunk(row, ch)) = make_uint4(pack16(p[0], p[1], bf), pack16(p[2], p[3], bf), pack16(p[4], p[5], bf), pack16(p[6], p[7], bf));
            *reinterpret_cast<uint4*>(atom + sw128_chunk(row, ch + 1)) =
                make_uint4(pack16(p[8], p[9], bf), pack16(p[10], p[11], bf), pack16(p[12], p[13], bf), pack16(p[14], p[15], bf));
        }
    }
    fence_proxy_async_smem();            // my generic-proxy stores of P -> visible to the tensor core (async proxy)
    fence_before_thread_sync();          // my TMEM reads of S are complete before O overwrites those columns
    __syncthreads();

    if (warp == 0 && elect_one()) {
        fence_after_thread_sync();
        // O[128 q x hd] = P V, contraction over j in 16-row steps; V is the MN-major B operand ([j rows x hd columns])
        const uint32_t idesc = make_idesc_f16(128, (uint32_t)hd, bf ? kFmtBF16 : kFmtF16, 0, 1);
        const uint64_t hiV = make_smem_desc_hi((uint32_t)kvAtom, 1024, kSwizzle128B);
        const uint32_t pa = smem_u32(Ps), va = smem_u32(Vs);
        for (int ks = 0; ks < LkP / 16; ++ks)
            umma_f16(tmem, smem_desc(hiK, pa + (ks >> 2) * qAtom + (ks & 3) * 32), smem_desc(hiV, va + ks * 2048), idesc, ks != 0);
        umma_commit(bar_mma);
    }
    __syncwarp();
    ok = mbar_wait(bar_mma, 1, kSpin) && ok;
    fence_after_thread_sync();

    // ---- O row -> global (16-bit) ----
    uint16_t* orow = static_cast<uint16_t*>(D.O) + ((size_t)b * Lq + row) * a.d + h * hd;
    for (int c0 = 0; c0 < hd; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
        tmem_ld_wait();
        if (row < Lq) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]), bf);
            *reinterpret_cast<uint4*>(orow + c0) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
    if (!__all_sync(0xffffffffu, ok) && lane == 0 && a.error_flag) atomicExch(a.error_flag, 1);
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
// Shared memory (all tiles hold only the rows that exist, padded to 16: two CTAs fit on an SM):
//   Q, dO [LqP x hd] | dS, P [LqP x 128] | K, V [LkP x hd] | barriers;  the bias-gradient scratch aliases Q (dead by then).
// Q, dO and dS are also A operands of M = 128 MMAs (S, dP, dQ): the tensor core reads 16 KB from each atom base, into the
// buffers behind it - rows >= LqP of those products are garbage and never used; these tiles come first so that the
// reads stay inside the CTA's allocation (the host pads it).
// TMEM (256 columns): S -> [0, LkP), dP -> [128, 128 + LkP); then dQ -> [0, hd), dK -> [128, 128 + hd); dV re-uses
// [0, hd) once every thread has read its dQ row (it is issued while the dK rows are being stored).
__global__ void __launch_bounds__(at::kThreads) k_attn_bwd_tc(const __grid_constant__ AttnTcDev a) {
    using namespace at;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const AttnTcDirDev& D = a.dir[blockIdx.y];
    const int b = blockIdx.x / a.heads, h = blockIdx.x - b * a.heads;
    const int hd = a.hd, na = (hd + 63) >> 6;
    const int Lq = D.Lq, Lk = D.Lk, LkP = D.LkP, LqP = D.LqP;
    const int kvAtom = LkP * 128, qAtom = LqP * 128;
    uint8_t* Qs = smem;
    uint8_t* dOs = Qs + na * qAtom;
    uint8_t* dSs = dOs + na * qAtom;                  // [LqP q x 128 j] = two atoms, always (MN-major A operand with M = 128 j)
    uint8_t* Ps = dSs + 2 * qAtom;
    uint8_t* Ks = Ps + 2 * qAtom;
    uint8_t* Vs = Ks + na * kvAtom;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(Vs + na * kvAtom);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma + 1);
    float* red = reinterpret_cast<float*>(tmem_ptr + 2);          // [4 warps][3 outputs][128 columns]
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int row = threadIdx.x;
    const bool bf = a.bf16 != 0;

    if (threadIdx.x == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;
    const uint64_t hiK = make_smem_desc_hi(0, 1024, kSwizzle128B);
    const uint32_t fmt = bf ? kFmtBF16 : kFmtF16;

    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(bar_tma, (uint32_t)(na * (2 * qAtom + 2 * kvAtom)));
        for (int t = 0; t < na; ++t) {
            tma_load_2d(Qs + t * qAtom, &D.tq, bar_tma, h * hd + t * 64, b * Lq);
            tma_load_2d(dOs + t * qAtom, &D.tdo, bar_tma, h * hd + t * 64, b * Lq);
            tma_load_2d(Ks + t * kvAtom, &D.tkv, bar_tma, h * hd + t * 64, b * Lk);
            tma_load_2d(Vs + t * kvAtom, &D.tkv, bar_tma, a.d + h * hd + t * 64, b * Lk);
        }
        ok = mbar_wait(bar_tma, 0, kSpin);
        fence_after_thread_sync();
        const uint32_t idesc = make_idesc_f16(128, (uint32_t)LkP, fmt, 0, 0);
        const uint32_t qa = smem_u32(Qs), doa = smem_u32(dOs), ka = smem_u32(Ks), va = smem_u32(Vs);
        for (int ks = 0; ks < hd / 16; ++ks) {       // S = Q K^T -> columns [0, LkP);  dP = dO V^T -> columns [128, 128 + LkP)
            const uint32_t offq = (ks >> 2) * qAtom + (ks & 3) * 32, offk = (ks >> 2) * kvAtom + (ks & 3) * 32;
            umma_f16(tmem, smem_desc(hiK, qa + offq), smem_desc(hiK, ka + offk), idesc, ks != 0);
            umma_f16(tmem + 128, smem_desc(hiK, doa + offq), smem_desc(hiK, va + offk), idesc, ks != 0);
        }
        umma_commit(bar_mma);
    }
    __syncwarp();
    ok = mbar_wait(bar_mma, 0, kSpin) && ok;
    fence_after_thread_sync();

    // ---- recompute P, form dS (my row): max; then sum and sum(e dP) in one pass; then write ----
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float c = a.scale * kLog2e;
    float mx = -INFINITY;
    for (int c0 = 0; c0 < LkP; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < Lk) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    const float mc = mx * c;
    float sum = 0.f, dot = 0.f;                       // sum_j e_j,  sum_j e_j dP_j
    for (int c0 = 0; c0 < LkP; c0 += 16) {
        uint32_t r[16], g[16];
        tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
        tmem_ld_32x32b_x16(tmem + lane_addr + 128 + c0, g);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < Lk) {
                const float e = ex2f(fmaf(__uint_as_float(r[i]), c, -mc));
                sum += e;
                dot = fmaf(e, __uint_as_float(g[i]), dot);
            }
    }
    const float inv = 1.0f / sum;
    dot *= inv;
    const bool row_ok = row < Lq;
    const float dsc = a.scale * a.ds_scale;
    if (warp * 32 < LqP) {               // warp-uniform (the TMEM loads are warp-collective); rows >= LqP of a straddling warp only skip the stores
        for (int c0 = 0; c0 < 128; c0 += 16) {
            float p[16], ds[16];
            if (c0 < LkP) {
                uint32_t r[16], g[16];
                tmem_ld_32x32b_x16(tmem + lane_addr + c0, r);
                tmem_ld_32x32b_x16(tmem + lane_addr + 128 + c0, g);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool v = row_ok && c0 + i < Lk;
                    p[i] = v ? ex2f(fmaf(__uint_as_float(r[i]), c, -mc)) * inv : 0.f;
                    ds[i] = v ? p[i] * (__uint_as_float(g[i]) - dot) * dsc : 0.f;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) { p[i] = 0.f; ds[i] = 0.f; }
            }
            if (row >= LqP) continue;          // no such row in the P / dS tiles
            const int ch = (c0 & 63) >> 3;
            uint8_t* pa = Ps + (c0 >> 6) * qAtom;
            uint8_t* da = dSs + (c0 >> 6) * qAtom;
            *reinterpret_cast<uint4*>(pa + sw128_chunk(row, ch)) = make_uint4(pack16(p[0], p[1], bf), pack16(p[2], p[3], bf), pack16(p[4], p[5], bf), pack16(p[6], p[7], bf));
            *reinterpret_cast<uint4*>(pa + sw128_chunk(row, ch + 1)) =
                make_uint4(pack16(p[8], p[9], bf), pack16(p[10], p[11], bf), pack16(p[12], p[13], bf), pack16(p[14], p[15], bf));
            *reinterpret_cast<uint4*>(da + sw128_chunk(row, ch)) =
                make_uint4(pack16(ds[0], ds[1], bf), pack16(ds[2], ds[3], bf), pack16(ds[4], ds[5], bf), pack16(ds[6], ds[7], bf));
            *reinterpret_cast<uint4*>(da + sw128_chunk(row, ch + 1)) =
                make_uint4(pack16(ds[8], ds[9], bf), pack16(ds[10], ds[11], bf), pack16(ds[12], ds[13], bf), pack16(ds[14], ds[15], bf));
        }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();

    const uint32_t qa = smem_u32(Qs), doa = smem_u32(dOs), ka = smem_u32(Ks), pa_ = smem_u32(Ps), dsa = smem_u32(dSs);
    const uint32_t idesc_t = make_idesc_f16(128, (uint32_t)hd, fmt, 1, 1);     // A MN-major (dS^T / P^T), B MN-major (Q / dO)
    const uint64_t hiQ = make_smem_desc_hi((uint32_t)qAtom, 1024, kSwizzle128B);         // MN-major, 64-column groups qAtom apart
    if (warp == 0 && elect_one()) {
        fence_after_thread_sync();
        const uint32_t idesc_q = make_idesc_f16(128, (uint32_t)hd, fmt, 0, 1);     // A K-major (dS), B MN-major (K)
        const uint64_t hiKV = make_smem_desc_hi((uint32_t)kvAtom, 1024, kSwizzle128B);   // MN-major, 64-column groups kvAtom apart
        // dQ[128 q x hd] = dS K  (contraction over j, LkP / 16 steps) -> [0, hd);  dK[128 j x hd] = dS^T Q (over q) -> [128, 128 + hd)
        for (int ks = 0; ks < LkP / 16; ++ks)
            umma_f16(tmem, smem_desc(hiK, dsa + (ks >> 2) * qAtom + (ks & 3) * 32), smem_desc(hiKV, ka + ks * 2048), idesc_q, ks != 0);
        for (int ks = 0; ks < LqP / 16; ++ks)
            umma_f16(tmem + 128, smem_desc(hiQ, dsa + ks * 2048), smem_desc(hiQ, qa + ks * 2048), idesc_t, ks != 0);
        umma_commit(bar_mma);
    }
    __syncwarp();
    ok = mbar_wait(bar_mma, 1, kSpin) && ok;
    fence_after_thread_sync();

    // ---- dQ / dK / dV rows -> global (16-bit) + per-warp column sums (bias gradients) ----
    const float inv_ds = 1.0f / a.ds_scale;
    uint16_t* dq = static_cast<uint16_t*>(D.dQ) + ((size_t)b * Lq + row) * a.d + h * hd;
    uint16_t* dk = static_cast<uint16_t*>(D.dKV) + ((size_t)b * Lk + row) * 2 * a.d + h * hd;
    uint16_t* dv = dk + a.d;
    auto drain = [&](int o, uint32_t tcol) {
        uint16_t* dst = o == 0 ? dq : o == 1 ? dk : dv;
        const bool valid = o == 0 ? row < Lq : row < Lk;
        const float sc = o == 2 ? 1.0f : inv_ds;
        for (int c0 = 0; c0 < hd; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + lane_addr + tcol + c0, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = valid ? __uint_as_float(r[i]) * sc : 0.f;     // rows past the sequence: garbage, masked
            if (valid) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    *reinterpret_cast<uint4*>(dst + c0 + q4 * 8) = make_uint4(pack16(v[q4 * 8], v[q4 * 8 + 1], bf), pack16(v[q4 * 8 + 2], v[q4 * 8 + 3], bf),
                                                                              pack16(v[q4 * 8 + 4], v[q4 * 8 + 5], bf), pack16(v[q4 * 8 + 6], v[q4 * 8 + 7], bf));
            }
            red[(warp * 3 + o) * 128 + c0 + lane] = warp_colsum32(v, lane);
        }
    };
    drain(0, 0);                                       // dQ
    fence_before_thread_sync();                        // every thread has read its dQ row: the columns may be overwritten
    __syncthreads();
    if (warp == 0 && elect_one()) {
        fence_after_thread_sync();
        // dV[128 j x hd] = P^T dO (contraction over q) -> [0, hd), issued while the dK rows drain
        for (int ks = 0; ks < LqP / 16; ++ks)
            umma_f16(tmem, smem_desc(hiQ, pa_ + ks * 2048), smem_desc(hiQ, doa + ks * 2048), idesc_t, ks != 0);
        umma_commit(bar_mma);
    }
    __syncwarp();
    drain(1, 128);                                     // dK
    ok = mbar_wait(bar_mma, 0, kSpin) && ok;           // third completion of the barrier: parity 0 again
    fence_after_thread_sync();
    drain(2, 0);                                       // dV
    __syncthreads();
    if (D.dbq) {
        for (int e = threadIdx.x; e < 3 * hd; e += kThreads) {
            const int o = e / hd, col = e - o * hd;
            const float sv = red[(0 * 3 + o) * 128 + col] + red[(1 * 3 + o) * 128 + col] + red[(2 * 3 + o) * 128 + col] + red[(3 * 3 + o) * 128 + col];
            if (o == 0) D.dbq[(size_t)b * a.d + h * hd + col] = sv;
            else D.dbkv[(size_t)b * 2 * a.d + (o - 1) * a.d + h * hd + col] = sv;
        }
    }
    if (!__all_sync(0xffffffffu, ok) && lane == 0 && a.error_flag) atomicExch(a.error_flag, 1);
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int* tc_error_flag();   // loss_tc.cu

bool attn_tc_shape_ok(int64_t T, int64_t R, int64_t d, int heads) {
    if (heads <= 0 || d % heads) return false;
    const int64_t hd = d / heads;
    return hd % 32 == 0 && hd >= 32 && hd <= 128 && d % 8 == 0 && T >= 1 && T <= 128 && R >= 1 && R <= 128;
}

static size_t attn_smem(int hd, int Lq, int Lk, bool bwd) {
    const size_t na = (hd + 63) / 64, LkP = (Lk + 15) / 16 * 16, LqP = (Lq + 15) / 16 * 16, kv = LkP * 128, qa = LqP * 128;
    size_t total, last_a;      // last_a: base of the last atom that is read as a 128-row A operand (16 KB from its base)
    if (!bwd) {
        const size_t nap = (LkP + 63) / 64;
        total = (na + nap) * qa + 2 * na * kv + 64;
        last_a = (na + nap - 1) * qa;
    } else {
        total = (2 * na + 4) * qa + 2 * na * kv + 64 + 4 * 3 * 128 * sizeof(float);
        last_a = (2 * na + 1) * qa;
    }
    return 1024 + max(total, last_a + (size_t)at::kQAtom);
}

// text side: rows = B*T, image side: rows = B*R.  dir 0 = text queries over image keys, dir 1 = the reverse.
int launch_attn_tc(const AttnTcProblem& p, bool backward, cudaStream_t st) {
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "attention tensor path needs an sm_100 device");
    if (!attn_tc_shape_ok(p.T, p.R, p.d, p.heads)) return fail(DCLIP_ERR_UNSUPPORTED, "attention tensor path: unsupported shape");
    AttnTcDev a{};
    const int dt = p.dtype16 == DCLIP_BF16 ? 1 : 0;
    a.B = p.B; a.heads = p.heads; a.hd = p.d / p.heads; a.d = p.d; a.bf16 = dt;
    a.scale = 1.0f / sqrtf((float)a.hd);
    a.ds_scale = dt ? 1.0f : 4096.0f;
    a.error_flag = tc_error_flag();
    for (int dir = 0; dir < 2; ++dir) {
        AttnTcDirDev& D = a.dir[dir];
        const int Lq = dir == 0 ? p.T : p.R, Lk = dir == 0 ? p.R : p.T;
        D.Lq = Lq; D.Lk = Lk; D.LkP = (Lk + 15) / 16 * 16; D.LqP = (Lq + 15) / 16 * 16;
        const void* Q = dir == 0 ? p.Qt : p.Qi;
        const void* KV = dir == 0 ? p.KVi : p.KVt;
        bool okm = make_tmap_2d_sw128_ld(&D.tq, Q, (uint64_t)p.B * Lq, (uint64_t)p.d, (uint64_t)p.d, (uint32_t)D.LqP, dt);
        okm = okm && make_tmap_2d_sw128_ld(&D.tkv, KV, (uint64_t)p.B * Lk, (uint64_t)2 * p.d, (uint64_t)2 * p.d, (uint32_t)D.LkP, dt);
        if (backward) {
            const void* dO = dir == 0 ? p.dOt : p.dOi;
            okm = okm && make_tmap_2d_sw128_ld(&D.tdo, dO, (uint64_t)p.B * Lq, (uint64_t)p.d, (uint64_t)p.d, (uint32_t)D.LqP, dt);
            D.dQ = dir == 0 ? p.dQt : p.dQi;
            D.dKV = dir == 0 ? p.dKVi : p.dKVt;
            D.dbq = dir == 0 ? p.dbq_t2i : p.dbq_i2t;
            D.dbkv = dir == 0 ? p.dbkv_t2i : p.dbkv_i2t;
            if (!D.dQ || !D.dKV || !dO) return fail(DCLIP_ERR_INVALID, "attention backward: null pointer");
        } else {
            D.tdo = D.tq;
            D.O = dir == 0 ? p.Ot : p.Oi;
            if (!D.O) return fail(DCLIP_ERR_INVALID, "attention forward: null pointer");
        }
        if (!okm) return fail(DCLIP_ERR_CUDA, "attention: cuTensorMapEncodeTiled failed");
    }
    const size_t smem = max(attn_smem(a.hd, p.T, p.R, backward), attn_smem(a.hd, p.R, p.T, backward));
    dim3 grid((unsigned)(p.B * p.heads), 2);
    if (backward) {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_attn_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_attn_bwd_tc<<<grid, at::kThreads, smem, st>>>(a);
        DCLIP_LAUNCH_OK("k_attn_bwd_tc");
    } else {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_attn_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_attn_fwd_tc<<<grid, at::kThreads, smem, st>>>(a);
        DCLIP_LAUNCH_OK("k_attn_fwd_tc");
    }
    return DCLIP_OK;
}

}  // namespace dclip
