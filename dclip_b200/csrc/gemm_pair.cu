// CTA-pair tcgen05 GEMM with 256 x 384 tiles (cta_group::2), fp32 accumulation in TMEM:
//     C[M, N] = alpha * A'(M, K) . B'(K, N)                    up to 2 problems per launch
// Built for the two gradient GEMMs of the distillation objective over the materialised coefficient matrix
// (loss_tc.cu::launch_infonce_bwd_gmat; autograd backward of training/CLIP_image_distillation.py:545-560):
//     dU_hat = G V_hat,  dV_hat = G^T U_hat        M = K = B, N = d
// Why this shape.  B = 4096, d = 768 is 2 x 16 x 2 = 64 tiles of 256 x 384: ONE wave on 74 CTA pairs, every pair
// walks the whole K range once (no split-K partials, no second wave at 30 % occupancy as with 128 x 256 tiles:
// 192 tiles on 148 SMs).  A pair tile needs two MMAs per K = 16 step, N = 256 (TMEM columns 0-255) and N = 128
// (columns 256-383); both CTAs supply their own 128 rows of A and half of each B operand, so a CTA moves
// 40 KB per 64-wide K stage (16 KB of A + 24 KB of B) for 802 tensor-pipe cycles.
//
// Operand storage (row-major 16-bit matrices, TMA boxes of 64 x 64 elements = 8 KB, SWIZZLE_128B):
//   A K-major  : stored [M, K]     2 boxes [64 m x 64 k] stacked (8-row groups 1024 B apart)
//   A MN-major : stored [K, M]     2 boxes [64 k x 64 m], LBO = 8 KB between the 64-wide M groups
//   B K-major  : stored [N, K]     3 boxes [64 n x 64 k]: two for the N = 256 MMA, one for the N = 128 MMA
//   B MN-major : stored [K, N]     3 boxes [64 k x 64 n]
// D column c of the tile is global column n0 + c: the leader's B halves are columns [0, 128) and [256, 320),
// the peer's [128, 256) and [320, 384).
// Warps: 0-4 TMA producers (one ring stage each, both CTAs), 5 MMA issuer (leader) + TMEM alloc, 6-21 epilogue
// (TMEM lane quadrant = warp & 3, the four warps of a quadrant take 96 columns each; output staged through the
// idle ring so that global stores are 128 contiguous bytes per row).
#include <cstdlib>

#include "common.cuh"
#include "gemm_tc.h"
#include "tcgen05.cuh"

namespace dclip {
using namespace tc;

namespace gp {
constexpr int TM = 256, TN = 384, KS = 64;
constexpr int kStages = 5;
constexpr int kBox = 8192;                       // one 64 x 64 16-bit box
constexpr int kAStage = 2 * kBox;                // my 128 rows of A
constexpr int kBStage = 3 * kBox;                // my halves of the two B operands (128 + 64 columns)
constexpr int kStageBytes = kAStage + kBStage;   // 40 KB
constexpr int kEpiWarps = 16;               // four per TMEM lane quadrant, 96 columns each
constexpr int kThreads = (kStages + 1 + kEpiWarps) * 32;
constexpr uint32_t kSpin = 1u << 24;
constexpr int kLateBlocks = 12;                  // ~10 k cycles of MMAs: long enough to pull 196 KB per SM from DRAM with every SM asking
}  // namespace gp

struct GemmPairDev {
    float* C; int M, N, K, ldc; int a_mn, b_mn; float alpha; int tiles_n, tile_begin;
    GemmPairTail tail;      // tail.out != null: the epilogue finishes the gradient instead of writing C
};
struct GemmPairArgs {
    CUtensorMap ta[2], tb[2];
    GemmPairDev p[2];
    int count, total_tiles, ab_fmt;
    float grad_scale; int io_dtype;      // fused tail: dclip_dtype of x / t / out
    int debug_skip;                      // stall-accounting build only (results are garbage): 1 no x / t loads, 2 no stores, 4 no L2 prefetch
    int* error_flag;
};

constexpr float kPairNormEps = 1e-12f;  // F.normalize eps (loss_simt.cu::kNormEps)

// 4 consecutive elements of a row in the I/O dtype (0 fp32, 1 bf16, 2 fp16) <-> fp32
__device__ __forceinline__ void load4_io(const void* base, size_t idx, int io, float (&v)[4]) {
    if (io == DCLIP_F32) {
        const float4 q = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        const uint2 q = *reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + idx);
        if (io == DCLIP_BF16) {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        }
    }
}
__device__ __forceinline__ void store4_io(void* base, size_t idx, int io, const float (&v)[4]) {
    if (io == DCLIP_F32) {
        *reinterpret_cast<float4*>(static_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        uint2 q;
        if (io == DCLIP_BF16) {
            const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
            q.x = *reinterpret_cast<const uint32_t*>(&a); q.y = *reinterpret_cast<const uint32_t*>(&b);
        } else {
            const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
            q.x = *reinterpret_cast<const uint32_t*>(&a); q.y = *reinterpret_cast<const uint32_t*>(&b);
        }
        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(base) + idx) = q;
    }
}


// 4 consecutive 16-bit elements (bf16 or fp16) of a row <-> fp32
template <bool kBf16> __device__ __forceinline__ void unpack4_16(const uint2& q, float (&v)[4]) {
    if (kBf16) {
        v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
        v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
    } else {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}
template <bool kBf16> __device__ __forceinline__ uint2 pack4_16(const float (&v)[4]) {
    uint2 q;
    if (kBf16) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        q.x = *reinterpret_cast<const uint32_t*>(&a); q.y = *reinterpret_cast<const uint32_t*>(&b);
    } else {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        q.x = *reinterpret_cast<const uint32_t*>(&a); q.y = *reinterpret_cast<const uint32_t*>(&b);
    }
    return q;
}

// Fused gradient tail for 16-bit I/O, one warp: 96 columns of 32 rows in three 32-column chunks.
//   out = acc * A_r + t * T_r - x * X_r     with the per-row float4 {A, T, X, -} = {alpha, ts, inv dot} * inv * grad_scale in shared memory
// The accumulator chunk goes through the warp's swizzled 4 KB staging buffer (thread = row on the way in, 8 lanes per row on
// the way out); the student / teacher elements a lane needs (8 rows x 8 B per operand and chunk) are requested BEFORE the chunk's
// TMEM load - loads issued where they are used serialise one L2 round trip per 4 rows (measured: +40 us).
template <bool kBf16>
__device__ __forceinline__ void fused_tail_epilogue16(uint32_t tmem_rows, uint32_t stage_s, const float4* row_sc4, int qd, int lane, int rbase, int n0,
                                                      int c_begin, int M, int N, const uint16_t* __restrict__ x16, const uint16_t* __restrict__ t16,
                                                      uint16_t* __restrict__ out16, bool skip_loads, bool skip_stores, bool skip_tmem = false,
                                                      bool skip_stage = false) {
    const int rb_row = lane >> 3, rb_chunk = lane & 7;
    int rowoff[8];          // element offset of my 8 rows (row k of a pass = rb_row + 4 k); -1 = past M
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int row = rbase + rb_row + 4 * k;
        rowoff[k] = row < M ? row * N : -1;
    }
    uint2 xq[8], tq[8];
    auto prefetch = [&](int c0) {
        const int ccol = n0 + c0 + rb_chunk * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            xq[k] = make_uint2(0u, 0u); tq[k] = make_uint2(0u, 0u);
            if (rowoff[k] >= 0 && ccol < N && !skip_loads) {
                xq[k] = __ldg(reinterpret_cast<const uint2*>(x16 + rowoff[k] + ccol));
                if (t16) tq[k] = __ldg(reinterpret_cast<const uint2*>(t16 + rowoff[k] + ccol));
            }
        }
    };
    prefetch(c_begin);
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_begin + 96; c0 += 32) {
        if (n0 + c0 >= N) break;
        uint32_t r[32];
        if (!skip_tmem) {
            tmem_ld_32x32b_x32(tmem_rows + c0, r);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = (uint32_t)(c0 + i + lane);
        }
        if (!skip_stage) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_s + lane * 128 + ((q ^ (lane & 7)) << 4)), "r"(r[q * 4]), "r"(r[q * 4 + 1]),
                         "r"(r[q * 4 + 2]), "r"(r[q * 4 + 3]) : "memory");
        }
        __syncwarp();
        const int ccol = n0 + c0 + rb_chunk * 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int rr = rb_row + 4 * k;
            uint4 a4;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a4.x), "=r"(a4.y), "=r"(a4.z), "=r"(a4.w)
                         : "r"(stage_s + rr * 128 + ((rb_chunk ^ (rr & 7)) << 4)) : "memory");
            if (rowoff[k] >= 0 && ccol < N) {
                const float4 sc = row_sc4[qd * 32 + rr];
                float xv[4], tv[4], o[4];
                unpack4_16<kBf16>(xq[k], xv);
                unpack4_16<kBf16>(tq[k], tv);
                o[0] = fmaf(__uint_as_float(a4.x), sc.x, fmaf(tv[0], sc.y, -xv[0] * sc.z));
                o[1] = fmaf(__uint_as_float(a4.y), sc.x, fmaf(tv[1], sc.y, -xv[1] * sc.z));
                o[2] = fmaf(__uint_as_float(a4.z), sc.x, fmaf(tv[2], sc.y, -xv[2] * sc.z));
                o[3] = fmaf(__uint_as_float(a4.w), sc.x, fmaf(tv[3], sc.y, -xv[3] * sc.z));
                if (!skip_stores) *reinterpret_cast<uint2*>(out16 + rowoff[k] + ccol) = pack4_16<kBf16>(o);
            }
        }
        __syncwarp();
        if (c0 + 32 < c_begin + 96 && n0 + c0 + 32 < N) prefetch(c0 + 32);      // next chunk (L2 hits: the tile was prefetched at kernel start)
    }
}

#ifdef DCLIP_GEMM_PROF
#define GP_SKIP(bit) (g.debug_skip & (bit))
__device__ long long g_gpair_prof[16];
__device__ __forceinline__ long long gtimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PPROF(k) do { if (blockIdx.x == 0) { g_gpair_prof[k] = clock64(); if ((k) == 0 || (k) == 6) g_gpair_prof[8 + (k)] = gtimer_ns(); } } while (0)
#else
#define GP_SKIP(bit) false
#define PPROF(k) do { } while (0)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gp::kThreads, 1) k_gemm_pair(const __grid_constant__ GemmPairArgs g) {
    using namespace gp;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kStages * kStageBytes);   // (leader) bytes of both CTAs
    uint64_t* empty = full + kStages;                                            // both CTAs (multicast commit)
    uint64_t* acc_full = empty + kStages;                                        // both CTAs
    uint64_t* late = acc_full + 1;                                               // both CTAs: kLateBlocks k-blocks before the end of the K loop
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(late + 1);
    float* row_sc = reinterpret_cast<float*>(ring + kStages * kStageBytes + 256);   // fused tail: float4[128] per-row scalars
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t cr = cluster_ctarank();

    const int tile = blockIdx.x >> 1;
    const int pi = (g.count > 1 && tile >= g.p[1].tile_begin) ? 1 : 0;
    const GemmPairDev P = g.p[pi];
    const int local = tile - P.tile_begin;
    const int tm = local / P.tiles_n, tn = local - tm * P.tiles_n;
    const int m0 = tm * TM + (int)cr * 128;       // my 128 rows of the tile
    const int n0 = tn * TN;
    const int nkb = (P.K + KS - 1) / KS;

    if (threadIdx.x == 0) PPROF(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        mbar_init(late, 1);
        fence_barrier_init();
    }
    if (warp == kStages) tmem_alloc_pair(tmem_ptr, 512);
    fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;
    if (threadIdx.x == 0) PPROF(1);

    if (warp < kStages) {
        if (elect_one()) {
            const CUtensorMap* ta = &g.ta[pi];
            const CUtensorMap* tb = &g.tb[pi];
            tma_prefetch_desc(ta);
            tma_prefetch_desc(tb);
            uint8_t* dst = ring + (size_t)warp * kStageBytes;
            const uint32_t fbar = mapa_u32(smem_u32(&full[warp]), 0);
            const int nb0 = n0 + (int)cr * 128, nb2 = n0 + 256 + (int)cr * 64;     // my columns of the two B operands
            uint32_t phase = 0;
            for (int kb = warp; kb < nkb && ok; kb += kStages, phase ^= 1) {
                const int k0 = kb * KS;
                ok = mbar_wait(&empty[warp], phase ^ 1, kSpin);
                if (cr == 0) mbar_arrive_expect_tx(&full[warp], 2 * kStageBytes);
                if (!P.a_mn) { tma_load_2d_pair(dst, ta, fbar, k0, m0); tma_load_2d_pair(dst + kBox, ta, fbar, k0, m0 + 64); }
                else { tma_load_2d_pair(dst, ta, fbar, m0, k0); tma_load_2d_pair(dst + kBox, ta, fbar, m0 + 64, k0); }
                uint8_t* db = dst + kAStage;
                if (!P.b_mn) {
                    tma_load_2d_pair(db, tb, fbar, k0, nb0); tma_load_2d_pair(db + kBox, tb, fbar, k0, nb0 + 64);
                    tma_load_2d_pair(db + 2 * kBox, tb, fbar, k0, nb2);
                } else {
                    tma_load_2d_pair(db, tb, fbar, nb0, k0); tma_load_2d_pair(db + kBox, tb, fbar, nb0 + 64, k0);
                    tma_load_2d_pair(db + 2 * kBox, tb, fbar, nb2, k0);
                }
            }
        }
    } else if (warp == kStages) {
        if (cr == 0 && elect_one()) {
            const uint32_t ring_addr = smem_u32(ring);
            const uint32_t idesc1 = make_idesc_f16(256, 256, (uint32_t)g.ab_fmt, (uint32_t)P.a_mn, (uint32_t)P.b_mn);
            const uint32_t idesc2 = make_idesc_f16(256, 128, (uint32_t)g.ab_fmt, (uint32_t)P.a_mn, (uint32_t)P.b_mn);
            const uint64_t hiA = make_smem_desc_hi(P.a_mn ? kBox : 0, 1024, kSwizzle128B);
            const uint64_t hiB = make_smem_desc_hi(P.b_mn ? kBox : 0, 1024, kSwizzle128B);
            const uint32_t stepA = P.a_mn ? 128u : 2u, stepB = P.b_mn ? 128u : 2u;     // per K = 16 step, >> 4 encoded
            uint32_t stage = 0, phase = 0;
            for (int kb = 0; kb < nkb && ok; ++kb) {
                ok = mbar_wait(&full[stage], phase, kSpin);
                if (kb == 0) PPROF(2);
                if (kb == nkb - 1) PPROF(3);
                fence_after_thread_sync();
                const uint32_t base = ring_addr + stage * kStageBytes;
                const uint64_t da0 = smem_desc(hiA, base);
                const uint64_t db1 = smem_desc(hiB, base + kAStage);
                const uint64_t db2 = smem_desc(hiB, base + kAStage + 2 * kBox);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    umma_f16_pair(tmem, da0 + (uint64_t)(ks * stepA), db1 + (uint64_t)(ks * stepB), idesc1, (kb | ks) != 0);
                    umma_f16_pair(tmem + 256, da0 + (uint64_t)(ks * stepA), db2 + (uint64_t)(ks * stepB), idesc2, (kb | ks) != 0);
                }
                umma_commit_pair(&empty[stage]);
                if (kb == max(0, nkb - 1 - kLateBlocks)) umma_commit_pair(late);      // wakes the epilogue warps' L2 prefetch (below)
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(acc_full);
        }
    } else {
        // ---- epilogue: thread = one output row (TMEM lane), 96 columns per warp in chunks of 32 (128 B of fp32 per row),
        // staged through this warp's 4 KB of the (now idle) ring: 8 lanes per row on the way out; 16 warps keep enough loads / stores
        // in flight for an epilogue that nothing overlaps (one tile per CTA)
        const int qd = warp & 3;
        const int we = warp - (kStages + 1);
        const int part = we >> 2;                 // my 96 of the tile's 384 columns
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const GemmPairTail& T = P.tail;
        const bool fused = T.out != nullptr;
        if (fused) {
            // per-row scalars of my 128 rows while the main loop runs (k_loss_grad_tail's math, loss_simt.cu):
            //   g = alpha acc + ts t,  dot = x_hat . g = alpha sum_j g_ij s_ij - w cos_i,  out = (g - x inv dot) inv grad_scale
            const int idx = we * 32 + lane;
            if (idx < 128) {
                const int row = m0 + idx;
                float sc = 0.f, ts = 0.f, dx = 0.f;      // inv grad_scale, -w / |t|, inv dot
                if (row < P.M) {
                    float r = 0.f;
                    for (int q = 0; q < T.dot_count; ++q) r += T.dot_part[(size_t)q * T.dot_stride + row];
                    const float inv = T.inv_norm[row];
                    float dot = P.alpha * r;
                    if (T.t) { ts = -T.w_cos_over_B * T.t_inv[row]; dot -= T.w_cos_over_B * T.cos[row]; }
                    if (inv >= 1.0f / kPairNormEps) dot = 0.f;      // clamped norm: F.normalize has no projection term
                    sc = inv * g.grad_scale; dx = inv * dot;
                }
                // out = acc * (alpha sc) + t * (ts sc) - x * (dx sc): three FMAs per element in the epilogue
                reinterpret_cast<float4*>(row_sc)[idx] = make_float4(P.alpha * sc, ts * sc, dx * sc, 0.f);
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            // my tile of the raw student / teacher rows was last touched by the prep kernel, four kernels and > 300 MB of L2 traffic
            // ago, and a prefetch issued at kernel start is evicted again by the operand stream (measured: no gain).  Pull it into
            // L2 when the K loop is kLateBlocks k-blocks from its end (128 rows x 384 columns x 2 arrays, one 128-byte line per
            // request): the epilogue's loads are then L2 hits instead of a 25 MB DRAM burst from every CTA at once.
            ok = mbar_wait(late, 0, kSpin);
            {
                const int esz = g.io_dtype == DCLIP_F32 ? 4 : 2;
                const int lines_per_row = (min(TN, P.N - n0) * esz + 127) / 128;
                const int total = 128 * lines_per_row;
                for (int i = we * 32 + lane; i < total; i += kEpiWarps * 32) {
                    const int r = i / lines_per_row, l = i - r * lines_per_row;
                    if (m0 + r < P.M && !GP_SKIP(4)) {
                        const size_t off = ((size_t)(m0 + r) * P.N + n0) * esz + (size_t)l * 128;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(T.x) + off));
                        if (T.t) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(T.t) + off));
                    }
                }
            }
        }
        const uint32_t stage_s = smem_u32(ring) + (uint32_t)we * 4096;
        const int rb_row = lane >> 3, rb_chunk = lane & 7;
        const int rbase = m0 + qd * 32;
        const bool io16 = g.io_dtype != DCLIP_F32;
        ok = mbar_wait(acc_full, 0, kSpin) && ok;
        ok = __all_sync(0xffffffffu, ok);
        if (we == 0 && lane == 0) PPROF(4);
        if (ok) {
            fence_after_thread_sync();
            if (fused && io16) {
                // (16-bit I/O: the epilogue that matters - the first chunk's loads could sit above the accumulator wait, but they are
                //  L2 hits by now and the wait is short: measured no difference)
                const float4* sc4 = reinterpret_cast<const float4*>(row_sc);
                if (g.io_dtype == DCLIP_BF16)
                    fused_tail_epilogue16<true>(tmem + lane_addr, stage_s, sc4, qd, lane, rbase, n0, part * 96, P.M, P.N, static_cast<const uint16_t*>(T.x),
                                                static_cast<const uint16_t*>(T.t), static_cast<uint16_t*>(T.out), GP_SKIP(1), GP_SKIP(2), GP_SKIP(8), GP_SKIP(16));
                else
                    fused_tail_epilogue16<false>(tmem + lane_addr, stage_s, sc4, qd, lane, rbase, n0, part * 96, P.M, P.N, static_cast<const uint16_t*>(T.x),
                                                 static_cast<const uint16_t*>(T.t), static_cast<uint16_t*>(T.out), GP_SKIP(1), GP_SKIP(2));
            } else {
                // plain fp32 C (alpha applied), or the fused tail with fp32 I/O (rare: tests)
#pragma unroll 1
                for (int c0 = part * 96; c0 < part * 96 + 96; c0 += 32) {
                    const int col = n0 + c0;
                    if (col >= P.N) break;
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem + lane_addr + c0, r);
                    tmem_ld_wait();
                    const float al = fused ? 1.0f : P.alpha;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint4 v = make_uint4(__float_as_uint(__uint_as_float(r[q * 4]) * al), __float_as_uint(__uint_as_float(r[q * 4 + 1]) * al),
                                                   __float_as_uint(__uint_as_float(r[q * 4 + 2]) * al), __float_as_uint(__uint_as_float(r[q * 4 + 3]) * al));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_s + lane * 128 + ((q ^ (lane & 7)) << 4)), "r"(v.x), "r"(v.y),
                                     "r"(v.z), "r"(v.w) : "memory");
                    }
                    __syncwarp();
                    const int ccol = col + rb_chunk * 4;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int rr = rb_row + 4 * k;
                        uint4 x;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                                     : "r"(stage_s + rr * 128 + ((rb_chunk ^ (rr & 7)) << 4)) : "memory");
                        if (rbase + rr < P.M && ccol < P.N) {
                            if (!fused) *reinterpret_cast<uint4*>(P.C + (size_t)(rbase + rr) * P.ldc + ccol) = x;
                            else {
                                const float4 sc = reinterpret_cast<const float4*>(row_sc)[qd * 32 + rr];
                                const size_t off = (size_t)(rbase + rr) * P.N + ccol;
                                const float acc[4] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)};
                                float xv[4], tv[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
                                load4_io(T.x, off, g.io_dtype, xv);
                                if (T.t) load4_io(T.t, off, g.io_dtype, tv);
#pragma unroll
                                for (int i = 0; i < 4; ++i) o[i] = fmaf(acc[i], sc.x, fmaf(tv[i], sc.y, -xv[i] * sc.z));
                                store4_io(T.out, off, g.io_dtype, o);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (!ok && lane == 0 && g.error_flag) atomicExch(g.error_flag, 1);
    if (warp == kStages + 1 && lane == 0) PPROF(5);
    fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == kStages) tmem_dealloc_pair(tmem, 512);
    if (threadIdx.x == 0) PPROF(6);
}

int* tc_error_flag();   // loss_tc.cu

bool gemm_pair_ok(const TcGemmProblem& q) {
    return !q.C16 && !q.bias && !q.residual && !q.accumulate && q.N % 4 == 0 && q.lda % 8 == 0 && q.ldb % 8 == 0 && q.ldc % 4 == 0;
}

// dtype16: 1 = bf16, 2 = fp16 (dclip_dtype values)
int launch_gemm_pair(const TcGemmProblem* probs, int count, int dtype16, cudaStream_t st, const GemmPairTail* tails, float grad_scale, int io_dtype) {
    using namespace gp;
    if (count < 1 || count > 2) return fail(DCLIP_ERR_INVALID, "gemm_pair: 1 or 2 problems per launch");
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "tensor path needs an sm_100 device");
    GemmPairArgs g{};
    g.count = count;
    g.ab_fmt = dtype16 == DCLIP_BF16 ? (int)kFmtBF16 : (int)kFmtF16;
    g.error_flag = tc_error_flag();
    const int dt = dtype16 == DCLIP_BF16 ? 1 : 0;
    int tiles = 0;
    for (int i = 0; i < count; ++i) {
        const TcGemmProblem& q = probs[i];
        if (q.M <= 0 || q.N <= 0 || q.K <= 0) return fail(DCLIP_ERR_INVALID, "gemm_pair: empty problem %d", i);
        if (!gemm_pair_ok(q)) return fail(DCLIP_ERR_INVALID, "gemm_pair: problem %d needs a plain fp32 output with 16-byte aligned rows", i);
        bool okm;
        if (!q.a_mn) okm = make_tmap_2d_sw128_ld(&g.ta[i], q.A, (uint64_t)q.M, (uint64_t)q.K, (uint64_t)q.lda, 64, dt);
        else okm = make_tmap_2d_sw128_ld(&g.ta[i], q.A, (uint64_t)q.K, (uint64_t)q.M, (uint64_t)q.lda, 64, dt);
        if (!q.b_mn) okm = okm && make_tmap_2d_sw128_ld(&g.tb[i], q.B, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.ldb, 64, dt);
        else okm = okm && make_tmap_2d_sw128_ld(&g.tb[i], q.B, (uint64_t)q.K, (uint64_t)q.N, (uint64_t)q.ldb, 64, dt);
        if (!okm) return fail(DCLIP_ERR_CUDA, "gemm_pair: cuTensorMapEncodeTiled failed for problem %d", i);
        GemmPairDev& d = g.p[i];
        d.C = q.C; d.M = q.M; d.N = q.N; d.K = q.K; d.ldc = q.ldc; d.a_mn = q.a_mn; d.b_mn = q.b_mn;
        d.alpha = q.alpha == 0.f ? 1.0f : q.alpha;
        if (tails && (long long)q.M * q.N >= (1ll << 31)) return fail(DCLIP_ERR_UNSUPPORTED, "gemm_pair: fused tail needs M * N < 2^31");
        if (tails) {
            d.tail = tails[i];
            if (!d.tail.out || !d.tail.x || !d.tail.inv_norm || !d.tail.dot_part || (d.tail.t && (!d.tail.cos || !d.tail.t_inv)))
                return fail(DCLIP_ERR_INVALID, "gemm_pair: incomplete gradient-tail description for problem %d", i);
        } else if (!q.C) return fail(DCLIP_ERR_INVALID, "gemm_pair: problem %d has no output", i);
        d.tiles_n = (q.N + TN - 1) / TN;
        d.tile_begin = tiles;
        tiles += ((q.M + TM - 1) / TM) * d.tiles_n;
    }
    g.total_tiles = tiles;
    g.grad_scale = grad_scale; g.io_dtype = io_dtype;
#ifdef DCLIP_GEMM_PROF
    { const char* e = getenv("DCLIP_GPAIR_SKIP"); g.debug_skip = e ? atoi(e) : 0; }
#endif
    const size_t smem = 1024 + (size_t)kStages * kStageBytes + 256 + 4 * 128 * 4;
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_gemm_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_pair<<<2 * tiles, kThreads, smem, st>>>(g);
    DCLIP_LAUNCH_OK("k_gemm_pair");
    return DCLIP_OK;
}

}  // namespace dclip

#ifdef DCLIP_GEMM_PROF
extern "C" int dclip_debug_gpair_prof(long long* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(out16, dclip::g_gpair_prof, 16 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
// stall-accounting build only (make prof): one 16-bit GEMM to fp32 through either kernel (tools/exp_gemm_pair.py)
extern "C" int dclip_debug_gemm16(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn, int b_mn,
                                  float alpha, int kernel, int dtype16, int count, void* stream) {
    dclip::TcGemmProblem p[2] = {};
    p[0].A = A; p[0].B = B; p[0].C = C; p[0].M = M; p[0].N = N; p[0].K = K; p[0].lda = lda; p[0].ldb = ldb; p[0].ldc = ldc;
    p[0].a_mn = a_mn; p[0].b_mn = b_mn; p[0].alpha = alpha;
    p[1] = p[0];
    return kernel == 0 ? dclip::launch_gemm_tc(p, count, dtype16, static_cast<cudaStream_t>(stream))
                       : dclip::launch_gemm_pair(p, count, dtype16, static_cast<cudaStream_t>(stream), nullptr, 1.0f, 0);
}
#endif
