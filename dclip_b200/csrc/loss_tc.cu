// Distillation objective — tcgen05 / TMEM / TMA tensor path (sm_100a): the InfoNCE BACKWARD
// (dX_hat = coef/tau * G . Y_hat with G recomputed; autograd backward of
// training/CLIP_image_distillation.py:545-560).  The forward row log-sum-exp runs on the 128 x 256
// streaming kernel in stream_tc.cu.  The B x B logits only ever exist as 128 x 64 fp32 tiles in TMEM.
//
// Work decomposition.  A CTA owns NI = 64 rows of X_hat (kept resident in shared memory for the
// whole kernel) and walks over 128-row tiles J of Y_hat:
//   phase 1   S^T[128 j x 64 i] = Y_J . X_I^T          (M=128, N=64, K=d;  K-major operands)
//   epilogue  thread = one j (TMEM lane), 64 i columns: p = exp2((s - 1) * log2e / tau),
//             g = p (a_i + b_j) - 2 [i == j], scaled by 2^13, rounded to fp16 and stored as a
//             128B-swizzled [j][i] tile (an MN-major B operand)
//   phase 2   dX^T[d x 64 i] += Y_J^T[d x 128 j] . G^T   (M=128 per d-block, N=64, K=128 j;
//             A = the very same TMA boxes as phase 1 read MN-major, B = the G tile)
// Accumulators: d/128 blocks x 64 columns (<= 384) + two 64-column S buffers = 512 TMEM columns: the
// gradient accumulator pins N to 64.  The fixed offset (s <= 1 for unit rows) replaces the running max
// of an online softmax, so the backward needs a single exp per logit:
//   softmax_row + softmax_col = p_ij (1/rowsum_i + 1/colsum_j).
// Valid while 2*log2e/tau stays inside the fp32 exponent range: tau >= 0.025 (reference: 0.05).
//
// Warps: 0-2 = TMA producers, one per ring slot (a single thread needs ~650 cycles per
// wait/expect_tx/issue round trip — measured with tools/tc_bench — so one producer cannot feed the
// MMAs), 3 = MMA issuer (+ TMEM alloc), 4..7 = epilogue (one TMEM lane quadrant each).
// Pipelines: 3-slot ring of 32 KB (two 128x64 boxes) between TMA and MMA; S double-buffered
// MMA<->epilogue; G double-buffered epilogue<->MMA.  Phase 1 of tile t+1 is issued before phase 2 of
// tile t so the tensor pipe works while the epilogue of tile t runs.
// Bound: shared-memory bandwidth — every 128x64x16 MMA re-reads 6 KB of operands and the TMA ring
// writes another 4 KB per MMA, ~80 cycles at 128 B/clk against 32 cycles of math (ncu: tensor pipe
// active 38 %, profiles/r1_ncu_full_loss_kernels.txt).
#include <cstdlib>

#include "common.cuh"

#include <mutex>
#include "gemm_tc.h"
#include "loss_internal.h"
#include "tcgen05.cuh"

namespace dclip {
using namespace tc;

constexpr int NI = 64;                 // X rows per CTA  (MMA N)
constexpr int TJ = 128;                // Y rows per tile (MMA M in phase 1, K in phase 2)
constexpr int kSlots = 3;
constexpr int kBoxBytes = TJ * 128;    // 128 rows x 64 fp16 = 16 KB
constexpr int kSlotBytes = 2 * kBoxBytes;
constexpr int kXBoxBytes = NI * 128;   // 64 rows x 64 fp16 = 8 KB
constexpr int kGBytes = TJ * 128;      // G tile: 128 j rows x 64 i fp16 = 16 KB
constexpr int kTcThreads = 256;        // warps 0-2: TMA producers (one ring slot each), 3: MMA issuer, 4-7: epilogue
constexpr int kAccCols = 384;          // S buffers live at columns 384 and 448
constexpr float kGScaleLog2 = 13.0f;   // G stored as fp16 * 2^13 (softmax sums <= 2 -> <= 16384)
constexpr uint32_t kSpin = 1u << 24;   // bounded waits: a protocol bug times out instead of hanging

struct Pipe {
    uint32_t stage = 0, phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == kSlots) { stage = 0; phase ^= 1; }
    }
};

// (A thread-block-cluster variant that TMA-multicast the Y boxes to 2 or 4 CTAs was built and measured:
// no gain — the kernel is bound by shared-memory operand bandwidth, not by L2 -> SM traffic — and dropped.
// The forward used to run on this 128 x 64 tiling too; it moved to the 128 x 256 kernel in stream_tc.cu.)
__global__ void __launch_bounds__(kTcThreads, 1)
k_infonce_bwd_tc(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmy0,
             const __grid_constant__ CUtensorMap tmx1, const __grid_constant__ CUtensorMap tmy1, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int side = blockIdx.z;
    const CUtensorMap* tmx = side ? &tmx1 : &tmx0;
    const CUtensorMap* tmy = side ? &tmy1 : &tmy0;
    const TcSide S = a.s[side];
    const int d = (int)a.d;
    const int nkb = d / 64;            // 64-column boxes along K
    const int NS = d / 128;            // ring slots per phase (= accumulator d-blocks)

    uint8_t* xs = smem;                                   // resident X_I: nkb boxes of 8 KB
    uint8_t* ring = xs + (size_t)nkb * kXBoxBytes;        // kSlots x 32 KB
    uint8_t* gs = ring + kSlots * kSlotBytes;             // 2 x 16 KB (bwd)
    uint8_t* tail = gs + 2 * kGBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);   // [kSlots]
    uint64_t* empty = full + kSlots;                      // [kSlots]
    uint64_t* s_full = empty + kSlots;                    // [2]
    uint64_t* s_empty = s_full + 2;                       // [2]
    uint64_t* g_full = s_empty + 2;                       // [2]
    uint64_t* g_empty = g_full + 2;                       // [2]
    uint64_t* x_full = g_empty + 2;                       // [1]
    uint64_t* acc_full = x_full + 1;                      // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* a_s = reinterpret_cast<float*>(tmem_ptr + 2);  // [NI] (bwd) column factors

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;     // warp-uniform role index
    const int64_t row0 = (int64_t)blockIdx.x * NI;
    const int64_t ntiles = (a.n + TJ - 1) / TJ;
    const int64_t t_begin = (int64_t)blockIdx.y * a.tiles_per_split;
    const int64_t t_end = min(ntiles, t_begin + a.tiles_per_split);
    const int T = (int)(t_end - t_begin);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128);
            mbar_init(&g_full[i], 128); mbar_init(&g_empty[i], 1);
        }
        mbar_init(x_full, 1); mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_ptr, 512);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;

    if (T > 0) {
        if (warp < kSlots) {
            // ===================== TMA producers: warp w owns ring slot w =====================
            if (elect_one()) {
                tma_prefetch_desc(tmy);
                if (warp == 0) {
                    tma_prefetch_desc(tmx);
                    mbar_arrive_expect_tx(x_full, (uint32_t)(nkb * kXBoxBytes));
                    for (int kb = 0; kb < nkb; ++kb) tma_load_2d(xs + (size_t)kb * kXBoxBytes, tmx, x_full, kb * 64, (int32_t)row0);
                }
                // Job q (in MMA consumption order) uses slot q % kSlots.  Order: P1(0), then per tile t:
                // P1(t+1) [if any], P2(t) [bwd].  Both phases of (tile, s) load the same two boxes.
                const int njobs = 2 * NS * T;
                uint8_t* dst = ring + (size_t)warp * kSlotBytes;
                uint32_t phase = 0;
                for (int q = warp; q < njobs && ok; q += kSlots, phase ^= 1) {
                    int tile, sidx;
                    if (q < NS) { tile = 0; sidx = q; }
                    else {
                        const int qq = q - NS, blk = min(qq / (2 * NS), T - 1), r = qq - blk * 2 * NS;
                        if (blk < T - 1 && r < NS) { tile = blk + 1; sidx = r; }
                        else { tile = blk; sidx = blk < T - 1 ? r - NS : r; }
                    }
                    const int32_t j0 = (int32_t)((t_begin + tile) * TJ);
                    ok = mbar_wait(&empty[warp], phase ^ 1, kSpin);
                    mbar_arrive_expect_tx(&full[warp], kSlotBytes);
                    tma_load_2d(dst, tmy, &full[warp], (2 * sidx) * 64, j0);
                    tma_load_2d(dst + kBoxBytes, tmy, &full[warp], (2 * sidx + 1) * 64, j0);
                }
            }
        } else if (warp == 3) {
            // ===================== MMA issuer =====================
            if (elect_one()) {
                const uint32_t idesc1 = make_idesc_f16(128, NI, kFmtF16, 0, 0);
                const uint32_t idesc2 = make_idesc_f16(128, NI, kFmtF16, 1, 1);
                const uint64_t hiK = make_smem_desc_hi(0, 1024, kSwizzle128B);
                const uint64_t hiA2 = make_smem_desc_hi(kBoxBytes, 1024, kSwizzle128B);   // MN-major A: LBO = next 64-d group
                const uint64_t hiB2 = make_smem_desc_hi(0, 1024, kSwizzle128B);           // MN-major B (G), N = 64: one group
                const uint32_t xs_addr = smem_u32(xs), ring_addr = smem_u32(ring), gs_addr = smem_u32(gs);
                Pipe p;
                ok = mbar_wait(x_full, 0, kSpin);
                auto phase1 = [&](int t) {
                    const int b = t & 1;
                    ok = ok && mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1, kSpin);
                    fence_after_thread_sync();
                    const uint32_t dcol = tmem + kAccCols + b * NI;
                    for (int s = 0; s < NS && ok; ++s) {
                        ok = mbar_wait(&full[p.stage], p.phase, kSpin);
                        fence_after_thread_sync();
                        const uint64_t da0 = smem_desc(hiK, ring_addr + p.stage * kSlotBytes);
                        const uint64_t db0 = smem_desc(hiK, xs_addr + (2 * s) * kXBoxBytes);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)      // +32 B per K step, next box at +16 KB (A) / +8 KB (B); >>4 encoded
                                umma_f16(dcol, da0 + (uint64_t)(h * (kBoxBytes >> 4) + ks * 2), db0 + (uint64_t)(h * (kXBoxBytes >> 4) + ks * 2), idesc1,
                                         (s | h | ks) != 0);
                        umma_commit(&empty[p.stage]);
                        p.advance();
                    }
                    umma_commit(&s_full[b]);
                };
                auto phase2 = [&](int t) {
                    const int b = t & 1;
                    ok = ok && mbar_wait(&g_full[b], (t >> 1) & 1, kSpin);
                    fence_after_thread_sync();
                    const uint32_t gaddr = gs_addr + b * kGBytes;
                    for (int s = 0; s < NS && ok; ++s) {
                        ok = mbar_wait(&full[p.stage], p.phase, kSpin);
                        fence_after_thread_sync();
                        const uint64_t da0 = smem_desc(hiA2, ring_addr + p.stage * kSlotBytes);
                        const uint64_t db0 = smem_desc(hiB2, gaddr);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)          // 16 j per step = two 8-row groups = +2048 B
                            umma_f16(tmem + s * NI, da0 + (uint64_t)(ks * 128), db0 + (uint64_t)(ks * 128), idesc2, (t | ks) != 0);
                        umma_commit(&empty[p.stage]);
                        p.advance();
                    }
                    umma_commit(&g_empty[b]);
                };
                if (ok) phase1(0);
                for (int t = 0; t < T && ok; ++t) {
                    if (t + 1 < T) phase1(t + 1);
                    if (ok) phase2(t);
                }
                umma_commit(acc_full);
            }
        } else {
            // ===================== epilogue warps (128 threads) =====================
            const int q = warp & 3;                       // TMEM lane quadrant this warp may access
            const int et = q * 32 + lane;                 // j row inside the tile == TMEM lane
            const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
            const float c = a.inv_tau * 1.4426950408889634f;    // log2(e) / tau
            if (et < NI) {
                const int64_t gi = row0 + et;
                // a'_i = 2^13 / rowsum_i = exp2((1/tau - lse_x) * log2e + 13)
                a_s[et] = gi < a.m ? exp2f((a.inv_tau - S.lse_x[gi]) * 1.4426950408889634f + kGScaleLog2) : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int t = 0; t < T && ok; ++t) {
                const int b = t & 1;
                const int64_t j0 = (t_begin + t) * TJ;
                const int64_t gj = j0 + et;
                ok = mbar_wait(&s_full[b], (t >> 1) & 1, kSpin);
                fence_after_thread_sync();
                ok = __all_sync(0xffffffffu, ok);         // keep the .aligned TMEM loads warp-uniform
                if (!ok) break;
                uint32_t r[NI];
                tmem_ld_32x32b_x32(tmem + lane_addr + kAccCols + b * NI, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                tmem_ld_32x32b_x32(tmem + lane_addr + kAccCols + b * NI + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                tmem_ld_wait();
                fence_before_thread_sync();
                mbar_arrive(&s_empty[b]);                 // S buffer may be overwritten by phase 1 of tile t+2
                const bool jvalid = gj < a.n;
                {
                    ok = ok && mbar_wait(&g_empty[b], ((t >> 1) & 1) ^ 1, kSpin);
                    const int64_t ly = S.lse_y_block > 0 ? (gj / S.lse_y_block) * S.lse_y_stride + gj % S.lse_y_block : gj;
                    const float bj = jvalid ? exp2f((a.inv_tau - S.lse_y[ly]) * 1.4426950408889634f + kGScaleLog2) : 0.f;
                    // diagonal element of this row, if it falls inside the CTA's 64 columns
                    const int64_t di = gj - a.row_offset - row0;      // column index i with global i == j
                    uint8_t* grow = gs + (size_t)b * kGBytes + (size_t)et * 128;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t packed[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float g2[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int i = ch * 8 + h * 2 + e;
                                float g = 0.f;
                                if (jvalid) {
                                    const float p = exp2f(fmaf(__uint_as_float(r[i]), c, -c));
                                    g = p * (a_s[i] + bj);
                                    if (di == i) g -= 16384.0f;      // 2 * 2^13
                                }
                                g2[e] = g;
                            }
                            const __half2 hh = __floats2half2_rn(g2[0], g2[1]);
                            packed[h] = *reinterpret_cast<const uint32_t*>(&hh);
                        }
                        // 128B swizzle: 16-byte chunk ch of row et lands at chunk (ch ^ (et & 7))
                        *reinterpret_cast<uint4*>(grow + ((ch ^ (et & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                    fence_proxy_async_smem();             // generic-proxy stores -> visible to the MMA (async proxy)
                    mbar_arrive(&g_full[b]);
                }
            }
            if (__all_sync(0xffffffffu, ok) && (ok = __all_sync(0xffffffffu, mbar_wait(acc_full, 0, kSpin)))) {
                fence_after_thread_sync();
                float* outp = S.out + (int64_t)blockIdx.y * a.m * a.d;
                for (int s = 0; s < NS; ++s) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(tmem + lane_addr + s * NI, r0);
                    tmem_ld_32x32b_x32(tmem + lane_addr + s * NI + 32, r1);
                    tmem_ld_wait();
                    const int dcol = s * 128 + et;            // TMEM lane = d index inside the block
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int64_t gi = row0 + i;
                        if (gi < a.m) outp[gi * a.d + dcol] = __uint_as_float(r0[i]) * a.out_scale;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int64_t gi = row0 + 32 + i;
                        if (gi < a.m) outp[gi * a.d + dcol] = __uint_as_float(r1[i]) * a.out_scale;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // empty column range (only possible with an over-split grid): contribute zeros
        const int et = (warp & 3) * 32 + lane;
        float* outp = S.out + (int64_t)blockIdx.y * a.m * a.d;
        for (int i = 0; i < NI; ++i)
            if (row0 + i < a.m)
                for (int c2 = et; c2 < d; c2 += 128) outp[(row0 + i) * a.d + c2] = 0.f;
    }
    if (!ok && lane == 0 && a.error_flag) atomicExch(a.error_flag, 1);
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 3) tmem_dealloc(tmem, 512);
}

// lse[i] = 1/tau + ln(sum over splits of the partial row sums)
__global__ void k_tc_lse_combine(const float* __restrict__ part, int nsplit, int64_t m, float inv_tau, float* __restrict__ lse) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float s = 0.f;
    for (int p = 0; p < nsplit; ++p) s += part[(int64_t)p * m + i];
    lse[i] = inv_tau + logf(s);
}

// dx[i, :] = sum over splits of the partial gradients (fixed order -> deterministic)
__global__ void k_tc_grad_combine(const float* __restrict__ part, int nsplit, int64_t count4, float* __restrict__ dx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count4) return;
    float4 acc = reinterpret_cast<const float4*>(part)[i];
    for (int p = 1; p < nsplit; ++p) {
        const float4 v = reinterpret_cast<const float4*>(part)[(int64_t)p * count4 + i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
}

__global__ void k_tc_grad_combine2(const float* __restrict__ part, int nsplit, int64_t count4, float* __restrict__ dx0, float* __restrict__ dx1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count4) return;
    const float4* src = reinterpret_cast<const float4*>(part) + (int64_t)blockIdx.y * nsplit * count4;
    float4 acc = src[i];
    for (int p = 1; p < nsplit; ++p) {
        const float4 v = src[(int64_t)p * count4 + i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(blockIdx.y ? dx1 : dx0)[i] = acc;
}

// diag[i] = x_hat_i . y_hat_{off+i} / tau from the fp16 operands (one warp per row)
__global__ void k_tc_diag(const __half* __restrict__ x, const __half* __restrict__ y, int64_t m, int64_t n, int64_t d, float inv_tau,
                          int64_t row_offset, float* __restrict__ diag) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= m) return;
    const int64_t j = row_offset + row;
    float acc = 0.f;
    if (j < n) {
        const __half2* xr = reinterpret_cast<const __half2*>(x + row * d);
        const __half2* yr = reinterpret_cast<const __half2*>(y + j * d);
        for (int64_t k = lane; k < d / 2; k += 32) {
            const float2 xa = __half22float2(xr[k]), ya = __half22float2(yr[k]);
            acc = fmaf(xa.x, ya.x, acc);
            acc = fmaf(xa.y, ya.y, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) diag[row] = acc * inv_tau;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_loss_shape_ok(int64_t d, float temperature) {
    return d % 128 == 0 && d >= 128 && d <= 768 && temperature >= 0.025f;
}

// The CTA-pair kernel (loss_tc_pair.cu) serves every d that splits into 256-wide blocks; DCLIP_BWD_SINGLE=1
// forces the single-CTA kernel above (A/B measurements, tests).
bool tc_pair_shape_ok(int64_t d);
int tc_pair_nsplit(int64_t m, int64_t n, int nsides);
int launch_infonce_bwd_pair(int nsides, const void* const x[2], const void* const y[2], const TcArgs& args, int nsplit, cudaStream_t st);
static bool use_pair_kernel(int64_t d) {
    static const bool forced_single = [] { const char* e = getenv("DCLIP_BWD_SINGLE"); return e && e[0] == '1'; }();
    return !forced_single && tc_pair_shape_ok(d);
}

int tc_nsplit(int64_t m, int64_t n, int64_t d, int nsides) {
    if (use_pair_kernel(d)) return tc_pair_nsplit(m, n, nsides);
    const int64_t row_ctas = ceil_div(m, NI) * nsides;
    const int64_t ntiles = ceil_div(n, TJ);
    int64_t want = sm_count() / (row_ctas > 0 ? row_ctas : 1);
    if (want < 1) want = 1;
    if (want > ntiles) want = ntiles;
    return (int)want;
}

static size_t tc_smem_bytes(int64_t d) {
    return 1024 + (size_t)((d + 63) / 64) * kXBoxBytes + (size_t)kSlots * kSlotBytes + 2 * kGBytes + 512;
}

// One pipeline-error int per device for the whole process (kernels launched from PyTorch's autograd thread report
// into the same flag the main thread reads).  Allocated on first use with cudaMalloc (legal while another stream is
// being captured in relaxed / thread-local mode; a global-mode capture must call dclip_debug_pipeline_error once
// before capturing) and cleared on a private non-blocking stream, never on the legacy stream.
int* tc_error_flag() {
    static std::mutex mu;
    static int* flag[16] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!flag[dev]) {
        int* p = nullptr;
        if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaStream_t s = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) {
            cudaMemsetAsync(p, 0, sizeof(int), s);
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        } else {
            cudaGetLastError();
            cudaMemset(p, 0, sizeof(int));
        }
        flag[dev] = p;
    }
    return flag[dev];
}

// x[side]: local rows [m, d] fp16; y[side]: all rows [n, d] fp16.  a.s[side].out = partial dx [nsplit][m][d].
int launch_infonce_bwd_tc(int nsides, const void* const x[2], const void* const y[2], const TcArgs& args_in, int nsplit, cudaStream_t st) {
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "tensor path needs an sm_100 device");
    if (use_pair_kernel(args_in.d)) return launch_infonce_bwd_pair(nsides, x, y, args_in, nsplit, st);
    TcArgs a = args_in;
    CUtensorMap tmx[2], tmy[2];
    for (int s = 0; s < 2; ++s) {
        const int src = s < nsides ? s : 0;
        if (!make_tmap_2d_sw128(&tmx[s], x[src], (uint64_t)a.m, (uint64_t)a.d, NI, 0) ||
            !make_tmap_2d_sw128(&tmy[s], y[src], (uint64_t)a.n, (uint64_t)a.d, TJ, 0))
            return fail(DCLIP_ERR_CUDA, "cuTensorMapEncodeTiled failed (m=%lld n=%lld d=%lld)", (long long)a.m, (long long)a.n, (long long)a.d);
    }
    const int64_t ntiles = ceil_div(a.n, TJ);
    a.tiles_per_split = (int)ceil_div(ntiles, nsplit);
    a.error_flag = tc_error_flag();
    const size_t smem = tc_smem_bytes(a.d);
    dim3 grid((unsigned)ceil_div(a.m, NI), (unsigned)nsplit, (unsigned)nsides);
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_infonce_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_infonce_bwd_tc<<<grid, kTcThreads, smem, st>>>(tmx[0], tmy[0], tmx[1], tmy[1], a);
    DCLIP_LAUNCH_OK("k_infonce_bwd_tc");
    return DCLIP_OK;
}

// ---- single-GPU backward through a MATERIALISED coefficient matrix ----------------------------------------------
// G = softmax_row + softmax_col - 2 I (B x B, fp16, scaled by 2^13) is written once by the streaming kernel (mode 3:
// the score tiles are recomputed once, 2 B^2 d) and both gradients are plain GEMMs over it,
//     dU_hat = alpha G V_hat        (A = G K-major,  B = V_hat stored [K, N])
//     dV_hat = alpha G^T U_hat      (A = G read MN-major, B = U_hat stored [K, N])
// 6 B^2 d executed instead of the 8 B^2 d of the fused CTA-pair kernel, which recomputes every score tile per side
// (and stays the kernel of the sharded path, where that recompute is what removes the [B, d] reduce-scatter).
// G (32 MB at B = 4096) lives in L2 between its writer and its two readers.
bool tc_gmat_ok(int64_t B, int64_t d) {
    const char* e = getenv("DCLIP_BWD_GMAT");      // "0": keep the fused CTA-pair kernel (A/B measurements, tests); read per call
    const bool off = e && e[0] == '0';
    return !off && B > 128 && d % 8 == 0 && tc_gmat_bytes(B) <= ((size_t)1 << 30);
}
size_t tc_gmat_bytes(int64_t B) {
    return (size_t)(ceil_div(B, 128) * 128) * (size_t)(ceil_div(B, 256) * 256) * 2;
}
size_t tc_gmat_dot_floats(int64_t B) { return (size_t)(stream_nsplit(B, B, 1) + ceil_div(B, 128)) * (size_t)B; }
int64_t tc_gmat_ld(int64_t B) { return ceil_div(B, 256) * 256; }
// stage 1: G and the row / column sums of g * s.  dot_ws: [ns][B] partial row sums (u side: x_hat_i . dU_hat_i / alpha),
// then [ceil(B / 128)][B] partial column sums (v side)
int launch_gmat_coeffs(const void* u_hat, const void* v_hat, int64_t B, int64_t d, const float* lse_u, const float* lse_v, float inv_tau,
                       void* g_buf, float* dot_ws, cudaStream_t st) {
    const int ns = stream_nsplit(B, B, 1);
    TcArgs a{};
    a.s[0] = TcSide{lse_u, lse_v, dot_ws};
    a.col_out = dot_ws + (size_t)ns * B;
    a.m = B; a.n = B; a.d = d; a.inv_tau = inv_tau; a.row_offset = 0;
    a.g_out = g_buf; a.ldg = tc_gmat_ld(B);
    const void* xs[2] = {u_hat, u_hat};
    const void* ys[2] = {v_hat, v_hat};
    return launch_stream_tc(3, 1, xs, ys, a, ns, st);
}
// stage 2: dU_hat = alpha G V_hat, dV_hat = alpha G^T U_hat; with `tail` the epilogue finishes the gradients (needs dot_ws)
int launch_gmat_gemms(const void* g_buf, const float* dot_ws, const void* u_hat, const void* v_hat, int64_t B, int64_t d, float out_scale,
                      float* du_hat, float* dv_hat, const GmatTail* tail, cudaStream_t st) {
    const int64_t ldg = tc_gmat_ld(B);
    const int ns = stream_nsplit(B, B, 1), nib = (int)ceil_div(B, 128);
    TcGemmProblem pr[2] = {};
    pr[0].A = g_buf; pr[0].lda = (int)ldg; pr[0].a_mn = 0; pr[0].B = v_hat; pr[0].ldb = (int)d; pr[0].b_mn = 1;
    pr[0].C = du_hat; pr[0].ldc = (int)d; pr[0].M = (int)B; pr[0].N = (int)d; pr[0].K = (int)B; pr[0].alpha = out_scale;
    pr[1] = pr[0];
    pr[1].a_mn = 1; pr[1].B = u_hat; pr[1].C = dv_hat;
    const char* e = getenv("DCLIP_GEMM_PAIR");      // "0": the 128 x 256 single-CTA grouped GEMM (A/B measurements; no fused tail)
    if (e && e[0] == '0') {
        if (tail) return fail(DCLIP_ERR_INVALID, "bwd_gmat: the fused gradient tail needs the CTA-pair GEMM");
        return launch_gemm_tc(pr, 2, DCLIP_F16, st);
    }
    if (!tail) return launch_gemm_pair(pr, 2, DCLIP_F16, st);
    GemmPairTail gt[2];
    for (int sd = 0; sd < 2; ++sd) {
        const GmatTailSide& t = tail->s[sd];
        gt[sd] = GemmPairTail{t.x, t.t, t.inv_norm, t.cos, t.t_inv, sd == 0 ? dot_ws : dot_ws + (size_t)ns * B, sd == 0 ? ns : nib, (long long)B,
                              t.d_x, t.w_cos_over_B};
        pr[sd].C = nullptr;
    }
    return launch_gemm_pair(pr, 2, DCLIP_F16, st, gt, tail->grad_scale, tail->dtype);
}
int launch_infonce_bwd_gmat(const void* u_hat, const void* v_hat, int64_t B, int64_t d, const float* lse_u, const float* lse_v, float inv_tau,
                            float out_scale, void* g_buf, float* dot_ws, float* du_hat, float* dv_hat, const GmatTail* tail, cudaStream_t st) {
    if (int e = launch_gmat_coeffs(u_hat, v_hat, B, d, lse_u, lse_v, inv_tau, g_buf, dot_ws, st)) return e;
    return launch_gmat_gemms(g_buf, dot_ws, u_hat, v_hat, B, d, out_scale, du_hat, dv_hat, tail, st);
}

int launch_tc_lse_combine(const float* part, int nsplit, int64_t m, float inv_tau, float* lse, cudaStream_t st) {
    k_tc_lse_combine<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(part, nsplit, m, inv_tau, lse);
    DCLIP_LAUNCH_OK("k_tc_lse_combine");
    return DCLIP_OK;
}
int launch_tc_grad_combine(const float* part, int nsplit, int64_t m, int64_t d, float* dx, cudaStream_t st) {
    const int64_t count4 = m * d / 4;
    k_tc_grad_combine<<<(unsigned)ceil_div(count4, 256), 256, 0, st>>>(part, nsplit, count4, dx);
    DCLIP_LAUNCH_OK("k_tc_grad_combine");
    return DCLIP_OK;
}
int launch_tc_grad_combine2(const float* part, int nsplit, int64_t m, int64_t d, float* dx0, float* dx1, cudaStream_t st) {
    const int64_t count4 = m * d / 4;
    k_tc_grad_combine2<<<dim3((unsigned)ceil_div(count4, 256), 2), 256, 0, st>>>(part, nsplit, count4, dx0, dx1);
    DCLIP_LAUNCH_OK("k_tc_grad_combine2");
    return DCLIP_OK;
}
int launch_tc_diag(const void* x, const void* y, int64_t m, int64_t n, int64_t d, float inv_tau, int64_t row_offset, float* diag,
                   cudaStream_t st) {
    k_tc_diag<<<(unsigned)ceil_div(m, 8), 256, 0, st>>>(static_cast<const __half*>(x), static_cast<const __half*>(y), m, n, d, inv_tau,
                                                        row_offset, diag);
    DCLIP_LAUNCH_OK("k_tc_diag");
    return DCLIP_OK;
}

// Reads (and clears) the pipeline-timeout flag.  Synchronises the stream: debug / test use only.
int tc_check_error_flag(cudaStream_t st) {
    int* f = tc_error_flag();
    if (!f) return 0;
    int h = 0;
    if (cudaMemcpyAsync(&h, f, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (h) cudaMemsetAsync(f, 0, sizeof(int), st);
    return h;
}

}  // namespace dclip
