// Streaming score-tile kernel (tcgen05 / TMEM / TMA): row log-sum-exp of the InfoNCE forward and the
// approximate top-32 candidate lists of the retrieval path.
//
// Both consume S = X . Y^T tile by tile without keeping it: a CTA owns 128 rows of X (one TMEM lane /
// epilogue thread each) and walks over 256-row tiles of Y:
//     S[128 i x 256 j] = X_I . Y_J^T        (M = 128, N = 256, K = d in 64-column stages, fp16 in, fp32 acc)
// The 128 x 256 tile shape is chosen for shared-memory bandwidth: tcgen05.mma re-reads both operands
// from shared memory for every K = 16 step ((128 + N) * 32 B), and TMA writes share the same 128 B/clk,
// so N = 256 moves 1.7x fewer shared-memory bytes per MAC than the 128 x 64 tiles of the backward
// kernel (which is pinned to N = 64 by its TMEM-resident gradient accumulator; measured with
// tools/tc_bench: N = 64 -> 53 cycles/MMA, N = 256 -> 132).  Neither operand is resident: a stage
// of the 4-deep ring holds one X box (128 x 64, 16 KB) and one Y box (256 x 64, 32 KB); X_I is only
// 128 x d and stays hot in L2.
//
// Warps: 0-3 TMA producers (one ring stage each), 4 MMA issuer (+ TMEM alloc), 5-8 epilogue (mode 1: 5-12,
// two warps per lane quadrant with 128 of the tile's 256 columns each).
// TMEM: two 256-column S buffers, so the epilogue of tile t overlaps the MMAs of tile t+1.
//   mode 0 (loss forward):  thread i adds exp2((s - 1) * log2e / tau) over the valid columns -> out[i]
//   mode 1 (loss forward, both directions from ONE pass over S; single-GPU objective where the column
//                           direction is the other side's row direction): row sums as mode 0, plus the column
//                           sums of the same p tile: a 32 x 32 butterfly (31 shuffles) leaves lane l with the
//                           sum of column l over the warp's 32 rows, the four warps meet in shared memory and
//                           the tile's 256 column sums go to col_out[row block][j] (summed over row blocks,
//                           in fixed order, by the finalize kernel)
//   mode 4 (Recall@K ranks): thread i holds the EXACT score ts_i of its query's target; a score tile entry counts when it beats
//                           ts_i by more than the operand-rounding bound eps, entries inside [ts_i - eps, ts_i + eps] go to a
//                           global band list and are decided by exact re-scoring (retrieval_tc.cu): ranks stay bit-exact
//   mode 2 (retrieval):     thread i caches the admission threshold of its query; a warp-wide max test
//                           rejects whole 32-score chunks, the rare hits are inserted warp-cooperatively
//                           into that query's 32-entry list in shared memory -> cand_s / cand_i[i][32]
#include <cstdlib>

#include "common.cuh"
#include "loss_internal.h"
#include "tcgen05.cuh"

namespace dclip {
using namespace tc;

// Optional stall accounting (build with -DDCLIP_PAIR_PROF): cycles CTA (0,0,0) spends in each wait.
#ifdef DCLIP_PAIR_PROF
__device__ unsigned long long g_stream_prof[16];
#define SPROF_BEGIN(v) const long long v = clock64()
#define SPROF_END(v, slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) atomicAdd(&g_stream_prof[(slot) + (kMode == 3 ? 8 : 0)], (unsigned long long)(clock64() - v)); } while (0)
#else
#define SPROF_BEGIN(v) do { } while (0)
#define SPROF_END(v, slot) do { } while (0)
#endif

constexpr int SM_ROWS = 128;             // X rows per CTA (MMA M)
constexpr int SN_COLS = 256;             // Y rows per tile (MMA N)
constexpr int kStagesSingle = 4;          // single-CTA ring: X box + 256-row Y box = 48 KB per stage
constexpr int kStagesPair = 6;            // CTA-pair ring: X box + MY 128 rows of the Y tile = 32 KB per stage
constexpr int kXStageBytes = SM_ROWS * 128;   // 16 KB
constexpr int kYStageBytes = SN_COLS * 128;   // 32 KB
// mode 3 gives one ring stage to the epilogue's staging buffers (8 warps x 4 KB)
constexpr int stream_stages(int mode, bool pair) { return pair ? (mode == 3 ? kStagesPair - 1 : kStagesPair) : kStagesSingle; }
constexpr int stream_stage_bytes(bool pair) { return kXStageBytes + (pair ? kYStageBytes / 2 : kYStageBytes); }
// mode 1: two warps per TMEM lane quadrant, 128 columns each; mode 3: four per quadrant, 64 columns each
constexpr int stream_epi_warps(int mode) { return mode == 3 ? 16 : mode == 1 ? 8 : 4; }
constexpr int stream_threads(int mode, bool pair) { return (stream_stages(mode, pair) + 1 + stream_epi_warps(mode)) * 32; }
constexpr int kGStageWarpBytes = 2048;    // mode 3: per-warp staging buffer (32 rows x 64 B, 16-byte chunks XOR-swizzled by the row pair)
constexpr int kGStageWarps = 16;
constexpr int stream_extra_bytes(int mode) {   // shared memory between the ring and the barriers
    return mode == 2 ? 2 * 32 * SM_ROWS * 4 : mode == 1 ? 2 * 4 * SN_COLS * 4 : mode == 3 ? 2 * SN_COLS * 4 + kGStageWarps * kGStageWarpBytes + 2 * 4 * SN_COLS * 4 : 0;
}
constexpr int kListLen = 32;
constexpr uint32_t kSpinS = 1u << 24;

// exponent arguments stay inside [-2 log2e / tau, 0] (tau >= 0.025): no denormal handling needed
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// kPair: the two CTAs of a (2,1,1) cluster share every 256-row Y tile (cta_group::2, M = 256 = both CTAs' 128 X rows,
// N = 256 = 128 Y rows from each CTA's shared memory): each CTA streams only HALF of Y, a stage shrinks from 48 to
// 32 KB and the ring deepens from 4 to 6 stages.  Row mapping is unchanged (CTA x owns X rows [128 x, 128 x + 128)).
// order-preserving float <-> uint32 map (for redux.sync.min on scores; -inf sorts below every finite value)
__device__ __forceinline__ unsigned float_to_ord(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord_to_float(unsigned u) { return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

template <int kMode, bool kPair>
__global__ void __launch_bounds__(stream_threads(kMode, kPair), 1)
k_stream_tc(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmy0,
            const __grid_constant__ CUtensorMap tmx1, const __grid_constant__ CUtensorMap tmy1, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int side = blockIdx.z;
    const CUtensorMap* tmx = side ? &tmx1 : &tmx0;
    const CUtensorMap* tmy = side ? &tmy1 : &tmy0;
    const TcSide S = a.s[side];
    const int nkb = (int)((a.d + 63) / 64);
    constexpr int kStages = stream_stages(kMode, kPair);
    constexpr int kStageBytes = stream_stage_bytes(kPair);
    const uint32_t cr = kPair ? cluster_ctarank() : 0u;

    uint8_t* ring = smem;                                             // kStages x (48 | 32) KB
    uint8_t* lists = ring + kStages * kStageBytes;                    // mode 2: [32][128] float + [32][128] int32 = 32 KB
    uint8_t* tail = lists + stream_extra_bytes(kMode);                // mode 1: colsum[2][4][256]; mode 3: b[2][256] + staging
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);               // [kStages]
    uint64_t* empty = full + kStages;                                 // [kStages]
    uint64_t* s_full = empty + kStages;                               // [2]
    uint64_t* s_empty = s_full + 2;                                   // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    SPROF_BEGIN(p_kernel);
    const int64_t row0 = (int64_t)blockIdx.x * SM_ROWS;
    const int64_t ntiles = (a.n + SN_COLS - 1) / SN_COLS;
    const int64_t t_begin = (int64_t)blockIdx.y * a.tiles_per_split;
    const int64_t t_end = min(ntiles, t_begin + a.tiles_per_split);
    const int T = (int)max((int64_t)0, t_end - t_begin);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        // s_empty: every epilogue thread (single CTA) / one lane per epilogue warp of BOTH CTAs (pair; leader's copy)
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], kPair ? 2 * stream_epi_warps(kMode) : stream_epi_warps(kMode) * 32); }
        fence_barrier_init();
    }
    if (warp == kStages) { if (kPair) tmem_alloc_pair(tmem_ptr, 512); else tmem_alloc(tmem_ptr, 512); }
    fence_before_thread_sync();
    __syncthreads();
    if (kPair) cluster_sync_all();          // peer barriers initialised, both TMEMs allocated
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;
    if (threadIdx.x == 0) SPROF_END(p_kernel, 6);

    if (warp < kStages) {
        // ===================== TMA producers: warp w owns ring stage w =====================
        if (T > 0 && elect_one()) {
            tma_prefetch_desc(tmx);
            tma_prefetch_desc(tmy);
            if (kMode == 0 && a.peers.world > 1) {
                // sharded forward: the Y rows of the other ranks are written into my buffer by THEIR prep kernels; wait for every
                // peer's arrival (system-scope), then order those generic-proxy writes before my async-proxy (TMA) reads
                const unsigned int target = p2p_ld_volatile(&p2p_flags(a.peers, a.peers.rank)->expected[kP2PKindOperands]) + 1;
                p2p_wait_all(a.peers, kP2PKindOperands, target, a.error_flag);
                fence_proxy_async_all();
            }
            const int njobs = T * nkb;
            uint8_t* dst = ring + (size_t)warp * kStageBytes;
            uint32_t phase = 0;
            // When Y does not fit in L2 (retrieval galleries) the row-block CTAs of a split all want the same
            // boxes, so only kStages distinct boxes per split are in flight and the stream runs at DRAM
            // LATENCY.  Each Y tile is therefore prefetched into L2 kPrefetchAhead tiles early by exactly one
            // of the gridDim.x CTAs that walk this split (the others then hit L2).
            constexpr int kPrefetchAhead = 16;
            const bool prefetch = a.n * a.d * 2 > (int64_t)32 * 1024 * 1024;
            for (int q = warp; q < njobs && ok; q += kStages, phase ^= 1) {
                const int tile = q / nkb, kb = q - tile * nkb;
                const int32_t j0 = (int32_t)((t_begin + tile) * SN_COLS);
                if (prefetch && kb == 0) {
                    const int first = tile == 0 ? 0 : kPrefetchAhead;        // the first tile also covers 0 .. ahead-1
                    for (int pt = tile + first; pt <= tile + kPrefetchAhead && pt < T; ++pt)
                        if (pt % (int)gridDim.x == (int)blockIdx.x)
                            for (int k2 = 0; k2 < nkb; ++k2) tma_prefetch_l2_2d(tmy, k2 * 64, (int32_t)((t_begin + pt) * SN_COLS));
                }
                if (a.sync_ctr && kb == 0 && tile % a.sync_every == 0) {
                    // soft lockstep: CTAs that run concurrently (same group of consecutive block ids) walk the gallery at
                    // the same pace, so a Y tile fetched from HBM once serves all of them out of L2.  Purely a
                    // performance device: the wait is bounded (~100 us) and the result never depends on it.
                    const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
                    const unsigned grp = lin / (unsigned)a.sync_group;
                    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
                    const unsigned members = min((unsigned)a.sync_group, total - grp * (unsigned)a.sync_group);
                    unsigned int* c = a.sync_ctr + (size_t)grp * a.sync_epochs + tile / a.sync_every;
                    atomicAdd(c, 1u);
                    const long long t0 = clock64();
                    while (*reinterpret_cast<volatile unsigned int*>(c) < members && clock64() - t0 < 200000) __nanosleep(200);
                }
                SPROF_BEGIN(p_em);
                ok = mbar_wait(&empty[warp], phase ^ 1, kSpinS);
                if (warp == 0) SPROF_END(p_em, 5);
                if (kPair) {
                    // both CTAs' bytes are credited to the leader's barrier; my B half = Y rows [j0 + 128 cr, + 128)
                    const uint32_t fbar = mapa_u32(smem_u32(&full[warp]), 0);
                    if (cr == 0) mbar_arrive_expect_tx(&full[warp], 2 * kStageBytes);
                    tma_load_2d_pair(dst, tmx, fbar, kb * 64, (int32_t)row0);
                    tma_load_2d_pair(dst + kXStageBytes, tmy, fbar, kb * 64, j0 + (int32_t)cr * 128);
                } else {
                    mbar_arrive_expect_tx(&full[warp], kStageBytes);
                    tma_load_2d(dst, tmx, &full[warp], kb * 64, (int32_t)row0);
                    tma_load_2d(dst + kXStageBytes, tmy, &full[warp], kb * 64, j0);
                }
            }
        }
    } else if (warp == kStages) {
        // ===================== MMA issuer =====================
        if (T > 0 && cr == 0 && elect_one()) {
            const uint32_t idesc = make_idesc_f16(kPair ? 2 * SM_ROWS : SM_ROWS, SN_COLS, kFmtF16, 0, 0);
            const uint64_t hiK = make_smem_desc_hi(0, 1024, kSwizzle128B);
            const uint32_t ring_addr = smem_u32(ring);
            uint32_t stage = 0, phase = 0;
            SPROF_BEGIN(p_all);
            for (int t = 0; t < T && ok; ++t) {
                const int b = t & 1;
                SPROF_BEGIN(p_se);
                ok = mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1, kSpinS);
                SPROF_END(p_se, 1);
                fence_after_thread_sync();
                const uint32_t dcol = tmem + b * SN_COLS;
                for (int kb = 0; kb < nkb && ok; ++kb) {
                    SPROF_BEGIN(p_fu);
                    ok = mbar_wait(&full[stage], phase, kSpinS);
                    SPROF_END(p_fu, 2);
                    fence_after_thread_sync();
                    const uint64_t da0 = smem_desc(hiK, ring_addr + stage * kStageBytes);
                    const uint64_t db0 = smem_desc(hiK, ring_addr + stage * kStageBytes + kXStageBytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (kPair) umma_f16_pair(dcol, da0 + (uint64_t)(ks * 2), db0 + (uint64_t)(ks * 2), idesc, (kb | ks) != 0);
                        else umma_f16(dcol, da0 + (uint64_t)(ks * 2), db0 + (uint64_t)(ks * 2), idesc, (kb | ks) != 0);
                    }
                    if (kPair) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (kPair) umma_commit_pair(&s_full[b]); else umma_commit(&s_full[b]);
            }
            SPROF_END(p_all, 0);
        }
    } else {
        // ===================== epilogue: thread = one X row =====================
        const int qd = warp & 3;                          // TMEM lane quadrant
        const int et = qd * 32 + lane;                    // row inside the CTA block
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const int64_t grow = row0 + et;
        const float c = a.inv_tau * 1.4426950408889634f;
        float sum = 0.f;
        float* colsum = reinterpret_cast<float*>(lists);  // mode 1
        const float rowvalid = grow < a.m ? 1.f : 0.f;    // rows past m are TMA zero-filled: keep them out of the column sums
        const int chalf = kMode == 1 ? (warp - (kStages + 1)) >> 2 : 0;     // mode 1: my 128-column half of the tile
        const int cpart = kMode == 3 ? (warp - (kStages + 1)) >> 2 : 0;     // mode 3: my 64-column quarter of the tile
        (void)colsum; (void)rowvalid;
        // retrieval state: per X row an unsorted list of its 32 best approximate scores in shared memory,
        // laid out [row][32] so that a whole warp can work on one row's list without bank conflicts; the
        // owning thread caches the list minimum (= admission threshold) in a register.
        float* ls = reinterpret_cast<float*>(lists);
        int32_t* li = reinterpret_cast<int32_t*>(lists + kListLen * SM_ROWS * 4);
        float mn = -INFINITY;
        // mode 3 (gradient coefficients): a_i = 2^13 / rowsum_i of my row; b_j per tile in shared memory
        constexpr float kLog2e = 1.4426950408889634f;
        float a_i = 0.f;
        float* bs = reinterpret_cast<float*>(lists);                                     // [2][256]
        const int we = warp - (kStages + 1);
        const uint32_t stage_s = smem_u32(lists + 2 * SN_COLS * 4) + (uint32_t)we * kGStageWarpBytes;
        if (kMode == 3 && grow < a.m) a_i = exp2f((a.inv_tau - S.lse_x[grow]) * kLog2e + 13.0f);
        float* csum3 = reinterpret_cast<float*>(lists + 2 * SN_COLS * 4 + kGStageWarps * kGStageWarpBytes);     // [2][4][256] column sums of g * s
        (void)a_i; (void)bs; (void)stage_s; (void)csum3; (void)cpart;
        // mode 4: my query's exact target score, the band around it and the (local) column of the target itself
        float rk_hi = INFINITY, rk_lo = INFINITY;
        int64_t rk_tcol = -1;
        int rk_cnt = 0;
        if (kMode == 4 && grow < a.m) {
            const float ts = a.rk_thr[grow];
            const float eps = a.rk_eps * (a.rk_norms ? a.rk_norms[0] * a.rk_norms[1] : 1.f);
            rk_hi = ts + eps; rk_lo = ts - eps;
            rk_tcol = a.rk_tgt[grow] - a.rk_goff;
        }
        (void)rk_hi; (void)rk_lo; (void)rk_tcol; (void)rk_cnt;
        if (kMode == 2) {
            for (int e = 0; e < kListLen; ++e) { ls[et * kListLen + e] = -INFINITY; li[et * kListLen + e] = -1; }
            __syncwarp();
        }
        for (int t = 0; t < T && ok; ++t) {
            const int b = t & 1;
            const int64_t j0 = (t_begin + t) * SN_COLS;
            const int ncols = (int)min((int64_t)SN_COLS, a.n - j0);
            if (kMode == 3) {
                // column factors of this tile: b_j = 2^13 / colsum_j (buffer b is rewritten two tiles later, after every
                // thread has passed the next tile's barrier)
                const int idx = we * 32 + lane;
                const int64_t gj = j0 + idx;
                if (idx < SN_COLS) bs[b * SN_COLS + idx] = gj < a.n ? exp2f((a.inv_tau - S.lse_y[gj]) * kLog2e + 13.0f) : 0.f;
                asm volatile("bar.sync 2, 512;" ::: "memory");
            }
            SPROF_BEGIN(p_sf);
            ok = mbar_wait(&s_full[b], (t >> 1) & 1, kSpinS);
            if (threadIdx.x == (kStages + 1) * 32) SPROF_END(p_sf, 3);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            fence_after_thread_sync();
            SPROF_BEGIN(p_ep);
            if (kMode == 3) {
                // G tile: g_ij = p_ij (a_i + b_j) - 2 [i == j], scaled by 2^13, rounded to fp16 and written to the B x B
                // coefficient matrix.  TMEM gives a thread one ROW; 32-column chunks (64 B of fp16 per row) are staged through
                // the warp's swizzled 2 KB buffer so that global stores are 4 lanes per row (two full sectors).  Sixteen
                // warps: the tile's epilogue (exp, scaling, rounding, the sums below) must stay under its 12 k-blocks of MMAs.
                const int rb_row = lane >> 2, rb_chunk = lane & 3;
                const int64_t rbase = row0 + qd * 32;
                uint16_t* gout = static_cast<uint16_t*>(a.g_out);
#pragma unroll 1
                for (int c0 = cpart * 64; c0 < cpart * 64 + 64; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem + lane_addr + b * SN_COLS + c0, r);
                    tmem_ld_wait();
                    const uint32_t b_addr = smem_u32(bs) + (uint32_t)(b * SN_COLS + c0) * 4;
                    // the diagonal of S crosses this warp's 32 x 32 chunk? (warp-uniform, rare)
                    const int64_t dc = a.row_offset + grow - (j0 + c0);       // my diagonal column inside the chunk, if in [0, 32)
                    const bool has_diag = a.row_offset + rbase < j0 + c0 + 32 && a.row_offset + rbase + 32 > j0 + c0;
                    float v[32];
                    uint32_t packed[16];
#pragma unroll
                    for (int q8 = 0; q8 < 4; ++q8) {
                        float4 b0, b1;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(b_addr + q8 * 32));
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(b_addr + q8 * 32 + 16));
                        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        float gv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) gv[i] = ex2_ftz(fmaf(__uint_as_float(r[q8 * 8 + i]), c, -c)) * (a_i + bv[i]);
                        if (has_diag) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) gv[i] += (dc == q8 * 8 + i) ? -16384.0f : 0.f;
                        }
                        // g * s with the ROUNDED coefficient (what the gradient GEMMs read): its row / column sums are the
                        // radial components x_hat_i . dX_hat_i of the two gradients (normalise backward, fused GEMM epilogue)
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const __half2 hv = __floats2half2_rn(gv[2 * h], gv[2 * h + 1]);
                            packed[q8 * 4 + h] = *reinterpret_cast<const uint32_t*>(&hv);
                            const float2 gr = __half22float2(hv);
                            v[q8 * 8 + 2 * h] = gr.x * __uint_as_float(r[q8 * 8 + 2 * h]);
                            v[q8 * 8 + 2 * h + 1] = gr.y * __uint_as_float(r[q8 * 8 + 2 * h + 1]);
                            sum += v[q8 * 8 + 2 * h] + v[q8 * 8 + 2 * h + 1];
                        }
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_s + lane * 64 + ((q4 ^ ((lane >> 1) & 3)) << 4)),
                                     "r"(packed[q4 * 4]), "r"(packed[q4 * 4 + 1]), "r"(packed[q4 * 4 + 2]), "r"(packed[q4 * 4 + 3]) : "memory");
                    __syncwarp();
                    // rows / columns past m / n land in the padding of the coefficient matrix (ld and row count are rounded
                    // up to whole tiles); its readers never look there (TMA bounds = m x n)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = rb_row + 8 * k;
                        uint4 x;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                                     : "r"(stage_s + rr * 64 + ((rb_chunk ^ ((rr >> 1) & 3)) << 4)) : "memory");
                        if (row0 < a.m)      // (the padded CTA of an odd pair owns no rows of the matrix at all)
                            *reinterpret_cast<uint4*>(gout + (size_t)(rbase + rr) * a.ldg + j0 + c0 + rb_chunk * 8) = x;
                    }
                    __syncwarp();
                    // column sums of g * s over my warp's 32 rows: a 32 x 32 transpose-reduce (as in mode 1)
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float send = upper ? v[i] : v[i + off];
                            const float keep = upper ? v[i + off] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    csum3[(b * 4 + qd) * SN_COLS + c0 + lane] = v[0];
                }
            }
#pragma unroll 1
            for (int c0 = chalf * 128; kMode != 3 && c0 < (kMode == 1 ? chalf * 128 + 128 : SN_COLS); c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + lane_addr + b * SN_COLS + c0, r);
                tmem_ld_wait();
                if (c0 >= ncols) continue;                // warp-uniform: whole chunk beyond the last Y row
                if (kMode == 1) {
                    // p for the row sum and, through a 32 x 32 transpose-reduce, for the column sums
                    float v[32];
                    if (c0 + 32 <= ncols) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float p = ex2_ftz(fmaf(__uint_as_float(r[i]), c, -c));
                            sum += p;
                            v[i] = p * rowvalid;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float p = (c0 + i < ncols) ? ex2_ftz(fmaf(__uint_as_float(r[i]), c, -c)) : 0.f;
                            sum += p;
                            v[i] = p * rowvalid;
                        }
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float send = upper ? v[i] : v[i + off];
                            const float keep = upper ? v[i + off] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    colsum[(b * 4 + qd) * SN_COLS + c0 + lane] = v[0];     // column c0 + lane over my warp's 32 rows
                } else if (kMode == 4) {
                    // definite wins are counted; anything the fp16 operands cannot decide is listed for the exact pass
                    const int lim = min(32, ncols - c0);
                    unsigned inband = 0u;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float sv = __uint_as_float(r[i]);
                        const bool valid = i < lim;
                        rk_cnt += (valid && sv > rk_hi) ? 1 : 0;
                        inband |= (valid && sv >= rk_lo && sv <= rk_hi) ? (1u << i) : 0u;
                    }
                    while (inband) {                          // rare (the target itself + ~0.2 neighbours per query on Flickr-shaped data)
                        const int i = __ffs(inband) - 1;
                        inband &= inband - 1;
                        const int64_t col = j0 + c0 + i;
                        if (col == rk_tcol) continue;         // the target never outranks itself
                        const unsigned pos = atomicAdd(a.rk_ctr, 1u);
                        if (pos < a.rk_cap) a.rk_band[pos] = make_int2((int)grow, (int)col);
                    }
                } else if (kMode == 0) {
                    if (c0 + 32 <= ncols) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum += exp2f(fmaf(__uint_as_float(r[i]), c, -c));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c0 + i < ncols) sum += exp2f(fmaf(__uint_as_float(r[i]), c, -c));
                    }
                } else {
                    // fast reject: does any of this thread's 32 scores beat its threshold?  (one max per
                    // score, no divergence; almost always "no" for the whole warp once the lists are warm)
                    float bg[4];                              // maxima of the four 8-score groups (the slow path skips clean groups)
                    if (c0 + 32 <= ncols) {                   // full chunk (all but the last tile): no column predicate
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            bg[g] = __uint_as_float(r[8 * g]);
#pragma unroll
                            for (int ii = 1; ii < 8; ++ii) bg[g] = fmaxf(bg[g], __uint_as_float(r[8 * g + ii]));
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            bg[g] = -INFINITY;
#pragma unroll
                            for (int ii = 0; ii < 8; ++ii) bg[g] = fmaxf(bg[g], (c0 + 8 * g + ii < ncols) ? __uint_as_float(r[8 * g + ii]) : -INFINITY);
                        }
                    }
                    const float best = fmaxf(fmaxf(bg[0], bg[1]), fmaxf(bg[2], bg[3]));
                    if (__any_sync(0xffffffffu, best > mn)) {
                        // slow path: the warp inserts the hits of one thread after the other, cooperatively
                        // (lane e looks at entry e of that row's list: replace the minimum, find the new one)
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                          if (!__any_sync(0xffffffffu, bg[g] > mn)) continue;          // whole 8-score group below every threshold
#pragma unroll
                          for (int ii = 0; ii < 8; ++ii) {    // (fully unrolled: r[] must stay in registers)
                            const int i = 8 * g + ii;
                            const float sv = __uint_as_float(r[i]);
                            unsigned hits = __ballot_sync(0xffffffffu, sv > mn && c0 + i < ncols);
                            while (hits) {
                                const int src = __ffs(hits) - 1;
                                hits &= hits - 1;
                                const float cs = __shfl_sync(0xffffffffu, sv, src);
                                const int row = qd * 32 + src;
                                float v = ls[row * kListLen + lane];
                                const float m = __shfl_sync(0xffffffffu, mn, src);     // the row's minimum = its owner's cached threshold
                                const int pos = __ffs(__ballot_sync(0xffffffffu, v == m)) - 1;
                                if (lane == pos) {
                                    v = cs;
                                    ls[row * kListLen + lane] = cs;
                                    li[row * kListLen + lane] = (int32_t)(j0 + c0 + i);
                                }
                                const float m2 = ord_to_float(__reduce_min_sync(0xffffffffu, float_to_ord(v)));     // one REDUX instead of 5 shuffles
                                if (lane == src) mn = m2;
                                __syncwarp();
                            }
                          }
                        }
                    }
                }
            }
            fence_before_thread_sync();
            if (kPair) {                                  // one (relaxed: no data published) arrival per warp on the leader's barrier
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&s_empty[b]), 0));
            } else {
                mbar_arrive(&s_empty[b]);                 // all of this thread's TMEM reads of buffer b are done
            }
            if (threadIdx.x == (kStages + 1) * 32) SPROF_END(p_ep, 4);
            if (kMode == 3) {
                asm volatile("bar.sync 2, 512;" ::: "memory");
                const int col = we * 32 + lane;
                if (col < ncols && row0 < a.m) {
                    const float* cs = csum3 + b * 4 * SN_COLS + col;
                    a.col_out[(int64_t)blockIdx.x * a.n + j0 + col] = (cs[0] + cs[SN_COLS]) + (cs[2 * SN_COLS] + cs[3 * SN_COLS]);
                }
            }
            if (kMode == 1) {
                // the four warps' partial column sums meet here (buffer b is rewritten two tiles later, after
                // every thread has passed the next tile's barrier)
                asm volatile("bar.sync 2, 256;" ::: "memory");
                const int col = chalf * 128 + et;
                if (col < ncols && row0 < a.m) {          // (a padded pair CTA past the last row block writes nothing)
                    const float* cs = colsum + b * 4 * SN_COLS + col;
                    a.col_out[(int64_t)blockIdx.x * a.n + j0 + col] = (cs[0] + cs[SN_COLS]) + (cs[2 * SN_COLS] + cs[3 * SN_COLS]);
                }
            }
        }
        if (kMode == 1) {
            // the two column halves of a row meet in shared memory (colsum is free after the last tile's barrier)
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (chalf == 1) colsum[et] = sum;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (chalf == 0 && grow < a.m) S.out[(int64_t)blockIdx.y * a.m + grow] = sum + colsum[et];
        } else if (kMode == 0) {
            if (grow < a.m) S.out[(int64_t)blockIdx.y * a.m + grow] = sum;
        } else if (kMode == 4) {
            if (grow < a.m && rk_cnt) atomicAdd(&a.rk_rank[grow], rk_cnt);      // (integer: the order of the column splits does not matter)
        } else if (kMode == 3) {
            // the four column quarters of a row meet in shared memory
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (cpart > 0) csum3[(cpart - 1) * SM_ROWS + et] = sum;
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (cpart == 0 && grow < a.m) S.out[(int64_t)blockIdx.y * a.m + grow] = (sum + csum3[et]) + (csum3[SM_ROWS + et] + csum3[2 * SM_ROWS + et]);
        } else {
            __syncwarp();
            if (grow < a.m) {
                const int64_t o = ((int64_t)blockIdx.y * a.m + grow) * kListLen;
                for (int e = 0; e < kListLen; ++e) {
                    a.cand_s[o + e] = ls[et * kListLen + e];
                    a.cand_i[o + e] = li[et * kListLen + e];
                }
            }
        }
    }
    if (!ok && lane == 0 && a.error_flag) atomicExch(a.error_flag, 1);
    if (a.sync_ctr && threadIdx.x == 0) {   // epochs this CTA never reaches (shorter column range): count it as arrived
        const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        unsigned int* c = a.sync_ctr + (size_t)(lin / (unsigned)a.sync_group) * a.sync_epochs;
        for (int e = (T + a.sync_every - 1) / a.sync_every; e < a.sync_epochs; ++e) atomicAdd(c + e, 1u);
    }
    fence_before_thread_sync();
    __syncthreads();
    if (kPair) cluster_sync_all();          // no CTA leaves while its peer can still signal its barriers / read its operands
    if (warp == kStages) { if (kPair) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
    if (threadIdx.x == 0) SPROF_END(p_kernel, 7);
}

// Column splits: enough CTAs to fill the SMs when there are few row blocks; with many row blocks the
// smallest split count (<= 4) whose grid fills >= 95 % of its last wave (782 row blocks x 1 split
// would leave the 6th wave 28 % full; x 2 fills 96 %).
int stream_nsplit(int64_t m, int64_t n, int nsides) {
    const int64_t row_ctas = ceil_div(m, SM_ROWS) * nsides;
    const int64_t ntiles = ceil_div(n, SN_COLS);
    const int64_t sms = sm_count();
    if (row_ctas <= 0) return 1;
    if (row_ctas < sms) {
        int64_t want = sms / row_ctas;
        if (want > ntiles) want = ntiles;
        return (int)(want < 1 ? 1 : want);
    }
    int best = 1;
    double best_eff = 0.0;
    for (int ns = 1; ns <= 4 && ns <= ntiles; ++ns) {
        const double ctas = (double)(row_ctas * ns);
        const double eff = ctas / (double)(ceil_div(row_ctas * ns, sms) * sms);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = ns; }
        if (eff >= 0.95) { best = ns; break; }
    }
    return best;
}

int* tc_error_flag();   // loss_tc.cu

// Lockstep is worth it when the Y operand does not fit in L2 and many CTAs stream it (measured: +21 % with a
// single wave of 128 CTAs over a 6 GB gallery, +37 % at cfg5): one group = one wave of SM-count consecutive CTAs,
// one epoch = sync_every() tiles.
static int sync_every() {      // tiles per lockstep epoch (DCLIP_STREAM_SYNC_EVERY overrides; 64 tiles = 25 MB of Y at d = 768)
    static const int v = [] { const char* e = getenv("DCLIP_STREAM_SYNC_EVERY"); const int x = e ? atoi(e) : 0; return x > 0 ? x : 64; }();
    return v;
}
size_t stream_sync_bytes(int64_t m, int64_t n, int64_t d, int nsplit) {
    if (n * d * 2 <= (int64_t)96 * 1024 * 1024) return 0;
    const int64_t ctas = (ceil_div(m, SM_ROWS) + 1) / 2 * 2 * nsplit;
    static const int min_ctas = [] { const char* e = getenv("DCLIP_STREAM_SYNC_MIN_CTAS"); return e ? atoi(e) : 64; }();
    if (ctas < min_ctas) return 0;
    const int64_t groups = ceil_div(ctas, sm_count());
    const int64_t epochs = ceil_div(ceil_div(ceil_div(n, SN_COLS), nsplit), sync_every());
    return (size_t)(groups * epochs) * sizeof(unsigned int);
}

// mode 0: x/y per side are the fp16 operands, a.s[side].out = partial row sums [nsplit][m]
// mode 1: one side; additionally a.col_out = partial column sums [ceil(m / 128)][n]
// mode 2: a.cand_s / a.cand_i = candidate lists [nsplit][m][32]
int launch_stream_tc(int mode, int nsides, const void* const x[2], const void* const y[2], const TcArgs& args_in, int nsplit, cudaStream_t st) {
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "tensor path needs an sm_100 device");
    TcArgs a = args_in;
    // CTA pairs whenever there is more than one 128-row block of X (DCLIP_STREAM_SINGLE=1 forces the single-CTA kernel)
    static const bool forced_single = [] { const char* e = getenv("DCLIP_STREAM_SINGLE"); return e && e[0] == '1'; }();
    const bool pair = !forced_single && a.m > SM_ROWS;
    CUtensorMap tmx[2], tmy[2];
    for (int s = 0; s < 2; ++s) {
        const int src = s < nsides ? s : 0;
        if (!make_tmap_2d_sw128(&tmx[s], x[src], (uint64_t)a.m, (uint64_t)a.d, SM_ROWS, 0) ||
            !make_tmap_2d_sw128(&tmy[s], y[src], (uint64_t)a.n, (uint64_t)a.d, pair ? SN_COLS / 2 : SN_COLS, 0))
            return fail(DCLIP_ERR_CUDA, "cuTensorMapEncodeTiled failed (m=%lld n=%lld d=%lld)", (long long)a.m, (long long)a.n, (long long)a.d);
    }
    const int64_t ntiles = ceil_div(a.n, SN_COLS);
    a.tiles_per_split = (int)ceil_div(ntiles, nsplit);
    a.error_flag = tc_error_flag();
    if (a.sync_ctr) {                       // caller provided stream_sync_bytes() of counters: zero them, fill in the geometry
        const size_t sb = stream_sync_bytes(a.m, a.n, a.d, nsplit);
        static const bool no_sync = [] { const char* e = getenv("DCLIP_STREAM_NOSYNC"); return e && e[0] == '1'; }();
        if (sb == 0 || !pair || nsides != 1 || no_sync) a.sync_ctr = nullptr;
        else {
            DCLIP_CUDA_OK(cudaMemsetAsync(a.sync_ctr, 0, sb, st));
            a.sync_every = sync_every(); a.sync_group = sm_count();
            a.sync_epochs = (int)ceil_div(a.tiles_per_split, sync_every());
        }
    }
    if (mode == 4 && (!a.rk_thr || !a.rk_tgt || !a.rk_rank || !a.rk_band || !a.rk_ctr || nsides != 1))
        return fail(DCLIP_ERR_INVALID, "stream_tc: mode 4 is single-sided and needs thresholds, targets, the rank output and a band list");
    if (mode == 1 && (!a.col_out || nsides != 1)) return fail(DCLIP_ERR_INVALID, "stream_tc: mode 1 is single-sided and needs col_out");
    if (mode == 3 && (!a.g_out || !a.col_out || !a.s[0].out || nsides != 1 || !pair || a.ldg % 8 || a.ldg < ceil_div(a.n, SN_COLS) * SN_COLS))
        return fail(DCLIP_ERR_INVALID, "stream_tc: mode 3 is single-sided, needs more than 128 rows and a tile-padded coefficient matrix");
    const size_t smem = 1024 + (size_t)stream_stages(mode, pair) * stream_stage_bytes(pair) +
                        (mode == 3 ? stream_extra_bytes(3) : mode == 2 ? stream_extra_bytes(2) : mode == 1 ? stream_extra_bytes(1) : 0) + 256;
    const int64_t row_blocks = ceil_div(a.m, SM_ROWS);
    dim3 grid((unsigned)(pair ? (row_blocks + 1) / 2 * 2 : row_blocks), (unsigned)nsplit, (unsigned)nsides);
#define DCLIP_STREAM_LAUNCH(MODE, PAIR)                                                                                   \
    do {                                                                                                                  \
        DCLIP_CUDA_OK(cudaFuncSetAttribute((k_stream_tc<MODE, PAIR>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        if (PAIR) {                                                                                                       \
            cudaLaunchConfig_t cfg = {};                                                                                  \
            cfg.gridDim = grid; cfg.blockDim = dim3(stream_threads(MODE, PAIR)); cfg.dynamicSmemBytes = smem; cfg.stream = st; \
            cudaLaunchAttribute at[1];                                                                                    \
            at[0].id = cudaLaunchAttributeClusterDimension;                                                               \
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                          \
            cfg.attrs = at; cfg.numAttrs = 1;                                                                             \
            DCLIP_CUDA_OK(cudaLaunchKernelEx(&cfg, k_stream_tc<MODE, PAIR>, tmx[0], tmy[0], tmx[1], tmy[1], a));          \
        } else {                                                                                                          \
            k_stream_tc<MODE, PAIR><<<grid, stream_threads(MODE, PAIR), smem, st>>>(tmx[0], tmy[0], tmx[1], tmy[1], a);   \
        }                                                                                                                 \
    } while (0)
    if (mode == 0) { if (pair) DCLIP_STREAM_LAUNCH(0, true); else DCLIP_STREAM_LAUNCH(0, false); }
    else if (mode == 1) { if (pair) DCLIP_STREAM_LAUNCH(1, true); else DCLIP_STREAM_LAUNCH(1, false); }
    else if (mode == 3) DCLIP_STREAM_LAUNCH(3, true);
    else if (mode == 4) { if (pair) DCLIP_STREAM_LAUNCH(4, true); else DCLIP_STREAM_LAUNCH(4, false); }
    else { if (pair) DCLIP_STREAM_LAUNCH(2, true); else DCLIP_STREAM_LAUNCH(2, false); }
#undef DCLIP_STREAM_LAUNCH
    DCLIP_LAUNCH_OK("k_stream_tc");
    return DCLIP_OK;
}

}  // namespace dclip

#ifdef DCLIP_PAIR_PROF
extern "C" int dclip_debug_stream_prof(unsigned long long* out16, int reset) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out16, dclip::g_stream_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(dclip::g_stream_prof, z, sizeof(z));
    }
    return 0;
}
#endif
