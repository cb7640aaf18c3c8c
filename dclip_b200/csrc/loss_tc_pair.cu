// Distillation objective — InfoNCE BACKWARD on CTA PAIRS (tcgen05 cta_group::2, sm_100a):
//   dX_hat = coef/tau * G . Y_hat,  G = softmax_row + softmax_col - 2 I  recomputed tile by tile
// (autograd backward of training/CLIP_image_distillation.py:545-560; SURVEY.md 8(a) "exact backward math").
//
// Why pairs.  The gradient accumulator dX^T[d x i] has to stay in TMEM for the whole walk over Y; with one
// CTA (loss_tc.cu) its 512 columns hold d/128 blocks of only 64 i-columns, so every MMA is 128 x 64 x 16 and
// re-reads 6 KB of shared-memory operands for 32 cycles of math (shared-memory bound, tensor pipe 38 %).
// A pair has two TMEMs: each CTA keeps HALF of d (d/256 blocks of 128 lanes) for 128 i-columns, and every
// MMA is a 256 x 128 x 16 cta_group::2 instruction: the same 6 KB of operand reads per CTA now feed 64
// cycles of math, and the 128 X rows are split between the two shared memories (64 rows = 96 KB each).
//
// Per pair: 128 rows of X_hat (I block), walking 256-row tiles J of Y_hat:
//   phase 1   S^T[256 j x 128 i] = Y_J . X_I^T     M = 256 (CTA c: its 128 j rows), N = 128 (CTA c: 64 i rows
//             of X, resident), K = d.  CTA c receives S^T rows [128c, 128c + 128) in TMEM columns 384-511.
//   epilogue  thread = one j (TMEM lane); g = p (a_i + b_j) - 2 [i == j], scaled by 2^13, fp16 (the diagonal term is
//             subtracted in fp32 before the rounding, in the one tile of the walk that faces the pair's rows).  Columns
//             i < 64 go to the LEADER's G tile, i >= 64 to the PEER's (st.shared::cluster): each CTA ends up
//             with G^T[256 j x 64 i] = its half of the phase-2 B operand.
//   phase 2   dX^T[256 d x 128 i] += Y_J^T[256 d x 256 j] . G^T   per 256-wide d block: M = 256 (CTA c: d
//             columns [256 blk + 128c, +128) read MN-major from the TMA boxes), N = 128, K = 256 j.
// TMEM per CTA: d/256 accumulators of 128 columns (<= 384) + one 128-column S buffer.
// Warps: 0-2 TMA producers (one 32 KB ring slot each, both CTAs), 3 MMA issuer (leader only; both allocate
// TMEM), 4-11 epilogue (lane quadrant = warp & 3, column half = (warp - (kSlots + 1)) >> 2).
// Barriers at identical offsets in both CTAs; `full`, `x_full`, `s_empty`, `g_full` are used in the leader's
// copy only (TMA bytes of both CTAs and remote arrivals are credited there), `empty`, `s_full`, `g_empty`,
// `acc_full` are signalled in both CTAs by multicast tcgen05.commit.
// Publishing G without GPU-scope membars.  Cluster-scope release / acquire (and fence.proxy.async on
// shared::cluster) compile to MEMBAR.ALL.GPU / CCTL.IVALL, ~3000 cycles per tile on the critical path next to
// the operand traffic.  Instead every G half has its completion barrier in the CTA that OWNS that half
// (`g_full` in the leader, `g_part` in the peer), and a warp arrives on it with a plain (CTA-scope release)
// arrive right after its stores: local stores + local arrive, or remote stores + remote arrive to the same
// destination CTA, which the SM-to-SM network delivers in order.  Each writer runs the cheap CTA-local
// generic->async proxy fence; the peer's otherwise idle MMA warp waits for `g_part`, fences, and forwards one
// arrival to the leader's `g_full` (4 local + 4 remote + 1 relay arrivals per tile).
#include <cstdlib>

#include "common.cuh"
#include "loss_internal.h"
#include "tcgen05.cuh"

namespace dclip {
using namespace tc;

// Optional stall accounting (build with -DDCLIP_PAIR_PROF): cycles the roles of cluster 0 spend in each wait.
#ifdef DCLIP_PAIR_PROF
#define DCLIP_DBG_SKIP(bit) (a.debug_skip & (bit))
__device__ unsigned long long g_pair_prof[32];
__device__ unsigned int g_pair_trace[8][2048];   // leader CTA of cluster 0: per-role clock stamps (low 32 bits)
#define PROF_ON (blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0)
#define TRACE(role, idx) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 2048) g_pair_trace[role][idx] = (unsigned int)clock64(); } while (0)
#define PROF_BEGIN(v) const long long v = clock64()
#define PROF_END(v, slot) do { if (PROF_ON) atomicAdd(&g_pair_prof[slot], (unsigned long long)(clock64() - v)); } while (0)
#else
#define DCLIP_DBG_SKIP(bit) false
#define PROF_BEGIN(v) do { } while (0)
#define PROF_END(v, slot) do { } while (0)
#define TRACE(role, idx) do { } while (0)
#endif

namespace pk {
constexpr int NI = 128;                // X rows per pair (MMA N)
constexpr int NH = 64;                 // X rows resident per CTA
constexpr int TJ = 256;                // Y rows per tile
constexpr int kBoxBytes = 128 * 128;   // [128 rows x 64 fp16] = 16 KB
constexpr int kXBoxBytes = NH * 128;   // [64 rows x 64 fp16] = 8 KB
constexpr int kGBytes = TJ * 128;      // G^T half: [256 j x 64 i] fp16 = 32 KB
constexpr int kEpiWarps = 8;
constexpr int kSCol = 384;
constexpr float kGScaleLog2 = 13.0f;   // G stored as fp16 * 2^13 (softmax sums <= 2 -> <= 16384)
constexpr uint32_t kSpin = 1u << 24;
}  // namespace pk

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// exponent arguments stay inside [-2 log2e / tau, 0] (tau >= 0.025): no denormal handling needed
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int kSlots>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kSlots + 9) * 32, 1)
k_infonce_bwd_pair(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmy0,
                   const __grid_constant__ CUtensorMap tmx1, const __grid_constant__ CUtensorMap tmy1, const TcArgs a) {
    using namespace pk;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int side = blockIdx.z;
    const CUtensorMap* tmx = side ? &tmx1 : &tmx0;
    const CUtensorMap* tmy = side ? &tmy1 : &tmy0;
    const TcSide S = a.s[side];
    const int d = (int)a.d;
    const int nkb = d / 64;            // 64-column boxes along d
    const int NS = d / 128;            // ring-slot jobs per phase (phase 1: 2 k-boxes each; phase 2: (d block, j half))

    uint8_t* xs = smem;                                   // resident X half: nkb boxes of 8 KB
    uint8_t* ring = xs + (size_t)nkb * kXBoxBytes;        // kSlots x 16 KB boxes
    uint8_t* gs = ring + kSlots * kBoxBytes;              // 32 KB
    uint8_t* tail = gs + kGBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(tail);   // [kSlots / 2]  (leader) one per PAIR of ring boxes
    uint64_t* empty = full + kSlots / 2;                  // [kSlots] one per box
    uint64_t* x_full = empty + kSlots;                    // (leader)
    uint64_t* s_full = x_full + 1;
    uint64_t* s_empty = s_full + 1;                       // (leader) 16 warp arrivals
    uint64_t* g_full = s_empty + 1;                       // (leader) my G half complete + relay: 9 arrivals
    uint64_t* g_part = g_full + 1;                        // (peer) the peer's G half complete: 8 warp arrivals
    uint64_t* g_empty = g_part + 1;
    uint64_t* acc_full = g_empty + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* a_s = reinterpret_cast<float*>(tail + 256);    // [NI] column factors 2^13 / rowsum_i (16-byte aligned: read as float4)

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t cr = cluster_ctarank();
    const int64_t row0 = (int64_t)(blockIdx.x >> 1) * NI;           // first X row of the pair
    const int64_t ntiles = (a.n + TJ - 1) / TJ;
    const int64_t t_begin = (int64_t)blockIdx.y * a.tiles_per_split;
    const int64_t t_end = min(ntiles, t_begin + a.tiles_per_split);
    const int T = (int)max((int64_t)0, t_end - t_begin);
    constexpr bool diag = false;       // the -2 I term of G is folded into the G tiles (epilogue, before the fp16 rounding): no extra MMA step

    PROF_BEGIN(p_kernel);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i >> 1], 1); mbar_init(&empty[i], 1); }
        mbar_init(x_full, 1);
        mbar_init(s_full, 1); mbar_init(s_empty, 2 * kEpiWarps);
        mbar_init(g_full, kEpiWarps + 1); mbar_init(g_part, kEpiWarps); mbar_init(g_empty, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == kSlots) tmem_alloc_pair(tmem_ptr, 512);
    fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();                // peer barriers initialised, both TMEMs allocated
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;
    if (threadIdx.x == 0) PROF_END(p_kernel, 28);

    if (T > 0) {
        if (warp < kSlots) {
            // ===================== TMA producers (both CTAs): warp w owns ring slot w =====================
            if (elect_one()) {
                tma_prefetch_desc(tmy);
                if (warp == 0) {
                    tma_prefetch_desc(tmx);
                    const uint32_t xbar = mapa_u32(smem_u32(x_full), 0);
                    if (cr == 0) mbar_arrive_expect_tx(x_full, (uint32_t)(2 * nkb * kXBoxBytes));
                    for (int kb = 0; kb < nkb; ++kb)
                        tma_load_2d_pair(xs + (size_t)kb * kXBoxBytes, tmx, xbar, kb * 64, (int32_t)(row0 + cr * NH));
                }
                // Box job q (MMA consumption order) uses slot q % kSlots.  Order: P1(0), then per tile t: P1(t+1) [if any],
                // P2(t); every phase consumes nkb boxes (phase 2 / diagonal step: in aligned pairs).
                const int NB = nkb;
                const int njobs = 2 * NB * T;
                const int ndiag = diag ? NB / 2 : 0;
                uint8_t* dst = ring + (size_t)warp * kBoxBytes;
                const uint32_t fbar = mapa_u32(smem_u32(&full[warp >> 1]), 0);
                uint32_t phase = 0;
                for (int q = warp; q < njobs + ndiag && ok; q += kSlots, phase ^= 1) {
                    int tile = 0, idx = 0, ph;
                    if (q >= njobs) { idx = q - njobs; ph = 3; }
                    else if (q < NB) { tile = 0; idx = q; ph = 1; }
                    else {
                        const int qq = q - NB, blk = min(qq / (2 * NB), T - 1), r = qq - blk * 2 * NB;
                        if (blk < T - 1 && r < NB) { tile = blk + 1; idx = r; ph = 1; }
                        else { tile = blk; idx = blk < T - 1 ? r - NB : r; ph = 2; }
                    }
                    const int32_t j0 = (int32_t)((t_begin + tile) * TJ);
                    int32_t c0, r0;
                    if (ph == 1) {              // K-major A: my 128 j rows, k box idx
                        c0 = idx * 64; r0 = j0 + (int32_t)cr * 128;
                    } else if (ph == 2) {       // MN-major A: box idx & 1 of my 128 columns of d block idx >> 2, j half (idx >> 1) & 1
                        c0 = (idx >> 2) * 256 + (int32_t)cr * 128 + (idx & 1) * 64; r0 = j0 + ((idx >> 1) & 1) * 128;
                    } else {                    // diagonal step: d block idx >> 1, the 128 Y rows facing the pair's X rows
                        c0 = (idx >> 1) * 256 + (int32_t)cr * 128 + (idx & 1) * 64; r0 = (int32_t)(a.row_offset + row0);
                    }
                    PROF_BEGIN(p_em);
                    ok = mbar_wait(&empty[warp], phase ^ 1, kSpin);
                    PROF_END(p_em, 6 + (int)cr);
                    TRACE(warp, 2 * (q / kSlots));
                    if (cr == 0 && (warp & 1) == 0) mbar_arrive_expect_tx(&full[warp >> 1], 4 * kBoxBytes);   // both boxes of the pair, both CTAs
                    tma_load_2d_pair(dst, tmy, fbar, c0, r0);
                    TRACE(warp, 2 * (q / kSlots) + 1);
                }
            }
        } else if (warp == kSlots) {
            // ===================== MMA issuer (leader only) =====================
            if (cr == 0 && elect_one()) {
                const uint32_t idesc1 = make_idesc_f16(256, NI, kFmtF16, 0, 0);
                const uint32_t idesc2 = make_idesc_f16(256, NI, kFmtF16, 1, 1);
                const uint64_t hiK = make_smem_desc_hi(0, 1024, kSwizzle128B);
                const uint64_t hiA2 = make_smem_desc_hi(kBoxBytes, 1024, kSwizzle128B);   // MN-major A: LBO = next 64-d group
                const uint64_t hiB2 = make_smem_desc_hi(0, 1024, kSwizzle128B);           // MN-major B (G half), 64 i = one group
                const uint32_t xs_addr = smem_u32(xs), ring_addr = smem_u32(ring), gs_addr = smem_u32(gs);
                uint32_t stage = 0, sphase = 0;
                PROF_BEGIN(p_all);
                ok = mbar_wait(x_full, 0, kSpin);
                PROF_END(p_all, 1);
                auto phase1 = [&](int t) {
                    PROF_BEGIN(p_se);
                    TRACE(7, 64 + 2 * t);
                    ok = ok && mbar_wait(s_empty, (t & 1) ^ 1, kSpin);
                    TRACE(7, 64 + 2 * t + 1);
                    PROF_END(p_se, 2);
                    fence_after_thread_sync();
                    for (int kb = 0; kb < nkb && ok; kb += 2) {      // two ring boxes per wait: the issuer's per-wait cost
                        PROF_BEGIN(p_f1);                              // (~300 cycles) must stay below the MMA time it feeds
                        ok = mbar_wait(&full[stage >> 1], sphase, kSpin);
                        PROF_END(p_f1, 3);
                        fence_after_thread_sync();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint64_t da0 = smem_desc(hiK, ring_addr + (stage + h) * kBoxBytes);
                            const uint64_t db0 = smem_desc(hiK, xs_addr + (kb + h) * kXBoxBytes);
                            if (!DCLIP_DBG_SKIP(1))
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_f16_pair(tmem + kSCol, da0 + (uint64_t)(ks * 2), db0 + (uint64_t)(ks * 2), idesc1, (kb | h | ks) != 0);
                            umma_commit_pair(&empty[stage + h]);          // box released as soon as ITS four MMAs are done
                        }
                        stage += 2;
                        if (stage == kSlots) { stage = 0; sphase ^= 1; }
                    }
                    umma_commit_pair(s_full);
                };
                auto phase2 = [&](int t) {
                    PROF_BEGIN(p_gf);
                    TRACE(7, 2 * t);
                    ok = ok && mbar_wait(g_full, t & 1, kSpin);
                    TRACE(7, 2 * t + 1);
                    PROF_END(p_gf, 4);
                    fence_proxy_async_smem();
                    fence_after_thread_sync();
                    for (int s = 0; s < NS && ok; ++s) {         // (d block s >> 1, j half s & 1): two adjacent ring boxes
                        PROF_BEGIN(p_f2);
                        ok = mbar_wait(&full[stage >> 1], sphase, kSpin);
                        PROF_END(p_f2, 5);
                        fence_after_thread_sync();
                        const int jh = s & 1;
                        const uint64_t da0 = smem_desc(hiA2, ring_addr + stage * kBoxBytes);
                        const uint64_t db0 = smem_desc(hiB2, gs_addr + jh * (kGBytes / 2));
                        if (!DCLIP_DBG_SKIP(2))
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)          // 16 j per step = two 8-row groups = +2048 B
                            umma_f16_pair(tmem + (s >> 1) * NI, da0 + (uint64_t)(ks * 128), db0 + (uint64_t)(ks * 128), idesc2, (t | jh | ks) != 0);
                        umma_commit_pair(&empty[stage]);
                        umma_commit_pair(&empty[stage + 1]);
                        stage += 2;
                        if (stage == kSlots) { stage = 0; sphase ^= 1; }
                    }
                    umma_commit_pair(g_empty);
                };
                if (ok) phase1(0);
                for (int t = 0; t < T && ok; ++t) {
                    if (t + 1 < T) phase1(t + 1);
                    if (ok) phase2(t);
                }
                if (diag && ok) {
                    // diagonal step: dX^T[256 d x 128 i] += Y_D^T[256 d x 128 j'] . (-2 * 2^13 I), Y_D = the Y rows facing X_I
                    ok = mbar_wait(g_full, T & 1, kSpin);
                    fence_proxy_async_smem();
                    fence_after_thread_sync();
                    for (int s = 0; s < NS / 2 && ok; ++s) {
                        ok = mbar_wait(&full[stage >> 1], sphase, kSpin);
                        fence_after_thread_sync();
                        const uint64_t da0 = smem_desc(hiA2, ring_addr + stage * kBoxBytes);
                        const uint64_t db0 = smem_desc(hiB2, gs_addr);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_f16_pair(tmem + s * NI, da0 + (uint64_t)(ks * 128), db0 + (uint64_t)(ks * 128), idesc2, 1);
                        umma_commit_pair(&empty[stage]);
                        umma_commit_pair(&empty[stage + 1]);
                        stage += 2;
                        if (stage == kSlots) { stage = 0; sphase ^= 1; }
                    }
                }
                umma_commit_pair(acc_full);
                PROF_END(p_all, 0);
            } else if (cr == 1 && elect_one()) {
                // relay: my G half is complete (8 warp arrivals, local and remote) -> one arrival on the leader's g_full
                const uint32_t g_full_l = mapa_u32(smem_u32(g_full), 0);
                for (int t = 0; t < T + (diag ? 1 : 0) && ok; ++t) {
                    ok = mbar_wait(g_part, t & 1, kSpin);
                    fence_proxy_async_smem();
                    mbar_arrive_cluster_relaxed(g_full_l);
                }
            }
        } else {
            // ===================== epilogue warps (256 threads per CTA) =====================
            const int q = warp & 3;                       // TMEM lane quadrant this warp may access
            const int hh = (warp - (kSlots + 1)) >> 2;               // column half: i in [64 hh, 64 hh + 64) -> G tile of CTA hh
            const int et = q * 32 + lane;                 // j row inside my 128-row half == TMEM lane
            const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
            const float c = a.inv_tau * 1.4426950408889634f;    // log2(e) / tau
            {
                const int idx = (warp - (kSlots + 1)) * 32 + lane;
                if (idx < NI) {
                    const int64_t gi = row0 + idx;
                    a_s[idx] = gi < a.m ? exp2f((a.inv_tau - S.lse_x[gi]) * 1.4426950408889634f + kGScaleLog2) : 0.f;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t s_empty_l = mapa_u32(smem_u32(s_empty), 0);
            // completion barrier of the G half I write: in the CTA that owns it (hh = 0: leader's g_full, hh = 1: peer's g_part)
            const uint32_t g_done = mapa_u32(smem_u32(hh == 0 ? g_full : g_part), (uint32_t)hh);
            const uint32_t grow_local = smem_u32(gs) + (uint32_t)(cr * 128 + et) * 128;
            const uint32_t grow_dst = mapa_u32(grow_local, (uint32_t)hh);
            const bool local_dst = (uint32_t)hh == cr;
            const uint32_t a_addr = smem_u32(a_s) + (uint32_t)hh * 256;
            for (int t = 0; t < T; ++t) {
                const int64_t j0 = (t_begin + t) * TJ;
                const int64_t gj = j0 + cr * 128 + et;
                PROF_BEGIN(p_sf);
                ok = mbar_wait(s_full, t & 1, kSpin);
                if (q == 0 && lane == 0) PROF_END(p_sf, 8 + 5 * ((int)cr * 2 + hh));
                ok = __all_sync(0xffffffffu, ok);         // keep the .aligned TMEM loads warp-uniform
                if (!ok) break;
                fence_after_thread_sync();
                PROF_BEGIN(p_ld);
                uint32_t r[64];
                tmem_ld_32x32b_x32(tmem + lane_addr + kSCol + hh * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                tmem_ld_32x32b_x32(tmem + lane_addr + kSCol + hh * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                tmem_ld_wait();
                fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(s_empty_l);      // S may be overwritten by phase 1 of tile t+1 (no data published)
                if (q == 0 && lane == 0) PROF_END(p_ld, 9 + 5 * ((int)cr * 2 + hh));
                PROF_BEGIN(p_cp);
                const bool jvalid = gj < a.n;
                const int64_t ly = S.lse_y_block > 0 ? (gj / S.lse_y_block) * S.lse_y_stride + gj % S.lse_y_block : gj;
                const float bj = jvalid ? exp2f((a.inv_tau - S.lse_y[ly]) * 1.4426950408889634f + kGScaleLog2) : 0.f;
                const float vm = jvalid ? 1.f : 0.f;      // rows past n: TMA zero-filled S, coefficient forced to 0
                // g = p (a_i + b_j) - 2 [i == j], all scaled by 2^13.  The diagonal term is subtracted in fp32 BEFORE the
                // fp16 rounding: p_ii (a_i + b_i) approaches 2 as training converges and G_ii is the small difference -
                // rounding the large value first (and subtracting an exact -2 afterwards, as a separate MMA step once
                // did) leaves an absolute error of 2^-11 * 2 on a quantity that can be 100x smaller.  Only the one tile
                // of the walk that faces the pair's own rows has a diagonal: a CTA-uniform branch keeps the common loop
                // free of the compare.
                const int64_t di = gj - (a.row_offset + row0 + hh * 64);      // my diagonal column inside this warp's 64, if in range
                const bool tile_has_diag = j0 + (int64_t)cr * 128 < a.row_offset + row0 + hh * 64 + 64 &&
                                           j0 + (int64_t)cr * 128 + 128 > a.row_offset + row0 + hh * 64;
                uint32_t packed[32];
                if (!tile_has_diag) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 a0 = lds_f4(a_addr + ch * 32), a1 = lds_f4(a_addr + ch * 32 + 16);
                        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float p0 = ex2_ftz(fmaf(__uint_as_float(r[ch * 8 + 2 * h]), c, -c));
                            const float p1 = ex2_ftz(fmaf(__uint_as_float(r[ch * 8 + 2 * h + 1]), c, -c));
                            const __half2 hv = __floats2half2_rn(p0 * fmaf(av[2 * h], vm, bj), p1 * fmaf(av[2 * h + 1], vm, bj));
                            packed[ch * 4 + h] = *reinterpret_cast<const uint32_t*>(&hv);
                        }
                    }
                } else {
                    const int dcol = (di >= 0 && di < 64 && jvalid) ? (int)di : -1;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 a0 = lds_f4(a_addr + ch * 32), a1 = lds_f4(a_addr + ch * 32 + 16);
                        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float p0 = ex2_ftz(fmaf(__uint_as_float(r[ch * 8 + 2 * h]), c, -c));
                            const float p1 = ex2_ftz(fmaf(__uint_as_float(r[ch * 8 + 2 * h + 1]), c, -c));
                            const float g0 = fmaf(p0, fmaf(av[2 * h], vm, bj), dcol == ch * 8 + 2 * h ? -16384.0f : 0.f);
                            const float g1 = fmaf(p1, fmaf(av[2 * h + 1], vm, bj), dcol == ch * 8 + 2 * h + 1 ? -16384.0f : 0.f);
                            const __half2 hv = __floats2half2_rn(g0, g1);
                            packed[ch * 4 + h] = *reinterpret_cast<const uint32_t*>(&hv);
                        }
                    }
                }
                if (q == 0 && lane == 0) PROF_END(p_cp, 10 + 5 * ((int)cr * 2 + hh));
                PROF_BEGIN(p_ge);
                ok = mbar_wait(g_empty, (t & 1) ^ 1, kSpin);      // phase 2 of tile t-1 has read G
                if (q == 0 && lane == 0) PROF_END(p_ge, 11 + 5 * ((int)cr * 2 + hh));
                PROF_BEGIN(p_st);
                // 128B swizzle: 16-byte chunk ch of row r lands at chunk (ch ^ (r & 7)); (128 cr + et) & 7 == et & 7
                if (!DCLIP_DBG_SKIP(4))
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint4 v = make_uint4(packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
                    const uint32_t off = (uint32_t)((ch ^ (et & 7)) << 4);
                    if (local_dst) *reinterpret_cast<uint4*>(gs + (size_t)(cr * 128 + et) * 128 + off) = v;
                    else st_shared_cluster_v4(grow_dst + off, v);
                }
                if (q == 0 && lane == 0) PROF_END(p_st, 12 + 5 * ((int)cr * 2 + hh));
                PROF_BEGIN(p_fe);
                fence_proxy_async_smem();                 // my generic-proxy stores -> async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(g_done);   // same destination CTA as the stores: delivered in order
                if (q == 0 && lane == 0 && cr == 0) PROF_END(p_fe, 30 + hh);
            }
            if (diag && __all_sync(0xffffffffu, ok)) {
                // my own B half of the diagonal step: G^T[128 j' x 64 i] = -2 * 2^13 [j' == 64 cr + i]  (fp16 0xF400 = -16384)
                ok = mbar_wait(g_empty, (T & 1) ^ 1, kSpin);
                const int e = (warp - (kSlots + 1)) * 32 + lane;      // 0..255: row e >> 1, chunks 4 (e & 1) .. +3
                const int jr = e >> 1, il = jr - (int)cr * NH;          // the one non-zero column of row jr, if in [0, 64)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ch = (e & 1) * 4 + k, pos = il - ch * 8;   // element position inside the 8-wide chunk
                    uint32_t w[4] = {0u, 0u, 0u, 0u};
                    if (pos >= 0 && pos < 8) w[pos >> 1] = (pos & 1) ? 0xF4000000u : 0x0000F400u;
                    *reinterpret_cast<uint4*>(gs + (size_t)jr * 128 + ((ch ^ (jr & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(cr == 0 ? g_full : g_part);
            }
            if (__all_sync(0xffffffffu, ok) && (ok = __all_sync(0xffffffffu, mbar_wait(acc_full, 0, kSpin)))) {
                fence_after_thread_sync();
                PROF_BEGIN(p_dr);
                float* outp = S.out + (int64_t)blockIdx.y * a.m * a.d;
                for (int s = 0; s < NS / 2; ++s) {
                    const int dcol = s * 256 + (int)cr * 128 + et;    // TMEM lane = d index inside my half of the block
                    float* op = outp + (row0 + hh * 64) * a.d + dcol;
                    const int rows_left = (int)min((int64_t)64, a.m - row0 - hh * 64);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t rr[32];
                        tmem_ld_32x32b_x32(tmem + lane_addr + s * NI + hh * 64 + h * 32, rr);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (h * 32 + i < rows_left) op[(size_t)(h * 32 + i) * d] = __uint_as_float(rr[i]) * a.out_scale;
                    }
                }
                if (q == 0 && lane == 0 && hh == 0) PROF_END(p_dr, 29);
            }
        }
    } else if (warp >= kSlots + 1) {
        // empty column range (only possible with an over-split grid): contribute zeros
        const int et = (warp - (kSlots + 1)) * 32 + lane;
        float* outp = S.out + (int64_t)blockIdx.y * a.m * a.d;
        for (int i = 0; i < NH; ++i) {
            const int64_t gi = row0 + cr * NH + i;
            if (gi < a.m)
                for (int c2 = et; c2 < d; c2 += 256) outp[gi * a.d + c2] = 0.f;
        }
    }
    if (!ok && lane == 0 && a.error_flag) atomicExch(a.error_flag, 1);
    fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();                // no CTA leaves while its peer can still touch its shared memory / TMEM
    if (threadIdx.x == 0 && cr == 1) PROF_END(p_kernel, 27);
    if (warp == kSlots) tmem_dealloc_pair(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_pair_shape_ok(int64_t d) { return d % 256 == 0 && d >= 256 && d <= 768; }

int tc_pair_nsplit(int64_t m, int64_t n, int nsides) {
    const int64_t pairs = ceil_div(m, pk::NI) * nsides;
    const int64_t ntiles = ceil_div(n, pk::TJ);
    int64_t want = (sm_count() / 2) / (pairs > 0 ? pairs : 1);
    if (want < 1) want = 1;
    if (want > ntiles) want = ntiles;
    return (int)want;
}

int* tc_error_flag();   // loss_tc.cu

int launch_infonce_bwd_pair(int nsides, const void* const x[2], const void* const y[2], const TcArgs& args_in, int nsplit, cudaStream_t st) {
    using namespace pk;
    TcArgs a = args_in;
    CUtensorMap tmx[2], tmy[2];
    for (int s = 0; s < 2; ++s) {
        const int src = s < nsides ? s : 0;
        if (!make_tmap_2d_sw128(&tmx[s], x[src], (uint64_t)a.m, (uint64_t)a.d, NH, 0) ||
            !make_tmap_2d_sw128(&tmy[s], y[src], (uint64_t)a.n, (uint64_t)a.d, 128, 0))
            return fail(DCLIP_ERR_CUDA, "cuTensorMapEncodeTiled failed (m=%lld n=%lld d=%lld)", (long long)a.m, (long long)a.n, (long long)a.d);
    }
    const int64_t ntiles = ceil_div(a.n, TJ);
    a.tiles_per_split = (int)ceil_div(ntiles, nsplit);
    a.error_flag = tc_error_flag();
#ifdef DCLIP_PAIR_PROF   // timing ablation (results are garbage): only in the stall-accounting build, never in the release library
    { static const int dbg = [] { const char* e = getenv("DCLIP_PAIR_DEBUG_SKIP"); return e ? atoi(e) : 0; }(); a.debug_skip = dbg; }
#else
    a.debug_skip = 0;
#endif
    // ring: 16 KB boxes, one producer warp each (finer slots keep more of the ring in flight: a slot is refilled
    // ~900-1000 cycles after its release, tools/tc_bench tmamp); 6 boxes fit next to X (96 KB) and G at d = 768
    static const int forced_slots = [] { const char* e = getenv("DCLIP_PAIR_SLOTS"); return e ? atoi(e) : 0; }();
    const size_t fixed = 1024 + (size_t)(a.d / 64) * kXBoxBytes + kGBytes + 1024;
    int slots = (int)((232448 - fixed) / kBoxBytes) & ~1;
    if (slots > 6) slots = 6;      // 8 boxes = 17 warps: the register cap (96) spills the epilogue, measured slower at d = 512
    if (forced_slots >= 2 && forced_slots <= slots && forced_slots % 2 == 0) slots = forced_slots;
    const size_t smem = fixed + (size_t)slots * kBoxBytes;
    dim3 grid((unsigned)(2 * ceil_div(a.m, NI)), (unsigned)nsplit, (unsigned)nsides);
    auto kern = slots == 6 ? k_infonce_bwd_pair<6> : slots == 4 ? k_infonce_bwd_pair<4> : k_infonce_bwd_pair<2>;
    DCLIP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, (slots + 9) * 32, smem, st>>>(tmx[0], tmy[0], tmx[1], tmy[1], a);
    DCLIP_LAUNCH_OK("k_infonce_bwd_pair");
    return DCLIP_OK;
}

}  // namespace dclip

#ifdef DCLIP_PAIR_PROF
extern "C" int dclip_debug_pair_trace(unsigned int* out, int n) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(out, dclip::g_pair_trace, (size_t)n * sizeof(unsigned int)) == cudaSuccess ? 0 : -1;
}
extern "C" int dclip_debug_pair_prof(unsigned long long* out32, int reset) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out32, dclip::g_pair_prof, 32 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[32] = {0};
        cudaMemcpyToSymbol(dclip::g_pair_prof, z, sizeof(z));
    }
    return 0;
}
#endif
