// Grouped tcgen05 GEMM for the cross-attention projections and their gradients (16-bit operands,
// fp32 accumulation in TMEM, fp32 output):
//     C[M, N] = A'(M, K) . B'(K, N) (+ bias[n]) (+ residual[m, n])        for up to 8 problems per launch
// replacing the cuBLAS calls under nn.MultiheadAttention's in-/out-projections
// (training/patch_text_aggregation.py:33,42) and the autograd GEMMs of their backward.
//
// Operand storage (row-major matrices, loaded by TMA with SWIZZLE_128B):
//   A K-major  : A stored [M, K]  (activations X, dY)          box [128 m  x 64 k]
//   A MN-major : A stored [K, M]  (dY^T for weight gradients)  2 boxes [64 k x 64 m], LBO = 8 KB
//   B K-major  : B stored [N, K]  (nn.Linear weights)          box [256 n  x 64 k]
//   B MN-major : B stored [K, N]  (X for dW, W for dX)         4 boxes [64 k x 64 n], LBO = 8 KB
// (the MN-major canonical layout is the same bytes TMA writes for a [k rows x 64 columns] box; both
// were validated against a CPU GEMM with tools/tc_probe).  Tile 128 x 256, K in 64-wide stages, 4-stage
// ring, 4 producer warps, 1 MMA-issuer warp, 8 epilogue warps (thread = output row; two warps per TMEM lane
// quadrant split the columns; bias / residual / conversion in registers, then a swizzled 4 KB staging buffer per warp so
// that global stores are 128 contiguous bytes per row - the epilogue of a K = 768 tile must not cost more than its
// 12-stage main loop).  Persistent CTAs.
// bias / residual: fp32 or the operands' 16-bit type; outputs: fp32 C and / or 16-bit C16.
#include "common.cuh"
#include "gemm_tc.h"
#include "tcgen05.cuh"

namespace dclip {
using namespace tc;

constexpr int GM = 128, GN = 256, GKS = 64;
constexpr int kGStages = 4;
constexpr int kGAStage = GM * 128;          // 16 KB
constexpr int kGBStage = GN * 128;          // 32 KB
constexpr int kGStageBytes = kGAStage + kGBStage;
constexpr int kGEpiWarps = 8;           // two per TMEM lane quadrant: each takes half of the tile's 256 columns
constexpr int kGThreads = (kGStages + 1 + kGEpiWarps) * 32;
constexpr uint32_t kSpinG = 1u << 24;

struct TcGemmDev {
    float* C; const void* bias; const void* residual; int bias_kind, res_kind;
    int M, N, K, ldc, ldr;
    int a_mn, b_mn;
    int tiles_n, tile_begin;      // tiles of this problem occupy [tile_begin, tile_begin + slices * tiles_m * tiles_n)
    int tiles_mn, kb_per_slice;   // split-K: slice = local tile / tiles_mn, k-blocks [slice * kb_per_slice, ...)
    long long split_stride;
    int seg_len; unsigned seg_map_a, seg_map_b;      // segmented contraction (gemm_tc.h)
    int accumulate;
    void* C16; int c16_kind, ldc16;   // optional 16-bit copy of C (0 none, 1 bf16, 2 fp16); C itself may then be null
    float alpha;
};
struct TcGemmArgs {
    CUtensorMap ta[kTcGemmMax], tb[kTcGemmMax];
    TcGemmDev p[kTcGemmMax];
    int count, total_tiles;
    int ab_fmt;                   // kFmtBF16 / kFmtF16
    int* error_flag;
    int prof_slot, debug_skip;    // stall-accounting build only (debug_skip: 1 no global stores, 2 no staging, 4 no TMEM load; results are garbage)
};

#ifdef DCLIP_GEMM_PROF
__device__ long long g_gemm_prof[16];
__device__ long long g_gemm_prof_ring[8][16];      // the last eight launches (slot = launch counter % 8), CTA 0
#define GPROF(k) do { if (blockIdx.x == 0) { const long long c_ = clock64(); g_gemm_prof[k] = c_; g_gemm_prof_ring[g.prof_slot][k] = c_; } } while (0)
#else
#define GPROF(k) do { } while (0)
#endif
#ifdef DCLIP_GEMM_PROF
#define GSKIPV (g.debug_skip)
#else
#define GSKIPV 0
#endif

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) { const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }
__device__ __forceinline__ uint32_t pack_f16(float a, float b) { const __half2 v = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }

struct TileCoord { int pi, m0, n0, nkb, kb0; long long c_off; };      // k-blocks [kb0, kb0 + nkb); c_off: split-K partial offset

// ---- epilogue.  TMEM hands each thread one output ROW (lane) and 32 consecutive columns per load; global memory wants a
// warp on few rows (a store in which every lane touches a different row costs ~32 memory-pipe cycles).  So a chunk goes
// through a private 4 KB shared-memory buffer (16-byte chunks XOR-swizzled by the row: conflict-free both ways) and comes
// back with 8 lanes per row = 128 contiguous bytes per row and store.  The epilogue is INSTRUCTION-bound (8 warps, ~13 k
// cycles per tile for a first version that tested the problem's optional features inside the store loop, against 5.6 k
// of main loop at K = 768), so it is specialised per tile:
//   MODE 0  16-bit output (+ bias)                          64-column chunks, packed to 16-bit BEFORE staging
//   MODE 1  fp32 output + bias + 16-bit residual            32-column passes staged as fp32; the residual is added at
//           (out-projection -> pre-LayerNorm sum)           read-back from coalesced 8-byte loads requested a pass ahead
//   MODE 2  anything else (accumulate, fp32 residual, both outputs): the general form of mode 1
// N % 8 == 0 (host-checked): a 4-/8-column group is inside C or not.
struct EpiCtx { int qd, we, half, lane; uint32_t stage_s, bias_s; };

__device__ __forceinline__ void epi_load_res16(const TcGemmDev& P, int row0, int ccol, uint2 (&raw)[8]) {
    // rows row0 + 4k, columns ccol .. ccol + 3 of a 16-bit residual
    const uint16_t* src = static_cast<const uint16_t*>(P.residual) + (size_t)row0 * P.ldr + ccol;
    const int rows_left = ccol < P.N ? P.M - row0 : 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        raw[k] = make_uint2(0u, 0u);
        if (4 * k < rows_left) raw[k] = *reinterpret_cast<const uint2*>(src + (size_t)(4 * k) * P.ldr);
    }
}
template <bool BF> __device__ __forceinline__ void widen2(uint32_t w, float& lo, float& hi) {
    if (BF) { lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u); }
    else { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w)); lo = f.x; hi = f.y; }
}
template <bool BF> __device__ __forceinline__ uint32_t pack2(float a, float b) { return BF ? pack_bf16(a, b) : pack_f16(a, b); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) { const uint4 v = lds128(addr); return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)); }

template <int MODE, bool BF>
__device__ __forceinline__ void epilogue_tile(const TcGemmDev& P, const TileCoord& tc, const EpiCtx& E, uint32_t tacc, uint2 (&res0)[8], int skip) {
    const int lane = E.lane;
    const int rb_row = lane >> 3, rb_chunk = lane & 7;           // read-back: 8 lanes per row, 4 rows per store
    const int row0 = tc.m0 + E.qd * 32 + rb_row;                 // my first read-back row
    const int col0 = tc.n0 + E.half * (GN / 2);                  // first column of this warp
    const bool has_bias = P.bias != nullptr;
    const uint32_t st_wr = E.stage_s + lane * 128, st_sw = (uint32_t)(lane & 7);
    if (MODE == 0) {
        uint16_t* c16 = static_cast<uint16_t*>(P.C16);
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            const int col = col0 + ch * 64;
            if (col >= P.N) break;                               // warp-uniform
            float v[64];
            tmem_ld_32x32b_x32(tacc + ch * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld_32x32b_x32(tacc + ch * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            // (columns past N hold whatever the MMA left there - TMA zero-fills the operands, so zeros; never stored)
            tmem_ld_wait();
            if (P.alpha != 1.0f) {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] *= P.alpha;
            }
            if (has_bias) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 bq = lds_f4(E.bias_s + (uint32_t)(E.half * (GN / 2) + ch * 64 + q * 4) * 4);      // broadcast
                    v[q * 4] += bq.x; v[q * 4 + 1] += bq.y; v[q * 4 + 2] += bq.z; v[q * 4 + 3] += bq.w;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                sts128(st_wr + ((q ^ st_sw) << 4), make_uint4(pack2<BF>(v[q * 8], v[q * 8 + 1]), pack2<BF>(v[q * 8 + 2], v[q * 8 + 3]),
                                                                pack2<BF>(v[q * 8 + 4], v[q * 8 + 5]), pack2<BF>(v[q * 8 + 6], v[q * 8 + 7])));
            __syncwarp();
            const int ccol = col + rb_chunk * 8;
            const int rows_left = (ccol < P.N && !(skip & 1)) ? P.M - row0 : 0;
            uint16_t* dst = c16 + (size_t)row0 * P.ldc16 + ccol;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int rr = rb_row + 4 * k;
                const uint4 x = lds128(E.stage_s + rr * 128 + ((rb_chunk ^ (rr & 7)) << 4));
                if (4 * k < rows_left) *reinterpret_cast<uint4*>(dst + (size_t)(4 * k) * P.ldc16) = x;
            }
            __syncwarp();
        }
        return;
    }
    // MODE 1 / 2: 32-column passes staged as fp32; lane (r, c) owns columns 4c .. 4c+3 of the rows r, r + 4, ...
    constexpr int kPasses = GN / 2 / 32;
    const bool res16 = MODE == 1 || (P.residual != nullptr && P.res_kind != 0);
#pragma unroll 1
    for (int ps = 0; ps < kPasses; ++ps) {
        const int col = col0 + ps * 32;
        if (col >= P.N) break;                                   // warp-uniform
        const int ccol = col + rb_chunk * 4;
        float v[32];
        tmem_ld_32x32b_x32(tacc + ps * 32, *reinterpret_cast<uint32_t(*)[32]>(v));
        tmem_ld_wait();
        if (P.alpha != 1.0f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= P.alpha;
        }
        if (has_bias) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 bq = lds_f4(E.bias_s + (uint32_t)(E.half * (GN / 2) + ps * 32 + q * 4) * 4);      // broadcast
                v[q * 4] += bq.x; v[q * 4 + 1] += bq.y; v[q * 4 + 2] += bq.z; v[q * 4 + 3] += bq.w;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            sts128(st_wr + ((q ^ st_sw) << 4), make_uint4(__float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]), __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3])));
        __syncwarp();
        uint2 cur[8];
        if (res16) {
            if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) cur[k] = res0[k];
                if (ps + 1 < kPasses && col + 32 < P.N) epi_load_res16(P, row0, ccol + 32, res0);      // next pass: in flight under this one
            } else epi_load_res16(P, row0, ccol, cur);
        }
        const int rows_left = (ccol < P.N && !(skip & 1)) ? P.M - row0 : 0;
        float* dst = MODE == 1 || P.C ? P.C + (size_t)row0 * P.ldc + ccol : nullptr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int rr = rb_row + 4 * k;
            float4 o = lds_f4(E.stage_s + rr * 128 + ((rb_chunk ^ (rr & 7)) << 4));
            if (res16) {
                float r0, r1, r2, r3;
                widen2<BF>(cur[k].x, r0, r1); widen2<BF>(cur[k].y, r2, r3);
                o.x += r0; o.y += r1; o.z += r2; o.w += r3;
            }
            if (4 * k < rows_left) {
                if (MODE == 2) {
                    if (P.residual && P.res_kind == 0) {
                        const float4 r = *reinterpret_cast<const float4*>(static_cast<const float*>(P.residual) + (size_t)(row0 + 4 * k) * P.ldr + ccol);
                        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                    }
                    if (P.C && P.accumulate) {
                        const float4 r = *reinterpret_cast<const float4*>(dst + (size_t)(4 * k) * P.ldc);
                        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                    }
                    if (P.c16_kind)
                        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(P.C16) + (size_t)(row0 + 4 * k) * P.ldc16 + ccol) = make_uint2(pack2<BF>(o.x, o.y), pack2<BF>(o.z, o.w));
                }
                if (MODE == 1 || P.C) *reinterpret_cast<float4*>(dst + (size_t)(4 * k) * P.ldc) = o;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ TileCoord tile_coord(const TcGemmArgs& g, int tile) {
    int pi = 0;
    for (int i = 1; i < g.count; ++i)
        if (tile >= g.p[i].tile_begin) pi = i;
    const TcGemmDev& P = g.p[pi];
    int local = tile - P.tile_begin;
    const int slice = local / P.tiles_mn;
    local -= slice * P.tiles_mn;
    const int tm = local / P.tiles_n, tn = local - tm * P.tiles_n;
    const int nkb_all = (P.K + GKS - 1) / GKS;
    const int kb0 = slice * P.kb_per_slice;
    return TileCoord{pi, tm * GM, tn * GN, min(P.kb_per_slice, nkb_all - kb0), kb0, (long long)slice * P.split_stride};
}

// Persistent: gridDim.x CTAs walk the tile list with stride gridDim.x.  The TMA ring keeps flowing across
// tile boundaries and the two 256-column TMEM accumulators let the epilogue of tile t overlap the main loop
// of tile t+1, so the per-tile prologue / epilogue latency is paid once per CTA, not once per tile.
__global__ void __launch_bounds__(kGThreads, 1) k_gemm_tc(const __grid_constant__ TcGemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kGStages * kGStageBytes);
    uint64_t* empty = full + kGStages;
    uint64_t* acc_full = empty + kGStages;       // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int total_tiles = g.total_tiles;

    if (threadIdx.x == 0) GPROF(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kGStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kGEpiWarps * 32); }
        fence_barrier_init();
    }
    if (warp == kGStages) tmem_alloc(tmem_ptr, 512);
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = *tmem_ptr;
    bool ok = true;
    if (threadIdx.x == 0) GPROF(1);

    if (warp < kGStages) {
        // ---- TMA producers: warp w owns ring stage w; stage job q = running k-block index over all my tiles
        if (elect_one()) {
            for (int i = 0; i < g.count; ++i) { tma_prefetch_desc(&g.ta[i]); tma_prefetch_desc(&g.tb[i]); }
            uint8_t* dst = ring + (size_t)warp * kGStageBytes;
            uint32_t phase = 0;
            int q0 = 0;                                    // job index of the first k-block of the current tile
            for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x) {
                const TileCoord tc = tile_coord(g, tile);
                const TcGemmDev& P = g.p[tc.pi];
                const CUtensorMap* ta = &g.ta[tc.pi];
                const CUtensorMap* tb = &g.tb[tc.pi];
                // my k-blocks of this tile: those with (q0 + kb) % kGStages == warp
                int kb = (warp - q0 % kGStages + kGStages) % kGStages;
                for (; kb < tc.nkb && ok; kb += kGStages, phase ^= 1) {
                    const int k0 = (tc.kb0 + kb) * GKS;
                    int ka = k0, kbb = k0;                 // physical k of the two operands
                    if (P.seg_len) {
                        const int sidx = k0 / P.seg_len, off = k0 - sidx * P.seg_len;
                        ka = (int)((P.seg_map_a >> (4 * sidx)) & 15u) * P.seg_len + off;
                        kbb = (int)((P.seg_map_b >> (4 * sidx)) & 15u) * P.seg_len + off;
                    }
                    ok = mbar_wait(&empty[warp], phase ^ 1, kSpinG);
                    if (warp == 0 && kb == 0 && tile == (int)blockIdx.x) GPROF(2);
                    mbar_arrive_expect_tx(&full[warp], kGStageBytes);
                    if (!P.a_mn) tma_load_2d(dst, ta, &full[warp], ka, tc.m0);                     // [128 m x 64 k]
                    else
                        for (int h = 0; h < 2; ++h) tma_load_2d(dst + h * 8192, ta, &full[warp], tc.m0 + h * 64, ka);   // [64 k x 64 m]
                    if (!P.b_mn) tma_load_2d(dst + kGAStage, tb, &full[warp], kbb, tc.n0);          // [256 n x 64 k]
                    else
                        for (int h = 0; h < 4; ++h) tma_load_2d(dst + kGAStage + h * 8192, tb, &full[warp], tc.n0 + h * 64, kbb);
                }
                q0 += tc.nkb;
            }
        }
    } else if (warp == kGStages) {
        // ---- MMA issuer
        if (elect_one()) {
            const uint32_t ring_addr = smem_u32(ring);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles && ok; tile += gridDim.x, ++it) {
                const TileCoord tc = tile_coord(g, tile);
                const TcGemmDev& P = g.p[tc.pi];
                const int b = it & 1;
                const uint32_t idesc = make_idesc_f16(GM, GN, (uint32_t)g.ab_fmt, (uint32_t)P.a_mn, (uint32_t)P.b_mn);
                const uint64_t hiA = make_smem_desc_hi(P.a_mn ? 8192 : 0, 1024, kSwizzle128B);
                const uint64_t hiB = make_smem_desc_hi(P.b_mn ? 8192 : 0, 1024, kSwizzle128B);
                const uint32_t stepA = P.a_mn ? 128u : 2u, stepB = P.b_mn ? 128u : 2u;   // per K=16 step, >>4 encoded
                ok = mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1, kSpinG);
                fence_after_thread_sync();
                for (int kb = 0; kb < tc.nkb && ok; ++kb) {
                    ok = mbar_wait(&full[stage], phase, kSpinG);
                    if (it == 0 && kb == 0) GPROF(3);
                    if (it == 0 && kb == tc.nkb - 1) GPROF(4);
                    fence_after_thread_sync();
                    const uint64_t da0 = smem_desc(hiA, ring_addr + stage * kGStageBytes);
                    const uint64_t db0 = smem_desc(hiB, ring_addr + stage * kGStageBytes + kGAStage);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tmem + b * GN, da0 + (uint64_t)(ks * stepA), db0 + (uint64_t)(ks * stepB), idesc, (kb | ks) != 0);
                    umma_commit(&empty[stage]);
                    if (++stage == kGStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[b]);
            }
        }
    } else {
        // ---- epilogue (epilogue_tile above), eight warps: warp % 4 is the TMEM lane quadrant it may read, the two warps of a
        // quadrant split the tile's columns.  The tile's bias row is staged once per tile in shared memory (before the
        // accumulator barrier: hidden under the main loop).
        EpiCtx E;
        E.qd = warp & 3;
        E.we = warp - (kGStages + 1);
        E.half = E.we >> 2;
        E.lane = lane;
        E.stage_s = smem_u32(ring + kGStages * kGStageBytes + 256) + (uint32_t)E.we * 4096;
        E.bias_s = smem_u32(ring + kGStages * kGStageBytes + 256 + kGEpiWarps * 4096);
        const bool bf = g.ab_fmt == (int)kFmtBF16;
        const int et = (int)threadIdx.x - (kGStages + 1) * 32;       // 0 .. 255: my bias column
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc = tile_coord(g, tile);
            TcGemmDev P = g.p[tc.pi];
            if (P.C) P.C += tc.c_off;            // split-K: this slice's partial product
            const int b = it & 1;
            if (P.bias) {      // uniform over the CTA's epilogue warps (same tile)
                float bv = 0.f;
                if (tc.n0 + et < P.N) {
                    if (P.bias_kind == 0) bv = static_cast<const float*>(P.bias)[tc.n0 + et];
                    else if (bf) bv = __bfloat162float(static_cast<const __nv_bfloat16*>(P.bias)[tc.n0 + et]);
                    else bv = __half2float(static_cast<const __half*>(P.bias)[tc.n0 + et]);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");      // the previous tile's readers are done
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(E.bias_s + et * 4), "f"(bv) : "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const int mode = (!P.C && P.c16_kind && !P.residual) ? 0 : (P.C && !P.c16_kind && P.residual && P.res_kind != 0 && !P.accumulate) ? 1 : 2;
            // mode 1: the first pass's residual words are requested before the accumulator barrier
            uint2 res0[8];
            if (mode == 1) epi_load_res16(P, tc.m0 + E.qd * 32 + (lane >> 3), tc.n0 + E.half * (GN / 2) + (lane & 7) * 4, res0);
            ok = mbar_wait(&acc_full[b], (it >> 1) & 1, kSpinG);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            if (it == 0 && warp == kGStages + 1 && lane == 0) GPROF(5);
            fence_after_thread_sync();
            const uint32_t tacc = tmem + (static_cast<uint32_t>(E.qd * 32) << 16) + b * GN + E.half * (GN / 2);
            if (mode == 0) { if (bf) epilogue_tile<0, true>(P, tc, E, tacc, res0, GSKIPV); else epilogue_tile<0, false>(P, tc, E, tacc, res0, GSKIPV); }
            else if (mode == 1) { if (bf) epilogue_tile<1, true>(P, tc, E, tacc, res0, GSKIPV); else epilogue_tile<1, false>(P, tc, E, tacc, res0, GSKIPV); }
            else { if (bf) epilogue_tile<2, true>(P, tc, E, tacc, res0, GSKIPV); else epilogue_tile<2, false>(P, tc, E, tacc, res0, GSKIPV); }
            fence_before_thread_sync();
            mbar_arrive(&acc_empty[b]);
            if (it == 0 && warp == kGStages + 1 && lane == 0) GPROF(6);
        }
    }
    if (!ok && lane == 0 && g.error_flag) atomicExch(g.error_flag, 1);
    fence_before_thread_sync();
    __syncthreads();
    if (warp == kGStages) tmem_dealloc(tmem, 512);
    if (threadIdx.x == 0) GPROF(7);
}

int* tc_error_flag();   // loss_tc.cu

int gemm_tc_slices(int K, int ksplit) {
    const int nkb = (K + GKS - 1) / GKS;
    const int ks = ksplit > 1 ? min(ksplit, nkb) : 1;
    const int per = (nkb + ks - 1) / ks;
    return (nkb + per - 1) / per;
}

// dtype16: 1 = bf16, 2 = fp16 (dclip_dtype values)
int launch_gemm_tc(const TcGemmProblem* probs, int count, int dtype16, cudaStream_t st) {
    if (count < 1 || count > kTcGemmMax) return fail(DCLIP_ERR_INVALID, "gemm_tc: 1..%d problems per launch", kTcGemmMax);
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "tensor path needs an sm_100 device");
    TcGemmArgs g{};
    g.count = count;
    g.ab_fmt = dtype16 == DCLIP_BF16 ? (int)kFmtBF16 : (int)kFmtF16;
    g.error_flag = tc_error_flag();
    const int dt = dtype16 == DCLIP_BF16 ? 1 : 0;
    int tiles = 0;
    for (int i = 0; i < count; ++i) {
        const TcGemmProblem& q = probs[i];
        if (q.M <= 0 || q.N <= 0 || q.K <= 0) return fail(DCLIP_ERR_INVALID, "gemm_tc: empty problem %d", i);
        if (!q.C && !q.C16) return fail(DCLIP_ERR_INVALID, "gemm_tc: problem %d has no output", i);
        if (q.accumulate && !q.C) return fail(DCLIP_ERR_INVALID, "gemm_tc: accumulate needs the fp32 output");
        if (q.lda % 8 || q.ldb % 8 || q.ldc % 4) return fail(DCLIP_ERR_INVALID, "gemm_tc: leading dimensions must keep 16-byte row alignment");
        if (q.N % 8) return fail(DCLIP_ERR_INVALID, "gemm_tc: N must be a multiple of 8 (got %d)", q.N);
        if ((q.C && ((reinterpret_cast<uintptr_t>(q.C) & 15) || q.ldc % 4)) || (q.C16 && ((reinterpret_cast<uintptr_t>(q.C16) & 15) || (q.ldc16 > 0 ? q.ldc16 : q.ldc) % 8)))
            return fail(DCLIP_ERR_INVALID, "gemm_tc: outputs must be 16-byte aligned with 16-byte aligned rows");
        if ((q.bias && (reinterpret_cast<uintptr_t>(q.bias) & 15)) || (q.residual && ((reinterpret_cast<uintptr_t>(q.residual) & 15) || q.ldr % (q.res_kind ? 8 : 4))))
            return fail(DCLIP_ERR_INVALID, "gemm_tc: bias / residual must be 16-byte aligned");
        bool okm;
        // A: K-major stored [M, K] (rows = M, cols = K); MN-major stored [K, M]
        if (q.seg_len && (q.seg_len % GKS || q.K % q.seg_len || q.K / q.seg_len > 8 || q.k_phys <= 0 || q.k_phys % q.seg_len))
            return fail(DCLIP_ERR_INVALID, "gemm_tc: problem %d: bad segmented contraction (seg_len %d, K %d, k_phys %d)", i, q.seg_len, q.K, q.k_phys);
        const uint64_t kext = q.seg_len ? (uint64_t)q.k_phys : (uint64_t)q.K;      // K extent of the STORED operands
        if (!q.a_mn) okm = make_tmap_2d_sw128_ld(&g.ta[i], q.A, (uint64_t)q.M, kext, (uint64_t)q.lda, GM, dt);
        else okm = make_tmap_2d_sw128_ld(&g.ta[i], q.A, kext, (uint64_t)q.M, (uint64_t)q.lda, GKS, dt);
        if (!q.b_mn) okm = okm && make_tmap_2d_sw128_ld(&g.tb[i], q.B, (uint64_t)q.N, kext, (uint64_t)q.ldb, GN, dt);
        else okm = okm && make_tmap_2d_sw128_ld(&g.tb[i], q.B, kext, (uint64_t)q.N, (uint64_t)q.ldb, GKS, dt);
        if (!okm) return fail(DCLIP_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled failed for problem %d", i);
        TcGemmDev& d = g.p[i];
        d.C = q.C; d.bias = q.bias; d.residual = q.residual; d.bias_kind = q.bias_kind; d.res_kind = q.res_kind;

        d.M = q.M; d.N = q.N; d.K = q.K; d.ldc = q.ldc; d.ldr = q.ldr;
        d.a_mn = q.a_mn; d.b_mn = q.b_mn; d.accumulate = q.accumulate;
        d.C16 = q.C16; d.c16_kind = q.C16 ? (dtype16 == DCLIP_BF16 ? 1 : 2) : 0; d.ldc16 = q.ldc16 > 0 ? q.ldc16 : q.ldc;
        d.alpha = q.alpha == 0.f ? 1.0f : q.alpha;
        d.seg_len = q.seg_len; d.seg_map_a = q.seg_map_a; d.seg_map_b = q.seg_map_b;
        d.tiles_n = (q.N + GN - 1) / GN;
        d.tile_begin = tiles;
        d.tiles_mn = ((q.M + GM - 1) / GM) * d.tiles_n;
        const int nkb = (q.K + GKS - 1) / GKS;
        int ks = q.ksplit > 1 ? min(q.ksplit, nkb) : 1;
        if (ks > 1 && (q.bias || q.residual || q.accumulate || q.C16 || !q.C)) return fail(DCLIP_ERR_INVALID, "gemm_tc: split-K problem %d must be a plain fp32 product", i);
        d.kb_per_slice = (nkb + ks - 1) / ks;
        ks = (nkb + d.kb_per_slice - 1) / d.kb_per_slice;            // no empty slice
        d.split_stride = ks > 1 ? q.split_stride : 0;
        tiles += ks * d.tiles_mn;
    }
    g.total_tiles = tiles;
#ifdef DCLIP_GEMM_PROF
    { static int launches = 0; g.prof_slot = launches++ & 7; const char* e = getenv("DCLIP_GEMM_SKIP"); g.debug_skip = e ? atoi(e) : 0; }
#endif
    const size_t smem = 1024 + (size_t)kGStages * kGStageBytes + 256 + kGEpiWarps * 4096 + GN * sizeof(float);
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_tc<<<min(tiles, sm_count()), kGThreads, smem, st>>>(g);
    DCLIP_LAUNCH_OK("k_gemm_tc");
    return DCLIP_OK;
}

}  // namespace dclip

#ifdef DCLIP_GEMM_PROF
extern "C" int dclip_debug_gemm_prof(long long* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(out16, dclip::g_gemm_prof, 16 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
extern "C" int dclip_debug_gemm_prof_ring(long long* out128) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(out128, dclip::g_gemm_prof_ring, 128 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif
