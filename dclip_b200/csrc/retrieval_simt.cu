// Retrieval — FP32 CUDA-core path implementing the bit-exact score contract.
//
// Replaces eval_scripts/flickr30k_eval.py:243-266 (similarity build) + :16-88
// (calculate_retrieval_metrics) and the faiss IndexFlatIP add/search behind
// training/compute_faiss.py:27,79 / training/image_tokenizer.py:262.
//
// Contract (identical to oracle/retrieval_oracle.c):
//   score(q, g) = acc_d, acc_{k+1} = fmaf(q[k], g[k], acc_k), fp32, k ascending, acc_0 = 0
//   order = (-score, +index)
// Every output element of the tile GEMM below is exactly that chain (one accumulator per
// output, k-chunks visited in ascending order), so ranks and top-k lists are bit-exact,
// ties included.  The score matrix never leaves registers / shared memory.
#include "common.cuh"
#include "loss_internal.h"

namespace dclip {

constexpr int RT = 64;          // tile: 64 queries x 64 gallery rows
constexpr int RK = 16;          // k-chunk
constexpr int RLD = RT + 4;     // smem row stride (floats)
constexpr int RThreads = 256;

template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// ---------------------------------------------------------------------------------------------
// The contract's score is a strictly SEQUENTIAL fp32 fmaf chain over k, so one thread has to walk it - but nothing says that
// thread must also wait for every load.  A warp stages 512-element chunks of the two rows in its slice of shared memory with
// coalesced 16-byte loads; lane 0 then walks the chunk out of shared memory (LDS.128 + 4 dependent FMAs).  One thread doing
// both was latency-bound: 17 us for 5000 pairs of d = 512, 31 us to normalise them.
// ---------------------------------------------------------------------------------------------
constexpr int kChainChunk = 512;                 // elements per staged chunk
constexpr int kChainWarps = 8;                   // warps per CTA (2 x 512 floats of shared memory each)

// returns (in lane 0) fmaf-chain sum_k a[k] * b[k];  `sa`, `sb`: this warp's two kChainChunk-float buffers (sb == sa when b == a)
template <typename TA, typename TB>
__device__ __forceinline__ float warp_chain_dot(const TA* __restrict__ a, const TB* __restrict__ b, int64_t d, float* sa, float* sb, int lane) {
    float acc = 0.f;
    const bool same = (const void*)a == (const void*)b;
    for (int64_t c0 = 0; c0 < d; c0 += kChainChunk) {
        const int cw = (int)min((int64_t)kChainChunk, d - c0);       // multiple of 4 (host-checked)
        __syncwarp();
        for (int k = lane * 4; k < cw; k += 128) {
            *reinterpret_cast<float4*>(sa + k) = ld4<TA>(a + c0 + k);
            if (!same) *reinterpret_cast<float4*>(sb + k) = ld4<TB>(b + c0 + k);
        }
        __syncwarp();
        if (lane == 0) {
            const float* pb = same ? sa : sb;
#pragma unroll 4
            for (int k = 0; k < cw; k += 4) {
                const float4 x = *reinterpret_cast<const float4*>(sa + k), y = *reinterpret_cast<const float4*>(pb + k);
                acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
            }
        }
    }
    return acc;
}

// row normalisation with the contract's sequential sum of squares (numpy semantics: no eps): one warp per row
template <typename T>
__global__ void __launch_bounds__(kChainWarps * 32) k_normalize_rows(const T* __restrict__ x, int64_t n, int64_t d, float* __restrict__ out) {
    __shared__ __align__(16) float buf[kChainWarps][kChainChunk];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * kChainWarps + w;
    if (row >= n) return;
    const T* xr = x + row * d;
    float ss = warp_chain_dot<T, T>(xr, xr, d, buf[w], buf[w], lane);
    const float nrm = sqrtf(__shfl_sync(0xffffffffu, ss, 0));
    for (int64_t k = lane * 4; k < d; k += 128) {
        const float4 v = ld4<T>(xr + k);
        *reinterpret_cast<float4*>(out + row * d + k) = make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
    }
}

// one warp per pair (staged chunks, lane 0 walks the chain)
__global__ void __launch_bounds__(kChainWarps * 32) k_pair_scores(const float* __restrict__ q, const float* __restrict__ g, int64_t d,
                                                                  const int64_t* __restrict__ qidx, const int64_t* __restrict__ gidx, int64_t npairs,
                                                                  float* __restrict__ out) {
    __shared__ __align__(16) float buf[kChainWarps][2][kChainChunk];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t p = (int64_t)blockIdx.x * kChainWarps + w;
    if (p >= npairs) return;
    const float acc = warp_chain_dot<float, float>(q + (qidx ? qidx[p] : p) * d, g + gidx[p] * d, d, buf[w][0], buf[w][1], lane);
    if (lane == 0) out[p] = acc;
}

// ---------------------------------------------------------------------------------------------
// 64x64 score tile, 4x4 per thread, operands staged k-major in smem
// ---------------------------------------------------------------------------------------------
template <typename TQ, typename TG>
struct ScoreTile {
    float acc[4][4];
    // rows: queries (optionally gathered through qidx), cols: gallery rows col0..col0+63
    __device__ __forceinline__ void compute(float (*As)[RLD], float (*Bs)[RLD], const TQ* __restrict__ q,
                                            const int64_t* __restrict__ qidx, int64_t row0, int64_t nrows,
                                            const TG* __restrict__ g, int64_t col0, int64_t ncols, int64_t d) {
        const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row lr, k offset lk
        const int64_t qrow = row0 + lr, grow = col0 + lr;
        const TQ* qp = nullptr;
        const TG* gp = nullptr;
        if (qrow < nrows) qp = q + (qidx ? qidx[qrow] : qrow) * d;
        if (grow < ncols) gp = g + grow * d;
        for (int64_t k0 = 0; k0 < d; k0 += RK) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (k0 + lk < d) {   // d % 4 == 0
                if (qp) a = ld4<TQ>(qp + k0 + lk);
                if (gp) b = ld4<TG>(gp + k0 + lk);
            }
            __syncthreads();
            As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
            Bs[lk + 0][lr] = b.x; Bs[lk + 1][lr] = b.y; Bs[lk + 2][lr] = b.z; Bs[lk + 3][lr] = b.w;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RK; ++k) {
                const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
        }
    }
};
// note: zero-padded k beyond d adds fmaf(0, 0, acc) == acc exactly, so the chain is unchanged.

// ---------------------------------------------------------------------------------------------
// rank counts: rank[p] += #{ j in shard : s_pj > ts[p] or (s_pj == ts[p] and goff + j < tgt[p]) }
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RThreads) k_rank_counts(const float* __restrict__ q, const float* __restrict__ g,
                                                          int64_t ng, int64_t d, const int64_t* __restrict__ qidx,
                                                          const float* __restrict__ ts, const int64_t* __restrict__ tgt,
                                                          int64_t npairs, int64_t goff, int tiles_per_split,
                                                          int32_t* __restrict__ rank, const int* __restrict__ gate) {
    __shared__ __align__(16) float As[RK][RLD];
    __shared__ __align__(16) float Bs[RK][RLD];
    if (gate && *gate == 0) return;      // fallback launch of the tensor path: only runs when its band list overflowed
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int64_t row0 = (int64_t)blockIdx.x * RT;
    float tsr[4];
    int64_t tgr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = row0 + ty * 4 + i;
        tsr[i] = p < npairs ? ts[p] : 0.f;
        tgr[i] = p < npairs ? tgt[p] : 0;
    }
    int cnt[4] = {0, 0, 0, 0};
    const int64_t ntiles = (ng + RT - 1) / RT;
    const int64_t t0 = (int64_t)blockIdx.y * tiles_per_split, t1 = min(ntiles, t0 + tiles_per_split);
    ScoreTile<float, float> tile;
    for (int64_t jt = t0; jt < t1; ++jt) {
        const int64_t col0 = jt * RT;
        tile.compute(As, Bs, q, qidx, row0, npairs, g, col0, ng, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = col0 + tx * 4 + j;
            if (col < ng) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float s = tile.acc[i][j];
                    cnt[i] += (s > tsr[i] || (s == tsr[i] && goff + col < tgr[i])) ? 1 : 0;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = cnt[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        const int64_t p = row0 + ty * 4 + i;
        if (tx == 0 && p < npairs && c) atomicAdd(&rank[p], c);
    }
}

// Several targets per query row (image -> text direction of the Recall@K evaluation, flickr30k_eval.py:48-64: the column of
// image i is ranked once for each of its ground-truth captions).  The pairs of one query are contiguous (CSR tptr); the
// score tile of a query row is computed ONCE and compared against all of its targets' scores - not once per pair as the
// gather-by-qidx form does (5x fewer FLOPs on the Flickr shape).  Up to kMaxTargets targets per row (host-checked).
constexpr int kMaxTargets = 8;
__global__ void __launch_bounds__(RThreads) k_rank_counts_multi(const float* __restrict__ q, const float* __restrict__ g, int64_t nq,
                                                                int64_t ng, int64_t d, const int64_t* __restrict__ tptr,
                                                                const float* __restrict__ ts, const int64_t* __restrict__ tgt,
                                                                int64_t goff, int tiles_per_split, int32_t* __restrict__ rank) {
    __shared__ __align__(16) float As[RK][RLD];
    __shared__ __align__(16) float Bs[RK][RLD];
    __shared__ float ts_s[RT][kMaxTargets];
    __shared__ long long tg_s[RT][kMaxTargets];
    __shared__ int cnt_s[RT][kMaxTargets];
    __shared__ int nt_s[RT];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int64_t row0 = (int64_t)blockIdx.x * RT;
    const unsigned half_mask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;     // the 16 lanes that share my rows
    for (int e = threadIdx.x; e < RT * kMaxTargets; e += RThreads) {
        const int r = e / kMaxTargets, k = e - r * kMaxTargets;
        const int64_t row = row0 + r;
        int n = 0;
        int64_t p0 = 0;
        if (row < nq) { p0 = tptr[row]; n = (int)(tptr[row + 1] - p0); }
        if (k == 0) nt_s[r] = n;
        ts_s[r][k] = k < n ? ts[p0 + k] : 0.f;
        tg_s[r][k] = k < n ? tgt[p0 + k] : 0;
        cnt_s[r][k] = 0;
    }
    __syncthreads();
    const int64_t ntiles = (ng + RT - 1) / RT;
    const int64_t t0 = (int64_t)blockIdx.y * tiles_per_split, t1 = min(ntiles, t0 + tiles_per_split);
    ScoreTile<float, float> tile;
    for (int64_t jt = t0; jt < t1; ++jt) {
        const int64_t col0 = jt * RT;
        tile.compute(As, Bs, q, nullptr, row0, nq, g, col0, ng, d);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            const int n = nt_s[r];                          // same for the 16 threads (half warp) that share the row
            for (int k = 0; k < n; ++k) {
                const float st = ts_s[r][k];
                const long long tg = tg_s[r][k];
                int c = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t col = col0 + tx * 4 + j;
                    const float sv = tile.acc[i][j];
                    c += (col < ng && (sv > st || (sv == st && goff + col < tg))) ? 1 : 0;
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(half_mask, c, o, 16);   // the other half-warp may run a different target count
                if (tx == 0 && c) cnt_s[r][k] += c;         // one writer per (row, target)
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < RT * kMaxTargets; e += RThreads) {
        const int r = e / kMaxTargets, k = e - r * kMaxTargets;
        if (k < nt_s[r] && cnt_s[r][k]) atomicAdd(&rank[tptr[row0 + r] + k], cnt_s[r][k]);
    }
}

// ---------------------------------------------------------------------------------------------
// warp-resident sorted list (lane l holds the l-th best), order (-score, +index)
// ---------------------------------------------------------------------------------------------
constexpr int64_t kNoIndex = INT64_MAX;
struct WarpList {
    float s;
    int64_t i;
    __device__ __forceinline__ void init() { s = -INFINITY; i = kNoIndex; }
    // true if (cs, ci) ranks strictly before this lane's element
    __device__ __forceinline__ bool cand_before(float cs, int64_t ci) const {
        return (i == kNoIndex) || cs > s || (cs == s && ci < i);
    }
    __device__ __forceinline__ void insert(float cs, int64_t ci, int lane) {
        const unsigned after = __ballot_sync(0xffffffffu, cand_before(cs, ci));   // lanes displaced by the candidate
        if (after == 0) return;
        const int pos = __ffs(after) - 1;
        const float ps = __shfl_up_sync(0xffffffffu, s, 1);
        const int64_t pi = __shfl_up_sync(0xffffffffu, i, 1);
        if (lane == pos) { s = cs; i = ci; }
        else if (lane > pos) { s = ps; i = pi; }
    }
};

template <typename TQ, typename TG>
__global__ void __launch_bounds__(RThreads) k_topk_ip(const TQ* __restrict__ q, const TG* __restrict__ g, int64_t nq,
                                                      int64_t ng, int64_t d, int k, int64_t goff, int tiles_per_split,
                                                      float* __restrict__ out_s, int64_t* __restrict__ out_i) {
    __shared__ __align__(16) float As[RK][RLD];
    __shared__ __align__(16) float Bs[RK][RLD];
    __shared__ float Ss[RT][RT + 1];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * RT;
    WarpList lists[8];   // warp w owns query rows w*8 .. w*8+7 of the tile
#pragma unroll
    for (int r = 0; r < 8; ++r) lists[r].init();
    const int64_t ntiles = (ng + RT - 1) / RT;
    const int64_t t0 = (int64_t)blockIdx.y * tiles_per_split, t1 = min(ntiles, t0 + tiles_per_split);
    ScoreTile<TQ, TG> tile;
    for (int64_t jt = t0; jt < t1; ++jt) {
        const int64_t col0 = jt * RT;
        tile.compute(As, Bs, q, nullptr, row0, nq, g, col0, ng, d);
        __syncthreads();   // previous epilogue finished reading Ss
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Ss[ty * 4 + i][tx * 4 + j] = tile.acc[i][j];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = warp * 8 + r;
            if (row0 + row >= nq) continue;   // warp-uniform
            // threshold = current k-th element (lane k-1)
            const float thr_s = __shfl_sync(0xffffffffu, lists[r].s, k - 1);
            const int64_t thr_i = __shfl_sync(0xffffffffu, lists[r].i, k - 1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                const int64_t col = col0 + c;
                const float s = Ss[row][c];
                const int64_t gi = goff + col;
                const bool hit = col < ng && (thr_i == kNoIndex || s > thr_s || (s == thr_s && gi < thr_i));
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float cs = __shfl_sync(0xffffffffu, s, src);
                    const int64_t ci = __shfl_sync(0xffffffffu, gi, src);
                    lists[r].insert(cs, ci, lane);
                }
            }
        }
    }
    // write this split's lists: [split][nq][k]
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t qrow = row0 + warp * 8 + r;
        if (qrow < nq && lane < k) {
            const int64_t o = ((int64_t)blockIdx.y * nq + qrow) * k + lane;
            out_s[o] = lists[r].s;
            out_i[o] = lists[r].i == kNoIndex ? -1 : lists[r].i;
        }
    }
}

// one warp per query: merge nlists sorted k-lists
__global__ void k_topk_merge(const float* __restrict__ in_s, const int64_t* __restrict__ in_i, int64_t nlists,
                             int64_t nq, int k, float* __restrict__ out_s, int64_t* __restrict__ out_i) {
    const int lane = threadIdx.x & 31;
    const int64_t qrow = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qrow >= nq) return;
    WarpList list;
    list.init();
    for (int64_t l = 0; l < nlists; ++l) {
        const int64_t base = (l * nq + qrow) * k;
        float s = -INFINITY;
        int64_t gi = -1;
        if (lane < k) { s = in_s[base + lane]; gi = in_i[base + lane]; }
        unsigned m = __ballot_sync(0xffffffffu, gi >= 0);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cs = __shfl_sync(0xffffffffu, s, src);
            const int64_t ci = __shfl_sync(0xffffffffu, gi, src);
            // lists are sorted: once a candidate fails to enter, the rest of this list fails too
            const float thr_s = __shfl_sync(0xffffffffu, list.s, k - 1);
            const int64_t thr_i = __shfl_sync(0xffffffffu, list.i, k - 1);
            if (!(thr_i == kNoIndex || cs > thr_s || (cs == thr_s && ci < thr_i))) break;
            list.insert(cs, ci, lane);
        }
    }
    if (lane < k) {
        out_s[qrow * k + lane] = list.s;
        out_i[qrow * k + lane] = list.i == kNoIndex ? -1 : list.i;
    }
}

// ---------------------------------------------------------------------------------------------
// Recall@K / MRR from ranks (single CTA, fixed-order reductions)
// ---------------------------------------------------------------------------------------------
__global__ void k_group_min_init(int32_t* best, int64_t n_groups) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups) best[i] = INT32_MAX;
}
__global__ void k_group_min(const int32_t* __restrict__ rank, const int64_t* __restrict__ group, int64_t n,
                            int64_t n_groups, int32_t* best) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && group[i] >= 0 && group[i] < n_groups) atomicMin(&best[group[i]], rank[i]);
}
__global__ void __launch_bounds__(1024) k_recall_metrics(const int32_t* __restrict__ rank, int64_t n, double* out4) {
    __shared__ double red[1024];
    __shared__ int cnt[4][32];
    int c1 = 0, c5 = 0, c10 = 0, cv = 0;
    double rr = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const int32_t r = rank[i];
        if (r == INT32_MAX) continue;   // empty group (image without captions): skipped, :55-56
        ++cv; c1 += r < 1; c5 += r < 5; c10 += r < 10;
        rr += 1.0 / ((double)r + 1.0);
    }
    // integer counts: warp reductions, then 32 partials per counter (4096 serialised 64-bit shared atomics cost 45 us here)
    c1 = __reduce_add_sync(0xffffffffu, c1); c5 = __reduce_add_sync(0xffffffffu, c5);
    c10 = __reduce_add_sync(0xffffffffu, c10); cv = __reduce_add_sync(0xffffffffu, cv);
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        cnt[0][w] = c1; cnt[1][w] = c5; cnt[2][w] = c10; cnt[3][w] = cv;
    }
    red[threadIdx.x] = rr;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        long long t[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; ++k)
            for (int w = 0; w < 32; ++w) t[k] += cnt[k][w];
        const double nv = (double)t[3];
        out4[0] = (double)t[0] / nv; out4[1] = (double)t[1] / nv; out4[2] = (double)t[2] / nv;
        out4[3] = red[0] / nv;
    }
}

// ranks inside a materialised matrix (compatibility path): one CTA per pair
template <typename T>
__global__ void __launch_bounds__(256) k_matrix_ranks(const T* __restrict__ sim, int64_t n_rows, int64_t n_cols,
                                                      const int64_t* __restrict__ line, const int64_t* __restrict__ tgt,
                                                      int along_rows, int32_t* __restrict__ rank) {
    __shared__ int red[8];
    const int64_t p = blockIdx.x;
    const int64_t l = line[p], t = tgt[p];
    const int64_t len = along_rows ? n_cols : n_rows;
    const int64_t stride = along_rows ? 1 : n_cols;
    const T* base = along_rows ? sim + l * n_cols : sim + l;
    const T st = base[t * stride];
    int c = 0;
    for (int64_t j = threadIdx.x; j < len; j += 256) {
        const T s = base[j * stride];
        c += (s > st || (s == st && j < t)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        rank[p] = tot;
    }
}

static int split_for(int64_t row_tiles, int64_t col_tiles) {
    int64_t want = ceil_div(2 * (int64_t)sm_count(), row_tiles > 0 ? row_tiles : 1);
    if (want < 1) want = 1;
    if (want > col_tiles) want = col_tiles;
    if (want > 1024) want = 1024;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace dclip

using namespace dclip;

extern "C" {

int dclip_normalize_rows(const void* x, int64_t n, int64_t d, int dtype, float* out, void* stream) {
    DCLIP_REQUIRE(x && out && n >= 0 && d > 0, "normalize_rows: bad arguments");
    if (n == 0) return DCLIP_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_REQUIRE(d % 4 == 0, "normalize_rows: d must be a multiple of 4");
    DCLIP_DISPATCH_DTYPE(dtype, (k_normalize_rows<T><<<(unsigned)ceil_div(n, kChainWarps), kChainWarps * 32, 0, st>>>(static_cast<const T*>(x), n, d, out)));
    DCLIP_LAUNCH_OK("k_normalize_rows");
    return DCLIP_OK;
}

int dclip_pair_scores(const float* q, const float* g, int64_t d, const int64_t* qidx, const int64_t* gidx, int64_t npairs,
                      float* out, void* stream) {
    DCLIP_REQUIRE(q && g && gidx && out && npairs >= 0, "pair_scores: bad arguments");
    DCLIP_REQUIRE(d > 0 && d % 4 == 0, "pair_scores: d must be a multiple of 4");
    if (npairs == 0) return DCLIP_OK;
    k_pair_scores<<<(unsigned)ceil_div(npairs, kChainWarps), kChainWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(q, g, d, qidx, gidx, npairs, out);
    DCLIP_LAUNCH_OK("k_pair_scores");
    return DCLIP_OK;
}

size_t dclip_rank_counts_workspace_bytes(int64_t npairs, int64_t ng, int64_t d, int path) {
    if (path == DCLIP_PATH_TENSOR) return rank_counts_tc_workspace_bytes(npairs, ng, d);
    return 256;
}

int dclip_rank_counts(const float* q, const float* g, int64_t ng, int64_t d, const int64_t* qidx, const float* target_score,
                      const int64_t* target_index, int64_t npairs, int64_t g_index_offset, int32_t* rank, void* workspace,
                      size_t workspace_bytes, int path, void* stream) {
    DCLIP_REQUIRE(q && g && target_score && target_index && rank, "rank_counts: null pointer");
    DCLIP_REQUIRE(d > 0 && d % 4 == 0 && ng >= 0 && npairs >= 0, "rank_counts: bad shape (d must be a multiple of 4)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (npairs == 0) return DCLIP_OK;
    DCLIP_CUDA_OK(cudaMemsetAsync(rank, 0, sizeof(int32_t) * npairs, st));
    if (ng == 0) return DCLIP_OK;
    const int* gate = nullptr;
    if (path == DCLIP_PATH_TENSOR) {
        // tcgen05 score tiles decide everything the fp16 operands can decide, the band around every target score is re-scored
        // exactly (retrieval_tc.cu); `gate` is set only when the band list overflowed, and then the exact kernel below redoes it all
        int* g_dev = nullptr;
        if (int e = launch_rank_counts_tc(q, g, ng, d, qidx, target_score, target_index, npairs, g_index_offset, rank, workspace, workspace_bytes,
                                          &g_dev, st))
            return e;
        gate = g_dev;
    }
    const int64_t row_tiles = ceil_div(npairs, RT), col_tiles = ceil_div(ng, RT);
    const int nsplit = split_for(row_tiles, col_tiles);
    const int tps = (int)ceil_div(col_tiles, nsplit);
    dim3 grid((unsigned)row_tiles, (unsigned)ceil_div(col_tiles, tps));
    k_rank_counts<<<grid, RThreads, 0, st>>>(q, g, ng, d, qidx, target_score, target_index, npairs, g_index_offset, tps, rank, gate);
    DCLIP_LAUNCH_OK("k_rank_counts");
    return DCLIP_OK;
}

int dclip_rank_counts_multi(const float* q, const float* g, int64_t nq, int64_t ng, int64_t d, const int64_t* target_ptr,
                            const float* target_score, const int64_t* target_index, int64_t npairs, int max_targets,
                            int64_t g_index_offset, int32_t* rank, void* stream) {
    DCLIP_REQUIRE(q && g && target_ptr && target_score && target_index && rank, "rank_counts_multi: null pointer");
    DCLIP_REQUIRE(d > 0 && d % 4 == 0 && nq >= 0 && ng >= 0 && npairs >= 0, "rank_counts_multi: bad shape (d must be a multiple of 4)");
    if (max_targets > kMaxTargets) return fail(DCLIP_ERR_UNSUPPORTED, "rank_counts_multi: at most %d targets per query row (got %d)", kMaxTargets, max_targets);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (npairs == 0) return DCLIP_OK;
    DCLIP_CUDA_OK(cudaMemsetAsync(rank, 0, sizeof(int32_t) * npairs, st));
    if (ng == 0 || nq == 0) return DCLIP_OK;
    const int64_t row_tiles = ceil_div(nq, RT), col_tiles = ceil_div(ng, RT);
    const int nsplit = split_for(row_tiles, col_tiles);
    const int tps = (int)ceil_div(col_tiles, nsplit);
    dim3 grid((unsigned)row_tiles, (unsigned)ceil_div(col_tiles, tps));
    k_rank_counts_multi<<<grid, RThreads, 0, st>>>(q, g, nq, ng, d, target_ptr, target_score, target_index, g_index_offset, tps, rank);
    DCLIP_LAUNCH_OK("k_rank_counts_multi");
    return DCLIP_OK;
}

int dclip_matrix_ranks(const void* sim, int is_f64, int64_t n_rows, int64_t n_cols, const int64_t* line_index,
                       const int64_t* target_index, int64_t npairs, int along_rows, int32_t* rank, void* stream) {
    DCLIP_REQUIRE(sim && line_index && target_index && rank && n_rows > 0 && n_cols > 0 && npairs >= 0, "matrix_ranks: bad arguments");
    if (npairs == 0) return DCLIP_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (is_f64) k_matrix_ranks<double><<<(unsigned)npairs, 256, 0, st>>>(static_cast<const double*>(sim), n_rows, n_cols, line_index, target_index, along_rows, rank);
    else k_matrix_ranks<float><<<(unsigned)npairs, 256, 0, st>>>(static_cast<const float*>(sim), n_rows, n_cols, line_index, target_index, along_rows, rank);
    DCLIP_LAUNCH_OK("k_matrix_ranks");
    return DCLIP_OK;
}

int dclip_recall_metrics(const int32_t* rank, const int64_t* group, int64_t n, int64_t n_groups, double* out4, void* workspace,
                         size_t workspace_bytes, void* stream) {
    DCLIP_REQUIRE(rank && out4 && n > 0, "recall_metrics: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (group) {
        DCLIP_REQUIRE(n_groups > 0, "recall_metrics: n_groups must be > 0 with groups");
        if (!workspace || workspace_bytes < (size_t)n_groups * 4) return fail(DCLIP_ERR_WORKSPACE, "recall_metrics: workspace too small");
        int32_t* best = static_cast<int32_t*>(workspace);
        k_group_min_init<<<(unsigned)ceil_div(n_groups, 256), 256, 0, st>>>(best, n_groups);
        k_group_min<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(rank, group, n, n_groups, best);
        k_recall_metrics<<<1, 1024, 0, st>>>(best, n_groups, out4);
    } else {
        k_recall_metrics<<<1, 1024, 0, st>>>(rank, n, out4);
    }
    DCLIP_LAUNCH_OK("k_recall_metrics");
    return DCLIP_OK;
}

static int topk_nsplit(int64_t nq, int64_t ng) { return split_for(ceil_div(nq, RT), ceil_div(ng, RT)); }

size_t dclip_topk_ip_workspace_bytes(int64_t nq, int64_t ng, int64_t d, int64_t k, int q_dtype, int g_dtype, int path) {
    (void)d; (void)q_dtype; (void)g_dtype; (void)path;
    const int ns = topk_nsplit(nq, ng);
    return 2 * align_up((size_t)ns * nq * k * 8, 256) + 256;
}

int dclip_topk_ip(const void* q, const void* g, int64_t nq, int64_t ng, int64_t d, int64_t k, int q_dtype, int g_dtype,
                  int64_t g_index_offset, float* out_scores, int64_t* out_index, void* workspace, size_t workspace_bytes, int path,
                  void* stream) {
    DCLIP_REQUIRE(q && g && out_scores && out_index, "topk_ip: null pointer");
    DCLIP_REQUIRE(k >= 1 && k <= 32, "topk_ip: k must be in [1, 32] (got %lld)", (long long)k);
    DCLIP_REQUIRE(d > 0 && d % 4 == 0 && nq >= 0 && ng >= 0, "topk_ip: bad shape (d must be a multiple of 4)");
    if (path == DCLIP_PATH_TENSOR) return fail(DCLIP_ERR_UNSUPPORTED, "topk_ip: tensor path not available");
    if (nq == 0) return DCLIP_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t row_tiles = ceil_div(nq, RT), col_tiles = ceil_div(ng > 0 ? ng : 1, RT);
    const int nsplit0 = topk_nsplit(nq, ng > 0 ? ng : 1);
    const int tps = (int)ceil_div(col_tiles, nsplit0);
    const int nsplit = (int)ceil_div(col_tiles, tps);
    float* ps = out_scores;
    int64_t* pi = out_index;
    if (nsplit > 1) {
        Workspace ws(workspace, workspace_bytes);
        ps = static_cast<float*>(ws.take((size_t)nsplit * nq * k * 4));
        pi = static_cast<int64_t*>(ws.take((size_t)nsplit * nq * k * 8));
        if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "topk_ip: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    }
    dim3 grid((unsigned)row_tiles, (unsigned)nsplit);
#define DCLIP_TOPK_LAUNCH(TQ, TG)                                                                                     \
    k_topk_ip<TQ, TG><<<grid, RThreads, 0, st>>>(static_cast<const TQ*>(q), static_cast<const TG*>(g), nq, ng, d, (int)k, \
                                                 g_index_offset, tps, ps, pi)
    if (q_dtype == DCLIP_F32 && g_dtype == DCLIP_F32) DCLIP_TOPK_LAUNCH(float, float);
    else if (q_dtype == DCLIP_F32 && g_dtype == DCLIP_F16) DCLIP_TOPK_LAUNCH(float, __half);
    else if (q_dtype == DCLIP_F16 && g_dtype == DCLIP_F16) DCLIP_TOPK_LAUNCH(__half, __half);
    else if (q_dtype == DCLIP_F16 && g_dtype == DCLIP_F32) DCLIP_TOPK_LAUNCH(__half, float);
    else if (q_dtype == DCLIP_BF16 && g_dtype == DCLIP_BF16) DCLIP_TOPK_LAUNCH(__nv_bfloat16, __nv_bfloat16);
    else return fail(DCLIP_ERR_UNSUPPORTED, "topk_ip: dtype combination q=%d g=%d not supported", q_dtype, g_dtype);
#undef DCLIP_TOPK_LAUNCH
    DCLIP_LAUNCH_OK("k_topk_ip");
    if (nsplit > 1) {
        k_topk_merge<<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(ps, pi, nsplit, nq, (int)k, out_scores, out_index);
        DCLIP_LAUNCH_OK("k_topk_merge");
    }
    return DCLIP_OK;
}

int dclip_topk_merge(const float* in_scores, const int64_t* in_index, int64_t nlists, int64_t nq, int64_t k, float* out_scores,
                     int64_t* out_index, void* stream) {
    DCLIP_REQUIRE(in_scores && in_index && out_scores && out_index, "topk_merge: null pointer");
    DCLIP_REQUIRE(k >= 1 && k <= 32 && nlists >= 1 && nq >= 0, "topk_merge: bad arguments");
    if (nq == 0) return DCLIP_OK;
    k_topk_merge<<<(unsigned)ceil_div(nq, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(in_scores, in_index, nlists, nq, (int)k,
                                                                                           out_scores, out_index);
    DCLIP_LAUNCH_OK("k_topk_merge");
    return DCLIP_OK;
}

}  // extern "C"
