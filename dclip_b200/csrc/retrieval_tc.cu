// Retrieval — tensor path: tcgen05 candidate generation + exact FP32 re-scoring with a certificate.
//
// Replaces the brute-force search of faiss IndexFlatIP (training/compute_faiss.py:27,79;
// training/image_tokenizer.py:262) at BASELINE configs[4] scale (10 M x 768 fp16 gallery sharded over
// 8 GPUs, 100 k queries, top-10).
//
//   1. k_stream_tc<2> (stream_tc.cu): 128 queries per CTA against 256-row gallery tiles streamed by TMA;
//      fp16 operands, fp32 accumulation in TMEM.  Each epilogue thread owns one query and keeps its 32
//      best APPROXIMATE scores (threshold filter); the score matrix never exists.
//   2. candidate lists of the gallery splits are merged (32 per query).
//   3. k_topk_rescore: every candidate is re-scored with the bit-exact contract of the FP32 path
//      (fp32 fmaf chain over k ascending on the stored values), sorted by (-score, +index), top-k kept.
//   4. certificate: any row outside the 32 candidates has approximate score <= a32 (the weakest kept),
//      hence exact score <= a32 + eps, eps = eps_rel * |q| * max_g |g|.  If exact_k > a32 + eps the exact
//      top-k is provably inside the candidates; otherwise the query is flagged and the caller re-runs
//      it on the exact FP32 path.  Results are therefore always bit-exact against the oracle.
#include "common.cuh"
#include "loss_internal.h"

namespace dclip {

constexpr int kCand = 32;

template <typename T>
__global__ void k_to_f16(const T* __restrict__ in, __half* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(to_f32<T>(in[i]));
}

// max_i |x_i|_2 over the rows of x (fp32 math), atomicMax on the float bits (norms are >= 0)
template <typename T>
__global__ void __launch_bounds__(256) k_max_row_norm(const T* __restrict__ x, int64_t n, int64_t d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    float best = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * 8) {
        float ss = 0.f;
        for (int64_t k = lane; k < d; k += 32) { const float v = to_f32<T>(x[row * d + k]); ss = fmaf(v, v, ss); }
        ss = warp_sum(ss);
        best = fmaxf(best, sqrtf(ss));
    }
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(best));
}

// merge the (unsorted) candidate lists of the gallery splits into one sorted list: one warp per query
__global__ void k_cand_merge(const float* __restrict__ in_s, const int32_t* __restrict__ in_i, int nsplit, int64_t nq,
                             float* __restrict__ out_s, int32_t* __restrict__ out_i) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    float ls = -INFINITY;
    int32_t li = INT32_MAX;
    for (int p = 0; p < nsplit; ++p) {
        const int64_t base = ((int64_t)p * nq + q) * kCand;
        const float s = in_s[base + lane];
        const int32_t gi = in_i[base + lane];
        unsigned m = __ballot_sync(0xffffffffu, gi >= 0);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cs = __shfl_sync(0xffffffffu, s, src);
            const int32_t ci = __shfl_sync(0xffffffffu, gi, src);
            const unsigned after = __ballot_sync(0xffffffffu, cs > ls || (cs == ls && ci < li));
            if (!after) continue;   // (per-split lists are unsorted: every entry is tried)
            const int pos = __ffs(after) - 1;
            const float ps = __shfl_up_sync(0xffffffffu, ls, 1);
            const int32_t pi = __shfl_up_sync(0xffffffffu, li, 1);
            if (lane == pos) { ls = cs; li = ci; }
            else if (lane > pos) { ls = ps; li = pi; }
        }
    }
    out_s[q * kCand + lane] = ls;
    out_i[q * kCand + lane] = li == INT32_MAX ? -1 : li;
}

template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <typename T> __device__ __forceinline__ void lds8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void lds8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void lds8<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void lds8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// exact re-scoring + certificate: one warp per query, lane l owns candidate l.
// kStage (16-bit galleries): the warp first copies its 32 candidate rows into shared memory with coalesced 16-byte
// loads, all in flight at once, and every lane then runs its k-ascending fma chain out of shared memory (rows
// padded by 16 bytes: conflict-free uint4 reads).  Without staging each lane walks its own row through 96
// dependent, uncoalesced global loads (measured: 0.8 ms per 4096 queries, 20 ms at the 100 k queries of cfg5).
template <typename TQ, typename TG, bool kStage>
__global__ void __launch_bounds__(256) k_topk_rescore(const TQ* __restrict__ q, const TG* __restrict__ g, int64_t nq, int64_t d, int k,
                                                      const float* __restrict__ cand_s, const int32_t* __restrict__ cand_i,
                                                      const float* __restrict__ gmax, float eps_rel, int64_t g_index_offset,
                                                      float* __restrict__ out_s, int64_t* __restrict__ out_i, int32_t* __restrict__ flags,
                                                      int32_t* __restrict__ n_flagged) {
    extern __shared__ __align__(16) float qs_all[];        // per warp: [d] query (+ kStage: 32 rows of d elements + 16 B pad)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int64_t qi = (int64_t)blockIdx.x * wpb + warp;
    if (qi >= nq) return;
    const size_t row_stride = (size_t)d * sizeof(TG) + 16;
    const size_t per_warp = (size_t)d * 4 + (kStage ? kCand * row_stride : 0);
    float* qs = reinterpret_cast<float*>(reinterpret_cast<char*>(qs_all) + (size_t)warp * per_warp);
    char* rows = reinterpret_cast<char*>(qs) + (size_t)d * 4;
    float qn = 0.f;
    for (int64_t c = lane; c < d; c += 32) { const float v = to_f32<TQ>(q[qi * d + c]); qs[c] = v; qn = fmaf(v, v, qn); }
    qn = sqrtf(warp_sum(qn));
    __syncwarp();
    const int32_t ci = cand_i[qi * kCand + lane];
    const float approx = cand_s[qi * kCand + lane];
    float s = -INFINITY;
    if (kStage) {
        const int chunks = (int)(d * sizeof(TG) / 16);    // 16-byte chunks per row
#pragma unroll 4
        for (int r = 0; r < kCand; ++r) {
            const int32_t cr = __shfl_sync(0xffffffffu, ci, r);
            if (cr < 0) continue;                          // warp-uniform
            const uint4* src = reinterpret_cast<const uint4*>(g + (int64_t)cr * d);
            uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)r * row_stride);
            for (int c = lane; c < chunks; c += 32) dst[c] = __ldg(src + c);
        }
        __syncwarp();
        if (ci >= 0) {
            const TG* gr = reinterpret_cast<const TG*>(rows + (size_t)lane * row_stride);
            float acc = 0.f;
            for (int64_t c = 0; c < d; c += 8) {      // the fma chain stays k-ascending (the contract)
                float gv[8];
                lds8<TG>(gr + c, gv);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc = fmaf(qs[c + e], gv[e], acc);
            }
            s = acc;
        }
    } else if (ci >= 0) {
        const TG* gr = g + (int64_t)ci * d;
        float acc = 0.f;
        for (int64_t c = 0; c < d; c += 8) {          // 16/32-byte loads; the fma chain itself stays k-ascending (the contract)
            float gv[8];
            ld8<TG>(gr + c, gv);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(qs[c + e], gv[e], acc);
        }
        s = acc;
    }
    const int64_t gi = ci >= 0 ? g_index_offset + ci : INT64_MAX;
    // rank among the 32 by (-score, +index)
    int rank = 0;
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const float so = __shfl_sync(0xffffffffu, s, o);
        const int64_t io = __shfl_sync(0xffffffffu, gi, o);
        rank += (so > s || (so == s && io < gi)) ? 1 : 0;
    }
    if (ci < 0) rank = 32 + lane;     // empty slots never rank inside k
    if (rank < k) {
        out_s[qi * k + rank] = s;
        out_i[qi * k + rank] = gi;
    }
    const int nvalid = __popc(__ballot_sync(0xffffffffu, ci >= 0));
    if (lane >= nvalid && lane < k) {                      // fewer than k candidates: faiss-style padding
        out_s[qi * k + lane] = -INFINITY;
        out_i[qi * k + lane] = -1;
    }
    // certificate
    const bool full = nvalid == kCand;
    float a32 = ci >= 0 ? approx : INFINITY;              // weakest kept approximate score (lists may be unsorted)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a32 = fminf(a32, __shfl_xor_sync(0xffffffffu, a32, o));
    const unsigned kth = __ballot_sync(0xffffffffu, rank == k - 1);
    float exact_k = -INFINITY;
    if (kth) exact_k = __shfl_sync(0xffffffffu, s, __ffs(kth) - 1);
    const bool certified = !full || (exact_k > a32 + eps_rel * qn * gmax[0]);
    if (lane == 0) {
        flags[qi] = certified ? 0 : 1;
        if (!certified) atomicAdd(n_flagged, 1);
    }
}

// ---------------------------------------------------------------------------------------------
// Recall@K rank counting on the tensor path (flickr30k_eval.py:29-64 via dclip_rank_counts):
//   rank[p] = #{ j : s_pj > ts_p  or  (s_pj == ts_p and goff + j < tgt_p) }      with s the contract's fp32 fmaf chain
// The fp16-operand score tiles (stream mode 4) differ from s by at most eps = 1.1e-3 |q|max |g|max (two operand roundings of
// 2^-11 each + both fp32 accumulations + fp16 subnormals), so every entry outside [ts - eps, ts + eps] is decided by the tile
// and the few inside are decided here by the exact chain - the ranks are bit-identical to the FP32 path.
// ---------------------------------------------------------------------------------------------
struct RankTcMisc { unsigned int band_ctr; int overflow; float qmax, gmax; };

// out[p][:] = fp16(x[idx ? idx[p] : p][:]) and the maximum row norm (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) k_rows_to_f16_norm(const float* __restrict__ x, const int64_t* __restrict__ idx, int64_t n, int64_t d,
                                                          __half* __restrict__ out, float* __restrict__ max_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= n) return;
    const float* src = x + (idx ? idx[p] : p) * d;
    float ss = 0.f;
    for (int64_t k = lane * 4; k < d; k += 128) {
        const float4 v = *reinterpret_cast<const float4*>(src + k);
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(out + p * d + k) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
    ss = warp_sum(ss);
    if (lane == 0) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(sqrtf(ss) * 1.000001f));
}

// exact decision of the band entries.  The chain is sequential by contract, so ONE lane walks an entry - but a warp takes 8
// entries at a time: all 32 lanes stage 128-element chunks of the 16 rows in shared memory (coalesced 512-byte segments, rows
// padded by 16 B against bank conflicts), lanes 0-7 walk one entry each.  (One thread per entry waits on every load itself: 64 us
// for the 30 k entries of the Flickr-shaped evaluation; one warp per entry leaves 31 lanes idle: 37 us.)  An overflowed list means
// the data defeats the band (e.g. a gallery of identical rows): the ranks are zeroed and the gated exact kernel recomputes them.
constexpr int kBandWarps = 4, kBandGroup = 8, kBandChunk = 128, kBandLd = kBandChunk + 4;
__global__ void __launch_bounds__(kBandWarps * 32) k_rank_band(const float* __restrict__ q, const float* __restrict__ g, int64_t d,
                                                               const int64_t* __restrict__ qidx, const float* __restrict__ ts,
                                                               const int64_t* __restrict__ tgt, int64_t goff, const int2* __restrict__ band,
                                                               RankTcMisc* misc, unsigned int cap, int64_t npairs, int32_t* __restrict__ rank) {
    __shared__ __align__(16) float buf[kBandWarps][kBandGroup][2][kBandLd];
    const unsigned int n = misc->band_ctr;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (n > cap) {
        const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (e == 0) misc->overflow = 1;
        for (int64_t i = e; i < npairs; i += (int64_t)gridDim.x * blockDim.x) rank[i] = 0;
        return;
    }
    const int64_t ngroups = ((int64_t)n + kBandGroup - 1) / kBandGroup;
    for (int64_t grp = (int64_t)blockIdx.x * kBandWarps + w; grp < ngroups; grp += (int64_t)gridDim.x * kBandWarps) {
        const int64_t e = grp * kBandGroup + lane;                 // lanes 0-7 own one entry each
        const bool mine = lane < kBandGroup && e < n;
        int2 b = make_int2(0, 0);
        const float* qr = q;                                       // (a lane without an entry stages row 0: never read back)
        const float* gr = g;
        if (mine) {
            b = band[e];
            qr = q + (qidx ? qidx[b.x] : (int64_t)b.x) * d;
            gr = g + (int64_t)b.y * d;
        }
        float acc = 0.f;
        for (int64_t c0 = 0; c0 < d; c0 += kBandChunk) {
            const int cw = (int)min((int64_t)kBandChunk, d - c0);                // multiple of 4
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 2 * kBandGroup; ++j) {                           // row j of the group: entry j >> 1, operand j & 1
                const float* src = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>((j & 1) ? gr : qr), j >> 1));
                if (lane * 4 < cw) *reinterpret_cast<float4*>(&buf[w][j >> 1][j & 1][lane * 4]) = __ldg(reinterpret_cast<const float4*>(src + c0 + lane * 4));
            }
            __syncwarp();
            if (mine) {
#pragma unroll 4
                for (int k = 0; k < cw; k += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(&buf[w][lane][0][k]), y = *reinterpret_cast<const float4*>(&buf[w][lane][1][k]);
                    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
                }
            }
        }
        if (mine) {
            const float t = ts[b.x];
            if (acc > t || (acc == t && goff + b.y < tgt[b.x])) atomicAdd(&rank[b.x], 1);
        }
    }
}

static unsigned int rank_band_cap(int64_t npairs) { return (unsigned int)min((int64_t)1 << 26, 8 * npairs + 8192); }

size_t rank_counts_tc_workspace_bytes(int64_t npairs, int64_t ng, int64_t d) {
    return align_up((size_t)npairs * d * 2, 256) + align_up((size_t)ng * d * 2, 256) + align_up((size_t)rank_band_cap(npairs) * sizeof(int2), 256) +
           256 + 1024;
}

// rank must be zeroed by the caller (stream-ordered).  *gate_out = device int that is 1 iff the exact fallback has to run.
int launch_rank_counts_tc(const float* q, const float* g, int64_t ng, int64_t d, const int64_t* qidx, const float* ts, const int64_t* tgt,
                          int64_t npairs, int64_t goff, int32_t* rank, void* workspace, size_t workspace_bytes, int** gate_out, cudaStream_t st) {
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "tensor path needs an sm_100 device");
    if (d % 8 != 0 || d < 64 || d > 1024) return fail(DCLIP_ERR_UNSUPPORTED, "rank_counts: tensor path needs d %% 8 == 0 and 64 <= d <= 1024 (got %lld)", (long long)d);
    if (npairs >= (1ll << 31) || ng >= (1ll << 31)) return fail(DCLIP_ERR_UNSUPPORTED, "rank_counts: tensor path needs fewer than 2^31 pairs / gallery rows");
    DCLIP_REQUIRE(aligned16(q) && aligned16(g) && aligned16(workspace), "rank_counts: pointers must be 16-byte aligned");
    const unsigned int cap = rank_band_cap(npairs);
    Workspace ws(workspace, workspace_bytes);
    __half* q16 = static_cast<__half*>(ws.take((size_t)npairs * d * 2));
    __half* g16 = static_cast<__half*>(ws.take((size_t)ng * d * 2));
    int2* band = static_cast<int2*>(ws.take((size_t)cap * sizeof(int2)));
    RankTcMisc* misc = static_cast<RankTcMisc*>(ws.take(256));
    if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "rank_counts: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    DCLIP_CUDA_OK(cudaMemsetAsync(misc, 0, sizeof(RankTcMisc), st));
    k_rows_to_f16_norm<<<(unsigned)ceil_div(npairs, 8), 256, 0, st>>>(q, qidx, npairs, d, q16, &misc->qmax);
    k_rows_to_f16_norm<<<(unsigned)ceil_div(ng, 8), 256, 0, st>>>(g, nullptr, ng, d, g16, &misc->gmax);
    DCLIP_LAUNCH_OK("k_rows_to_f16_norm");
    TcArgs a{};
    a.m = npairs; a.n = ng; a.d = d; a.inv_tau = 1.f;
    a.rk_thr = ts; a.rk_tgt = tgt; a.rk_rank = rank; a.rk_band = band; a.rk_ctr = &misc->band_ctr; a.rk_cap = cap; a.rk_goff = goff;
    a.rk_eps = 1.1e-3f; a.rk_norms = &misc->qmax;
    const void* xs[2] = {q16, q16};
    const void* ys[2] = {g16, g16};
    if (int e = launch_stream_tc(4, 1, xs, ys, a, stream_nsplit(npairs, ng, 1), st)) return e;
    k_rank_band<<<(unsigned)(8 * sm_count()), kBandWarps * 32, 0, st>>>(q, g, d, qidx, ts, tgt, goff, band, misc, cap, npairs, rank);
    DCLIP_LAUNCH_OK("k_rank_band");
    *gate_out = &misc->overflow;
    return DCLIP_OK;
}

int tc_topk_nsplit(int64_t nq, int64_t ng) { return stream_nsplit(nq, ng, 1); }

}  // namespace dclip

using namespace dclip;

extern "C" {

int dclip_max_row_norm(const void* x, int64_t n, int64_t d, int dtype, float* out, void* stream) {
    DCLIP_REQUIRE(x && out && n >= 0 && d > 0, "max_row_norm: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), st));
    if (n == 0) return DCLIP_OK;
    const unsigned grid = (unsigned)min((int64_t)sm_count() * 8, ceil_div(n, 8));
    DCLIP_DISPATCH_DTYPE(dtype, (k_max_row_norm<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x), n, d, out)));
    DCLIP_LAUNCH_OK("k_max_row_norm");
    return DCLIP_OK;
}

size_t dclip_topk_ip_tc_workspace_bytes(int64_t nq, int64_t ng, int64_t d, int q_dtype) {
    const int ns = tc_topk_nsplit(nq, ng);
    size_t total = 0;
    total += align_up((size_t)nq * d * 2, 256);                       // fp16 queries
    total += 2 * align_up((size_t)ns * nq * kCand * 4, 256);          // per-split candidates
    total += 2 * align_up((size_t)nq * kCand * 4, 256);               // merged candidates
    total += align_up(stream_sync_bytes(nq, ng, d, ns), 256);         // lockstep counters of the streaming kernel
    (void)q_dtype;
    return total + 1024;
}

/* Tensor-path exact top-k.  q: [nq, d] (any dtype), g_exact: [ng, d] in g_dtype = the stored gallery the
 * scores are defined on, g_fp16: the same gallery in fp16 for the tensor pass (== g_exact when g_dtype is
 * DCLIP_F16).  gmax: device float = max gallery row norm (dclip_max_row_norm).  flags: int32[nq] (1 = the
 * certificate failed: re-run that query with dclip_topk_ip on the FP32 path), n_flagged: device int32. */
int dclip_topk_ip_tc(const void* q, int q_dtype, const void* g_exact, int g_dtype, const void* g_fp16, int64_t nq, int64_t ng,
                     int64_t d, int64_t k, int64_t g_index_offset, const float* gmax, float* out_scores, int64_t* out_index,
                     int32_t* flags, int32_t* n_flagged, void* workspace, size_t workspace_bytes, void* stream) {
    DCLIP_REQUIRE(q && g_exact && g_fp16 && gmax && out_scores && out_index && flags && n_flagged, "topk_ip_tc: null pointer");
    DCLIP_REQUIRE(k >= 1 && k <= 16, "topk_ip_tc: k must be in [1, 16] (32 candidates are re-scored)");
    DCLIP_REQUIRE(d % 8 == 0 && d >= 64 && d <= 1024, "topk_ip_tc: tensor path needs d %% 8 == 0 and 64 <= d <= 1024 (got %lld)", (long long)d);
    DCLIP_REQUIRE(ng > 0 && ng < (1ll << 26), "topk_ip_tc: gallery shard must have 1 .. 2^26-1 rows (got %lld)", (long long)ng);
    DCLIP_REQUIRE(aligned16(g_fp16) && aligned16(workspace), "topk_ip_tc: pointers must be 16-byte aligned");
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "tensor path needs an sm_100 device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nq == 0) return DCLIP_OK;
    const int ns = tc_topk_nsplit(nq, ng);
    Workspace ws(workspace, workspace_bytes);
    __half* q16 = static_cast<__half*>(ws.take((size_t)nq * d * 2));
    float* ps = static_cast<float*>(ws.take((size_t)ns * nq * kCand * 4));
    int32_t* pi = static_cast<int32_t*>(ws.take((size_t)ns * nq * kCand * 4));
    float* ms = static_cast<float*>(ws.take((size_t)nq * kCand * 4));
    int32_t* mi = static_cast<int32_t*>(ws.take((size_t)nq * kCand * 4));
    const size_t sync_bytes = stream_sync_bytes(nq, ng, d, ns);
    unsigned int* sync_ctr = sync_bytes ? static_cast<unsigned int*>(ws.take(sync_bytes)) : nullptr;
    if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "topk_ip_tc: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    const void* qh = q;
    if (q_dtype != DCLIP_F16) {
        DCLIP_DISPATCH_DTYPE(q_dtype, (k_to_f16<T><<<(unsigned)ceil_div(nq * d, 256), 256, 0, st>>>(static_cast<const T*>(q), q16, nq * d)));
        DCLIP_LAUNCH_OK("k_to_f16");
        qh = q16;
    }
    TcArgs a{};
    a.m = nq; a.n = ng; a.d = d; a.inv_tau = 1.f; a.row_offset = 0;
    a.cand_s = ps; a.cand_i = pi;
    a.sync_ctr = sync_ctr;
    const void* xs[2] = {qh, qh};
    const void* ys[2] = {g_fp16, g_fp16};
    if (int e = launch_stream_tc(2, 1, xs, ys, a, ns, st)) return e;
    const float* cs = ps;
    const int32_t* ci = pi;
    if (ns > 1) {
        k_cand_merge<<<(unsigned)ceil_div(nq, 8), 256, 0, st>>>(ps, pi, ns, nq, ms, mi);
        DCLIP_LAUNCH_OK("k_cand_merge");
        cs = ms; ci = mi;
    }
    DCLIP_CUDA_OK(cudaMemsetAsync(n_flagged, 0, sizeof(int32_t), st));
    // eps_rel: accumulation-order difference of exact fp16 products (d * 2^-22, worst case) plus the fp16
    // rounding of operands that were not fp16 to begin with (2^-11 each)
    float eps_rel = (float)d * 2.384185791015625e-07f;
    if (q_dtype != DCLIP_F16) eps_rel += 4.8828125e-4f;
    if (g_dtype != DCLIP_F16) eps_rel += 4.8828125e-4f;
    // 16-bit galleries: candidate rows staged in shared memory, as many warps (= queries) per block as fit
    const bool stage = g_dtype != DCLIP_F32 && d % 8 == 0;
    const size_t per_warp = (size_t)d * 4 + (stage ? (size_t)kCand * ((size_t)d * 2 + 16) : 0);
    int wpb = stage ? (int)((size_t)220 * 1024 / per_warp) : 8;
    if (wpb > 8) wpb = 8;
    if (wpb < 1) return fail(DCLIP_ERR_UNSUPPORTED, "topk_ip_tc: d too large for the re-scoring kernel");
    const size_t smem = (size_t)wpb * per_warp;
    const unsigned grid = (unsigned)ceil_div(nq, wpb);
#define DCLIP_RESCORE(TQ, TG, ST)                                                                                                \
    do {                                                                                                                         \
        DCLIP_CUDA_OK(cudaFuncSetAttribute((k_topk_rescore<TQ, TG, ST>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_topk_rescore<TQ, TG, ST><<<grid, wpb * 32, smem, st>>>(static_cast<const TQ*>(q), static_cast<const TG*>(g_exact), nq, d, (int)k, cs, \
                                                                ci, gmax, eps_rel, g_index_offset, out_scores, out_index, flags, n_flagged); \
    } while (0)
    if (q_dtype == DCLIP_F32 && g_dtype == DCLIP_F32) DCLIP_RESCORE(float, float, false);
    else if (q_dtype == DCLIP_F32 && g_dtype == DCLIP_F16) DCLIP_RESCORE(float, __half, true);
    else if (q_dtype == DCLIP_F16 && g_dtype == DCLIP_F16) DCLIP_RESCORE(__half, __half, true);
    else if (q_dtype == DCLIP_F16 && g_dtype == DCLIP_F32) DCLIP_RESCORE(__half, float, false);
    else if (q_dtype == DCLIP_BF16 && g_dtype == DCLIP_BF16) DCLIP_RESCORE(__nv_bfloat16, __nv_bfloat16, true);
    else return fail(DCLIP_ERR_UNSUPPORTED, "topk_ip_tc: dtype combination q=%d g=%d not supported", q_dtype, g_dtype);
#undef DCLIP_RESCORE
    DCLIP_LAUNCH_OK("k_topk_rescore");
    return DCLIP_OK;
}

}  // extern "C"
