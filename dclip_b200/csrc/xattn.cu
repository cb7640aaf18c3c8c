// Meta-teacher bidirectional cross-attention + aggregation — forward and backward.
//
// Replaces (paths relative to the reference tree):
//   CrossModalAttention.forward              training/patch_text_aggregation.py:21-46
//     = 2 x nn.MultiheadAttention (packed in-proj, no masks, need_weights path) + residual + LayerNorm
//   PatchTextAggregation.aggregation         training/patch_text_aggregation.py:243-265
//   combine + NaN/Inf guard                  training/patch_text_aggregation.py:640-656
//   and the autograd backward the teacher step runs through them
//   (training/train_contrastive_teacher.py:347-362).
//
// Two tiers.  fp32 inputs: everything below runs in fp32 on CUDA cores (exact tier).  16-bit inputs
// (bf16 / fp16 parameters and activations, BASELINE configs[2]): the ten projection GEMMs of the forward
// and backward (97 % of the FLOPs) run on tcgen05 through gemm_tc.cu with the 16-bit tensors as operands
// and fp32 accumulation; attention cores, LayerNorm, aggregation and all statistics stay fp32.
// This file is the fp32 CUDA-core implementation plus the orchestration of both tiers.  All intermediates are fp32 in a
// caller-provided `saved` buffer that the backward re-reads (Q, K|V, attention output, pre-LN sums,
// LN statistics); attention probabilities are never stored — the backward recomputes them per
// (batch, head) in shared memory.  16-bit inputs/parameters are widened once on entry.
//
// Row layout: text rows = b*T + t, image rows = b*R + r (batch-first, as the callers hold them; the
// reference's transpose to [L, B, d] only permutes rows of the same per-(b, head) problems).
#include <cooperative_groups.h>

#include "common.cuh"
#include "gemm_tc.h"

#include <cstdlib>

namespace cg = cooperative_groups;

namespace dclip {

// ---------------------------------------------------------------------------------------------
// grouped fp32 GEMM:  C[m,n] (+)= alpha * sum_k A'(m,k) B'(k,n) + bias[n] + residual[m,n]
// ---------------------------------------------------------------------------------------------
struct GemmProblem {
    const float* A; const float* B; float* C; const float* bias; const float* residual;
    int M, N, K;
    int lda, ldb, ldc, ldr;
    int transA, transB;      // A'(m,k) = transA ? A[k*lda+m] : A[m*lda+k];  B'(k,n) = transB ? B[n*ldb+k] : B[k*ldb+n]
    float alpha;
    int accumulate;
};
struct GemmGroup { GemmProblem p[8]; int count; };

constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256) k_gemm_group(const __grid_constant__ GemmGroup g) {
    const GemmProblem& P = g.p[blockIdx.z];
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    if (m0 >= P.M || n0 >= P.N) return;
    __shared__ __align__(16) float As[GK][GT + 4];
    __shared__ __align__(16) float Bs[GK][GT + 4];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < P.K; k0 += GK) {
        __syncthreads();
        // 64x16 elements per operand, 4 per thread; index so that the contiguous dim is the fast one
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + e * 256;
            int m, k;
            if (P.transA) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            float v = 0.f;
            if (m0 + m < P.M && k0 + k < P.K)
                v = P.transA ? P.A[(size_t)(k0 + k) * P.lda + m0 + m] : P.A[(size_t)(m0 + m) * P.lda + k0 + k];
            As[k][m] = v;
            int n, kb;
            if (P.transB) { kb = idx & 15; n = idx >> 4; } else { n = idx & 63; kb = idx >> 6; }
            float w = 0.f;
            if (n0 + n < P.N && k0 + kb < P.K)
                w = P.transB ? P.B[(size_t)(n0 + n) * P.ldb + k0 + kb] : P.B[(size_t)(k0 + kb) * P.ldb + n0 + n];
            Bs[kb][n] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= P.N) continue;
            float v = P.alpha * acc[i][j];
            if (P.bias) v += P.bias[n];
            if (P.residual) v += P.residual[(size_t)m * P.ldr + n];
            float* c = P.C + (size_t)m * P.ldc + n;
            *c = P.accumulate ? *c + v : v;
        }
    }
}

static int launch_gemm_group(const GemmGroup& g, cudaStream_t st) {
    int maxM = 0, maxN = 0;
    for (int i = 0; i < g.count; ++i) { maxM = max(maxM, g.p[i].M); maxN = max(maxN, g.p[i].N); }
    if (maxM == 0 || maxN == 0) return DCLIP_OK;
    dim3 grid((maxN + GT - 1) / GT, (maxM + GT - 1) / GT, g.count);
    k_gemm_group<<<grid, 256, 0, st>>>(g);
    DCLIP_LAUNCH_OK("k_gemm_group");
    return DCLIP_OK;
}

static GemmProblem gemm(const float* A, int lda, bool tA, const float* B, int ldb, bool tB, float* C, int ldc, int M, int N, int K,
                        const float* bias = nullptr, const float* residual = nullptr, int ldr = 0, bool accumulate = false) {
    GemmProblem p{};
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.residual = residual;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = ldr;
    p.transA = tA; p.transB = tB; p.alpha = 1.f; p.accumulate = accumulate;
    return p;
}

// ---------------------------------------------------------------------------------------------
// The same grouped fp32 GEMM on tcgen05: bf16 x 3 operand splitting.
// An fp32 value is x = h + m + l with three bf16 terms (8 + 8 + 8 mantissa bits: exact to 2^-27 |x|, fp32's full exponent
// range - gradients of 1e-8 keep their precision, which an fp16 split would lose).  The product of two split operands
// needs six of the nine term products (h l, l h, m m, h m, m h, h h; the rest is below 2^-26), and a sum of products over k
// of six term pairs is ONE contraction over 6 K.  Every operand is stored once as [h | m | l] along K (segments of Kp = K
// rounded up to the 64-wide k-block) and the grouped tcgen05 kernel (gemm_tc.cu, "segmented contraction") walks the six
// logical segments
//     A: h l m h m h      B: l h m m h h      (smallest products first: they are accumulated before the large h h term
//                                              can swamp them)
// by mapping each to the stored term it reads - 16-bit MMAs, fp32 accumulation in TMEM, nothing else changes.
// Measured against an fp64 product (tools/exp_split_gemm.py, N(0,1) operands): 4.7e-7 at K = 512, 7.3e-7 at K = 768, 2.6e-6 at
// K = 2464, torch's fp32 matmul (TF32 off) 4.1e-7 / 5.0e-7 / 8.9e-7 - inside the FP32 tier's 1e-5, at 6x the MMA work of a
// 16-bit GEMM instead of CUDA-core FFMA (configs[0] teacher step: 1287 us of k_gemm_group per step).
// This is what stands in for the TF32 tier north_star mentions: kind::tf32 keeps 10 mantissa bits (1e-3); the reference trains
// its teacher in fp32 with TF32 off (torch default), so the drop-in keeps fp32 accuracy.
// ---------------------------------------------------------------------------------------------
constexpr int kSplitSegs = 6;          // logical segments of the contraction (term products)
constexpr int kSplitTerms = 3;         // stored terms per operand: h | m | l
constexpr int kSplitK = 6;           // slices of a long contraction (weight gradients)
constexpr int kSplitKMinBlocks = 96; // ... when it has at least this many 64-wide k-blocks
struct SplitJob {
    const float* src; __nv_bfloat16* dst;
    int rows, cols, ld;        // source matrix [rows, cols], row stride ld
    int out_rows, out_cols;    // iteration space: K-major operand rows x Kp;  MN-major operand Kp x ldd
    int ldd, Kp, kmajor;
};
struct SplitGroup { SplitJob j[16]; int count; };

__global__ void __launch_bounds__(256) k_split_bf16x3(const __grid_constant__ SplitGroup g) {
    const SplitJob& J = g.j[blockIdx.y];
    const int ncg = J.out_cols >> 3;
    const long long total = (long long)J.out_rows * ncg;
    const bool vec = (J.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(J.src) & 15) == 0;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int r = (int)(e / ncg), c = (int)(e - (long long)r * ncg) * 8;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < J.rows) {
            const float* sp = J.src + (size_t)r * J.ld + c;
            if (vec && c + 8 <= J.cols) {
                const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (c + i < J.cols) x[i] = sp[i];
            }
        }
        uint32_t t[3][4];      // terms h, m, l: 8 bf16 each
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float r0 = x[i], r1 = x[i + 1];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const __nv_bfloat16 b0 = __float2bfloat16_rn(r0), b1 = __float2bfloat16_rn(r1);
                r0 -= __bfloat162float(b0); r1 -= __bfloat162float(b1);       // exact: the remainder of a rounding fits fp32
                t[q][i >> 1] = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
            }
        }
#pragma unroll
        for (int q = 0; q < kSplitTerms; ++q) {
            const size_t off = J.kmajor ? (size_t)r * J.ldd + (size_t)q * J.Kp + c : ((size_t)q * J.Kp + r) * J.ldd + c;
            *reinterpret_cast<uint4*>(J.dst + off) = make_uint4(t[q][0], t[q][1], t[q][2], t[q][3]);
        }
    }
}

// split-K partial products summed in slice order; bias / residual / the accumulate input of the problem are applied here.
// Two kinds of problem are cut: weight gradients (d x d outputs = a handful of tiles under a contraction over every row of the
// batch) and launches whose tiles would fill less than half of the SMs (out-projections at d = 512: 44 tiles of 48 k-blocks).
struct SumJob { const float* part; float* out; const float* bias; const float* residual; int M, N, ldc, ldr, count, accumulate; long long stride; };
struct SumGroup { SumJob j[8]; int count; };
__global__ void __launch_bounds__(256) k_sum_partials(const __grid_constant__ SumGroup g) {
    const SumJob& J = g.j[blockIdx.y];
    const int n4 = J.N >> 2;
    const long long total = (long long)J.M * n4;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int m = (int)(e / n4), n = (int)(e - (long long)m * n4) * 4;
        const float* p = J.part + (size_t)m * J.N + n;
        float4 s = *reinterpret_cast<const float4*>(p);
        for (int q = 1; q < J.count; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(p + (size_t)q * J.stride);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        if (J.bias) { const float4 v = *reinterpret_cast<const float4*>(J.bias + n); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        if (J.residual) { const float4 v = *reinterpret_cast<const float4*>(J.residual + (size_t)m * J.ldr + n); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        float* o = J.out + (size_t)m * J.ldc + n;
        if (J.accumulate) { const float4 v = *reinterpret_cast<const float4*>(o); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        *reinterpret_cast<float4*>(o) = s;
    }
}

struct SplitScratch { char* base; size_t bytes; };

// fp32 tier on tensor cores?  (sm_100 device; DCLIP_XATTN_SPLIT=0 keeps the CUDA-core GEMM for A/B timing and tests)
static bool gemm_split_enabled() {
    if (!device_is_sm100()) return false;
    const char* e = getenv("DCLIP_XATTN_SPLIT");
    return !(e && e[0] == '0');
}
// scratch of the split path: the six-term operands of the largest launch + split-K partials.  Partials exist for weight gradients
// (d x d outputs, always cut) and for launches of fewer than sm_count() / 2 tiles (<= 74 tiles of 128 x 256 outputs, cut up to 6 ways).
static size_t split_partial_cap(size_t out_elems) { return (out_elems < (size_t)74 * 128 * 256 ? out_elems : (size_t)74 * 128 * 256) * kSplitK * sizeof(float); }
static size_t split_fwd_bytes(int64_t B, int64_t T, int64_t R, int64_t d) {
    const size_t dp = (size_t)(d + 63) / 64 * 64, rows = (size_t)B * (T + R);
    return (rows + 6 * d + 16) * kSplitTerms * dp * 2 + 16 * 256 + split_partial_cap((rows + 256) * 3 * dp) + 8 * 256;
}
static size_t split_bwd_bytes(int64_t B, int64_t T, int64_t R, int64_t d) {
    // the largest launch: weight gradients (contraction over the rows, padded to 64 per side) with their 4 + 2 operand matrices
    const size_t dp = (size_t)(d + 63) / 64 * 64, rows = (size_t)B * (T + R) + 128;
    return ((rows + 32) * 12 * dp + 12 * dp * dp) * 2 + 16 * 256 + (size_t)kSplitK * 6 * dp * dp * sizeof(float) + split_partial_cap((rows + 256) * 2 * dp) + 8 * 256;
}

static int launch_gemm_group_split(const GemmGroup& g, const SplitScratch& sc, cudaStream_t st) {
    SplitGroup sg{};
    SumGroup sums{};
    TcGemmProblem pr[8] = {};
    size_t used = 0;
    auto operand = [&](const float* src, bool kmajor, int mn, int K, int ld, int Kp, const __nv_bfloat16** out, int* ldd) -> int {
        const int rows = kmajor ? mn : K, cols = kmajor ? K : mn;
        for (int i = 0; i < sg.count; ++i) {        // a matrix that feeds several problems (either operand side) is split once
            const SplitJob& J = sg.j[i];
            if (J.src == src && J.kmajor == (int)kmajor && J.rows == rows && J.cols == cols && J.ld == ld) { *out = J.dst; *ldd = J.ldd; return DCLIP_OK; }
        }
        if (sg.count == 16) return fail(DCLIP_ERR_INVALID, "gemm split: too many distinct operands");
        SplitJob J{};
        J.src = src; J.rows = rows; J.cols = cols; J.ld = ld; J.Kp = Kp; J.kmajor = kmajor;
        if (kmajor) { J.ldd = kSplitTerms * Kp; J.out_rows = mn; J.out_cols = Kp; }
        else { J.ldd = (mn + 7) / 8 * 8; J.out_rows = Kp; J.out_cols = J.ldd; }
        const size_t bytes = (kmajor ? (size_t)mn * J.ldd : (size_t)kSplitTerms * Kp * J.ldd) * 2;
        if (used + bytes > sc.bytes) return fail(DCLIP_ERR_WORKSPACE, "gemm split: scratch too small (%zu + %zu > %zu)", used, bytes, sc.bytes);
        J.dst = reinterpret_cast<__nv_bfloat16*>(sc.base + used);
        used += align_up(bytes, 256);
        sg.j[sg.count++] = J;
        *out = J.dst; *ldd = J.ldd;
        return DCLIP_OK;
    };
    // logical segment -> stored term (0 = h, 1 = m, 2 = l):  A: h l m h m h,  B: l h m m h h
    constexpr unsigned kMapA = 0u | (2u << 4) | (1u << 8) | (0u << 12) | (1u << 16) | (0u << 20);
    constexpr unsigned kMapB = 2u | (0u << 4) | (1u << 8) | (1u << 12) | (0u << 16) | (0u << 20);
    for (int i = 0; i < g.count; ++i) {
        const GemmProblem& P = g.p[i];
        const int Kp = (P.K + 63) / 64 * 64;          // segments are whole k-blocks
        const __nv_bfloat16* A2 = nullptr; const __nv_bfloat16* B2 = nullptr;
        int lda2 = 0, ldb2 = 0;
        if (int e = operand(P.A, !P.transA, P.M, P.K, P.lda, Kp, &A2, &lda2)) return e;
        if (int e = operand(P.B, P.transB != 0, P.N, P.K, P.ldb, Kp, &B2, &ldb2)) return e;
        TcGemmProblem q{};
        q.A = A2; q.B = B2; q.C = P.C; q.bias = P.bias; q.residual = P.residual; q.C16 = nullptr;
        q.M = P.M; q.N = P.N; q.K = kSplitSegs * Kp; q.lda = lda2; q.ldb = ldb2; q.ldc = P.ldc; q.ldr = P.ldr;
        q.a_mn = P.transA ? 1 : 0; q.b_mn = P.transB ? 0 : 1; q.accumulate = P.accumulate; q.alpha = P.alpha;
        q.bias_kind = 0; q.res_kind = 0;
        q.seg_len = Kp; q.k_phys = kSplitTerms * Kp; q.seg_map_a = kMapA; q.seg_map_b = kMapB;
        pr[i] = q;
    }
    // split-K (see k_sum_partials): long contractions always, everything when the launch would leave most SMs idle
    // long contractions: as many slices (<= kSplitK) as still leave the launch within one wave of CTAs
    auto tiles_of = [&](int i) { return ((g.p[i].M + 127) / 128) * ((g.p[i].N + 255) / 256); };
    auto is_long = [&](int i) { return (pr[i].K + 63) / 64 >= kSplitKMinBlocks; };
    int long_slices = kSplitK;
    for (int ks = kSplitK; ks >= 2; --ks) {
        int tot = 0;
        for (int i = 0; i < g.count; ++i) tot += tiles_of(i) * (is_long(i) ? gemm_tc_slices(pr[i].K, ks) : 1);
        if (tot <= sm_count()) { long_slices = ks; break; }
    }
    int total_tiles = 0;      // with the long contractions already cut
    for (int i = 0; i < g.count; ++i) total_tiles += tiles_of(i) * (is_long(i) ? gemm_tc_slices(pr[i].K, long_slices) : 1);
    const int fill = total_tiles > 0 ? sm_count() / total_tiles : 1;
    for (int i = 0; i < g.count; ++i) {
        const GemmProblem& P = g.p[i];
        TcGemmProblem& q = pr[i];
        const int nkb = (q.K + 63) / 64;
        int want = nkb >= kSplitKMinBlocks ? long_slices : 1;
        if (want == 1 && fill >= 2) want = min(min(fill, kSplitK), nkb / 8);
        if (want < 2 || sums.count == 8) continue;
        const int ks = gemm_tc_slices(q.K, want);
        const size_t bytes = (size_t)ks * P.M * P.N * sizeof(float);
        if (ks < 2 || used + bytes > sc.bytes) continue;
        float* part = reinterpret_cast<float*>(sc.base + used);
        used += align_up(bytes, 256);
        sums.j[sums.count++] = SumJob{part, P.C, P.bias, P.residual, P.M, P.N, P.ldc, P.ldr, ks, P.accumulate, (long long)P.M * P.N};
        q.C = part; q.ldc = P.N; q.bias = nullptr; q.residual = nullptr; q.accumulate = 0; q.ksplit = want; q.split_stride = (long long)P.M * P.N;
    }

    long long mx = 0;
    for (int i = 0; i < sg.count; ++i) mx = max(mx, (long long)sg.j[i].out_rows * (sg.j[i].out_cols >> 3));
    if (mx > 0) {
        k_split_bf16x3<<<dim3((unsigned)min((long long)1184, (mx + 255) / 256), (unsigned)sg.count), 256, 0, st>>>(sg);
        DCLIP_LAUNCH_OK("k_split_bf16x3");
    }
    if (int e = launch_gemm_tc(pr, g.count, DCLIP_BF16, st)) return e;
    if (sums.count) {
        long long mxn = 0;
        for (int i = 0; i < sums.count; ++i) mxn = max(mxn, (long long)sums.j[i].M * sums.j[i].N);
        k_sum_partials<<<dim3((unsigned)min((long long)592, (mxn / 4 + 255) / 256), (unsigned)sums.count), 256, 0, st>>>(sums);
        DCLIP_LAUNCH_OK("k_sum_partials");
    }
    return DCLIP_OK;
}

// the fp32 grouped GEMM: tcgen05 through operand splitting where a scratch area was provided, CUDA cores otherwise
static int run_gemm_group(const GemmGroup& g, const SplitScratch& sc, cudaStream_t st) {
    if (sc.base) {
        bool ok = true;
        for (int i = 0; i < g.count; ++i) {
            const GemmProblem& P = g.p[i];
            ok = ok && P.N % 8 == 0 && P.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(P.C) & 15) == 0 && (!P.residual || (P.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(P.residual) & 15) == 0)) &&
                 (!P.bias || (reinterpret_cast<uintptr_t>(P.bias) & 15) == 0);
        }
        if (ok) return launch_gemm_group_split(g, sc, st);
    }
    return launch_gemm_group(g, st);
}

// ---------------------------------------------------------------------------------------------
// casts and column sums
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_to_f32(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = to_f32<T>(in[i]);
}
template <typename T>
__global__ void k_from_f32(const float* __restrict__ in, T* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = from_f32<T>(in[i]);
}

// grouped casts: up to 12 tensors per launch (blockIdx.y = tensor)
struct CastGroup { const void* src[12]; void* dst[12]; long long n[12]; int count; };
template <typename T>
__global__ void k_widen_group(const __grid_constant__ CastGroup g) {
    const T* in = static_cast<const T*>(g.src[blockIdx.y]);
    float* out = static_cast<float*>(g.dst[blockIdx.y]);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.n[blockIdx.y]; i += (long long)gridDim.x * blockDim.x)
        out[i] = to_f32<T>(in[i]);
}
template <typename T>
__global__ void k_narrow_group(const __grid_constant__ CastGroup g) {
    const float* in = static_cast<const float*>(g.src[blockIdx.y]);
    T* out = static_cast<T*>(g.dst[blockIdx.y]);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.n[blockIdx.y]; i += (long long)gridDim.x * blockDim.x)
        out[i] = from_f32<T>(in[i]);
}

// out[c] = sum_rows M[row, c]   (and optionally out2[c] = sum_rows M[row,c] * W[row,c]); fixed order.
// Up to 8 problems per launch (blockIdx.y); a block owns 32 columns, 32 row lanes.
struct ColsumProblem { const float* Mx; const float* W; int rows, cols, ld, ldw; float* out; float* out2; };
struct ColsumGroup { ColsumProblem p[8]; int count; };
// The rows are split over the 8 CTAs of a cluster (blockIdx.z): a thread walks ~rows / 256 of them instead of rows / 32 (the loop
// is a chain of L2 round trips), the per-CTA sums meet in CTA 0 through distributed shared memory, rank order: deterministic.
constexpr int kColsumCluster = 8;
__global__ void __cluster_dims__(1, 1, kColsumCluster) __launch_bounds__(1024) k_colsum(const __grid_constant__ ColsumGroup g) {
    const ColsumProblem& P = g.p[blockIdx.y];
    if ((int)blockIdx.x * 32 >= P.cols) return;        // the whole cluster (same x, y) leaves together
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float r1[32][33], r2[32][33];
    __shared__ float fin[2][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int rank = (int)blockIdx.z;
    const int per = (P.rows + kColsumCluster - 1) / kColsumCluster;
    const int r_begin = rank * per, r_end = min(P.rows, r_begin + per);
    float s1 = 0.f, s2 = 0.f;
    if (c < P.cols) {
#pragma unroll 4
        for (int r = r_begin + grp; r < r_end; r += 32) {
            const float v = P.Mx[(size_t)r * P.ld + c];
            s1 += v;
            if (P.W) s2 = fmaf(v, P.W[(size_t)r * P.ldw + c], s2);
        }
    }
    r1[grp][lane] = s1; r2[grp][lane] = s2;
    __syncthreads();
    if (grp == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 32; ++i) { a += r1[i][lane]; b += r2[i][lane]; }
        fin[0][lane] = a; fin[1][lane] = b;
    }
    cluster.sync();
    if (rank == 0 && grp == 0 && c < P.cols) {
        float a = 0.f, b = 0.f;
        for (int q = 0; q < kColsumCluster; ++q) {
            const float* f = cluster.map_shared_rank(&fin[0][0], q);
            a += f[lane]; b += f[32 + lane];
        }
        if (P.out) P.out[c] = a;
        if (P.out2) P.out2[c] = b;
    }
    cluster.sync();        // nobody leaves while CTA 0 still reads its shared memory
}
static int launch_colsum(const ColsumGroup& g, cudaStream_t st) {
    int maxc = 0;
    for (int i = 0; i < g.count; ++i) maxc = max(maxc, g.p[i].cols);
    if (maxc == 0) return DCLIP_OK;
    k_colsum<<<dim3((maxc + 31) / 32, g.count, kColsumCluster), 1024, 0, st>>>(g);
    DCLIP_LAUNCH_OK("k_colsum");
    return DCLIP_OK;
}

// ---------------------------------------------------------------------------------------------
// attention core, forward: one CTA per (batch, head, direction)
//   O[q, :] = softmax_j(scale * Q[q] . K[j]) V[j]
// ---------------------------------------------------------------------------------------------
struct AttnDir {
    const float* Q;  int ldq;     // [B*Lq, d] (head h at columns h*hd)
    const float* KV; int ldkv;    // [B*Lk, 2d]: K at column h*hd, V at column d + h*hd
    float* O;                     // [B*Lq, d]
    int Lq, Lk;
    // backward only
    const float* dO; float* dQ; float* dKV;
};
struct AttnArgs { AttnDir dir[2]; int B, heads, hd, d; float scale; };

// C[m][n] = sum_k a(m,k) b(k,n) for small operands living in shared memory, all kAttnThreads threads,
// 4 x 4 outputs per thread, a warp covers a 16 x 32 block of C (4 x 8 micro-tiles: operand reads are
// broadcasts or conflict-free).  a(m,k) = TA ? A[k*lda+m] : A[m*lda+k];  b(k,n) = TB ? B[n*ldb+k] : B[k*ldb+n].
constexpr int kAttnThreads = 256;
template <bool TA, bool TB, typename F>
__device__ __forceinline__ void smem_gemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M, int N, int K,
                                          F&& store) {
    const int tiles_m = (M + 3) >> 2, tiles_n = (N + 3) >> 2;
    // a warp covers (32 / TN) x TN micro-tiles; 4 x 8 is the conflict-free default, but an 8-region image side (N = 8: two tile
    // columns, or M = 8: two tile rows) would leave 75 % / 50 % of the lanes without a tile: take the shape that pads least
    int tn_shift = 3;
    {
        int best = ((tiles_m + 3) >> 2) * ((tiles_n + 7) >> 3);
        for (int sh = 0; sh <= 5; ++sh) {
            const int TNc = 1 << sh, TMc = 32 >> sh;
            const int cost = ((tiles_m + TMc - 1) / TMc) * ((tiles_n + TNc - 1) / TNc);
            if (cost < best) { best = cost; tn_shift = sh; }
        }
    }
    const int TN = 1 << tn_shift, TM = 32 >> tn_shift;
    const int blocks_n = (tiles_n + TN - 1) >> tn_shift, blocks_m = (tiles_m + TM - 1) / TM;
    for (int t = threadIdx.x; t < blocks_m * blocks_n * 32; t += kAttnThreads) {
        const int blk = t >> 5, inb = t & 31;
        const int tm = (blk / blocks_n) * TM + (inb >> tn_shift), tn = (blk % blocks_n) * TN + (inb & (TN - 1));
        if (tm >= tiles_m || tn >= tiles_n) continue;
        const int m0 = tm * 4, n0 = tn * 4;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = min(m0 + i, M - 1), n = min(n0 + i, N - 1);     // clamped reads, masked stores
                av[i] = TA ? A[k * lda + m] : A[m * lda + k];
                bv[i] = TB ? B[n * ldb + k] : B[k * ldb + n];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (m0 + i < M && n0 + j < N) store(m0 + i, n0 + j, acc[i][j]);
    }
}

// row softmax of P[Lq][lp] in place (one warp per row); optionally dS = P o (dP - rowsum(dP o P)) * scale
__device__ __forceinline__ void smem_softmax_rows(float* P, int lp, int Lq, int Lk, float* dS, float scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < Lq; q += kAttnThreads / 32) {
        float mx = -INFINITY;
        for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, P[q * lp + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < Lk; j += 32) { const float e = expf(P[q * lp + j] - mx); P[q * lp + j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float dot = 0.f;
        for (int j = lane; j < Lk; j += 32) {
            const float pv = P[q * lp + j] * inv;
            P[q * lp + j] = pv;
            if (dS) dot = fmaf(pv, dS[q * lp + j], dot);
        }
        if (dS) {
            dot = warp_sum(dot);
            for (int j = lane; j < Lk; j += 32) dS[q * lp + j] = P[q * lp + j] * (dS[q * lp + j] - dot) * scale;
        }
    }
}

__device__ __forceinline__ void load_head(float* dst, int ld, const float* src, int64_t row0, int rows, int64_t ldsrc, int col0, int hd) {
    const int q4 = hd >> 2;
    if ((hd & 3) == 0 && (ldsrc & 3) == 0 && (col0 & 3) == 0 && q4 <= kAttnThreads && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // a thread keeps its 4-column group and walks rows: one index split per call instead of one per element (the per-element
        // idx / hd was as many instructions as the FMAs of the five micro-GEMMs together), 16-byte global loads
        const int c = (threadIdx.x % q4) * 4, r0 = threadIdx.x / q4, rstep = kAttnThreads / q4;
        if (r0 < rstep)            // threads past the last whole row group (hd / 4 does not divide the block) sit out
            for (int r = r0; r < rows; r += rstep) {
                const float4 v = *reinterpret_cast<const float4*>(src + (row0 + r) * ldsrc + col0 + c);
                float* o = dst + r * ld + c;
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            }
        return;
    }
    for (int idx = threadIdx.x; idx < rows * hd; idx += kAttnThreads) {
        const int r = idx / hd, e = idx - r * hd;
        dst[r * ld + e] = src[(row0 + r) * ldsrc + col0 + e];
    }
}

__global__ void __launch_bounds__(kAttnThreads) k_attn_fwd(const __grid_constant__ AttnArgs a) {
    extern __shared__ float sm[];
    const AttnDir& D = a.dir[blockIdx.y];
    const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads, hd = a.hd, ld = hd + 1;
    const int Lq = D.Lq, Lk = D.Lk, lp = Lk + 1;
    float* Qs = sm;                   // [Lq][ld]
    float* Ks = Qs + Lq * ld;         // [Lk][ld]
    float* Vs = Ks + Lk * ld;         // [Lk][ld]
    float* Ps = Vs + Lk * ld;         // [Lq][lp]
    load_head(Qs, ld, D.Q, (int64_t)b * Lq, Lq, D.ldq, h * hd, hd);
    load_head(Ks, ld, D.KV, (int64_t)b * Lk, Lk, D.ldkv, h * hd, hd);
    load_head(Vs, ld, D.KV, (int64_t)b * Lk, Lk, D.ldkv, a.d + h * hd, hd);
    __syncthreads();
    const float scale = a.scale;
    smem_gemm<false, true>(Qs, ld, Ks, ld, Lq, Lk, hd, [&](int q, int j, float v) { Ps[q * lp + j] = v * scale; });
    __syncthreads();
    smem_softmax_rows(Ps, lp, Lq, Lk, nullptr, 0.f);
    __syncthreads();
    float* O = D.O + (size_t)b * Lq * a.d + h * hd;
    const int dd = a.d;
    smem_gemm<false, false>(Ps, lp, Vs, ld, Lq, hd, Lk, [&](int q, int e, float v) { O[(size_t)q * dd + e] = v; });
}

// backward: recompute P, then dV = P^T dO, dS = P o (dP - rowsum(dP o P)) scale, dQ = dS K, dK = dS^T Q
__global__ void __launch_bounds__(kAttnThreads) k_attn_bwd(const __grid_constant__ AttnArgs a) {
    extern __shared__ float sm[];
    const AttnDir& D = a.dir[blockIdx.y];
    const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads, hd = a.hd, ld = hd + 1;
    const int Lq = D.Lq, Lk = D.Lk, lp = Lk + 1;
    float* Qs = sm;                     // [Lq][ld]
    float* dOs = Qs + Lq * ld;          // [Lq][ld]
    float* Ks = dOs + Lq * ld;          // [Lk][ld]
    float* Vs = Ks + Lk * ld;           // [Lk][ld]
    float* Ps = Vs + Lk * ld;           // [Lq][lp]
    float* dSs = Ps + Lq * lp;          // [Lq][lp]
    load_head(Qs, ld, D.Q, (int64_t)b * Lq, Lq, D.ldq, h * hd, hd);
    load_head(dOs, ld, D.dO, (int64_t)b * Lq, Lq, a.d, h * hd, hd);
    load_head(Ks, ld, D.KV, (int64_t)b * Lk, Lk, D.ldkv, h * hd, hd);
    load_head(Vs, ld, D.KV, (int64_t)b * Lk, Lk, D.ldkv, a.d + h * hd, hd);
    __syncthreads();
    const float scale = a.scale;
    smem_gemm<false, true>(Qs, ld, Ks, ld, Lq, Lk, hd, [&](int q, int j, float v) { Ps[q * lp + j] = v * scale; });
    smem_gemm<false, true>(dOs, ld, Vs, ld, Lq, Lk, hd, [&](int q, int j, float v) { dSs[q * lp + j] = v; });      // dP
    __syncthreads();
    smem_softmax_rows(Ps, lp, Lq, Lk, dSs, scale);
    __syncthreads();
    const int dd = a.d;
    float* dQ = D.dQ + (size_t)b * Lq * dd + h * hd;
    float* dKV = D.dKV + (size_t)b * Lk * 2 * dd + h * hd;
    smem_gemm<false, false>(dSs, lp, Ks, ld, Lq, hd, Lk, [&](int q, int e, float v) { dQ[(size_t)q * dd + e] = v; });
    smem_gemm<true, false>(dSs, lp, Qs, ld, Lk, hd, Lq, [&](int j, int e, float v) { dKV[(size_t)j * 2 * dd + e] = v; });         // dK
    smem_gemm<true, false>(Ps, lp, dOs, ld, Lk, hd, Lq, [&](int j, int e, float v) { dKV[(size_t)j * 2 * dd + dd + e] = v; });   // dV
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, biased variance, affine): one warp per row, d <= 4096
// ---------------------------------------------------------------------------------------------
constexpr float kLnEps = 1e-5f;
// both sides (text rows, image rows) in one launch: blocks [0, nblk0) serve side 0
struct LnFwdSide32 { const float* z; const float* gamma; const float* beta; int rows; float* out; float* stats; };
struct LnFwdArgs32 { LnFwdSide32 s[2]; int d, nblk0; };
__global__ void __launch_bounds__(256) k_ln_fwd(const __grid_constant__ LnFwdArgs32 a) {
    const int side = (int)blockIdx.x >= a.nblk0 ? 1 : 0;
    const LnFwdSide32& S = a.s[side];
    const int d = a.d;
    const int lane = threadIdx.x & 31;
    const int row = ((int)blockIdx.x - (side ? a.nblk0 : 0)) * 8 + (threadIdx.x >> 5);
    if (row >= S.rows) return;
    const float* zr = S.z + (size_t)row * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += zr[c];
    const float mu = warp_sum(s) / d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = zr[c] - mu; v = fmaf(t, t, v); }
    const float rstd = rsqrtf(warp_sum(v) / d + kLnEps);
    for (int c = lane; c < d; c += 32) S.out[(size_t)row * d + c] = (zr[c] - mu) * rstd * S.gamma[c] + S.beta[c];
    if (lane == 0) { S.stats[2 * row] = mu; S.stats[2 * row + 1] = rstd; }
}
static int launch_ln_fwd32(const LnFwdSide32& s0, const LnFwdSide32& s1, int d, cudaStream_t st) {
    LnFwdArgs32 a{};
    a.s[0] = s0; a.s[1] = s1; a.d = d; a.nblk0 = (int)ceil_div(s0.rows, 8);
    k_ln_fwd<<<(unsigned)(a.nblk0 + ceil_div(s1.rows, 8)), 256, 0, st>>>(a);
    DCLIP_LAUNCH_OK("k_ln_fwd");
    return DCLIP_OK;
}
// dz = rstd (dy g - mean(dy g) - zhat mean(dy g zhat));  also writes zhat for the dgamma column sums
struct LnBwdSide32 { const float* dy; const float* z; const float* gamma; const float* stats; int rows; float* dz; float* zhat; };
struct LnBwdArgs32 { LnBwdSide32 s[2]; int d, nblk0; };
__global__ void __launch_bounds__(256) k_ln_bwd(const __grid_constant__ LnBwdArgs32 a) {
    const int side = (int)blockIdx.x >= a.nblk0 ? 1 : 0;
    const LnBwdSide32& S = a.s[side];
    const int d = a.d;
    const int lane = threadIdx.x & 31;
    const int row = ((int)blockIdx.x - (side ? a.nblk0 : 0)) * 8 + (threadIdx.x >> 5);
    if (row >= S.rows) return;
    const float mu = S.stats[2 * row], rstd = S.stats[2 * row + 1];
    const float* zr = S.z + (size_t)row * d;
    const float* dyr = S.dy + (size_t)row * d;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float zh = (zr[c] - mu) * rstd, g = dyr[c] * S.gamma[c];
        s1 += g;
        s2 = fmaf(g, zh, s2);
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
    for (int c = lane; c < d; c += 32) {
        const float zh = (zr[c] - mu) * rstd, g = dyr[c] * S.gamma[c];
        S.dz[(size_t)row * d + c] = rstd * (g - s1 - zh * s2);
        S.zhat[(size_t)row * d + c] = zh;
    }
}
static int launch_ln_bwd32(const LnBwdSide32& s0, const LnBwdSide32& s1, int d, cudaStream_t st) {
    LnBwdArgs32 a{};
    a.s[0] = s0; a.s[1] = s1; a.d = d; a.nblk0 = (int)ceil_div(s0.rows, 8);
    k_ln_bwd<<<(unsigned)(a.nblk0 + ceil_div(s1.rows, 8)), 256, 0, st>>>(a);
    DCLIP_LAUNCH_OK("k_ln_bwd");
    return DCLIP_OK;
}

// ---------------------------------------------------------------------------------------------
// aggregation (patch_text_aggregation.py:243-265): one CTA per batch element
// ---------------------------------------------------------------------------------------------
constexpr float kCosEps = 1e-8f;
constexpr int kAggThreads = 1024;
__device__ __forceinline__ float block_sum(float v, float* red) {   // kAggThreads threads
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < kAggThreads / 32; ++i) t += red[i];
    return t;
}

// The kernels make several passes over x[b] (L x d); staged once into shared memory when it fits (16-byte
// coalesced loads), every later pass reads shared memory instead of L dependent global loads per column.
template <typename T>
__device__ __forceinline__ const T* agg_stage(const T* xb, int L, int d, void* stage_raw, bool staged) {
    if (!staged) return xb;
    const size_t n16 = (size_t)L * d * sizeof(T) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(xb);
    uint4* dst = reinterpret_cast<uint4*>(stage_raw);
    for (size_t i = threadIdx.x; i < n16; i += kAggThreads) dst[i] = __ldg(src + i);
    __syncthreads();
    return reinterpret_cast<const T*>(stage_raw);
}

template <typename T>
__global__ void __launch_bounds__(kAggThreads) k_aggregate_fwd(const T* __restrict__ x, int L, int d, float temperature, T* __restrict__ out,
                                                       float* __restrict__ w_saved, int staged) {
    extern __shared__ __align__(16) float sm[];
    float* mean = sm;            // [d]
    float* cs = mean + d;        // [L]
    float* red = cs + L;         // [32]
    const int b = blockIdx.x;
    const T* xb = agg_stage<T>(x + (size_t)b * L * d, L, d, sm + (((size_t)d + L + 32 + 3) & ~(size_t)3), staged != 0);
    float part = 0.f;
    for (int c = threadIdx.x; c < d; c += kAggThreads) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += to_f32<T>(xb[(size_t)l * d + c]);
        s /= L;
        mean[c] = s;
        part = fmaf(s, s, part);
    }
    const float nm = fmaxf(sqrtf(block_sum(part, red)), kCosEps);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int l = warp; l < L; l += kAggThreads / 32) {
        float dot = 0.f, nx = 0.f;
        for (int c = lane; c < d; c += 32) { const float v = to_f32<T>(xb[(size_t)l * d + c]); dot = fmaf(v, mean[c], dot); nx = fmaf(v, v, nx); }
        dot = warp_sum(dot); nx = warp_sum(nx);
        if (lane == 0) cs[l] = dot / (fmaxf(sqrtf(nx), kCosEps) * nm) / temperature;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int l = 0; l < L; ++l) mx = fmaxf(mx, cs[l]);
    float den = 0.f;
    for (int l = 0; l < L; ++l) den += expf(cs[l] - mx);
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += kAggThreads) {
        const float w = expf(cs[l] - mx) / den;
        cs[l] = w;
        if (w_saved) w_saved[(size_t)b * L + l] = w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += kAggThreads) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s = fmaf(cs[l], to_f32<T>(xb[(size_t)l * d + c]), s);
        out[(size_t)b * d + c] = from_f32<T>(s);
    }
}

template <typename T>
__global__ void __launch_bounds__(kAggThreads) k_aggregate_bwd(const T* __restrict__ x, const float* __restrict__ w_saved, const T* __restrict__ dout,
                                                       int L, int d, float temperature, T* __restrict__ dx, int staged) {
    extern __shared__ __align__(16) float sm[];
    float* mean = sm;            // [d]
    float* dmean = mean + d;     // [d]
    float* cosv = dmean + d;     // [L]  cosine similarities
    float* nxs = cosv + L;       // [L]  clamped row norms
    float* dcs = nxs + L;        // [L]  d loss / d cos
    float* red = dcs + L;        // [32]
    const int b = blockIdx.x;
    const T* xb = agg_stage<T>(x + (size_t)b * L * d, L, d, sm + (((size_t)2 * d + 3 * L + 32 + 3) & ~(size_t)3), staged != 0);
    const T* dg = dout + (size_t)b * d;
    const float* w = w_saved + (size_t)b * L;
    float part = 0.f;
    for (int c = threadIdx.x; c < d; c += kAggThreads) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += to_f32<T>(xb[(size_t)l * d + c]);
        s /= L;
        mean[c] = s;
        part = fmaf(s, s, part);
    }
    const float nm_raw = sqrtf(block_sum(part, red));
    const float nm = fmaxf(nm_raw, kCosEps);
    const bool m_clamped = nm_raw < kCosEps;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int l = warp; l < L; l += kAggThreads / 32) {
        float dot = 0.f, nx = 0.f, dw = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float v = to_f32<T>(xb[(size_t)l * d + c]);
            dot = fmaf(v, mean[c], dot); nx = fmaf(v, v, nx); dw = fmaf(v, to_f32<T>(dg[c]), dw);
        }
        dot = warp_sum(dot); nx = warp_sum(nx); dw = warp_sum(dw);
        if (lane == 0) {
            const float nxr = sqrtf(nx);
            nxs[l] = nxr < kCosEps ? -kCosEps : nxr;     // negative marks "clamped"
            cosv[l] = dot / (fmaxf(nxr, kCosEps) * nm);
            dcs[l] = dw;                                  // dw_l for now
        }
    }
    __syncthreads();
    float sdot = 0.f;
    for (int l = 0; l < L; ++l) sdot = fmaf(w[l], dcs[l], sdot);
    __syncthreads();
    // per-row coefficients (no division left in the two L x d passes below):
    //   dc_l = w_l / T (dw_l - sum w dw);  e_l = dc_l / (nx_l nm);  f_l = dc_l cos_l / nx_l^2 (0 if the row norm is clamped)
    //   dmean_c = (sum_l e_l x_lc - mean_c S2) / L,  S2 = sum_l dc_l cos_l / nm^2 (0 if the mean norm is clamped)
    //   dx_lc   = w_l dg_c + e_l mean_c - f_l x_lc + dmean_c
    float* el = nxs;             // reuse: e_l   (after reading nxs below)
    float* fl = cosv;            // reuse: f_l
    float s2part = 0.f;
    for (int l = threadIdx.x; l < L; l += kAggThreads) {
        const float dc = w[l] / temperature * (dcs[l] - sdot);
        const float nx = fabsf(nxs[l]);
        const bool row_ok = nxs[l] > 0.f;
        const float cosl = cosv[l];
        s2part += m_clamped ? 0.f : dc * cosl / (nm * nm);
        dcs[l] = dc;
        el[l] = dc / (nx * nm);
        fl[l] = row_ok ? dc * cosl / (nx * nx) : 0.f;
    }
    const float S2 = block_sum(s2part, red);
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += kAggThreads) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) s = fmaf(el[l], to_f32<T>(xb[(size_t)l * d + c]), s);
        dmean[c] = (s - mean[c] * S2) / L;
    }
    __syncthreads();
    const int warp2 = threadIdx.x >> 5, lane2 = threadIdx.x & 31;
    for (int l = warp2; l < L; l += kAggThreads / 32) {
        const float wl = w[l], e = el[l], f = fl[l];
        const T* xr = xb + (size_t)l * d;
        T* dr = dx + ((size_t)b * L + l) * d;
        for (int c = lane2; c < d; c += 32)
            dr[c] = from_f32<T>(fmaf(wl, to_f32<T>(dg[c]), fmaf(e, mean[c], fmaf(-f, to_f32<T>(xr[c]), dmean[c]))));
    }
}

// combine + whole-tensor NaN/Inf guard (patch_text_aggregation.py:647-651)
template <typename T>
__global__ void k_combine_fwd(const T* __restrict__ a, const T* __restrict__ b, int64_t n, T* __restrict__ out, int* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = 0.5f * to_f32<T>(a[i]) + 0.5f * to_f32<T>(b[i]);
    out[i] = from_f32<T>(v);
    if (!isfinite(v)) atomicOr(flag, 1);
}
template <typename T>
__global__ void k_guard_zero(T* __restrict__ out, int64_t n, const int* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && *flag) out[i] = from_f32<T>(0.f);
}
template <typename T>
__global__ void k_combine_bwd(const T* __restrict__ dout, int64_t n, const int* __restrict__ flag, T* __restrict__ da, T* __restrict__ db) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = *flag ? 0.f : 0.5f * to_f32<T>(dout[i]);
    da[i] = from_f32<T>(v);
    db[i] = from_f32<T>(v);
}

// ---------------------------------------------------------------------------------------------
// saved-buffer layout (floats)
// ---------------------------------------------------------------------------------------------
struct XLayout {
    size_t Xt, Xi, Qt, KVi, Qi, KVt, Ot, Oi, Zt, Zi, St, Si, O16t, O16i, W, total;
    // W: fp32 copies of the 12 parameter tensors (only when dtype != f32)
};
static const int kParamCount = 12;
// order of the params[] / grads[] arrays == state_dict order of the reference module
//  0 t2i.in_proj_weight [3d,d]  1 t2i.in_proj_bias [3d]  2 t2i.out_proj.weight [d,d]  3 t2i.out_proj.bias [d]
//  4 i2t.in_proj_weight         5 i2t.in_proj_bias       6 i2t.out_proj.weight        7 i2t.out_proj.bias
//  8 norm_text.weight [d]  9 norm_text.bias  10 norm_image.weight  11 norm_image.bias
static size_t param_numel(int i, int64_t d) {
    switch (i % 4 + (i >= 8 ? 4 : 0)) {
        case 0: return (size_t)3 * d * d;
        case 1: return (size_t)3 * d;
        case 2: return (size_t)d * d;
        case 3: return (size_t)d;
        default: return (size_t)d;
    }
}
static XLayout xlayout(int64_t B, int64_t T, int64_t R, int64_t d, bool widen) {
    XLayout L{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
    const size_t nt = (size_t)B * T, ni = (size_t)B * R;
    L.Xt = take(widen ? nt * d : 0); L.Xi = take(widen ? ni * d : 0);
    L.Qt = take(nt * d); L.KVi = take(ni * 2 * d); L.Qi = take(ni * d); L.KVt = take(nt * 2 * d);
    L.Ot = take(nt * d); L.Oi = take(ni * d); L.Zt = take(nt * d); L.Zi = take(ni * d);
    L.St = take(nt * 2); L.Si = take(ni * 2);
    L.O16t = take(widen ? (nt * d + 1) / 2 : 0); L.O16i = take(widen ? (ni * d + 1) / 2 : 0);   // 16-bit copies of O (tensor tier)
    L.W = o;
    if (widen) for (int i = 0; i < kParamCount; ++i) take(param_numel(i, d));
    L.total = o;
    return L;
}

template <typename T>
static int widen(const void* src, float* dst, size_t n, cudaStream_t st) {
    if (n == 0) return DCLIP_OK;
    k_to_f32<T><<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, st>>>(static_cast<const T*>(src), dst, (int64_t)n);
    DCLIP_LAUNCH_OK("k_to_f32");
    return DCLIP_OK;
}
template <typename T>
static int narrow(const float* src, void* dst, size_t n, cudaStream_t st) {
    if (n == 0) return DCLIP_OK;
    k_from_f32<T><<<(unsigned)ceil_div((int64_t)n, 256), 256, 0, st>>>(src, static_cast<T*>(dst), (int64_t)n);
    DCLIP_LAUNCH_OK("k_from_f32");
    return DCLIP_OK;
}

static int launch_cast_group(const CastGroup& g, bool widen_dir, int dtype, cudaStream_t st) {
    if (g.count == 0) return DCLIP_OK;
    long long mx = 0;
    for (int i = 0; i < g.count; ++i) mx = g.n[i] > mx ? g.n[i] : mx;
    if (mx == 0) return DCLIP_OK;
    const dim3 grid((unsigned)min((long long)1184, (mx + 255) / 256), g.count);
    if (widen_dir) { DCLIP_DISPATCH_DTYPE(dtype, (k_widen_group<T><<<grid, 256, 0, st>>>(g))); }
    else { DCLIP_DISPATCH_DTYPE(dtype, (k_narrow_group<T><<<grid, 256, 0, st>>>(g))); }
    DCLIP_LAUNCH_OK("k_cast_group");
    return DCLIP_OK;
}
static void cast_add(CastGroup& g, const void* src, void* dst, size_t n) {
    g.src[g.count] = src; g.dst[g.count] = dst; g.n[g.count] = (long long)n; ++g.count;
}

// tensor tier available?  (16-bit dtype on an sm_100 device; DCLIP_XATTN_TENSOR=0 forces the fp32 tier)
static bool xattn_use_tc(int dtype, int64_t d) {
    if (dtype == DCLIP_F32 || d % 8 != 0 || !device_is_sm100()) return false;
    const char* e = getenv("DCLIP_XATTN_TENSOR");
    return !(e && e[0] == '0');
}

static TcGemmProblem tcp(const void* A, int lda, bool a_mn, const void* B, int ldb, bool b_mn, float* C, int ldc, int M, int N, int K,
                         const float* bias = nullptr, const float* residual = nullptr, int ldr = 0, bool accumulate = false) {
    TcGemmProblem p{};
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.residual = residual; p.C16 = nullptr;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldr = ldr;
    p.a_mn = a_mn; p.b_mn = b_mn; p.accumulate = accumulate;
    return p;
}

// 16-bit tier with every contraction on tcgen05 (xattn_tc.cu)
bool xattn_tc_shape_ok(int64_t T, int64_t R, int64_t d, int heads);
size_t xattn_tc_saved_bytes(int64_t B, int64_t T, int64_t R, int64_t d);
size_t xattn_tc_scratch_bytes(int64_t B, int64_t T, int64_t R, int64_t d);
int xattn_fwd_tc(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype, const void* const* params,
                 void* text_out, void* image_out, void* saved, size_t saved_bytes, cudaStream_t st);
int xattn_bwd_tc(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype, const void* const* params,
                 const void* saved, const void* d_text_out, const void* d_image_out, void* const* grads, void* d_text, void* d_image,
                 void* scratch, size_t scratch_bytes, cudaStream_t st);
// DCLIP_XATTN_ATTN=simt keeps the round-1 hybrid (tcgen05 projections, fp32 CUDA-core attention) for A/B timing
static bool xattn_full_tc(int dtype, int64_t T, int64_t R, int64_t d, int heads) {
    if (!xattn_use_tc(dtype, d) || !xattn_tc_shape_ok(T, R, d, heads)) return false;
    const char* e = getenv("DCLIP_XATTN_ATTN");
    return !(e && e[0] == 's');
}

static size_t attn_fwd_smem(int Lq, int Lk, int hd) { return ((size_t)(Lq + 2 * Lk) * (hd + 1) + (size_t)Lq * (Lk + 1)) * 4; }
static size_t attn_bwd_smem(int Lq, int Lk, int hd) { return ((size_t)(2 * Lq + 2 * Lk) * (hd + 1) + (size_t)2 * Lq * (Lk + 1)) * 4; }

}  // namespace dclip

using namespace dclip;

extern "C" {

size_t dclip_xattn_saved_bytes(int64_t B, int64_t T, int64_t R, int64_t d, int dtype) {
    const size_t base = xlayout(B, T, R, d, dtype != DCLIP_F32).total * sizeof(float) + 256;
    if (dtype == DCLIP_F32) return base + 256 + split_fwd_bytes(B, T, R, d);      // + the split operands of the tensor-core fp32 GEMMs
    return max(base, xattn_tc_saved_bytes(B, T, R, d));     // either 16-bit tier may be chosen at call time
}
// scratch for the backward: dZ, zhat, dO (text+image), dQ, dKV (both) + widened output grads
size_t dclip_xattn_bwd_scratch_bytes(int64_t B, int64_t T, int64_t R, int64_t d, int dtype) {
    const size_t nt = (size_t)B * T, ni = (size_t)B * R;
    size_t f = (nt + ni) * d * 6 + (nt + ni) * 2 * d + 1024;       // dY, dZ, zhat, dO, dQ, dX + dKV
    if (dtype != DCLIP_F32) { f += (nt + ni) * d; for (int i = 0; i < kParamCount; ++i) f += param_numel(i, d) + 64; f += (nt + ni) * d * 2 + 1024; }
    const size_t base = f * sizeof(float) + 4096;
    if (dtype == DCLIP_F32) return base + 256 + split_bwd_bytes(B, T, R, d);
    return max(base, xattn_tc_scratch_bytes(B, T, R, d));
}

int dclip_xattn_fwd(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype,
                    const void* const* params, void* text_out, void* image_out, void* saved, size_t saved_bytes, void* stream) {
    DCLIP_REQUIRE(text && image && params && text_out && image_out && saved, "xattn_fwd: null pointer");
    DCLIP_REQUIRE(B > 0 && T > 0 && R > 0 && d > 0 && heads > 0 && d % heads == 0, "xattn_fwd: bad shape B=%lld T=%lld R=%lld d=%lld heads=%d",
                  (long long)B, (long long)T, (long long)R, (long long)d, heads);
    if (xattn_full_tc(dtype, T, R, d, heads))
        return xattn_fwd_tc(text, image, B, T, R, d, heads, dtype, params, text_out, image_out, saved, saved_bytes, static_cast<cudaStream_t>(stream));
    const bool wide = dtype != DCLIP_F32;
    const XLayout L = xlayout(B, T, R, d, wide);
    if (saved_bytes < L.total * sizeof(float)) return fail(DCLIP_ERR_WORKSPACE, "xattn_fwd: saved buffer too small (%zu < %zu)", saved_bytes, L.total * sizeof(float));
    const int hd = (int)(d / heads);
    // fp32 inputs on an sm_100 device: the four GEMM launches run on tcgen05 through bf16 x 3 operand splitting (scratch behind the saved tensors)
    SplitScratch sc{nullptr, 0};
    if (!wide && d % 8 == 0 && gemm_split_enabled()) {
        const size_t off = align_up(L.total * sizeof(float), 256);
        if (saved_bytes >= off + split_fwd_bytes(B, T, R, d)) sc = SplitScratch{static_cast<char*>(saved) + off, split_fwd_bytes(B, T, R, d)};
    }
    const size_t sm_t2i = attn_fwd_smem((int)T, (int)R, hd), sm_i2t = attn_fwd_smem((int)R, (int)T, hd);
    if (max(sm_t2i, sm_i2t) > 220 * 1024) return fail(DCLIP_ERR_UNSUPPORTED, "xattn_fwd: sequence x head_dim too large for shared memory");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* S = static_cast<float*>(saved);
    const int nt = (int)(B * T), ni = (int)(B * R), D = (int)d;
    const float* Xt = static_cast<const float*>(text);
    const float* Xi = static_cast<const float*>(image);
    const float* Wp[kParamCount];
    const bool tc = xattn_use_tc(dtype, d);
    if (wide) {
        // fp32 copies: parameters (the tensor tier consumes the four weight matrices in 16 bit directly and
        // only needs the small bias / LayerNorm vectors widened) and the two inputs (residual path)
        CastGroup cg{};
        size_t o = L.W;
        for (int i = 0; i < kParamCount; ++i) {
            const size_t n = param_numel(i, d);
            const bool is_matrix = (i % 4 == 0 || i % 4 == 2) && i < 8;
            if (!(tc && is_matrix)) cast_add(cg, params[i], S + o, n);
            Wp[i] = S + o;
            o += (n + 63) / 64 * 64;
        }
        if (int e = launch_cast_group(cg, true, dtype, st)) return e;
        CastGroup cx{};
        cast_add(cx, text, S + L.Xt, (size_t)nt * d);
        cast_add(cx, image, S + L.Xi, (size_t)ni * d);
        if (int e = launch_cast_group(cx, true, dtype, st)) return e;
        Xt = S + L.Xt; Xi = S + L.Xi;
    } else {
        for (int i = 0; i < kParamCount; ++i) Wp[i] = static_cast<const float*>(params[i]);
    }
    // (1) in-projections: Q_t, K|V_i (text_to_image) and Q_i, K|V_t (image_to_text)
    const size_t esz = 2;   // 16-bit element size (tensor tier only)
    const char* w_t2i = static_cast<const char*>(params[0]);
    const char* w_i2t = static_cast<const char*>(params[4]);
    if (tc) {
        TcGemmProblem pr[4];
        pr[0] = tcp(text, D, false, w_t2i, D, false, S + L.Qt, D, nt, D, D, Wp[1]);
        pr[1] = tcp(image, D, false, w_t2i + (size_t)d * d * esz, D, false, S + L.KVi, 2 * D, ni, 2 * D, D, Wp[1] + d);
        pr[2] = tcp(image, D, false, w_i2t, D, false, S + L.Qi, D, ni, D, D, Wp[5]);
        pr[3] = tcp(text, D, false, w_i2t + (size_t)d * d * esz, D, false, S + L.KVt, 2 * D, nt, 2 * D, D, Wp[5] + d);
        if (int e = launch_gemm_tc(pr, 4, dtype, st)) return e;
    } else {
        GemmGroup g{};
        g.p[0] = gemm(Xt, D, false, Wp[0], D, true, S + L.Qt, D, nt, D, D, Wp[1]);
        g.p[1] = gemm(Xi, D, false, Wp[0] + (size_t)d * d, D, true, S + L.KVi, 2 * D, ni, 2 * D, D, Wp[1] + d);
        g.p[2] = gemm(Xi, D, false, Wp[4], D, true, S + L.Qi, D, ni, D, D, Wp[5]);
        g.p[3] = gemm(Xt, D, false, Wp[4] + (size_t)d * d, D, true, S + L.KVt, 2 * D, nt, 2 * D, D, Wp[5] + d);
        g.count = 4;
        if (int e = run_gemm_group(g, sc, st)) return e;
    }
    // (2) attention cores
    AttnArgs a{};
    a.B = (int)B; a.heads = heads; a.hd = hd; a.d = D; a.scale = 1.0f / sqrtf((float)hd);
    a.dir[0] = AttnDir{S + L.Qt, D, S + L.KVi, 2 * D, S + L.Ot, (int)T, (int)R, nullptr, nullptr, nullptr};
    a.dir[1] = AttnDir{S + L.Qi, D, S + L.KVt, 2 * D, S + L.Oi, (int)R, (int)T, nullptr, nullptr, nullptr};
    const size_t smf = max(sm_t2i, sm_i2t);
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_attn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
    k_attn_fwd<<<dim3((unsigned)(B * heads), 2), kAttnThreads, smf, st>>>(a);
    DCLIP_LAUNCH_OK("k_attn_fwd");
    // (3) out-projection + bias + residual -> Z
    if (tc) {
        void* O16t = const_cast<float*>(S + L.O16t);
        void* O16i = const_cast<float*>(S + L.O16i);
        CastGroup co{};
        cast_add(co, S + L.Ot, O16t, (size_t)nt * d);
        cast_add(co, S + L.Oi, O16i, (size_t)ni * d);
        if (int e = launch_cast_group(co, false, dtype, st)) return e;
        TcGemmProblem pr[2];
        pr[0] = tcp(O16t, D, false, params[2], D, false, S + L.Zt, D, nt, D, D, Wp[3], Xt, D);
        pr[1] = tcp(O16i, D, false, params[6], D, false, S + L.Zi, D, ni, D, D, Wp[7], Xi, D);
        if (int e = launch_gemm_tc(pr, 2, dtype, st)) return e;
    } else {
        GemmGroup g2{};
        g2.p[0] = gemm(S + L.Ot, D, false, Wp[2], D, true, S + L.Zt, D, nt, D, D, Wp[3], Xt, D);
        g2.p[1] = gemm(S + L.Oi, D, false, Wp[6], D, true, S + L.Zi, D, ni, D, D, Wp[7], Xi, D);
        g2.count = 2;
        if (int e = run_gemm_group(g2, sc, st)) return e;
    }
    // (4) LayerNorm
    if (!wide) {
        if (int e = launch_ln_fwd32(LnFwdSide32{S + L.Zt, Wp[8], Wp[9], nt, static_cast<float*>(text_out), S + L.St},
                                    LnFwdSide32{S + L.Zi, Wp[10], Wp[11], ni, static_cast<float*>(image_out), S + L.Si}, D, st)) return e;
    } else {
        // fp32 result staged in the widened-input slots (X copies are re-derived in the backward from the originals)
        if (int e = launch_ln_fwd32(LnFwdSide32{S + L.Zt, Wp[8], Wp[9], nt, S + L.Xt, S + L.St},
                                    LnFwdSide32{S + L.Zi, Wp[10], Wp[11], ni, S + L.Xi, S + L.Si}, D, st)) return e;
        CastGroup cy{};
        cast_add(cy, S + L.Xt, text_out, (size_t)nt * d);
        cast_add(cy, S + L.Xi, image_out, (size_t)ni * d);
        if (int e = launch_cast_group(cy, false, dtype, st)) return e;
    }
    return DCLIP_OK;
}

int dclip_xattn_bwd(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype,
                    const void* const* params, const void* saved, const void* d_text_out, const void* d_image_out,
                    void* const* grads, void* d_text, void* d_image, void* scratch, size_t scratch_bytes, void* stream) {
    DCLIP_REQUIRE(text && image && params && saved && d_text_out && d_image_out && grads && scratch, "xattn_bwd: null pointer");
    DCLIP_REQUIRE((d_text == nullptr) == (d_image == nullptr), "xattn_bwd: give both input gradients or neither");
    if (scratch_bytes < dclip_xattn_bwd_scratch_bytes(B, T, R, d, dtype)) return fail(DCLIP_ERR_WORKSPACE, "xattn_bwd: scratch too small");
    if (xattn_full_tc(dtype, T, R, d, heads))
        return xattn_bwd_tc(text, image, B, T, R, d, heads, dtype, params, saved, d_text_out, d_image_out, grads, d_text, d_image, scratch,
                            scratch_bytes, static_cast<cudaStream_t>(stream));
    const bool wide = dtype != DCLIP_F32;
    const XLayout L = xlayout(B, T, R, d, wide);
    const int hd = (int)(d / heads);
    const size_t smb = max(attn_bwd_smem((int)T, (int)R, hd), attn_bwd_smem((int)R, (int)T, hd));
    if (smb > 220 * 1024) return fail(DCLIP_ERR_UNSUPPORTED, "xattn_bwd: sequence x head_dim too large for shared memory");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* S = static_cast<const float*>(saved);
    const int nt = (int)(B * T), ni = (int)(B * R), D = (int)d;
    // scratch carve (floats)
    float* W = static_cast<float*>(scratch);
    size_t o = 0;
    auto take = [&](size_t n) { float* r = W + o; o += (n + 63) / 64 * 64; return r; };
    float* dYt = take((size_t)nt * d); float* dYi = take((size_t)ni * d);     // widened upstream grads
    float* dZt = take((size_t)nt * d); float* dZi = take((size_t)ni * d);
    float* zht = take((size_t)nt * d); float* zhi = take((size_t)ni * d);
    float* dOt = take((size_t)nt * d); float* dOi = take((size_t)ni * d);
    float* dQt = take((size_t)nt * d); float* dQi = take((size_t)ni * d);
    float* dKVi = take((size_t)ni * 2 * d); float* dKVt = take((size_t)nt * 2 * d);
    float* dXt = take((size_t)nt * d); float* dXi = take((size_t)ni * d);
    const bool tc = xattn_use_tc(dtype, d);
    // 16-bit GEMM operands of the tensor tier (element counts halved: two per float slot)
    void* dZ16t = tc ? take(((size_t)nt * d + 1) / 2) : nullptr; void* dZ16i = tc ? take(((size_t)ni * d + 1) / 2) : nullptr;
    void* dQ16t = tc ? take(((size_t)nt * d + 1) / 2) : nullptr; void* dQ16i = tc ? take(((size_t)ni * d + 1) / 2) : nullptr;
    void* dKV16i = tc ? take((size_t)ni * d) : nullptr;           void* dKV16t = tc ? take((size_t)nt * d) : nullptr;
    float* Xt_w = nullptr; float* Xi_w = nullptr;
    float* G[kParamCount];
    const float* Wp[kParamCount];
    const float* Xt = static_cast<const float*>(text);
    const float* Xi = static_cast<const float*>(image);
    const float* dyt = static_cast<const float*>(d_text_out);
    const float* dyi = static_cast<const float*>(d_image_out);
    if (wide) {
        Xt_w = take((size_t)nt * d); Xi_w = take((size_t)ni * d);
        size_t wo = L.W;
        for (int i = 0; i < kParamCount; ++i) {
            G[i] = take(param_numel(i, d));
            Wp[i] = S + wo;
            wo += (param_numel(i, d) + 63) / 64 * 64;
        }
        CastGroup cw{};
        if (!tc) {   // the tensor tier feeds the 16-bit inputs to the GEMMs directly
            cast_add(cw, text, Xt_w, (size_t)nt * d);
            cast_add(cw, image, Xi_w, (size_t)ni * d);
        }
        cast_add(cw, d_text_out, dYt, (size_t)nt * d);
        cast_add(cw, d_image_out, dYi, (size_t)ni * d);
        if (int e = launch_cast_group(cw, true, dtype, st)) return e;
        Xt = Xt_w; Xi = Xi_w; dyt = dYt; dyi = dYi;
    } else {
        for (int i = 0; i < kParamCount; ++i) { G[i] = static_cast<float*>(grads[i]); Wp[i] = static_cast<const float*>(params[i]); }
    }
    if (o * sizeof(float) > scratch_bytes) return fail(DCLIP_ERR_WORKSPACE, "xattn_bwd: scratch too small (%zu < %zu)", scratch_bytes, o * sizeof(float));
    SplitScratch sc{nullptr, 0};      // fp32 inputs: tensor-core GEMMs through operand splitting (see run_gemm_group)
    if (!wide && d % 8 == 0 && gemm_split_enabled()) {
        const size_t off = align_up(o * sizeof(float), 256);
        if (scratch_bytes >= off + split_bwd_bytes(B, T, R, d)) sc = SplitScratch{static_cast<char*>(scratch) + off, split_bwd_bytes(B, T, R, d)};
    }

    // (4') LayerNorm backward: dZ, zhat; dgamma = colsum(dY * zhat), dbeta = colsum(dY)
    if (int e = launch_ln_bwd32(LnBwdSide32{dyt, S + L.Zt, Wp[8], S + L.St, nt, dZt, zht}, LnBwdSide32{dyi, S + L.Zi, Wp[10], S + L.Si, ni, dZi, zhi}, D, st))
        return e;
    ColsumGroup cg{};
    cg.p[0] = ColsumProblem{dyt, zht, nt, D, D, D, G[9], G[8]};       // dbeta = colsum(dY), dgamma = colsum(dY * zhat)
    cg.p[1] = ColsumProblem{dyi, zhi, ni, D, D, D, G[11], G[10]};
    cg.p[2] = ColsumProblem{dZt, nullptr, nt, D, D, 0, G[3], nullptr};   // out-proj bias
    cg.p[3] = ColsumProblem{dZi, nullptr, ni, D, D, 0, G[7], nullptr};
    cg.count = 4;
    if (int e = launch_colsum(cg, st)) return e;
    // (3') out-projection backward: dWo = dZ^T O, dbo = colsum(dZ), dO = dZ Wo
    if (tc) {
        CastGroup cz{};
        cast_add(cz, dZt, dZ16t, (size_t)nt * d);
        cast_add(cz, dZi, dZ16i, (size_t)ni * d);
        if (int e = launch_cast_group(cz, false, dtype, st)) return e;
        TcGemmProblem pr[4];
        pr[0] = tcp(dZ16t, D, true, S + L.O16t, D, true, G[2], D, D, D, nt);             // dWo = dZ^T O
        pr[1] = tcp(dZ16i, D, true, S + L.O16i, D, true, G[6], D, D, D, ni);
        pr[2] = tcp(dZ16t, D, false, params[2], D, true, dOt, D, nt, D, D);              // dO = dZ Wo
        pr[3] = tcp(dZ16i, D, false, params[6], D, true, dOi, D, ni, D, D);
        if (int e = launch_gemm_tc(pr, 4, dtype, st)) return e;
    } else {
        GemmGroup g{};
        g.p[0] = gemm(dZt, D, true, S + L.Ot, D, false, G[2], D, D, D, nt);
        g.p[1] = gemm(dZi, D, true, S + L.Oi, D, false, G[6], D, D, D, ni);
        g.p[2] = gemm(dZt, D, false, Wp[2], D, false, dOt, D, nt, D, D);
        g.p[3] = gemm(dZi, D, false, Wp[6], D, false, dOi, D, ni, D, D);
        g.count = 4;
        if (int e = run_gemm_group(g, sc, st)) return e;
    }
    // (2') attention cores backward
    AttnArgs a{};
    a.B = (int)B; a.heads = heads; a.hd = hd; a.d = D; a.scale = 1.0f / sqrtf((float)hd);
    a.dir[0] = AttnDir{S + L.Qt, D, S + L.KVi, 2 * D, nullptr, (int)T, (int)R, dOt, dQt, dKVi};
    a.dir[1] = AttnDir{S + L.Qi, D, S + L.KVt, 2 * D, nullptr, (int)R, (int)T, dOi, dQi, dKVt};
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_attn_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    k_attn_bwd<<<dim3((unsigned)(B * heads), 2), kAttnThreads, smb, st>>>(a);
    DCLIP_LAUNCH_OK("k_attn_bwd");
    // (1') in-projection backward: weight grads (rows 0..d = q, d..3d = k|v) and bias grads
    ColsumGroup cg2{};
    cg2.p[0] = ColsumProblem{dQt, nullptr, nt, D, D, 0, G[1], nullptr};
    cg2.p[1] = ColsumProblem{dKVi, nullptr, ni, 2 * D, 2 * D, 0, G[1] + d, nullptr};
    cg2.p[2] = ColsumProblem{dQi, nullptr, ni, D, D, 0, G[5], nullptr};
    cg2.p[3] = ColsumProblem{dKVt, nullptr, nt, 2 * D, 2 * D, 0, G[5] + d, nullptr};
    cg2.count = 4;
    if (int e = launch_colsum(cg2, st)) return e;
    const char* w_t2i = static_cast<const char*>(params[0]);
    const char* w_i2t = static_cast<const char*>(params[4]);
    if (tc) {
        CastGroup cq{};
        cast_add(cq, dQt, dQ16t, (size_t)nt * d);
        cast_add(cq, dQi, dQ16i, (size_t)ni * d);
        cast_add(cq, dKVi, dKV16i, (size_t)ni * 2 * d);
        cast_add(cq, dKVt, dKV16t, (size_t)nt * 2 * d);
        if (int e = launch_cast_group(cq, false, dtype, st)) return e;
        TcGemmProblem pr[4];
        pr[0] = tcp(dQ16t, D, true, text, D, true, G[0], D, D, D, nt);                              // dWq = dQ^T X
        pr[1] = tcp(dKV16i, 2 * D, true, image, D, true, G[0] + (size_t)d * d, D, 2 * D, D, ni);     // dWkv = dKV^T X
        pr[2] = tcp(dQ16i, D, true, image, D, true, G[4], D, D, D, ni);
        pr[3] = tcp(dKV16t, 2 * D, true, text, D, true, G[4] + (size_t)d * d, D, 2 * D, D, nt);
        if (int e = launch_gemm_tc(pr, 4, dtype, st)) return e;
        if (d_text) {
            TcGemmProblem px[2];
            px[0] = tcp(dQ16t, D, false, w_t2i, D, true, dXt, D, nt, D, D, nullptr, dZt, D);       // residual path + dQ Wq
            px[1] = tcp(dQ16i, D, false, w_i2t, D, true, dXi, D, ni, D, D, nullptr, dZi, D);
            if (int e = launch_gemm_tc(px, 2, dtype, st)) return e;
            TcGemmProblem py[2];
            py[0] = tcp(dKV16t, 2 * D, false, w_i2t + (size_t)d * d * 2, D, true, dXt, D, nt, D, 2 * D, nullptr, nullptr, 0, true);
            py[1] = tcp(dKV16i, 2 * D, false, w_t2i + (size_t)d * d * 2, D, true, dXi, D, ni, D, 2 * D, nullptr, nullptr, 0, true);
            if (int e = launch_gemm_tc(py, 2, dtype, st)) return e;
        }
    } else {
        GemmGroup g2{};
        g2.p[0] = gemm(dQt, D, true, Xt, D, false, G[0], D, D, D, nt);
        g2.p[1] = gemm(dKVi, 2 * D, true, Xi, D, false, G[0] + (size_t)d * d, D, 2 * D, D, ni);
        g2.p[2] = gemm(dQi, D, true, Xi, D, false, G[4], D, D, D, ni);
        g2.p[3] = gemm(dKVt, 2 * D, true, Xt, D, false, G[4] + (size_t)d * d, D, 2 * D, D, nt);
        g2.count = 4;
        if (int e = run_gemm_group(g2, sc, st)) return e;
        if (d_text) {
            // input gradients: residual path + query projection + key/value projection of the other direction
            GemmGroup g3{};
            g3.p[0] = gemm(dQt, D, false, Wp[0], D, false, dXt, D, nt, D, D, nullptr, dZt, D);
            g3.p[1] = gemm(dQi, D, false, Wp[4], D, false, dXi, D, ni, D, D, nullptr, dZi, D);
            g3.count = 2;
            if (int e = run_gemm_group(g3, sc, st)) return e;
            GemmGroup g4{};
            g4.p[0] = gemm(dKVt, 2 * D, false, Wp[4] + (size_t)d * d, D, false, dXt, D, nt, D, 2 * D, nullptr, nullptr, 0, true);
            g4.p[1] = gemm(dKVi, 2 * D, false, Wp[0] + (size_t)d * d, D, false, dXi, D, ni, D, 2 * D, nullptr, nullptr, 0, true);
            g4.count = 2;
            if (int e = run_gemm_group(g4, sc, st)) return e;
        }
    }
    if (d_text) {
        if (wide) {
            CastGroup cdx{};
            cast_add(cdx, dXt, d_text, (size_t)nt * d);
            cast_add(cdx, dXi, d_image, (size_t)ni * d);
            if (int e = launch_cast_group(cdx, false, dtype, st)) return e;
        } else {
            DCLIP_CUDA_OK(cudaMemcpyAsync(d_text, dXt, (size_t)nt * d * 4, cudaMemcpyDeviceToDevice, st));
            DCLIP_CUDA_OK(cudaMemcpyAsync(d_image, dXi, (size_t)ni * d * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (wide) {
        CastGroup cgr{};
        for (int i = 0; i < kParamCount; ++i) cast_add(cgr, G[i], grads[i], param_numel(i, d));
        if (int e = launch_cast_group(cgr, false, dtype, st)) return e;
    }
    return DCLIP_OK;
}

// out[n, m] = x[n, k] W[m, k]^T (+ bias[m]) - the per-token `text_projection` loop of training/text_tokenizer.py:202-207
// as ONE GEMM over all rows (nn.Linear layout: W is [out_features, in_features]).  16-bit inputs: tcgen05 (gemm_tc.cu,
// fp32 accumulation, output rounded once); fp32 inputs: the exact fp32 GEMM.
int dclip_linear_rows(const void* x, const void* W, const float* bias, int64_t n, int64_t k, int64_t m, int dtype, void* out, void* stream) {
    DCLIP_REQUIRE(x && W && out && n > 0 && k > 0 && m > 0, "linear_rows: bad arguments");
    DCLIP_REQUIRE(n < (1ll << 31) && k < (1ll << 31) && m < (1ll << 31), "linear_rows: dimension too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DCLIP_F32) {
        GemmGroup g{};
        g.p[0] = gemm(static_cast<const float*>(x), (int)k, false, static_cast<const float*>(W), (int)k, true, static_cast<float*>(out), (int)m,
                      (int)n, (int)m, (int)k, bias);
        g.count = 1;
        return launch_gemm_group(g, st);
    }
    DCLIP_REQUIRE(k % 8 == 0 && m % 8 == 0, "linear_rows: the 16-bit path needs in / out features in multiples of 8 (got %lld, %lld)", (long long)k, (long long)m);
    if (!device_is_sm100()) return fail(DCLIP_ERR_ARCH, "linear_rows: the 16-bit path needs an sm_100 device");
    TcGemmProblem p = tcp(x, (int)k, false, W, (int)k, false, nullptr, (int)m, (int)n, (int)m, (int)k, bias);
    p.C16 = out; p.ldc16 = (int)m;
    return launch_gemm_tc(&p, 1, dtype, st);
}

int dclip_aggregate_fwd(const void* x, int64_t B, int64_t L, int64_t d, int dtype, float temperature, void* out, float* weights,
                        void* stream) {
    DCLIP_REQUIRE(x && out && B > 0 && L > 0 && d > 0 && temperature > 0.f, "aggregate_fwd: bad arguments");
    size_t sm = (((size_t)d + L + 32 + 3) & ~(size_t)3) * 4;
    if (sm > 200 * 1024) return fail(DCLIP_ERR_UNSUPPORTED, "aggregate_fwd: d + L too large");
    const size_t rows = (size_t)L * d * dtype_size(dtype);
    const int staged = rows % 16 == 0 && aligned16(x) && sm + rows <= 220 * 1024 ? 1 : 0;
    if (staged) sm += rows;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_DISPATCH_DTYPE(dtype, {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_aggregate_fwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_aggregate_fwd<T><<<(unsigned)B, kAggThreads, sm, st>>>(static_cast<const T*>(x), (int)L, (int)d, temperature, static_cast<T*>(out), weights,
                                                         staged);
    });
    DCLIP_LAUNCH_OK("k_aggregate_fwd");
    return DCLIP_OK;
}

int dclip_aggregate_bwd(const void* x, const float* weights, const void* d_out, int64_t B, int64_t L, int64_t d, int dtype,
                        float temperature, void* d_x, void* stream) {
    DCLIP_REQUIRE(x && weights && d_out && d_x && B > 0 && L > 0 && d > 0 && temperature > 0.f, "aggregate_bwd: bad arguments");
    size_t sm = (((size_t)2 * d + 3 * L + 32 + 3) & ~(size_t)3) * 4;
    if (sm > 200 * 1024) return fail(DCLIP_ERR_UNSUPPORTED, "aggregate_bwd: d + L too large");
    const size_t rows = (size_t)L * d * dtype_size(dtype);
    const int staged = rows % 16 == 0 && aligned16(x) && sm + rows <= 220 * 1024 ? 1 : 0;
    if (staged) sm += rows;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_DISPATCH_DTYPE(dtype, {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_aggregate_bwd<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_aggregate_bwd<T><<<(unsigned)B, kAggThreads, sm, st>>>(static_cast<const T*>(x), weights, static_cast<const T*>(d_out), (int)L, (int)d,
                                                         temperature, static_cast<T*>(d_x), staged);
    });
    DCLIP_LAUNCH_OK("k_aggregate_bwd");
    return DCLIP_OK;
}

int dclip_combine_guard_fwd(const void* a, const void* b, int64_t n, int dtype, void* out, int* flag, void* stream) {
    DCLIP_REQUIRE(a && b && out && flag && n > 0, "combine_guard_fwd: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
    DCLIP_DISPATCH_DTYPE(dtype, {
        k_combine_fwd<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), n, static_cast<T*>(out), flag);
        k_guard_zero<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(static_cast<T*>(out), n, flag);
    });
    DCLIP_LAUNCH_OK("k_combine_fwd");
    return DCLIP_OK;
}

int dclip_combine_guard_bwd(const void* d_out, int64_t n, int dtype, const int* flag, void* d_a, void* d_b, void* stream) {
    DCLIP_REQUIRE(d_out && flag && d_a && d_b && n > 0, "combine_guard_bwd: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_DISPATCH_DTYPE(dtype, {
        k_combine_bwd<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(static_cast<const T*>(d_out), n, flag, static_cast<T*>(d_a), static_cast<T*>(d_b));
    });
    DCLIP_LAUNCH_OK("k_combine_bwd");
    return DCLIP_OK;
}

}  // extern "C"
