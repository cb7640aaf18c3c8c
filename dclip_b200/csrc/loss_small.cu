// Distillation objective for SMALL batches (B <= 64) in ONE launch — the teacher step's loss
// (training/train_contrastive_teacher.py:251-261, :357) and the reference's own distillation batch size (32,
// README.md:51-59; training/CLIP_image_distillation.py:532-576, :594-634).
//
// At B = 32 the B x B logits are 4 KB and the whole step is a few MFLOP: the generic path's six launches
// (normalise, row/column log-sum-exp, combine, loss sums, gradient tiles, gradient tail) are pure launch latency.
// Here one cluster of 8 CTAs does everything; CTA r owns the column slice [r w, (r+1) w) of the embedding dimension:
//   1. loads its slice of the four [B, d] inputs to shared memory (fp32),
//   2. forms the PARTIAL raw logits  sum_c u_ic v_jc,  squared norms and student-teacher dots over its slice,
//   3. cluster exchange #1: every CTA adds the 8 partial blocks in rank order (distributed shared memory reads), so all
//      CTAs hold bit-identical full sums; normalisation (F.normalize eps 1e-12) is applied to the logits afterwards,
//   4. every CTA computes the B x B softmax statistics redundantly (row / column log-sum-exp, G = the logits'
//      gradient); CTA 0 writes the four losses,
//   5. gradient of the normalised rows on the CTA's slice (G V_hat / tau + cosine term), partial projection dots
//      u_hat . d_u_hat -> cluster exchange #2 -> normalise-backward on the slice, scaled and rounded to the I/O dtype.
// fp32 arithmetic end to end (both tiers use it: inputs may be fp32, bf16 or fp16); fixed summation order.
#include <cooperative_groups.h>

#include "common.cuh"
#include "loss_internal.h"

namespace cg = cooperative_groups;

namespace dclip {

namespace sm {
constexpr int kCluster = 8;
constexpr int kThreads = 1024;        // one CTA per SM, 32 warps: the per-thread loops are shared-memory latency chains
constexpr int kMaxB = 64;
constexpr float kNormEps = 1e-12f;
}  // namespace sm

#ifdef DCLIP_SMALL_PROF
__device__ long long g_small_prof[16];
#define SPROF(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_small_prof[k] = clock64(); } while (0)
#else
#define SPROF(k) do { } while (0)
#endif

struct SmallLossArgs {
    const void* s_img; const void* s_txt; const void* t_img; const void* t_txt;
    int B, d, w;                       // w = columns per CTA (multiple of 8)
    float inv_tau, w_img, w_txt, w_con, grad_scale;
    float* losses; void* d_img; void* d_txt; float* stats;
};

// Shared-memory layouts are chosen for the access patterns of the three micro-GEMMs (the kernel is shared-memory
// bandwidth bound, not FLOP bound):
//   slices U, V, TU, TV  [B][ws], ws = w + 4 (an odd multiple of 4 words): 16-byte aligned rows, and 8 consecutive ROWS read
//                        with LDS.128 at the same column fall into 8 different bank groups
//   logits / G           [B][B + 1]: row walks and column walks are both conflict-free
// partial block (exchange #1, floats): S[B*B] | nu[B] nv[B] ntu[B] ntv[B] dut[B] dvt[B];  exchange #2: ru[B] rv[B]
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// e / n for the index splits of the phase loops: the kernel is instruction-bound on its 8 SMs and a run-time integer division is
// ~30 instructions (ncu source view: 19 % of all executed instructions were these splits) - batch sizes are powers of two in practice
struct FastDiv {      // exact for 0 <= e < 2^16, 1 <= n < 2^16 (multiply-high by ceil(2^32 / n))
    int n; unsigned magic;
    __device__ __forceinline__ explicit FastDiv(int n_) : n(n_), magic(n_ > 1 ? 0xFFFFFFFFu / (unsigned)n_ + 1u : 0u) {}
    __device__ __forceinline__ int div(int e) const { return n > 1 ? (int)__umulhi((unsigned)e, magic) : e; }
};

template <typename T>
__global__ void __cluster_dims__(sm::kCluster, 1, 1) __launch_bounds__(sm::kThreads) k_distill_small(const SmallLossArgs a) {
    using namespace sm;
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int B = a.B, d = a.d, w = a.w;
    const int ws = w + 4, BP = B + 1, hB = (B + 1) >> 1;
    const int c0 = rank * w, wc = max(0, min(w, d - c0));            // my column slice (may be empty for the last ranks)
    const bool has_t = a.t_img != nullptr;
    const int NP = B * B + 6 * B;
    float* U = smem;                       // [B][ws] raw student image rows, later u_hat
    float* V = U + (size_t)B * ws;
    float* TU = V + (size_t)B * ws;        // teacher rows, later the gradients of the normalised rows
    float* TV = TU + (size_t)B * ws;
    float* part = TV + (size_t)B * ws;     // my partial block [NP]
    float* full = part + NP;               // summed block [NP]
    float* Sg = full + NP;                 // logits [B][BP]
    float* G = Sg + B * BP;                // [B][BP]
    float* vec = G + B * BP;               // inv_u, inv_v, inv_tu, inv_tv, lse_r, lse_c, ru, rv : 8 x [B]
    float* part2 = vec + 8 * B;            // [2B] projection dots, exchanged
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;

    SPROF(0);
    // 1. load my slice (8 columns per thread and tensor; columns past d are zeros)
#pragma unroll 1
    for (int e = tid; e < B * (w >> 3); e += kThreads) {
        const int i = FastDiv(w >> 3).div(e), c = (e - i * (w >> 3)) * 8;
        const bool in = c < wc;          // wc is a multiple of 8 (d % 8 == 0, w % 8 == 0)
        const size_t g = (size_t)i * d + c0 + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const void* src = q == 0 ? a.s_img : q == 1 ? a.s_txt : q == 2 ? a.t_img : a.t_txt;
            float* dst = (q == 0 ? U : q == 1 ? V : q == 2 ? TU : TV) + (size_t)i * ws + c;
            float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (in && src) load8<T>(static_cast<const T*>(src) + g, x);
            *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
        }
    }
    __syncthreads();
    SPROF(1);
    // 2. partial raw logits: thread (ti, tj) owns rows {ti, ti + hB} x columns {tj, tj + hB}
#pragma unroll 1
    for (int e = tid; e < hB * hB; e += kThreads) {
        const int ti = FastDiv(hB).div(e), tj = e - ti * hB;
        const int i1 = min(ti + hB, B - 1), j1 = min(tj + hB, B - 1);
        const float* u0 = U + (size_t)ti * ws; const float* u1 = U + (size_t)i1 * ws;
        const float* v0 = V + (size_t)tj * ws; const float* v1 = V + (size_t)j1 * ws;
        float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll 2
        for (int c = 0; c < w; c += 4) {
            const float4 a0 = lds4(u0 + c), a1 = lds4(u1 + c), b0 = lds4(v0 + c), b1 = lds4(v1 + c);
            s00 = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, s00))));
            s01 = fmaf(a0.x, b1.x, fmaf(a0.y, b1.y, fmaf(a0.z, b1.z, fmaf(a0.w, b1.w, s01))));
            s10 = fmaf(a1.x, b0.x, fmaf(a1.y, b0.y, fmaf(a1.z, b0.z, fmaf(a1.w, b0.w, s10))));
            s11 = fmaf(a1.x, b1.x, fmaf(a1.y, b1.y, fmaf(a1.z, b1.z, fmaf(a1.w, b1.w, s11))));
        }
        part[ti * B + tj] = s00;
        if (tj + hB < B) part[ti * B + tj + hB] = s01;
        if (ti + hB < B) {
            part[(ti + hB) * B + tj] = s10;
            if (tj + hB < B) part[(ti + hB) * B + tj + hB] = s11;
        }
    }
    // squared norms and student-teacher dots: one warp per (quantity, row)
    const FastDiv divB(B);
#pragma unroll 1
    for (int q = 0; q < 6; ++q) {
        const float* xq = (q == 0 || q == 4) ? U : (q == 1 || q == 5) ? V : q == 2 ? TU : TV;
        const float* yq = q == 4 ? TU : q == 5 ? TV : xq;
        if (!has_t && q >= 2) {                        // no teachers: those sums are zero (no warp walks them)
            for (int i = tid; i < B; i += kThreads) part[B * B + q * B + i] = 0.f;
            continue;
        }
#pragma unroll 1
        for (int i = warp; i < B; i += kWarps) {
            const float* x = xq + (size_t)i * ws;
            const float* y = yq + (size_t)i * ws;
            float s = 0.f;
            for (int c = lane * 4; c < w; c += 128) {
                const float4 p = lds4(x + c), r = lds4(y + c);
                s = fmaf(p.x, r.x, fmaf(p.y, r.y, fmaf(p.z, r.z, fmaf(p.w, r.w, s))));
            }
            s = warp_sum(s);
            if (lane == 0) part[B * B + q * B + i] = s;
        }
    }
    SPROF(2);
    cluster.sync();
    SPROF(3);
    // 3. exchange #1: fixed rank order -> identical sums in every CTA
#pragma unroll 1
    for (int e = tid; e < NP; e += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < kCluster; ++r) s += *cluster.map_shared_rank(part + e, r);
        full[e] = s;
    }
    __syncthreads();
    SPROF(4);
    float* inv_u = vec; float* inv_v = vec + B; float* inv_tu = vec + 2 * B; float* inv_tv = vec + 3 * B;
    float* lse_r = vec + 4 * B; float* lse_c = vec + 5 * B;
#pragma unroll 1
    for (int e = tid; e < 4 * B; e += kThreads) vec[e] = 1.0f / fmaxf(sqrtf(full[B * B + e]), kNormEps);
    __syncthreads();
    // 4. logits, statistics, G
#pragma unroll 1
    for (int e = tid; e < B * B; e += kThreads) {
        const int i = divB.div(e), j = e - i * B;
        Sg[i * BP + j] = full[e] * inv_u[i] * inv_v[j] * a.inv_tau;
    }
    __syncthreads();
    SPROF(5);
#pragma unroll 1
    for (int e = warp; e < 2 * B; e += kWarps) {        // row / column log-sum-exp: one warp per (side, row), fixed-order reductions
        const int side = e >= B ? 1 : 0, i = e - side * B;
        float mx = -INFINITY;
        for (int j = lane; j < B; j += 32) mx = fmaxf(mx, side ? Sg[j * BP + i] : Sg[i * BP + j]);
        mx = warp_max(mx);
        float s = 0.f;
        for (int j = lane; j < B; j += 32) s += __expf((side ? Sg[j * BP + i] : Sg[i * BP + j]) - mx);
        s = warp_sum(s);
        if (lane == 0) vec[(4 + side) * B + i] = mx + __logf(s);
    }
    __syncthreads();
    SPROF(6);
    if (rank == 0 && warp == 0) {
        // losses (one warp, fixed order): L_con = (sum lse_r + sum lse_c - 2 sum diag) / 2B ; L_img = mean(1 - cos)
        float sc = 0.f, si = 0.f, st = 0.f;
        for (int i = lane; i < B; i += 32) {
            sc += lse_r[i] + lse_c[i] - 2.0f * Sg[i * BP + i];
            if (has_t) {
                si += 1.0f - full[B * B + 4 * B + i] * inv_u[i] * inv_tu[i];
                st += 1.0f - full[B * B + 5 * B + i] * inv_v[i] * inv_tv[i];
            }
        }
        sc = warp_sum(sc); si = warp_sum(si); st = warp_sum(st);
        if (lane == 0 && a.losses) {
            const float l_con = sc / (2.0f * B), l_img = si / B, l_txt = st / B;
            a.losses[0] = a.w_img * l_img + a.w_txt * l_txt + a.w_con * l_con;
            a.losses[1] = l_img; a.losses[2] = l_txt; a.losses[3] = l_con;
        }
    }
    if (rank == 0 && a.stats)
#pragma unroll 1
        for (int e = tid; e < 2 * B; e += kThreads) a.stats[e] = vec[4 * B + e];
    if (!a.d_img) { cluster.sync(); return; }      // forward only (no CTA may exit while a peer still reads its shared memory)
    SPROF(7);
    const float coef = a.w_con / (2.0f * B);
#pragma unroll 1
    for (int e = tid; e < B * B; e += kThreads) {
        const int i = divB.div(e), j = e - i * B;
        const float s = Sg[i * BP + j];
        G[i * BP + j] = coef * (__expf(s - lse_r[i]) + __expf(s - lse_c[j]) - (i == j ? 2.0f : 0.f));
    }
    // normalised rows of my slice (in place)
#pragma unroll 1
    for (int e = tid; e < B * (w >> 2); e += kThreads) {
        const int i = FastDiv(w >> 2).div(e), c = (e - i * (w >> 2)) * 4;
        float4 x = lds4(U + (size_t)i * ws + c), y = lds4(V + (size_t)i * ws + c);
        const float su = inv_u[i], sv = inv_v[i];
        *reinterpret_cast<float4*>(U + (size_t)i * ws + c) = make_float4(x.x * su, x.y * su, x.z * su, x.w * su);
        *reinterpret_cast<float4*>(V + (size_t)i * ws + c) = make_float4(y.x * sv, y.y * sv, y.z * sv, y.w * sv);
    }
    __syncthreads();
    SPROF(8);
    // 5. d_u_hat / d_v_hat on my slice.  A warp owns (side, 4-column group, 32 rows): lane = row i reads its own G row /
    // column (conflict-free) and every lane the same 16 bytes of y_hat_j (broadcast): 2 shared loads per 4 FMAs.
    float* DU = TU;                       // the teacher slices are consumed on the fly and overwritten with the gradients
    float* DV = TV;
    const int ncg = w >> 2, nrb = (B + 31) >> 5;
    const FastDiv divR(nrb);
#pragma unroll 1
    for (int e = warp; e < 2 * ncg * nrb; e += kWarps) {
        const int side = e >= ncg * nrb ? 1 : 0, r = e - side * ncg * nrb;
        const int cgi = divR.div(r), i = (r - cgi * nrb) * 32 + lane;
        const int c = cgi * 4;
        const int ic = min(i, B - 1);
        const float* Y = (side ? U : V) + c;
        const float* g = side ? G + ic : G + ic * BP;
        const int gs = side ? BP : 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int j = 0; j < B; ++j) {
            const float gv = g[j * gs];
            const float4 y = lds4(Y + (size_t)j * ws);
            acc.x = fmaf(gv, y.x, acc.x); acc.y = fmaf(gv, y.y, acc.y); acc.z = fmaf(gv, y.z, acc.z); acc.w = fmaf(gv, y.w, acc.w);
        }
        if (i < B) {
            float* tptr = (side ? TV : TU) + (size_t)i * ws + c;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            float ts = 0.f;
            if (has_t) { t = lds4(tptr); ts = -(side ? a.w_txt : a.w_img) / B * (side ? inv_tv[i] : inv_tu[i]); }
            *reinterpret_cast<float4*>(tptr) = make_float4(fmaf(acc.x, a.inv_tau, ts * t.x), fmaf(acc.y, a.inv_tau, ts * t.y),
                                                           fmaf(acc.z, a.inv_tau, ts * t.z), fmaf(acc.w, a.inv_tau, ts * t.w));
        }
    }
    __syncthreads();
    SPROF(9);
    // partial projection dots u_hat_i . d_u_hat_i over my slice: one warp per (side, row)
#pragma unroll 1
    for (int e = warp; e < 2 * B; e += kWarps) {
        const int side = e >= B ? 1 : 0, i = e - side * B;
        const float* x = (side ? V : U) + (size_t)i * ws;
        const float* gq = (side ? DV : DU) + (size_t)i * ws;
        float s = 0.f;
        for (int c = lane * 4; c < w; c += 128) {
            const float4 p = lds4(x + c), r = lds4(gq + c);
            s = fmaf(p.x, r.x, fmaf(p.y, r.y, fmaf(p.z, r.z, fmaf(p.w, r.w, s))));
        }
        s = warp_sum(s);
        if (lane == 0) part2[e] = s;
    }
    SPROF(10);
    cluster.sync();
    SPROF(11);
#pragma unroll 1
    for (int e = tid; e < 2 * B; e += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < kCluster; ++r) s += *cluster.map_shared_rank(part2 + e, r);
        const float inv = e < B ? inv_u[e] : inv_v[e - B];
        vec[6 * B + e] = inv >= 1.0f / kNormEps ? 0.f : s;          // clamped norm: F.normalize has no projection term
    }
    __syncthreads();
    const int nch = w >> 3;
#pragma unroll 1
    for (int e = tid; e < 2 * B * nch; e += kThreads) {
        const int side = e >= B * nch ? 1 : 0, r = e - side * B * nch;
        const int i = FastDiv(nch).div(r), c = (r - i * nch) * 8;
        if (c >= wc) continue;
        const float* x = (side ? V : U) + (size_t)i * ws + c;
        const float* gq = (side ? DV : DU) + (size_t)i * ws + c;
        const float dot = vec[(6 + side) * B + i];
        const float sc = (side ? inv_v[i] : inv_u[i]) * a.grad_scale;
        const float4 x0 = lds4(x), x1 = lds4(x + 4), g0 = lds4(gq), g1 = lds4(gq + 4);
        const float o[8] = {(g0.x - x0.x * dot) * sc, (g0.y - x0.y * dot) * sc, (g0.z - x0.z * dot) * sc, (g0.w - x0.w * dot) * sc,
                            (g1.x - x1.x * dot) * sc, (g1.y - x1.y * dot) * sc, (g1.z - x1.z * dot) * sc, (g1.w - x1.w * dot) * sc};
        store8<T>(static_cast<T*>(side ? a.d_txt : a.d_img) + (size_t)i * d + c0 + c, o);
    }
    SPROF(12);
    cluster.sync();          // peers may still be reading my part2
    SPROF(13);
}

bool small_loss_shape_ok(int64_t B, int64_t d) { return B >= 1 && B <= sm::kMaxB && d % 8 == 0 && d >= 8 && d <= 2048; }

static size_t small_loss_smem(int B, int w) {
    const size_t NP = (size_t)B * B + 6 * B;
    return ((size_t)4 * B * (w + 4) + 2 * NP + (size_t)2 * B * (B + 1) + 8 * B + 2 * B) * sizeof(float) + 16;
}

int launch_distill_small(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t B, int64_t d, int dtype,
                         float temperature, float w_img, float w_txt, float w_con, float grad_scale, float* losses, void* d_img, void* d_txt,
                         float* stats, cudaStream_t st) {
    if (!small_loss_shape_ok(B, d)) return fail(DCLIP_ERR_UNSUPPORTED, "small-batch loss kernel: B <= 64, d %% 8 == 0, d <= 2048");
    SmallLossArgs a{};
    a.s_img = s_img; a.s_txt = s_txt; a.t_img = t_img; a.t_txt = t_txt;
    a.B = (int)B; a.d = (int)d;
    a.w = (int)(ceil_div(ceil_div(d, sm::kCluster), 8) * 8);
    a.inv_tau = 1.0f / temperature; a.w_img = w_img; a.w_txt = w_txt; a.w_con = w_con; a.grad_scale = grad_scale;
    a.losses = losses; a.d_img = d_img; a.d_txt = d_txt; a.stats = stats;
    const size_t smem = small_loss_smem(a.B, a.w);
    if (smem > 220 * 1024) return fail(DCLIP_ERR_UNSUPPORTED, "small-batch loss kernel: B x d too large for shared memory");
    DCLIP_DISPATCH_DTYPE(dtype, {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_distill_small<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_distill_small<T><<<sm::kCluster, sm::kThreads, smem, st>>>(a);
    });
    DCLIP_LAUNCH_OK("k_distill_small");
    return DCLIP_OK;
}

}  // namespace dclip

#ifdef DCLIP_SMALL_PROF
extern "C" int dclip_debug_small_prof(long long* out16) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    return cudaMemcpyFromSymbol(out16, dclip::g_small_prof, 16 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif
