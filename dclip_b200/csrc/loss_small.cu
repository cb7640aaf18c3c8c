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

// partial block layout (floats): S[B*B] | nu[B] nv[B] ntu[B] ntv[B] dut[B] dvt[B]  (exchange #1), ru[B] rv[B] (exchange #2)
template <typename T>
__global__ void __cluster_dims__(sm::kCluster, 1, 1) __launch_bounds__(sm::kThreads) k_distill_small(const SmallLossArgs a) {
    using namespace sm;
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int B = a.B, d = a.d, w = a.w;
    const int ws = w + 1;                  // padded row stride: rows of a slice start in different banks
    const int c0 = rank * w, wc = max(0, min(w, d - c0));            // my column slice (may be empty for the last ranks)
    const bool has_t = a.t_img != nullptr;
    const int NP = B * B + 6 * B;
    float* U = smem;                       // [B][ws] raw student image rows, later u_hat
    float* V = U + (size_t)B * ws;
    float* TU = V + (size_t)B * ws;        // teacher rows (unused without teachers)
    float* TV = TU + (size_t)B * ws;
    float* part = TV + (size_t)B * ws;     // my partial block [NP]
    float* full = part + NP;               // summed block [NP]
    float* G = full + NP;                  // [B][B]
    float* vec = G + B * B;                // inv_u, inv_v, inv_tu, inv_tv, lse_r, lse_c, ru, rv : 8 x [B]
    float* part2 = vec + 8 * B;            // [2B] projection dots, exchanged
    const int tid = threadIdx.x;

    SPROF(0);
    // 1. load my slice
    for (int e = tid; e < B * (w >> 3); e += kThreads) {
        const int i = e / (w >> 3), ch = e - i * (w >> 3);
        const int c = ch * 8;
        float x[8];
        const bool in = c < wc;          // wc is a multiple of 8 (d % 8 == 0, w % 8 == 0)
        const size_t g = (size_t)i * d + c0 + c;
        auto put = [&](const void* src, float* dst) {
            if (in && src) load8<T>(static_cast<const T*>(src) + g, x);
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[(size_t)i * ws + c + k] = x[k];
        };
        put(a.s_img, U); put(a.s_txt, V);
        if (has_t) { put(a.t_img, TU); put(a.t_txt, TV); }
    }
    __syncthreads();
    SPROF(1);
    // 2. partial sums over my slice
    for (int e = tid; e < B * B; e += kThreads) {
        const int i = e / B, j = e - i * B;
        const float* u = U + (size_t)i * ws;
        const float* v = V + (size_t)j * ws;
        float s = 0.f;
        for (int c = 0; c < w; ++c) s = fmaf(u[c], v[c], s);
        part[e] = s;
    }
    for (int e = tid; e < 6 * B; e += kThreads) {
        const int q = e / B, i = e - q * B;
        const float* x = (q == 0 || q == 4) ? U + (size_t)i * ws : (q == 1 || q == 5) ? V + (size_t)i * ws : q == 2 ? TU + (size_t)i * ws : TV + (size_t)i * ws;
        const float* y = q == 4 ? TU + (size_t)i * ws : q == 5 ? TV + (size_t)i * ws : x;
        float s = 0.f;
        if (has_t || q < 2)
            for (int c = 0; c < w; ++c) s = fmaf(x[c], y[c], s);
        part[B * B + e] = s;
    }
    SPROF(2);
    cluster.sync();
    SPROF(3);
    // 3. exchange #1: fixed rank order -> identical sums in every CTA
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
    for (int e = tid; e < 4 * B; e += kThreads) vec[e] = 1.0f / fmaxf(sqrtf(full[B * B + e]), kNormEps);
    __syncthreads();
    // 4. logits, statistics, G
    for (int e = tid; e < B * B; e += kThreads) {
        const int i = e / B, j = e - i * B;
        full[e] = full[e] * inv_u[i] * inv_v[j] * a.inv_tau;
    }
    __syncthreads();
    SPROF(5);
    {   // row / column log-sum-exp: one warp per (side, row), lanes over j, fixed-order shuffle reductions
        const int lane = tid & 31;
        for (int e = tid >> 5; e < 2 * B; e += kThreads / 32) {
            const int side = e / B, i = e - side * B;
            float mx = -INFINITY;
            for (int j = lane; j < B; j += 32) mx = fmaxf(mx, side ? full[j * B + i] : full[i * B + j]);
            mx = warp_max(mx);
            float s = 0.f;
            for (int j = lane; j < B; j += 32) s += expf((side ? full[j * B + i] : full[i * B + j]) - mx);
            s = warp_sum(s);
            if (lane == 0) vec[(4 + side) * B + i] = mx + logf(s);
        }
    }
    __syncthreads();
    SPROF(6);
    if (rank == 0 && tid < 32) {
        // losses (one warp, fixed order): L_con = (sum lse_r + sum lse_c - 2 sum diag) / 2B ; L_img = mean(1 - cos)
        float sc = 0.f, si = 0.f, st = 0.f;
        for (int i = tid; i < B; i += 32) {
            sc += lse_r[i] + lse_c[i] - 2.0f * full[i * B + i];
            if (has_t) {
                si += 1.0f - full[B * B + 4 * B + i] * inv_u[i] * inv_tu[i];
                st += 1.0f - full[B * B + 5 * B + i] * inv_v[i] * inv_tv[i];
            }
        }
        sc = warp_sum(sc); si = warp_sum(si); st = warp_sum(st);
        if (tid == 0 && a.losses) {
            const float l_con = sc / (2.0f * B), l_img = si / B, l_txt = st / B;
            a.losses[0] = a.w_img * l_img + a.w_txt * l_txt + a.w_con * l_con;
            a.losses[1] = l_img; a.losses[2] = l_txt; a.losses[3] = l_con;
        }
    }
    if (rank == 0 && a.stats)
        for (int e = tid; e < 2 * B; e += kThreads) a.stats[e] = vec[4 * B + e];
    if (!a.d_img) { cluster.sync(); return; }      // forward only (no CTA may exit while a peer still reads its shared memory)
    SPROF(7);
    const float coef = a.w_con / (2.0f * B);
    for (int e = tid; e < B * B; e += kThreads) {
        const int i = e / B, j = e - i * B;
        const float s = full[e];
        G[e] = coef * (expf(s - lse_r[i]) + expf(s - lse_c[j]) - (i == j ? 2.0f : 0.f));
    }
    // normalised rows of my slice (in place)
    for (int e = tid; e < B * w; e += kThreads) {
        const int i = e / w, c = e - i * w;
        U[(size_t)i * ws + c] *= inv_u[i]; V[(size_t)i * ws + c] *= inv_v[i];
    }
    __syncthreads();
    SPROF(8);
    // 5. d_u_hat / d_v_hat on my slice: thread = (row i, 8-column chunk); accumulate over j
    const int nch = w >> 3;
    float* DU = TU;                       // the teacher slices are consumed on the fly and overwritten with the gradients
    float* DV = TV;
    for (int e = tid; e < 2 * B * nch; e += kThreads) {
        const int side = e / (B * nch), r = e - side * B * nch;
        const int i = r / nch, c = (r - i * nch) * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* Y = side ? U : V;
        for (int j = 0; j < B; ++j) {
            const float g = side ? G[j * B + i] : G[i * B + j];
            const float* y = Y + (size_t)j * ws + c;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(g, y[k], acc[k]);
        }
        float* tptr = (side ? TV : TU) + (size_t)i * ws + c;
        const float ts = has_t ? -(side ? a.w_txt : a.w_img) / B * (side ? inv_tv[i] : inv_tu[i]) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tptr[k] = fmaf(acc[k], a.inv_tau, has_t ? ts * tptr[k] : 0.f);
    }
    __syncthreads();
    SPROF(9);
    for (int e = tid; e < 2 * B; e += kThreads) {
        const int side = e / B, i = e - side * B;
        const float* x = (side ? V : U) + (size_t)i * ws;
        const float* g = (side ? DV : DU) + (size_t)i * ws;
        float s = 0.f;
        for (int c = 0; c < w; ++c) s = fmaf(x[c], g[c], s);
        part2[e] = s;
    }
    SPROF(10);
    cluster.sync();
    SPROF(11);
    for (int e = tid; e < 2 * B; e += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < kCluster; ++r) s += *cluster.map_shared_rank(part2 + e, r);
        const float inv = e < B ? inv_u[e] : inv_v[e - B];
        vec[6 * B + e] = inv >= 1.0f / kNormEps ? 0.f : s;          // clamped norm: F.normalize has no projection term
    }
    __syncthreads();
    for (int e = tid; e < 2 * B * nch; e += kThreads) {
        const int side = e / (B * nch), r = e - side * B * nch;
        const int i = r / nch, c = (r - i * nch) * 8;
        if (c >= wc) continue;
        const float* x = (side ? V : U) + (size_t)i * ws + c;
        const float* g = (side ? DV : DU) + (size_t)i * ws + c;
        const float dot = vec[(6 + side) * B + i];
        const float sc = (side ? inv_v[i] : inv_u[i]) * a.grad_scale;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (g[k] - x[k] * dot) * sc;
        store8<T>(static_cast<T*>(side ? a.d_txt : a.d_img) + (size_t)i * d + c0 + c, o);
    }
    SPROF(12);
    cluster.sync();          // peers may still be reading my part2
    SPROF(13);
}

bool small_loss_shape_ok(int64_t B, int64_t d) { return B >= 1 && B <= sm::kMaxB && d % 8 == 0 && d >= 8 && d <= 2048; }

static size_t small_loss_smem(int B, int w) {
    const size_t NP = (size_t)B * B + 6 * B;
    return ((size_t)4 * B * (w + 1) + 2 * NP + (size_t)B * B + 8 * B + 2 * B) * sizeof(float) + 16;
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
