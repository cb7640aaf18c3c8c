// Global embedding of the meta-teacher — the tail of compute_global_embedding_batch fused into one launch each way:
//   aggregation(attended_text), aggregation(attended_image)   training/patch_text_aggregation.py:243-265, :643-644
//   global = 0.5 * text_global + 0.5 * image_global            :647
//   whole-tensor NaN / Inf guard (-> zeros)                    :649-651
// and the autograd backward of the same.
//
// Forward: one 2-CTA cluster per batch element (CTA 0 = text side, CTA 1 = image side).  A CTA stages its [L, d] block
// in shared memory once (16-byte loads; straight from L2 when the block does not fit) and makes three passes over it with
// one WARP PER ROW and lane-owned 8-element column chunks, so every pass is coalesced and the column reductions are
// register accumulators merged across the 8 warps through shared memory:
//   pass 1  column sums -> mean; row norms          pass 2  cos(x_l, mean) -> softmax over l (temperature 2.0)
//   pass 3  weighted column sum -> the side's global embedding
// The image CTA hands its vector to the text CTA through distributed shared memory; the text CTA writes
// 0.5 a + 0.5 b and the per-element non-finite flag.  The last cluster to finish zeroes the whole [B, d] output if any
// flag is set (device-side: the reference's `.any()` is a host sync).
// Backward: same cluster shape, no exchange: each side turns 0.5 * d_global (0 when the guard fired) into d_attended.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dclip {

namespace ge {
constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr float kCosEps = 1e-8f;
}  // namespace ge

struct GlobalSide { const void* x; int L; float* w; void* dx; };
struct GlobalArgs {
    GlobalSide s[2];
    int B, d;
    float temperature;
    void* out;                 // fwd: [B, d]
    const void* dout;          // bwd: [B, d]
    int* flags;                // [B] per-element non-finite flags, [B] = completion counter
    int staged[2];             // side's [L, d] block fits in shared memory
};

template <typename T>
__device__ __forceinline__ const T* ge_stage(const T* xb, int L, int d, void* stage, bool staged) {
    if (!staged) return xb;
    const size_t n16 = (size_t)L * d * sizeof(T) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(xb);
    uint4* dst = reinterpret_cast<uint4*>(stage);
    for (size_t i = threadIdx.x; i < n16; i += ge::kThreads) dst[i] = __ldg(src + i);
    __syncthreads();
    return reinterpret_cast<const T*>(stage);
}

// Per-column fp32 vectors in shared memory (mean, dmean, the merged accumulators) are kept CHUNK-TRANSPOSED: element i of
// 8-element chunk ch lives at [i * nch + ch], so the lanes of a warp (consecutive chunks) touch consecutive words.
// merge per-warp column accumulators: red[warp][d] -> dst[d] (sum over warps in order) * scale, same layout
template <int NC>
__device__ __forceinline__ void ge_merge(float (&acc)[NC][8], float* red, float* dst, int d, float scale) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nch = d >> 3;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NC; ++k)
        if (lane + 32 * k < nch)
#pragma unroll
            for (int i = 0; i < 8; ++i) red[(size_t)warp * d + i * nch + lane + 32 * k] = acc[k][i];       // chunk-transposed: lanes hit consecutive words
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += ge::kThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < ge::kWarps; ++w) s += red[(size_t)w * d + c];
        dst[c] = s * scale;
    }
    __syncthreads();
}

__device__ __forceinline__ float ge_block_sum(float v, float* scratch) {      // all threads get the sum
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < ge::kWarps; ++i) t += scratch[i];
    return t;
}

// shared memory (floats): mean[d] | vecb[d] | red[8][d] | rowv[4][L] | scratch[32] | stage
template <typename T, int NC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ge::kThreads) k_global_fwd(const __grid_constant__ GlobalArgs a) {
    using namespace ge;
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int side = (int)cluster.block_rank();
    const int b = blockIdx.x >> 1;
    const GlobalSide& S = a.s[side];
    const int L = S.L, d = a.d, nch = d >> 3;
    const int Lmax = max(a.s[0].L, a.s[1].L);
    float* mean = sm;
    float* vecb = mean + d;                  // text CTA: the image side's vector lands here (DSMEM)
    float* red = vecb + d;
    float* rowv = red + (size_t)kWarps * d;  // nx[L], cs[L]
    float* scratch = rowv + 4 * Lmax;
    void* stage = scratch + 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* xb = ge_stage<T>(static_cast<const T*>(S.x) + (size_t)b * L * d, L, d, stage, a.staged[side] != 0);
    float* nx = rowv;
    float* cs = rowv + Lmax;

    // pass 1: column sums and row norms
    float acc[NC][8];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    for (int l = warp; l < L; l += kWarps) {
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc[k][i] += v[i]; q = fmaf(v[i], v[i], q); }
            }
        }
        q = warp_sum(q);
        if (lane == 0) nx[l] = sqrtf(q);
    }
    ge_merge<NC>(acc, red, mean, d, 1.0f / L);
    float part = 0.f;
    for (int c = threadIdx.x; c < d; c += kThreads) part = fmaf(mean[c], mean[c], part);
    const float nm = fmaxf(sqrtf(ge_block_sum(part, scratch)), kCosEps);
    // pass 2: cosine to the mean
    for (int l = warp; l < L; l += kWarps) {
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) dot = fmaf(v[i], mean[i * nch + ch], dot);
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) cs[l] = dot / (fmaxf(nx[l], kCosEps) * nm) / a.temperature;
    }
    __syncthreads();
    // softmax over the L rows (warp 0)
    if (warp == 0) {
        float mx = -INFINITY;
        for (int l = lane; l < L; l += 32) mx = fmaxf(mx, cs[l]);
        mx = warp_max(mx);
        float den = 0.f;
        for (int l = lane; l < L; l += 32) den += expf(cs[l] - mx);
        den = warp_sum(den);
        for (int l = lane; l < L; l += 32) {
            const float w = expf(cs[l] - mx) / den;
            cs[l] = w;
            if (S.w) S.w[(size_t)b * L + l] = w;
        }
    }
    __syncthreads();
    // pass 3: weighted column sum
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    for (int l = warp; l < L; l += kWarps) {
        const float w = cs[l];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(w, v[i], acc[k][i]);
            }
        }
    }
    ge_merge<NC>(acc, red, mean, d, 1.0f);           // mean[] now holds this side's global embedding
    // combine: image side -> text CTA's vecb
    if (side == 1) {
        float* remote = cluster.map_shared_rank(vecb, 0);
        for (int c = threadIdx.x; c < d; c += kThreads) remote[c] = mean[c];
    }
    cluster.sync();
    if (side == 1) return;
    bool bad = false;
    T* out = static_cast<T*>(a.out) + (size_t)b * d;
    for (int c = threadIdx.x; c < d; c += kThreads) {
        const int t = (c & 7) * nch + (c >> 3);              // chunk-transposed position of column c
        const float v = 0.5f * mean[t] + 0.5f * vecb[t];
        bad |= !isfinite(v);
        out[c] = from_f32<T>(v);
    }
    const int any_bad = __syncthreads_or(bad);
    __shared__ int last;
    if (threadIdx.x == 0) {
        a.flags[b] = any_bad;
        __threadfence();
        last = atomicAdd(&a.flags[a.B], 1) == a.B - 1;
    }
    __syncthreads();
    if (!last) return;
    // the last text CTA: every element's output and flag is visible; zero everything if any flag is set
    __threadfence();
    int any = 0;
    for (int i = threadIdx.x; i < a.B; i += kThreads) any |= __ldcg(a.flags + i);
    any = __syncthreads_or(any);
    if (any) {
        T* all = static_cast<T*>(a.out);
        for (size_t i = threadIdx.x; i < (size_t)a.B * d; i += kThreads) all[i] = from_f32<T>(0.f);
    }
    if (threadIdx.x == 0) a.flags[a.B] = 0;          // counter ready for the next launch
}

// backward of one side: x [L, d], w [L] (saved softmax weights), dg = 0.5 * d_global (0 if guarded)
//   dc_l = w_l / T (dw_l - sum w dw), dw_l = x_l . dg;  e_l = dc_l / (nx_l nm);  f_l = dc_l cos_l / nx_l^2 (0 if clamped)
//   dmean = (sum_l e_l x_l - mean S2) / L,  S2 = sum_l dc_l cos_l / nm^2 (0 if the mean's norm is clamped)
//   dx_l = w_l dg + e_l mean - f_l x_l + dmean
template <typename T, int NC>
__global__ void __launch_bounds__(ge::kThreads) k_global_bwd(const __grid_constant__ GlobalArgs a) {
    using namespace ge;
    extern __shared__ __align__(16) float sm[];
    const int side = blockIdx.x & 1, b = blockIdx.x >> 1;
    const GlobalSide& S = a.s[side];
    const int L = S.L, d = a.d, nch = d >> 3;
    const int Lmax = max(a.s[0].L, a.s[1].L);
    float* mean = sm;
    float* dmean = mean + d;
    float* red = dmean + d;
    float* rowv = red + (size_t)kWarps * d;  // nx[L] (negative = clamped), cosv[L], dcs[L], wl[L]
    float* scratch = rowv + 4 * Lmax;
    void* stage = scratch + 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* xb = ge_stage<T>(static_cast<const T*>(S.x) + (size_t)b * L * d, L, d, stage, a.staged[side] != 0);
    float* nx = rowv; float* cosv = rowv + Lmax; float* dcs = rowv + 2 * Lmax; float* wl = rowv + 3 * Lmax;
    int guarded = 0;
    for (int i = threadIdx.x; i < a.B; i += kThreads) guarded |= a.flags[i];
    guarded = __syncthreads_or(guarded);
    const float gscale = guarded ? 0.f : 0.5f;
    const T* dgp = static_cast<const T*>(a.dout) + (size_t)b * d;
    float dg[NC][8];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nch) {
            load8<T>(dgp + ch * 8, dg[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) dg[k][i] *= gscale;
        }
    }
    for (int l = threadIdx.x; l < L; l += kThreads) wl[l] = S.w[(size_t)b * L + l];

    float acc[NC][8];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    for (int l = warp; l < L; l += kWarps) {
        float q = 0.f, dw = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc[k][i] += v[i]; q = fmaf(v[i], v[i], q); dw = fmaf(v[i], dg[k][i], dw); }
            }
        }
        q = warp_sum(q); dw = warp_sum(dw);
        if (lane == 0) { const float r = sqrtf(q); nx[l] = r < kCosEps ? -kCosEps : r; dcs[l] = dw; }
    }
    ge_merge<NC>(acc, red, mean, d, 1.0f / L);
    float part = 0.f;
    for (int c = threadIdx.x; c < d; c += kThreads) part = fmaf(mean[c], mean[c], part);
    const float nm_raw = sqrtf(ge_block_sum(part, scratch));
    const float nm = fmaxf(nm_raw, kCosEps);
    const bool m_clamped = nm_raw < kCosEps;
    for (int l = warp; l < L; l += kWarps) {
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) dot = fmaf(v[i], mean[i * nch + ch], dot);
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) cosv[l] = dot / (fabsf(nx[l]) * nm);
    }
    __syncthreads();
    float sp = 0.f;
    for (int l = threadIdx.x; l < L; l += kThreads) sp = fmaf(wl[l], dcs[l], sp);
    const float sdot = ge_block_sum(sp, scratch);
    float s2p = 0.f;
    // per-row coefficients; el -> nx[], fl -> cosv[] (reuse)
    for (int l = threadIdx.x; l < L; l += kThreads) {
        const float dc = wl[l] / a.temperature * (dcs[l] - sdot);
        const float n = fabsf(nx[l]);
        const bool row_ok = nx[l] > 0.f;
        const float cl = cosv[l];
        s2p += m_clamped ? 0.f : dc * cl / (nm * nm);
        nx[l] = dc / (n * nm);
        cosv[l] = row_ok ? dc * cl / (n * n) : 0.f;
    }
    const float S2 = ge_block_sum(s2p, scratch);
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    for (int l = warp; l < L; l += kWarps) {
        const float e = nx[l];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(e, v[i], acc[k][i]);
            }
        }
    }
    ge_merge<NC>(acc, red, dmean, d, 1.0f);
    for (int c = threadIdx.x; c < d; c += kThreads) dmean[c] = (dmean[c] - mean[c] * S2) / L;
    __syncthreads();
    T* dxb = static_cast<T*>(S.dx) + (size_t)b * L * d;
    for (int l = warp; l < L; l += kWarps) {
        const float w = wl[l], e = nx[l], f = cosv[l];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float v[8], o[8];
                load8<T>(xb + (size_t)l * d + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(w, dg[k][i], fmaf(e, mean[i * nch + ch], fmaf(-f, v[i], dmean[i * nch + ch])));
                store8<T>(dxb + (size_t)l * d + ch * 8, o);
            }
        }
    }
}

static size_t ge_smem_base(int d, int Lmax) { return ((size_t)2 * d + (size_t)ge::kWarps * d + 4 * Lmax + 32) * sizeof(float) + 16; }

template <typename T, int NC>
static int launch_ge(const GlobalArgs& a, size_t smem, bool bwd, cudaStream_t st) {
    if (bwd) {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_global_bwd<T, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_global_bwd<T, NC><<<2 * a.B, ge::kThreads, smem, st>>>(a);
    } else {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_global_fwd<T, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_global_fwd<T, NC><<<2 * a.B, ge::kThreads, smem, st>>>(a);
    }
    DCLIP_LAUNCH_OK("k_global_embedding");
    return DCLIP_OK;
}
template <typename T>
static int launch_ge_nc(const GlobalArgs& a, size_t smem, bool bwd, cudaStream_t st) {
    switch ((a.d + 255) / 256) {
        case 1: return launch_ge<T, 1>(a, smem, bwd, st);
        case 2: return launch_ge<T, 2>(a, smem, bwd, st);
        case 3: return launch_ge<T, 3>(a, smem, bwd, st);
        default: return launch_ge<T, 4>(a, smem, bwd, st);
    }
}

static int global_embedding(GlobalArgs a, int dtype, bool bwd, cudaStream_t st) {
    const int Lmax = max(a.s[0].L, a.s[1].L);
    const size_t base = ge_smem_base(a.d, Lmax);
    size_t stage = 0;
    for (int s = 0; s < 2; ++s) {
        const size_t bytes = (size_t)a.s[s].L * a.d * dtype_size(dtype);
        a.staged[s] = (bytes % 16 == 0 && aligned16(a.s[s].x) && base + bytes <= 220 * 1024) ? 1 : 0;
        if (a.staged[s]) stage = max(stage, bytes);
    }
    const size_t smem = base + stage;
    DCLIP_DISPATCH_DTYPE(dtype, return launch_ge_nc<T>(a, smem, bwd, st));
    return DCLIP_OK;
}

}  // namespace dclip

using namespace dclip;

extern "C" {

size_t dclip_global_embedding_flags_bytes(int64_t B) { return (size_t)(B + 1) * sizeof(int); }

int dclip_global_embedding_fwd(const void* attended_text, const void* attended_image, int64_t B, int64_t T, int64_t R, int64_t d, int dtype,
                               float temperature, void* out, float* w_text, float* w_image, int* flags, void* stream) {
    DCLIP_REQUIRE(attended_text && attended_image && out && flags && B > 0 && T > 0 && R > 0 && d > 0 && temperature > 0.f,
                  "global_embedding_fwd: bad arguments");
    DCLIP_REQUIRE(d % 8 == 0 && d <= 1024 && aligned16(attended_text) && aligned16(attended_image) && aligned16(out),
                  "global_embedding_fwd: d must be a multiple of 8, <= 1024, pointers 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_CUDA_OK(cudaMemsetAsync(flags + B, 0, sizeof(int), st));       // completion counter (the flags themselves are always written)
    GlobalArgs a{};
    a.s[0] = GlobalSide{attended_text, (int)T, w_text, nullptr};
    a.s[1] = GlobalSide{attended_image, (int)R, w_image, nullptr};
    a.B = (int)B; a.d = (int)d; a.temperature = temperature; a.out = out; a.flags = flags;
    return global_embedding(a, dtype, false, st);
}

int dclip_global_embedding_bwd(const void* attended_text, const void* attended_image, const float* w_text, const float* w_image,
                               const void* d_out, const int* flags, int64_t B, int64_t T, int64_t R, int64_t d, int dtype, float temperature,
                               void* d_attended_text, void* d_attended_image, void* stream) {
    DCLIP_REQUIRE(attended_text && attended_image && w_text && w_image && d_out && flags && d_attended_text && d_attended_image,
                  "global_embedding_bwd: null pointer");
    DCLIP_REQUIRE(B > 0 && T > 0 && R > 0 && d % 8 == 0 && d <= 1024 && temperature > 0.f, "global_embedding_bwd: bad shape");
    DCLIP_REQUIRE(aligned16(attended_text) && aligned16(attended_image) && aligned16(d_out) && aligned16(d_attended_text) && aligned16(d_attended_image),
                  "global_embedding_bwd: pointers must be 16-byte aligned");
    GlobalArgs a{};
    a.s[0] = GlobalSide{attended_text, (int)T, const_cast<float*>(w_text), d_attended_text};
    a.s[1] = GlobalSide{attended_image, (int)R, const_cast<float*>(w_image), d_attended_image};
    a.B = (int)B; a.d = (int)d; a.temperature = temperature; a.dout = d_out; a.flags = const_cast<int*>(flags);
    return global_embedding(a, dtype, true, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
