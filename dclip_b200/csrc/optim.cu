// Fused multi-tensor Adam / AdamW with global-norm gradient clipping — SURVEY.md 8(f) row 3.
//
// Replaces, for the teacher step, `torch.optim.Adam(trainable_params, lr=1e-5)` + `optimizer.step()`
// (training/train_contrastive_teacher.py:245-248, :360-362: 12 tensors, one foreach launch group per tensor) and,
// for the distillation step, `AdamW(self.parameters(), lr)` + Lightning's `gradient_clip_val=0.5`
// (training/CLIP_image_distillation.py:679-682, training/CLIP_image_distill_training.py:41-42).
//
// Two launches for ANY number of tensors (one without clipping), graph-capturable, no host synchronisation:
//   k_grad_sumsq   every block reduces one chunk of one tensor to a partial sum of squares (fp64 accumulation of
//                  fp32 squares); the last block to finish adds the partials IN CHUNK ORDER (bit-reproducible) and
//                  publishes  clip_coef = min(1, max_norm / (sqrt(total) + 1e-6))  - torch.nn.utils.clip_grad_norm_
//   k_adam_step    p, m, v updated in one pass (HBM-bound: reads g, p, m, v; writes p, m, v); the step counter lives
//                  in device memory and is advanced by the kernel, bias corrections are computed in double like
//                  torch's Python scalars, so a captured graph replays correct steps.
// Update rule (torch.optim.Adam / AdamW, amsgrad = False, maximize = False):
//   g' = clip_coef * g (+ weight_decay * p   for Adam's L2 form)
//   p  = p * (1 - lr * weight_decay)          (AdamW's decoupled form)
//   m  = b1 m + (1 - b1) g' ;  v = b2 v + (1 - b2) g'^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Parameters and gradients may be fp32, bf16 or fp16; the moments are always fp32.
#include "common.cuh"

namespace dclip {

constexpr int kOptMaxTensors = 64;
constexpr int kOptChunk = 8192;          // elements per block
constexpr int kOptThreads = 256;

struct OptTensors {
    void* p[kOptMaxTensors];
    const void* g[kOptMaxTensors];
    float* m[kOptMaxTensors];
    float* v[kOptMaxTensors];
    long long n[kOptMaxTensors];
    int chunk_begin[kOptMaxTensors + 1];  // prefix sum of ceil(n / kOptChunk)
    int count;
};

__device__ __forceinline__ int find_tensor(const OptTensors& t, int chunk) {
    int lo = 0, hi = t.count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.chunk_begin[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, long long i, long long n, float (&out)[4], bool vec) {
    if (vec) {
        if constexpr (sizeof(T) == 4) {
            const float4 x = *reinterpret_cast<const float4*>(p + i);
            out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = x.w;
        } else {
            const uint2 raw = *reinterpret_cast<const uint2*>(p + i);
            const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = to_f32<T>(e[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = i + j < n ? to_f32<T>(p[i + j]) : 0.f;
    }
}
template <typename T>
__device__ __forceinline__ void store4(T* p, long long i, long long n, const float (&in)[4], bool vec) {
    if (vec) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(p + i) = make_float4(in[0], in[1], in[2], in[3]);
        } else {
            uint2 raw;
            T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = from_f32<T>(in[j]);
            *reinterpret_cast<uint2*>(p + i) = raw;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i + j < n) p[i + j] = from_f32<T>(in[j]);
    }
}

// device-resident optimizer state shared by the two kernels
struct OptState {
    double total_sumsq;        // last computed sum of squares of all gradients
    float clip_coef;           // multiplies every gradient in k_adam_step (1 when clipping is off)
    float grad_norm;           // sqrt(total_sumsq), for logging (clip_grad_norm_'s return value)
    int step;                  // number of optimizer steps taken so far
    unsigned int done_ctr;     // last-block detection
};

template <typename T>
__global__ void __launch_bounds__(kOptThreads) k_grad_sumsq(const __grid_constant__ OptTensors t, double* __restrict__ partial, int nchunks,
                                                            float max_norm, OptState* st) {
    const int chunk = blockIdx.x;
    const int ti = find_tensor(t, chunk);
    const long long n = t.n[ti];
    const long long base = (long long)(chunk - t.chunk_begin[ti]) * kOptChunk;
    const T* g = static_cast<const T*>(t.g[ti]);
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    double acc = 0.0;
    for (int it = 0; it < kOptChunk / (kOptThreads * 4); ++it) {
        const long long i = base + ((long long)it * kOptThreads + threadIdx.x) * 4;
        if (i >= n) break;
        float x[4];
        load4<T>(g, i, n, x, vec && i + 4 <= n);
        acc += (double)x[0] * x[0] + (double)x[1] * x[1] + (double)x[2] * x[2] + (double)x[3] * x[3];
    }
    __shared__ double red[kOptThreads / 32];
    __shared__ bool last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < kOptThreads / 32; ++i) s += red[i];
        partial[chunk] = s;
        __threadfence();
        last = atomicAdd(&st->done_ctr, 1u) == (unsigned)nchunks - 1;
    }
    __syncthreads();
    if (!last) return;
    // last block: fixed-order sum of the chunk partials (every thread a strided slice, then a fixed tree)
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < nchunks; i += kOptThreads) s += __ldcg(partial + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < kOptThreads / 32; ++i) tot += red[i];
        const float norm = (float)sqrt(tot);
        st->total_sumsq = tot;
        st->grad_norm = norm;
        // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1.0
        st->clip_coef = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
        st->done_ctr = 0;
    }
}

// hyper-parameters stay in double until the point where torch casts its Python scalars to the tensor dtype
struct AdamHyper { double lr, beta1, beta2, eps, weight_decay; int decoupled; const float* lr_dev; };

template <typename T>
__global__ void __launch_bounds__(kOptThreads) k_adam_step(const __grid_constant__ OptTensors t, const AdamHyper h, OptState* st, int nchunks) {
    const int chunk = blockIdx.x;
    const int ti = find_tensor(t, chunk);
    const long long n = t.n[ti];
    const long long base = (long long)(chunk - t.chunk_begin[ti]) * kOptChunk;
    T* p = static_cast<T*>(t.p[ti]);
    const T* g = static_cast<const T*>(t.g[ti]);
    float* m = t.m[ti];
    float* v = t.v[ti];
    const bool vecT = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    const bool vecF = ((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    // scalars: the step this launch performs is st->step + 1 (every block reads the same, the last block advances it)
    const int step = st->step + 1;
    const double lr = h.lr_dev ? (double)*h.lr_dev : h.lr;
    const double bc1 = 1.0 - pow(h.beta1, (double)step);
    const double bc2 = 1.0 - pow(h.beta2, (double)step);
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float coef = st->clip_coef;
    const float decay = h.decoupled ? (float)(1.0 - lr * h.weight_decay) : 1.0f;
    const float l2 = h.decoupled ? 0.f : (float)h.weight_decay;
    const float b2 = (float)h.beta2, omb1 = (float)(1.0 - h.beta1), omb2 = (float)(1.0 - h.beta2), eps = (float)h.eps;
    for (int it = 0; it < kOptChunk / (kOptThreads * 4); ++it) {
        const long long i = base + ((long long)it * kOptThreads + threadIdx.x) * 4;
        if (i >= n) break;
        const bool full = i + 4 <= n;
        float pv[4], gv[4], mv[4], vv[4];
        load4<T>(p, i, n, pv, vecT && full);
        load4<T>(g, i, n, gv, vecT && full);
        load4<float>(m, i, n, mv, vecF && full);
        load4<float>(v, i, n, vv, vecF && full);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gj = gv[j] * coef;
            float pj = pv[j] * decay;
            gj = fmaf(l2, pv[j], gj);
            mv[j] = fmaf(omb1, gj - mv[j], mv[j]);                        // exp_avg.lerp_(grad, 1 - beta1)
            vv[j] = fmaf(omb2 * gj, gj, b2 * vv[j]);                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
            const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pv[j] = pj - step_size * (mv[j] / denom);
        }
        store4<T>(p, i, n, pv, vecT && full);
        store4<float>(m, i, n, mv, vecF && full);
        store4<float>(v, i, n, vv, vecF && full);
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&st->done_ctr, 1u) == (unsigned)nchunks - 1;
        if (last) { st->step = step; st->done_ctr = 0; }
    }
}

static int fill_tensors(OptTensors& t, void* const* params, const void* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                        const int64_t* numel, int count, int begin, int n) {
    t.count = n;
    int chunks = 0;
    for (int i = 0; i < n; ++i) {
        const int s = begin + i;
        if (s >= count) return -1;
        t.p[i] = params ? params[s] : nullptr;
        t.g[i] = grads[s];
        t.m[i] = exp_avg ? exp_avg[s] : nullptr;
        t.v[i] = exp_avg_sq ? exp_avg_sq[s] : nullptr;
        t.n[i] = numel[s];
        t.chunk_begin[i] = chunks;
        chunks += (int)ceil_div(numel[s], kOptChunk);
    }
    t.chunk_begin[n] = chunks;
    return chunks;
}

}  // namespace dclip

using namespace dclip;

extern "C" {

size_t dclip_optim_state_bytes(void) { return align_up(sizeof(OptState), 256); }

size_t dclip_optim_workspace_bytes(const int64_t* numel, int count) {
    int64_t chunks = 0;
    for (int i = 0; i < count; ++i) chunks += ceil_div(numel[i], kOptChunk);
    return align_up((size_t)chunks * sizeof(double), 256) + 256;
}

int dclip_optim_init_state(void* state, void* stream) {
    DCLIP_REQUIRE(state, "optim_init_state: null pointer");
    OptState s{};
    s.clip_coef = 1.0f;
    // pageable-host -> device copy of a 32-byte struct: staged by the runtime before the call returns
    DCLIP_CUDA_OK(cudaMemcpyAsync(state, &s, sizeof(s), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return DCLIP_OK;
}

int dclip_grad_clip_coef(const void* const* grads, const int64_t* numel, int count, int dtype, float max_norm, void* state,
                         void* workspace, size_t workspace_bytes, void* stream) {
    DCLIP_REQUIRE(grads && numel && state && workspace && count > 0 && count <= kOptMaxTensors, "grad_clip_coef: 1..%d tensors", kOptMaxTensors);
    if (workspace_bytes < dclip_optim_workspace_bytes(numel, count)) return fail(DCLIP_ERR_WORKSPACE, "grad_clip_coef: workspace too small");
    for (int i = 0; i < count; ++i) DCLIP_REQUIRE(grads[i] && numel[i] > 0, "grad_clip_coef: tensor %d is null or empty", i);
    OptTensors t{};
    const int chunks = fill_tensors(t, nullptr, grads, nullptr, nullptr, numel, count, 0, count);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_DISPATCH_DTYPE(dtype, (k_grad_sumsq<T><<<chunks, kOptThreads, 0, st>>>(t, static_cast<double*>(workspace), chunks, max_norm,
                                                                               static_cast<OptState*>(state))));
    DCLIP_LAUNCH_OK("k_grad_sumsq");
    return DCLIP_OK;
}

int dclip_adam_step(void* const* params, const void* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                    int count, int dtype, double lr, const float* lr_dev, double beta1, double beta2, double eps, double weight_decay,
                    int decoupled_weight_decay, void* state, void* stream) {
    DCLIP_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && state && count > 0 && count <= kOptMaxTensors,
                  "adam_step: 1..%d tensors", kOptMaxTensors);
    DCLIP_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0 && weight_decay >= 0.0, "adam_step: bad hyper-parameters");
    for (int i = 0; i < count; ++i)
        DCLIP_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] > 0, "adam_step: tensor %d is null or empty", i);
    OptTensors t{};
    const int chunks = fill_tensors(t, params, grads, exp_avg, exp_avg_sq, numel, count, 0, count);
    AdamHyper h{lr, beta1, beta2, eps, weight_decay, decoupled_weight_decay, lr_dev};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_DISPATCH_DTYPE(dtype, (k_adam_step<T><<<chunks, kOptThreads, 0, st>>>(t, h, static_cast<OptState*>(state), chunks)));
    DCLIP_LAUNCH_OK("k_adam_step");
    return DCLIP_OK;
}

int dclip_optim_read_state(const void* state, float* grad_norm, float* clip_coef, int* step, void* stream) {
    DCLIP_REQUIRE(state, "optim_read_state: null pointer");
    OptState s{};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DCLIP_CUDA_OK(cudaMemcpyAsync(&s, state, sizeof(s), cudaMemcpyDeviceToHost, st));
    DCLIP_CUDA_OK(cudaStreamSynchronize(st));
    if (grad_norm) *grad_norm = s.grad_norm;
    if (clip_coef) *clip_coef = s.clip_coef;
    if (step) *step = s.step;
    return DCLIP_OK;
}

}  // extern "C"
