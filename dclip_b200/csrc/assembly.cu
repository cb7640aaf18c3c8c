// Ragged batch assembly feeding the cross-attention — SURVEY.md 8(f) row 2.
//
// Replaces the per-sample Python loops that sit immediately upstream of CrossModalAttention.forward:
//   * zero-pad / truncate / stack of the per-sample region embeddings to [B, max_patches, d] and of the per-sample
//     token embeddings to [B, max_tokens, d]               training/patch_text_aggregation.py:555-620
//     (padding rows are EXACT zeros and stay attended: no mask exists; a sample without boxes is one zero row :489-491;
//      a sample whose text embedding holds a NaN / Inf is replaced by zeros as a whole :542-544)
//   * the token filter + per-token `text_projection` loop  training/text_tokenizer.py:196-213
//     (keep token i iff token_id > 0, 0 < i < L - 1 and attention_mask == 1; with no token left the sample falls
//      back to its pooled sentence embedding)
// as a varlen (cu_seqlens) layout: the projection runs as ONE GEMM over all B x L hidden rows (dclip_linear_rows:
// tcgen05 for 16-bit inputs, the exact fp32 GEMM otherwise - in xattn.cu) and a single gather kernel writes the
// padded batch.  Pure data movement is bit-exact; the kernels are HBM-bound (16-byte accesses, one warp per row).
#include "common.cuh"

namespace dclip {

// ---- token filter: one warp per sample, ballot compaction in token order ----------------------------------------
// row_index[cu[b] + t] = b * L + (position of the t-th kept token);  lengths[b] = kept tokens (0 -> fallback row)
__global__ void __launch_bounds__(128) k_token_select(const long long* __restrict__ ids, const long long* __restrict__ mask, int B, int L,
                                                      int* __restrict__ lengths, int* __restrict__ positions) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    int kept = 0;
    for (int i0 = 0; i0 < L; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        if (i < L) keep = ids[(size_t)b * L + i] > 0 && i > 0 && i < L - 1 && mask[(size_t)b * L + i] == 1;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) positions[(size_t)b * L + kept + __popc(m & ((1u << lane) - 1))] = i;
        kept += __popc(m);
    }
    if (lane == 0) lengths[b] = kept;
}

// exclusive scan of max(lengths, min_len) -> cu[B + 1] and the batch maximum; one block (B is a batch size)
__global__ void __launch_bounds__(1024) k_cu_seqlens(const int* __restrict__ lengths, int B, int min_len, int* __restrict__ cu,
                                                     int* __restrict__ max_len) {
    __shared__ int warp_tot[32];
    __shared__ int carry, best;
    if (threadIdx.x == 0) { carry = 0; best = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = b < B ? max(lengths[b], min_len) : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[w] = x;
        int mx = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) atomicMax(&best, mx);
        __syncthreads();
        int off = carry;
        for (int i = 0; i < w; ++i) off += warp_tot[i];
        if (b < B) cu[b] = off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { cu[B] = carry; if (max_len) *max_len = best; }
}

// ---- per-sample non-finite guard (reference :542-544): flag[b] = 1 if any element of sample b's rows is NaN / Inf --
template <typename T>
__global__ void __launch_bounds__(256) k_sample_nonfinite(const T* __restrict__ src, const int* __restrict__ cu, const int* __restrict__ lengths,
                                                          const int* __restrict__ positions, int L_src, int d, int* __restrict__ flag) {
    const int b = blockIdx.x;
    const int n = lengths ? lengths[b] : cu[b + 1] - cu[b];
    bool bad = false;
    for (long long e = threadIdx.x; e < (long long)n * d; e += 256) {
        const int t = (int)(e / d), c = (int)(e - (long long)t * d);
        const long long row = positions ? (long long)b * L_src + positions[(size_t)b * L_src + t] : (long long)cu[b] + t;
        const float v = to_f32<T>(src[row * d + c]);
        bad |= !isfinite(v);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) flag[b] = 1;
}

// ---- gather + zero-pad: out[b, t, :] = src[row(b, t), :] for t < min(len_b, max_len), exact zeros otherwise --------
//  positions == null : flat varlen source, row(b, t) = cu[b] + t
//  positions != null : [B, L_src] source (e.g. the projected hidden states), row(b, t) = b * L_src + positions[b][t]
//  lengths[b] == 0 with a fallback: row 0 of the sample = fallback[b] (the pooled sentence embedding)
template <typename T>
__global__ void __launch_bounds__(256) k_pad_gather(const T* __restrict__ src, const int* __restrict__ cu, const int* __restrict__ lengths,
                                                    const int* __restrict__ positions, int L_src, const T* __restrict__ fallback,
                                                    const int* __restrict__ zero_flag, int B, int max_len, int d, T* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B * max_len) return;
    const int b = row / max_len, t = row - b * max_len;
    int len = lengths ? lengths[b] : cu[b + 1] - cu[b];
    const T* s = nullptr;
    if (t < len) {
        const long long r = positions ? (long long)b * L_src + positions[(size_t)b * L_src + t] : (long long)cu[b] + t;
        s = src + r * d;
    } else if (len == 0 && t == 0 && fallback) {
        s = fallback + (size_t)b * d;
    }
    if (zero_flag && zero_flag[b]) s = nullptr;
    T* o = out + (size_t)row * d;
    const int per16 = 16 / (int)sizeof(T);
    if ((d % per16) == 0 && ((reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(s)) & 15) == 0) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int c = lane; c < d / per16; c += 32)
            reinterpret_cast<uint4*>(o)[c] = s ? __ldg(reinterpret_cast<const uint4*>(s) + c) : z;
    } else {
        for (int c = lane; c < d; c += 32) o[c] = s ? s[c] : from_f32<T>(0.f);
    }
}

}  // namespace dclip

using namespace dclip;

extern "C" {

int dclip_token_select(const int64_t* input_ids, const int64_t* attention_mask, int64_t B, int64_t L, int* lengths, int* positions,
                       void* stream) {
    DCLIP_REQUIRE(input_ids && attention_mask && lengths && positions && B > 0 && L > 0, "token_select: bad arguments");
    k_token_select<<<(unsigned)ceil_div(B, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(input_ids), reinterpret_cast<const long long*>(attention_mask), (int)B, (int)L, lengths, positions);
    DCLIP_LAUNCH_OK("k_token_select");
    return DCLIP_OK;
}

int dclip_cu_seqlens(const int* lengths, int64_t B, int min_len, int* cu_seqlens, int* max_len, void* stream) {
    DCLIP_REQUIRE(lengths && cu_seqlens && B > 0 && min_len >= 0, "cu_seqlens: bad arguments");
    k_cu_seqlens<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(lengths, (int)B, min_len, cu_seqlens, max_len);
    DCLIP_LAUNCH_OK("k_cu_seqlens");
    return DCLIP_OK;
}

int dclip_pad_gather(const void* src, int dtype, const int* cu_seqlens, const int* lengths, const int* positions, int64_t L_src,
                     const void* fallback, int nan_guard, int* sample_flags, int64_t B, int64_t max_len, int64_t d, void* out, void* stream) {
    DCLIP_REQUIRE(src && out && B > 0 && max_len > 0 && d > 0, "pad_gather: bad arguments");
    DCLIP_REQUIRE(cu_seqlens || (lengths && positions), "pad_gather: give cu_seqlens (flat source) or lengths + positions ([B, L_src] source)");
    DCLIP_REQUIRE(!positions || (L_src > 0 && lengths), "pad_gather: positions need L_src and lengths");
    DCLIP_REQUIRE(!nan_guard || sample_flags, "pad_gather: the non-finite guard needs sample_flags[B]");
    DCLIP_REQUIRE(B * max_len < (1ll << 31), "pad_gather: batch too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nan_guard) DCLIP_CUDA_OK(cudaMemsetAsync(sample_flags, 0, (size_t)B * sizeof(int), st));
    DCLIP_DISPATCH_DTYPE(dtype, {
        if (nan_guard)
            k_sample_nonfinite<T><<<(unsigned)B, 256, 0, st>>>(static_cast<const T*>(src), cu_seqlens, lengths, positions, (int)L_src, (int)d, sample_flags);
        k_pad_gather<T><<<(unsigned)ceil_div(B * max_len, 8), 256, 0, st>>>(static_cast<const T*>(src), cu_seqlens, lengths, positions, (int)L_src,
                                                                            static_cast<const T*>(fallback), nan_guard ? sample_flags : nullptr,
                                                                            (int)B, (int)max_len, (int)d, static_cast<T*>(out));
    });
    DCLIP_LAUNCH_OK("k_pad_gather");
    return DCLIP_OK;
}

}  // extern "C"
