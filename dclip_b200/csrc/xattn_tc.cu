// Meta-teacher bidirectional cross-attention, 16-bit tier (bf16 / fp16 parameters and activations; BASELINE configs[2]):
// everything between the module inputs and outputs stays 16-bit in HBM, all contractions run on tcgen05.
//
//   forward  (training/patch_text_aggregation.py:21-46), 4 launches:
//     k_gemm_tc        in-projections of both directions, 4 problems, bias fused, 16-bit Q / K|V out          (K1)
//     k_attn_fwd_tc    softmax(Q K^T / sqrt(hd)) V per (batch, head, direction), scores in TMEM              (K2)
//     k_gemm_tc        out-projections, bias + residual (the 16-bit module input) fused, fp32 pre-LN sum Z    (K3)
//     k_ln_fwd16       LayerNorm of both sides, 16-bit module outputs + (mean, rstd)                           (K4)
//   backward (autograd of the same, training/train_contrastive_teacher.py:361), 5 launches (+2 for input gradients):
//     k_ln_bwd16       dZ (16-bit) + per-block column partials of dgamma, dbeta and the out-projection bias gradient
//     k_gemm_tc        dO = dZ Wo                                                                              2 problems
//     k_attn_bwd_tc    dQ, dK|dV (16-bit) + per-batch column partials of the in-projection bias gradients
//     k_gemm_tc        dWq = dQ^T X, dWkv = dKV^T X, dWo = dZ^T O (straight into the parameter gradients)      6 problems
//     k_vec_finalize   the eight small gradient vectors from their partials (fixed order: bit-reproducible)
// Statistics (softmax, LayerNorm, all partial sums) are fp32.  Nothing is widened to fp32 in HBM except Z.
#include "attn_tc.h"
#include "common.cuh"
#include "gemm_tc.h"

namespace dclip {

constexpr float kLnEps16 = 1e-5f;
constexpr int kLnRowsPerBlock = 16;       // backward: 8 warps x 2 rows (latency-bound: many small blocks beat few long ones)

// ---- LayerNorm forward: one warp per row; lane l owns the 8-element chunks l, l + 32, ... (NC = ceil(d / 256) of them) ----
struct LnSide { const float* Z; const void* gamma; const void* beta; void* Y; float* stats; int rows; };
struct LnFwdArgs { LnSide s[2]; int d, nblk0; };

template <typename T, int NC>
__global__ void __launch_bounds__(256) k_ln_fwd16(const __grid_constant__ LnFwdArgs a) {
    const int side = (int)blockIdx.x >= a.nblk0 ? 1 : 0;
    const LnSide& S = a.s[side];
    const int lane = threadIdx.x & 31;
    const int row = ((int)blockIdx.x - (side ? a.nblk0 : 0)) * 8 + (threadIdx.x >> 5);
    if (row >= S.rows) return;
    const int d = a.d, nch = d >> 3;
    const float* z = S.Z + (size_t)row * d;
    float v[NC][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nch) {
            load8<float>(z + ch * 8, v[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[k][i];
        }
    }
    const float mu = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k)
        if (lane + 32 * k < nch)
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float t = v[k][i] - mu; q = fmaf(t, t, q); }
    const float rstd = rsqrtf(warp_sum(q) / d + kLnEps16);
    T* y = static_cast<T*>(S.Y) + (size_t)row * d;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nch) {
            float g[8], b[8], o[8];
            load8<T>(static_cast<const T*>(S.gamma) + ch * 8, g);
            load8<T>(static_cast<const T*>(S.beta) + ch * 8, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaf((v[k][i] - mu) * rstd, g[i], b[i]);
            store8<T>(y + ch * 8, o);
        }
    }
    if (lane == 0) { S.stats[2 * row] = mu; S.stats[2 * row + 1] = rstd; }
}

// ---- LayerNorm backward: dz = rstd (dy g - mean(dy g) - zhat mean(dy g zhat)), 16-bit dZ out; each block leaves the column
// sums over its rows of dy (-> dbeta), dy zhat (-> dgamma) and dz (-> out-projection bias gradient) in part[block][3][d] ----
struct LnBwdSide { const void* dY; const float* Z; const void* gamma; const float* stats; void* dZ; int rows; };
struct LnBwdArgs { LnBwdSide s[2]; int d, nblk0; float* part; };

template <typename T, int NC>
__global__ void __launch_bounds__(256, 2) k_ln_bwd16(const __grid_constant__ LnBwdArgs a) {
    // The column sums of a warp's rows are accumulated in the warp's own slice of shared memory, not in registers (72 of them:
    // 171 registers = one block per SM = 226 blocks in two waves at the configs[2] shape; now two blocks per SM, one wave,
    // twice the rows in flight per SM for a kernel that is a chain of global-memory round trips).
    extern __shared__ float red[];                     // [8 warps][3][d], chunk-transposed columns
    const int side = (int)blockIdx.x >= a.nblk0 ? 1 : 0;
    const LnBwdSide& S = a.s[side];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = ((int)blockIdx.x - (side ? a.nblk0 : 0)) * kLnRowsPerBlock;
    const int d = a.d, nch = d >> 3;
    float* mine = red + (size_t)warp * 3 * d;
    bool first = true;
    for (int r = warp; r < kLnRowsPerBlock; r += 8) {
        const int row = row0 + r;
        if (row >= S.rows) break;
        const float mu = S.stats[2 * row], rstd = S.stats[2 * row + 1];
        const float* z = S.Z + (size_t)row * d;
        const T* dy = static_cast<const T*>(S.dY) + (size_t)row * d;
        float zh[NC][8], g[NC][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float dv[8], gam[8];
                load8<float>(z + ch * 8, zh[k]);
                load8<T>(dy + ch * 8, dv);
                load8<T>(static_cast<const T*>(S.gamma) + ch * 8, gam);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    zh[k][i] = (zh[k][i] - mu) * rstd;
                    g[k][i] = dv[i] * gam[i];
                    s1 += g[k][i];
                    s2 = fmaf(g[k][i], zh[k][i], s2);
                    float* p0 = mine + i * nch + ch;             // column sums of dy (-> dbeta) and dy zhat (-> dgamma)
                    const float c1 = dv[i] * zh[k][i];
                    if (first) { p0[0] = dv[i]; p0[d] = c1; } else { p0[0] += dv[i]; p0[d] += c1; }
                }
            }
        }
        s1 = warp_sum(s1) / d;
        s2 = warp_sum(s2) / d;
        T* dz = static_cast<T*>(S.dZ) + (size_t)row * d;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            if (ch < nch) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    o[i] = rstd * (g[k][i] - s1 - zh[k][i] * s2);
                    float* p2 = mine + 2 * d + i * nch + ch;     // column sums of dz (-> out-projection bias gradient)
                    if (first) p2[0] = o[i]; else p2[0] += o[i];
                }
                store8<T>(dz + ch * 8, o);
            }
        }
        first = false;
    }
    if (first)                                         // no row for this warp: its slice still takes part in the block sum
        for (int e = lane; e < 3 * d; e += 32) mine[e] = 0.f;
    __syncthreads();
    float* out = a.part + (size_t)blockIdx.x * 3 * d;
    for (int e = threadIdx.x; e < 3 * d; e += 256) {
        const int q = e / d, c = e - q * d;
        const int t = q * d + (c & 7) * nch + (c >> 3);      // chunk-transposed position of column c
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[(size_t)w * 3 * d + t];
        out[e] = s;
    }
}

// ---- small gradient vectors from their partials: dst[i] = sum_p src[p * stride + i], fixed order ----
constexpr int kFinMaxJobs = 12;
struct FinJob { const float* src; int nparts; long long stride; int n; void* dst; };
struct FinArgs { FinJob j[kFinMaxJobs]; int count; };
template <typename T>
__global__ void __launch_bounds__(256) k_vec_finalize(const __grid_constant__ FinArgs a) {
    // block = 32 columns x 8 part lanes: the loads of a column are spread over 8 threads (the loop is L2-latency bound),
    // merged in fixed order through shared memory
    __shared__ float red[8][33];
    const FinJob& J = a.j[blockIdx.y];
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (i < J.n) {
#pragma unroll 4
        for (int p = py; p < J.nparts; p += 8) s += J.src[(size_t)p * J.stride + i];
    }
    red[py][cx] = s;
    __syncthreads();
    if (py == 0 && i < J.n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][cx];
        static_cast<T*>(J.dst)[i] = from_f32<T>(t);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
struct XTcLayout {      // byte offsets
    size_t Qt, KVi, Qi, KVt, Ot, Oi, Zt, Zi, St, Si, total;
};
static XTcLayout xtc_layout(int64_t B, int64_t T, int64_t R, int64_t d) {
    XTcLayout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    const size_t nt = (size_t)B * T, ni = (size_t)B * R;
    L.Qt = take(nt * d * 2); L.KVi = take(ni * 2 * d * 2); L.Qi = take(ni * d * 2); L.KVt = take(nt * 2 * d * 2);
    L.Ot = take(nt * d * 2); L.Oi = take(ni * d * 2);
    L.Zt = take(nt * d * 4); L.Zi = take(ni * d * 4);
    L.St = take(nt * 2 * 4); L.Si = take(ni * 2 * 4);
    L.total = o;
    return L;
}
struct XTcScratch {     // byte offsets of the backward scratch
    size_t dZt, dZi, dOt, dOi, dQt, dQi, dKVi, dKVt, lnpart, dbq0, dbkv0, dbq1, dbkv1, dXt, dXi, total;
    int nblk_t, nblk_i;
};
static XTcScratch xtc_scratch(int64_t B, int64_t T, int64_t R, int64_t d) {
    XTcScratch S{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    const size_t nt = (size_t)B * T, ni = (size_t)B * R;
    S.nblk_t = (int)ceil_div((int64_t)nt, kLnRowsPerBlock);
    S.nblk_i = (int)ceil_div((int64_t)ni, kLnRowsPerBlock);
    S.dZt = take(nt * d * 2); S.dZi = take(ni * d * 2); S.dOt = take(nt * d * 2); S.dOi = take(ni * d * 2);
    S.dQt = take(nt * d * 2); S.dQi = take(ni * d * 2); S.dKVi = take(ni * 2 * d * 2); S.dKVt = take(nt * 2 * d * 2);
    S.lnpart = take((size_t)(S.nblk_t + S.nblk_i) * 3 * d * 4);
    S.dbq0 = take((size_t)B * d * 4); S.dbkv0 = take((size_t)B * 2 * d * 4); S.dbq1 = take((size_t)B * d * 4); S.dbkv1 = take((size_t)B * 2 * d * 4);
    S.dXt = take(nt * d * 4); S.dXi = take(ni * d * 4);
    S.total = o;
    return S;
}

bool xattn_tc_shape_ok(int64_t T, int64_t R, int64_t d, int heads) { return attn_tc_shape_ok(T, R, d, heads) && d <= 1024; }
size_t xattn_tc_saved_bytes(int64_t B, int64_t T, int64_t R, int64_t d) { return xtc_layout(B, T, R, d).total + 256; }
size_t xattn_tc_scratch_bytes(int64_t B, int64_t T, int64_t R, int64_t d) { return xtc_scratch(B, T, R, d).total + 256; }

static TcGemmProblem tcg(const void* A, int lda, bool a_mn, const void* Bm, int ldb, bool b_mn, int M, int N, int K) {
    TcGemmProblem p{};
    p.A = A; p.B = Bm; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = N; p.a_mn = a_mn; p.b_mn = b_mn;
    return p;
}
static TcGemmProblem out16(TcGemmProblem p, void* C16, int ldc16, const void* bias16 = nullptr) {
    p.C16 = C16; p.ldc16 = ldc16; p.bias = bias16; p.bias_kind = 1;
    return p;
}

template <typename T>
static int launch_ln_fwd(const LnFwdArgs& a, int nblk, cudaStream_t st) {
    const int nc = (a.d + 255) / 256;
    if (nc == 1) k_ln_fwd16<T, 1><<<nblk, 256, 0, st>>>(a);
    else if (nc == 2) k_ln_fwd16<T, 2><<<nblk, 256, 0, st>>>(a);
    else if (nc == 3) k_ln_fwd16<T, 3><<<nblk, 256, 0, st>>>(a);
    else k_ln_fwd16<T, 4><<<nblk, 256, 0, st>>>(a);
    DCLIP_LAUNCH_OK("k_ln_fwd16");
    return DCLIP_OK;
}
template <typename T, int NC>
static int launch_ln_bwd_nc(const LnBwdArgs& a, int nblk, cudaStream_t st) {
    const size_t smem = (size_t)8 * 3 * a.d * sizeof(float);
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_ln_bwd16<T, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ln_bwd16<T, NC><<<nblk, 256, smem, st>>>(a);
    DCLIP_LAUNCH_OK("k_ln_bwd16");
    return DCLIP_OK;
}
template <typename T>
static int launch_ln_bwd(const LnBwdArgs& a, int nblk, cudaStream_t st) {
    const int nc = (a.d + 255) / 256;
    if (nc == 1) return launch_ln_bwd_nc<T, 1>(a, nblk, st);
    if (nc == 2) return launch_ln_bwd_nc<T, 2>(a, nblk, st);
    if (nc == 3) return launch_ln_bwd_nc<T, 3>(a, nblk, st);
    return launch_ln_bwd_nc<T, 4>(a, nblk, st);
}

// params / grads: the reference's state-dict order (see xattn.cu), all 16-bit.
int xattn_fwd_tc(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype, const void* const* params,
                 void* text_out, void* image_out, void* saved, size_t saved_bytes, cudaStream_t st) {
    const XTcLayout L = xtc_layout(B, T, R, d);
    if (saved_bytes < L.total) return fail(DCLIP_ERR_WORKSPACE, "xattn_fwd: saved buffer too small (%zu < %zu)", saved_bytes, L.total);
    DCLIP_REQUIRE(aligned16(text) && aligned16(image) && aligned16(saved) && aligned16(text_out) && aligned16(image_out), "xattn_fwd: pointers must be 16-byte aligned");
    char* S = static_cast<char*>(saved);
    const int nt = (int)(B * T), ni = (int)(B * R), D = (int)d;
    const char* w_t2i = static_cast<const char*>(params[0]);
    const char* w_i2t = static_cast<const char*>(params[4]);
    const char* b_t2i = static_cast<const char*>(params[1]);
    const char* b_i2t = static_cast<const char*>(params[5]);
    const size_t dd2 = (size_t)d * d * 2;
    // (1) in-projections: Q_t, K|V_i (text_to_image) and Q_i, K|V_t (image_to_text)
    TcGemmProblem pr[4];
    pr[0] = out16(tcg(text, D, false, w_t2i, D, false, nt, D, D), S + L.Qt, D, b_t2i);
    pr[1] = out16(tcg(image, D, false, w_t2i + dd2, D, false, ni, 2 * D, D), S + L.KVi, 2 * D, b_t2i + (size_t)d * 2);
    pr[2] = out16(tcg(image, D, false, w_i2t, D, false, ni, D, D), S + L.Qi, D, b_i2t);
    pr[3] = out16(tcg(text, D, false, w_i2t + dd2, D, false, nt, 2 * D, D), S + L.KVt, 2 * D, b_i2t + (size_t)d * 2);
    if (int e = launch_gemm_tc(pr, 4, dtype, st)) return e;
    // (2) attention cores
    AttnTcProblem ap{};
    ap.B = (int)B; ap.T = (int)T; ap.R = (int)R; ap.d = D; ap.heads = heads; ap.dtype16 = dtype;
    ap.Qt = S + L.Qt; ap.KVi = S + L.KVi; ap.Qi = S + L.Qi; ap.KVt = S + L.KVt; ap.Ot = S + L.Ot; ap.Oi = S + L.Oi;
    if (int e = launch_attn_tc(ap, false, st)) return e;
    // (3) out-projection + bias + residual -> Z (fp32)
    TcGemmProblem po[2];
    po[0] = tcg(S + L.Ot, D, false, params[2], D, false, nt, D, D);
    po[0].C = reinterpret_cast<float*>(S + L.Zt); po[0].bias = params[3]; po[0].bias_kind = 1; po[0].residual = text; po[0].res_kind = 1; po[0].ldr = D;
    po[1] = tcg(S + L.Oi, D, false, params[6], D, false, ni, D, D);
    po[1].C = reinterpret_cast<float*>(S + L.Zi); po[1].bias = params[7]; po[1].bias_kind = 1; po[1].residual = image; po[1].res_kind = 1; po[1].ldr = D;
    if (int e = launch_gemm_tc(po, 2, dtype, st)) return e;
    // (4) LayerNorm, both sides
    LnFwdArgs la{};
    la.s[0] = LnSide{reinterpret_cast<const float*>(S + L.Zt), params[8], params[9], text_out, reinterpret_cast<float*>(S + L.St), nt};
    la.s[1] = LnSide{reinterpret_cast<const float*>(S + L.Zi), params[10], params[11], image_out, reinterpret_cast<float*>(S + L.Si), ni};
    la.d = D; la.nblk0 = (int)ceil_div(nt, 8);
    const int nblk = la.nblk0 + (int)ceil_div(ni, 8);
    if (dtype == DCLIP_BF16) return launch_ln_fwd<__nv_bfloat16>(la, nblk, st);
    return launch_ln_fwd<__half>(la, nblk, st);
}

int xattn_bwd_tc(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype, const void* const* params,
                 const void* saved, const void* d_text_out, const void* d_image_out, void* const* grads, void* d_text, void* d_image,
                 void* scratch, size_t scratch_bytes, cudaStream_t st) {
    const XTcLayout L = xtc_layout(B, T, R, d);
    const XTcScratch W = xtc_scratch(B, T, R, d);
    if (scratch_bytes < W.total) return fail(DCLIP_ERR_WORKSPACE, "xattn_bwd: scratch too small (%zu < %zu)", scratch_bytes, W.total);
    DCLIP_REQUIRE(aligned16(d_text_out) && aligned16(d_image_out) && aligned16(scratch), "xattn_bwd: pointers must be 16-byte aligned");
    for (int i = 0; i < 12; ++i) DCLIP_REQUIRE(grads[i] && aligned16(grads[i]), "xattn_bwd: gradient %d is null or misaligned", i);
    const char* S = static_cast<const char*>(saved);
    char* X = static_cast<char*>(scratch);
    const int nt = (int)(B * T), ni = (int)(B * R), D = (int)d;
    const size_t dd2 = (size_t)d * d * 2;
    // (4') LayerNorm backward
    LnBwdArgs lb{};
    lb.s[0] = LnBwdSide{d_text_out, reinterpret_cast<const float*>(S + L.Zt), params[8], reinterpret_cast<const float*>(S + L.St), X + W.dZt, nt};
    lb.s[1] = LnBwdSide{d_image_out, reinterpret_cast<const float*>(S + L.Zi), params[10], reinterpret_cast<const float*>(S + L.Si), X + W.dZi, ni};
    lb.d = D; lb.nblk0 = W.nblk_t; lb.part = reinterpret_cast<float*>(X + W.lnpart);
    if (dtype == DCLIP_BF16) { if (int e = launch_ln_bwd<__nv_bfloat16>(lb, W.nblk_t + W.nblk_i, st)) return e; }
    else { if (int e = launch_ln_bwd<__half>(lb, W.nblk_t + W.nblk_i, st)) return e; }
    // (3') out-projection backward: dO = dZ Wo.  The weight gradients dWo = dZ^T O are leaves of the dependency graph with a long
    // contraction (K = B T rows: a tile takes 3x as long as a dO tile and would set this launch's critical path), so they ride
    // with the in-projection weight gradients below, whose tiles are as long and leave SMs idle (108 of 148).
    TcGemmProblem pr[2];
    pr[0] = out16(tcg(X + W.dZt, D, false, params[2], D, true, nt, D, D), X + W.dOt, D);
    pr[1] = out16(tcg(X + W.dZi, D, false, params[6], D, true, ni, D, D), X + W.dOi, D);
    if (int e = launch_gemm_tc(pr, 2, dtype, st)) return e;
    // (2') attention cores backward (+ in-projection bias gradient partials per batch element)
    AttnTcProblem ap{};
    ap.B = (int)B; ap.T = (int)T; ap.R = (int)R; ap.d = D; ap.heads = heads; ap.dtype16 = dtype;
    ap.Qt = S + L.Qt; ap.KVi = S + L.KVi; ap.Qi = S + L.Qi; ap.KVt = S + L.KVt;
    ap.dOt = X + W.dOt; ap.dOi = X + W.dOi;
    ap.dQt = X + W.dQt; ap.dKVi = X + W.dKVi; ap.dQi = X + W.dQi; ap.dKVt = X + W.dKVt;
    ap.dbq_t2i = reinterpret_cast<float*>(X + W.dbq0); ap.dbkv_t2i = reinterpret_cast<float*>(X + W.dbkv0);
    ap.dbq_i2t = reinterpret_cast<float*>(X + W.dbq1); ap.dbkv_i2t = reinterpret_cast<float*>(X + W.dbkv1);
    if (int e = launch_attn_tc(ap, true, st)) return e;
    // (1') in-projection weight gradients (rows 0..d = q, d..3d = k|v), straight into the parameter gradients
    char* g_t2i = static_cast<char*>(grads[0]);
    char* g_i2t = static_cast<char*>(grads[4]);
    TcGemmProblem pw[6];
    pw[0] = out16(tcg(X + W.dQt, D, true, text, D, true, D, D, nt), g_t2i, D);
    pw[1] = out16(tcg(X + W.dKVt, 2 * D, true, text, D, true, 2 * D, D, nt), g_i2t + dd2, D);
    pw[2] = out16(tcg(X + W.dZt, D, true, S + L.Ot, D, true, D, D, nt), grads[2], D);              // dWo (text side)
    pw[3] = out16(tcg(X + W.dKVi, 2 * D, true, image, D, true, 2 * D, D, ni), g_t2i + dd2, D);
    pw[4] = out16(tcg(X + W.dQi, D, true, image, D, true, D, D, ni), g_i2t, D);
    pw[5] = out16(tcg(X + W.dZi, D, true, S + L.Oi, D, true, D, D, ni), grads[6], D);              // dWo (image side)
    if (int e = launch_gemm_tc(pw, 6, dtype, st)) return e;      // long contractions first: a persistent CTA that gets two tiles gets a short second one
    if (d_text) {
        // input gradients: residual path (dZ) + query projection, then + key/value projection of the other direction
        const char* w_t2i = static_cast<const char*>(params[0]);
        const char* w_i2t = static_cast<const char*>(params[4]);
        TcGemmProblem px[2];
        px[0] = tcg(X + W.dQt, D, false, w_t2i, D, true, nt, D, D);
        px[0].C = reinterpret_cast<float*>(X + W.dXt); px[0].residual = X + W.dZt; px[0].res_kind = 1; px[0].ldr = D;
        px[1] = tcg(X + W.dQi, D, false, w_i2t, D, true, ni, D, D);
        px[1].C = reinterpret_cast<float*>(X + W.dXi); px[1].residual = X + W.dZi; px[1].res_kind = 1; px[1].ldr = D;
        if (int e = launch_gemm_tc(px, 2, dtype, st)) return e;
        TcGemmProblem py[2];
        py[0] = tcg(X + W.dKVt, 2 * D, false, w_i2t + dd2, D, true, nt, D, 2 * D);
        py[0].C = reinterpret_cast<float*>(X + W.dXt); py[0].accumulate = 1; py[0].C16 = d_text; py[0].ldc16 = D;
        py[1] = tcg(X + W.dKVi, 2 * D, false, w_t2i + dd2, D, true, ni, D, 2 * D);
        py[1].C = reinterpret_cast<float*>(X + W.dXi); py[1].accumulate = 1; py[1].C16 = d_image; py[1].ldc16 = D;
        if (int e = launch_gemm_tc(py, 2, dtype, st)) return e;
    }
    // small vectors: LayerNorm affine gradients, out-projection and in-projection bias gradients
    FinArgs fa{};
    int nj = 0;
    const float* lp = reinterpret_cast<const float*>(X + W.lnpart);
    const float* lpi = lp + (size_t)W.nblk_t * 3 * d;
    auto job = [&](const float* src, int nparts, long long stride, int n, void* dst) { fa.j[nj++] = FinJob{src, nparts, stride, n, dst}; };
    job(lp, W.nblk_t, 3 * d, D, grads[9]);            // norm_text.bias   = colsum(dY)
    job(lp + d, W.nblk_t, 3 * d, D, grads[8]);        // norm_text.weight = colsum(dY zhat)
    job(lp + 2 * d, W.nblk_t, 3 * d, D, grads[3]);    // text_to_image.out_proj.bias = colsum(dZ)
    job(lpi, W.nblk_i, 3 * d, D, grads[11]);
    job(lpi + d, W.nblk_i, 3 * d, D, grads[10]);
    job(lpi + 2 * d, W.nblk_i, 3 * d, D, grads[7]);
    job(reinterpret_cast<const float*>(X + W.dbq0), (int)B, d, D, grads[1]);
    job(reinterpret_cast<const float*>(X + W.dbkv0), (int)B, 2 * d, 2 * D, static_cast<char*>(grads[1]) + (size_t)d * 2);
    job(reinterpret_cast<const float*>(X + W.dbq1), (int)B, d, D, grads[5]);
    job(reinterpret_cast<const float*>(X + W.dbkv1), (int)B, 2 * d, 2 * D, static_cast<char*>(grads[5]) + (size_t)d * 2);
    fa.count = nj;
    const dim3 grid((unsigned)ceil_div(2 * D, 32), (unsigned)nj);
    if (dtype == DCLIP_BF16) k_vec_finalize<__nv_bfloat16><<<grid, 256, 0, st>>>(fa);
    else k_vec_finalize<__half><<<grid, 256, 0, st>>>(fa);
    DCLIP_LAUNCH_OK("k_vec_finalize");
    return DCLIP_OK;
}

}  // namespace dclip
