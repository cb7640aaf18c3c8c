// Distillation objective — FP32 CUDA-core path (exact tier: 1e-5 vs the fp32 reference).
//
// Reference math: training/CLIP_image_distillation.py:532-576 (compute_contrastive_loss,
// cosine_distillation_loss) and :594-634 (combine).  The B x B logits are never stored: the
// forward keeps per-row log-sum-exp vectors of S and of S^T (each side computes the row block
// it owns against all rows of the other side — the same kernel serves both sides and every
// data-parallel rank), the backward recomputes 32 x TJ score tiles and feeds them straight
// into the second contraction (dX = G . Y) from shared memory.
//
// Layout: X_hat tile (32 rows x d, fp32) stays resident in shared memory for the whole CTA
// life; Y_hat tiles (TJ rows x d) stream through; row stride d+4 floats keeps 128-bit
// shared loads conflict-free.  Grid: (row tiles, column splits, sides).
#include "common.cuh"
#include "loss_internal.h"
#include "tcgen05.cuh"

namespace dclip {

constexpr int kTI = 32;          // rows of X per CTA
constexpr int kThreads = 256;
constexpr float kNormEps = 1e-12f;  // F.normalize eps

// ---------------------------------------------------------------------------------------------
// prep: one warp per row.  x_hat = x / max(|x|, eps); cos = x_hat . t_hat
// ---------------------------------------------------------------------------------------------

// x / max(|x|, eps): the FP32 tier divides like F.normalize (1e-5 parity); 16-bit OPERANDS (rounded to 2^-11 right after) take the
// reciprocal the kernel stores as inv_norm anyway - an fp32 division is ~12 instructions with a branch, and the register-resident
// prep kernel was bound by exactly those (48 per lane and row pair: 17 -> see DESIGN 4.3)
template <typename OT> __device__ __forceinline__ float unit_component(float x, float nx, float inv_nx) { return x * inv_nx; }
template <> __device__ __forceinline__ float unit_component<float>(float x, float nx, float inv_nx) { return x / nx; }

template <typename T, typename OT>
__global__ void __launch_bounds__(kThreads) k_loss_prep(PrepArgs a) {
    const PrepSide s = a.s[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= a.m) return;
    const int64_t d = a.d;
    const T* x = static_cast<const T*>(s.x) + row * d;
    const T* t = s.t ? static_cast<const T*>(s.t) + row * d : nullptr;
    float ss = 0.f, tt = 0.f, xt = 0.f;
    for (int64_t k = lane * 8; k < d; k += 256) {
        float xv[8];
        load8<T>(x + k, xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(xv[i], xv[i], ss);
        if (t) {
            float tv[8];
            load8<T>(t + k, tv);
#pragma unroll
            for (int i = 0; i < 8; ++i) { tt = fmaf(tv[i], tv[i], tt); xt = fmaf(xv[i], tv[i], xt); }
        }
    }
    ss = warp_sum(ss);
    const float nx = fmaxf(sqrtf(ss), kNormEps);
    if (t) { tt = warp_sum(tt); xt = warp_sum(xt); }
    if (lane == 0) {
        s.inv_norm[row] = 1.0f / nx;
        if (s.cos) s.cos[row] = t ? xt / (nx * fmaxf(sqrtf(tt), kNormEps)) : 0.f;
    }
    OT* out = static_cast<OT*>(s.xhat) + row * d;
    const float inv_nx = 1.0f / nx;
    for (int64_t k = lane * 8; k < d; k += 256) {
        float xv[8];
        load8<T>(x + k, xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = unit_component<OT>(xv[i], nx, inv_nx);
        store8<OT>(out + k, xv);
    }
}

// both sides of one row per warp (single-GPU composite): also yields diag[i] = u_hat_i . v_hat_i / tau
// computed from the operand values as rounded to OT, i.e. the S_ii the tensor tiles see.
template <typename T, typename OT>
__global__ void __launch_bounds__(kThreads) k_loss_prep_pair(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= a.m) return;
    const int64_t d = a.d;
    float nrm[2], inv_nrm[2];
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        const PrepSide s = a.s[sd];
        const T* x = static_cast<const T*>(s.x) + row * d;
        const T* t = s.t ? static_cast<const T*>(s.t) + row * d : nullptr;
        float ss = 0.f, tt = 0.f, xt = 0.f;
        for (int64_t k = lane * 8; k < d; k += 256) {
            float xv[8];
            load8<T>(x + k, xv);
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(xv[i], xv[i], ss);
            if (t) {
                float tv[8];
                load8<T>(t + k, tv);
#pragma unroll
                for (int i = 0; i < 8; ++i) { tt = fmaf(tv[i], tv[i], tt); xt = fmaf(xv[i], tv[i], xt); }
            }
        }
        ss = warp_sum(ss);
        const float nx = fmaxf(sqrtf(ss), kNormEps);
        nrm[sd] = nx; inv_nrm[sd] = 1.0f / nx;
        if (t) { tt = warp_sum(tt); xt = warp_sum(xt); }
        if (lane == 0) {
            s.inv_norm[row] = 1.0f / nx;
            if (s.cos) s.cos[row] = t ? xt / (nx * fmaxf(sqrtf(tt), kNormEps)) : 0.f;
            if (s.t_inv) s.t_inv[row] = t ? 1.0f / fmaxf(sqrtf(tt), kNormEps) : 0.f;
        }
    }
    const T* x0 = static_cast<const T*>(a.s[0].x) + row * d;
    const T* x1 = static_cast<const T*>(a.s[1].x) + row * d;
    OT* o0 = static_cast<OT*>(a.s[0].xhat) + row * d;
    OT* o1 = static_cast<OT*>(a.s[1].xhat) + row * d;
    float dot = 0.f;
    for (int64_t k = lane * 8; k < d; k += 256) {
        float u[8], v[8];
        load8<T>(x0 + k, u);
        load8<T>(x1 + k, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            u[i] = to_f32<OT>(from_f32<OT>(unit_component<OT>(u[i], nrm[0], inv_nrm[0])));
            v[i] = to_f32<OT>(from_f32<OT>(unit_component<OT>(v[i], nrm[1], inv_nrm[1])));
            dot = fmaf(u[i], v[i], dot);
        }
        store8<OT>(o0 + k, u);
        store8<OT>(o1 + k, v);
    }
    dot = warp_sum(dot);
    if (lane == 0 && a.diag) a.diag[row] = dot * a.inv_tau;
}

// Same, for d <= 1024: every load of the row (both students, both teachers) is issued before the first use, so a warp has
// 4 NC 16-byte loads in flight instead of one dependent round trip per chunk (the generic kernel runs at ~40 % of the HBM
// rate for exactly that reason), and the second pass works on registers.
// (rows are held as the RAW 16-byte words they were loaded as - 4 NC of them per lane, all requested before the first use;
// keeping fp32 copies instead costs 2 x the registers and makes the compiler serialise load -> convert -> load)
template <typename T> __device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& raw, float (&v)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& raw, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <typename T, typename OT, int NC>
__global__ void __launch_bounds__(128) k_loss_prep_pair_nc(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const bool px = a.peers.world > 1;
    if (row >= a.m && !px) return;
    const int64_t d = a.d;
    const int nch = row < a.m ? (int)(d >> 3) : 0;       // (a warp past the last row only takes part in the block's signalling)
    const bool has_t = a.s[0].t != nullptr;
    uint4 xr[2][NC], tr[2][NC];
#pragma unroll
    for (int sd = 0; sd < 2; ++sd)
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int ch = lane + 32 * k;
            xr[sd][k] = make_uint4(0u, 0u, 0u, 0u); tr[sd][k] = make_uint4(0u, 0u, 0u, 0u);
            if (ch < nch) {
                xr[sd][k] = ldg_nc16(static_cast<const T*>(a.s[sd].x) + row * d + ch * 8);
                if (has_t) tr[sd][k] = ldg_nc16(static_cast<const T*>(a.s[sd].t) + row * d + ch * 8);
            }
        }
    if (px) {
        // my rows go straight into every peer's operand buffer: not before the peer has finished reading the previous step's
        const unsigned int target = p2p_ld_volatile(&p2p_flags(a.peers, a.peers.rank)->expected[kP2PKindDone]) + 1;
        p2p_wait_all_warp(a.peers, kP2PKindDone, target, a.error_flag, lane);
    }
    float nrm[2], inv_nrm[2];
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        float ss = 0.f, tt = 0.f, xt = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (lane + 32 * k < nch) {
                float x[8], t[8];
                unpack8<T>(xr[sd][k], x);
                unpack8<T>(tr[sd][k], t);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    ss = fmaf(x[i], x[i], ss);
                    if (has_t) { tt = fmaf(t[i], t[i], tt); xt = fmaf(x[i], t[i], xt); }
                }
            }
        ss = warp_sum(ss);
        const float nx = fmaxf(sqrtf(ss), kNormEps);
        nrm[sd] = nx; inv_nrm[sd] = 1.0f / nx;
        if (has_t) { tt = warp_sum(tt); xt = warp_sum(xt); }
        if (lane == 0 && row < a.m) {
            a.s[sd].inv_norm[row] = 1.0f / nx;
            if (a.s[sd].cos) a.s[sd].cos[row] = has_t ? xt / (nx * fmaxf(sqrtf(tt), kNormEps)) : 0.f;
            if (a.s[sd].t_inv) a.s[sd].t_inv[row] = has_t ? 1.0f / fmaxf(sqrtf(tt), kNormEps) : 0.f;
        }
    }
    OT* o0 = static_cast<OT*>(a.s[0].xhat) + row * d;
    OT* o1 = static_cast<OT*>(a.s[1].xhat) + row * d;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nch) {
            float u[8], v[8];
            unpack8<T>(xr[0][k], u);
            unpack8<T>(xr[1][k], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                u[i] = to_f32<OT>(from_f32<OT>(unit_component<OT>(u[i], nrm[0], inv_nrm[0])));
                v[i] = to_f32<OT>(from_f32<OT>(unit_component<OT>(v[i], nrm[1], inv_nrm[1])));
                dot = fmaf(u[i], v[i], dot);
            }
            store8<OT>(o0 + ch * 8, u);
            store8<OT>(o1 + ch * 8, v);
            if (px)
                for (int r = 0; r < a.peers.world; ++r)
                    if (r != a.peers.rank) {         // 512 contiguous bytes per warp and store: full NVLink write packets
                        store8<OT>(p2p_remote(a.peers, o0 + ch * 8, r), u);
                        store8<OT>(p2p_remote(a.peers, o1 + ch * 8, r), v);
                    }
        }
    }
    dot = warp_sum(dot);
    if (lane == 0 && a.diag && row < a.m) a.diag[row] = dot * a.inv_tau;
    if (px) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && p2p_last_block(a.peers, gridDim.x)) p2p_signal_all(a.peers, kP2PKindOperands);
    }
}

// cos(x_i, t_i) for both sides with the students' clamped inverse norms from the prep kernel (one warp per row)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_cos_rows(CosArgs a) {
    const CosSide s = a.s[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= a.m) return;
    const int64_t d = a.d;
    const T* x = static_cast<const T*>(s.x) + row * d;
    const T* t = static_cast<const T*>(s.t) + row * d;
    float tt = 0.f, xt = 0.f;
    for (int64_t k = lane * 8; k < d; k += 256) {
        float xv[8], tv[8];
        load8<T>(x + k, xv);
        load8<T>(t + k, tv);
#pragma unroll
        for (int i = 0; i < 8; ++i) { tt = fmaf(tv[i], tv[i], tt); xt = fmaf(xv[i], tv[i], xt); }
    }
    tt = warp_sum(tt); xt = warp_sum(xt);
    if (lane == 0) s.cos[row] = xt * s.inv_norm[row] / fmaxf(sqrtf(tt), kNormEps);
}

// tensor path finalize: both lse vectors from the partial sums + the loss sums.  One cluster of 8 CTAs (a
// single CTA spends its time on load latency); every reduction runs in a fixed order, so the result is
// bit-reproducible: thread -> rows i = g, g + 4096, ...; CTA tree; then cluster rank 0 adds the 8 CTA sums
// it received through distributed shared memory in rank order.
constexpr int kFinCtas = 8, kFinThreads = 512;
__global__ void __cluster_dims__(kFinCtas, 1, 1) __launch_bounds__(kFinThreads) k_tc_finalize(FinalizeArgs f) {
    __shared__ float red[3][kFinThreads];
    __shared__ float gather[kFinCtas][4];
    const SumsArgs& a = f.sums;
    const uint32_t rank = tc::cluster_ctarank();
    const int nsv = f.nsplit_v > 0 ? f.nsplit_v : f.nsplit;
    const bool px = f.peers.world > 1;
    if (px && rank == 0 && threadIdx.x == 0) {
        // the kernels before this one have read their targets: retire this step's "operands" and "step done" waits
        P2PFlags* fl = p2p_flags(f.peers, f.peers.rank);
        fl->expected[kP2PKindOperands] += 1;
        fl->expected[kP2PKindDone] += 1;
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int64_t i = (int64_t)rank * kFinThreads + threadIdx.x; i < f.m; i += kFinCtas * kFinThreads) {
        // a thread has one or two rows: every load of the row is requested before the first add (a batch of 8 at a time was four
        // dependent L2 round trips for the 32 column partials of B = 4096); the sums keep their order
        const float dg = a.diag[i];
        const float c_img = a.cos_img ? a.cos_img[i] : 0.f, c_txt = a.cos_txt ? a.cos_txt[i] : 0.f;
        float su = 0.f, sv = 0.f;
#pragma unroll 4
        for (int p = 0; p < f.nsplit; ++p) su += f.part_u[(int64_t)p * f.m + i];
        for (int p0 = 0; p0 < nsv; p0 += 32) {
            float pv[32];
#pragma unroll
            for (int p = 0; p < 32; ++p) pv[p] = p0 + p < nsv ? __ldg(f.part_v + (int64_t)(p0 + p) * f.m + i) : 0.f;
#pragma unroll
            for (int p = 0; p < 32; ++p) sv += pv[p];
        }
        const float lu = f.inv_tau + logf(su), lv = f.inv_tau + logf(sv);
        f.lse_u[i] = lu; f.lse_v[i] = lv;
        s0 += lu + lv - 2.0f * dg;
        if (a.cos_img) s1 += 1.0f - c_img;
        if (a.cos_txt) s2 += 1.0f - c_txt;
    }
    red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2;
    __syncthreads();
    for (int o = kFinThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
            red[2][threadIdx.x] += red[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) tc::st_shared_cluster_f32(tc::mapa_u32(tc::smem_u32(&gather[rank][threadIdx.x]), 0), red[threadIdx.x][0]);
    if (px) __threadfence();               // my lse stores are visible to the CTAs that forward them
    tc::cluster_sync_all();
    if (px) {
        // the statistics row {lse_u[m], lse_v[m]} (contiguous in the symmetric buffer) goes to every peer as 16-byte stores, 512
        // contiguous bytes per warp (scattered 4-byte remote stores cost one NVLink packet each): CTA c forwards to peers c, c + 8
        const int64_t n16 = (2 * f.m * 4) / 16, tail4 = (2 * f.m) - n16 * 4;
        for (int r = (int)rank; r < f.peers.world; r += kFinCtas) {
            if (r == f.peers.rank) continue;
            const uint4* src = reinterpret_cast<const uint4*>(f.lse_u);
            uint4* dst = reinterpret_cast<uint4*>(p2p_remote(f.peers, f.lse_u, r));
            for (int64_t i = threadIdx.x; i < n16; i += kFinThreads) dst[i] = __ldcg(src + i);
            if ((int64_t)threadIdx.x < tail4) p2p_remote(f.peers, f.lse_u, r)[n16 * 4 + threadIdx.x] = __ldcg(f.lse_u + n16 * 4 + threadIdx.x);
        }
        __threadfence_system();
        tc::cluster_sync_all();
    }
    if (rank == 0 && threadIdx.x == 0) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int r = 0; r < kFinCtas; ++r) { t0 += gather[r][0]; t1 += gather[r][1]; t2 += gather[r][2]; }
        if (a.out3) { a.out3[0] = t0; a.out3[1] = t1; a.out3[2] = t2; }   // sharded / students phase: raw local sums
        if (a.zero_counter) *a.zero_counter = 0u;
        if (px && a.out3) {
            for (int r = 0; r < f.peers.world; ++r)
                if (r != f.peers.rank) { float* o = p2p_remote(f.peers, a.out3, r); o[0] = t0; o[1] = t1; o[2] = t2; }
            p2p_signal_all(f.peers, kP2PKindStats);      // (fences first) every CTA's row copies were fenced before the second cluster barrier
        }
        if (a.losses) {
            const float l_con = t0 / (2.0f * (float)f.m);
            const float l_img = a.cos_img ? t1 / (float)f.m : 0.f;
            const float l_txt = a.cos_txt ? t2 / (float)f.m : 0.f;
            a.losses[0] = a.w_img * l_img + a.w_txt * l_txt + a.w_con * l_con;
            a.losses[1] = l_img; a.losses[2] = l_txt; a.losses[3] = l_con;
        }
    }
}

// global losses from the per-rank raw sums (fixed rank order: every rank gets the same bits)
__global__ void k_sharded_losses(const float* stat, int world, int64_t stride, int64_t off, float B, int has_img, int has_txt,
                                 float w_img, float w_txt, float w_con, float* __restrict__ losses, const PeerCtx peers, int* error_flag) {
    if (blockIdx.x != 0) return;
    if (peers.world > 1) {
        // one warp: lane r waits for rank r's statistics (written into my buffer by its finalize kernel), then the wait is retired
        P2PFlags* fl = p2p_flags(peers, peers.rank);
        const unsigned int target = p2p_ld_volatile(&fl->expected[kP2PKindStats]) + 1;
        p2p_wait_all_warp(peers, kP2PKindStats, target, error_flag, (int)threadIdx.x);
        if (threadIdx.x == 0) fl->expected[kP2PKindStats] = target;
    }
    if (threadIdx.x != 0) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < world; ++r) {
        const volatile float* q = stat + r * stride + off;
        s0 += q[0]; s1 += q[1]; s2 += q[2];
    }
    const float l_con = s0 / (2.0f * B), l_img = has_img ? s1 / B : 0.f, l_txt = has_txt ? s2 / B : 0.f;
    losses[0] = w_img * l_img + w_txt * l_txt + w_con * l_con;
    losses[1] = l_img; losses[2] = l_txt; losses[3] = l_con;
}

// ---------------------------------------------------------------------------------------------
// score tile: 2 x CPT micro-tile per thread from smem-resident rows, k ascending (fma chain)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_rows(float* dst, int ld, const float* src, int64_t row0, int64_t nrows,
                                          int rows, int64_t d) {
    const int vec_per_row = (int)(d / 4);
    for (int idx = threadIdx.x; idx < rows * vec_per_row; idx += kThreads) {
        const int r = idx / vec_per_row, c = idx - r * vec_per_row;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * d) + c);
        *reinterpret_cast<float4*>(dst + r * ld + c * 4) = v;
    }
}

template <int CPT>
__device__ __forceinline__ void score_tile(const float* Xs, const float* Ys, int ld, int d, int ty, int tx,
                                           float (&acc)[2][CPT]) {
    const float4* x0 = reinterpret_cast<const float4*>(Xs + ty * ld);
    const float4* x1 = reinterpret_cast<const float4*>(Xs + (ty + 16) * ld);
    const float4* y0 = reinterpret_cast<const float4*>(Ys + tx * ld);
    const float4* y1 = reinterpret_cast<const float4*>(Ys + (tx + 16) * ld);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int k = 0; k < d / 4; ++k) {
        const float4 a0 = x0[k], a1 = x1[k], b0 = y0[k];
        acc[0][0] = fmaf(a0.x, b0.x, acc[0][0]); acc[0][0] = fmaf(a0.y, b0.y, acc[0][0]);
        acc[0][0] = fmaf(a0.z, b0.z, acc[0][0]); acc[0][0] = fmaf(a0.w, b0.w, acc[0][0]);
        acc[1][0] = fmaf(a1.x, b0.x, acc[1][0]); acc[1][0] = fmaf(a1.y, b0.y, acc[1][0]);
        acc[1][0] = fmaf(a1.z, b0.z, acc[1][0]); acc[1][0] = fmaf(a1.w, b0.w, acc[1][0]);
        if (CPT == 2) {
            const float4 b1 = y1[k];
            acc[0][CPT - 1] = fmaf(a0.x, b1.x, acc[0][CPT - 1]); acc[0][CPT - 1] = fmaf(a0.y, b1.y, acc[0][CPT - 1]);
            acc[0][CPT - 1] = fmaf(a0.z, b1.z, acc[0][CPT - 1]); acc[0][CPT - 1] = fmaf(a0.w, b1.w, acc[0][CPT - 1]);
            acc[1][CPT - 1] = fmaf(a1.x, b1.x, acc[1][CPT - 1]); acc[1][CPT - 1] = fmaf(a1.y, b1.y, acc[1][CPT - 1]);
            acc[1][CPT - 1] = fmaf(a1.z, b1.z, acc[1][CPT - 1]); acc[1][CPT - 1] = fmaf(a1.w, b1.w, acc[1][CPT - 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward: per-row online (max, sum) over a column split; diag pick
// ---------------------------------------------------------------------------------------------
template <int CPT>
__global__ void __launch_bounds__(kThreads) k_rowlse_partial(LseArgs a) {
    constexpr int TJ = 16 * CPT;
    extern __shared__ __align__(16) float smem[];
    const LseSide s = a.s[blockIdx.z];
    const int d = (int)a.d, ld = d + 4;
    float* Xs = smem;
    float* Ys = smem + kTI * ld;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int64_t row0 = (int64_t)blockIdx.x * kTI;
    load_rows(Xs, ld, s.x, row0, a.m, kTI, a.d);

    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int64_t ntiles = (a.n + TJ - 1) / TJ;
    const int64_t t_begin = (int64_t)blockIdx.y * a.tiles_per_split;
    const int64_t t_end = min(ntiles, t_begin + a.tiles_per_split);
    for (int64_t jt = t_begin; jt < t_end; ++jt) {
        const int64_t col0 = jt * TJ;
        __syncthreads();
        load_rows(Ys, ld, s.y, col0, a.n, TJ, a.d);
        __syncthreads();
        float acc[2][CPT];
        score_tile<CPT>(Xs, Ys, ld, d, ty, tx, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t grow = row0 + ty + 16 * r;
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int64_t gcol = col0 + tx + 16 * c;
                float v = acc[r][c] / a.temperature;
                if (gcol >= a.n) v = -INFINITY;
                else if (s.diag && gcol == a.row_offset + grow && grow < a.m) s.diag[grow] = v;
                acc[r][c] = v;
                tmax = fmaxf(tmax, v);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            const float m_new = fmaxf(m_run[r], tmax);   // finite: every tile has >= 1 valid column
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; ++c) part += expf(acc[r][c] - m_new);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            l_run[r] = l_run[r] * expf(m_run[r] - m_new) + part;
            m_run[r] = m_new;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t grow = row0 + ty + 16 * r;
            if (grow < a.m) {
                s.part_m[(int64_t)blockIdx.y * a.m + grow] = m_run[r];
                s.part_l[(int64_t)blockIdx.y * a.m + grow] = l_run[r];
            }
        }
    }
}

__global__ void k_lse_combine(LseArgs a) {
    const LseSide s = a.s[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.m) return;
    float M = -INFINITY;
    for (int p = 0; p < a.nsplit; ++p) M = fmaxf(M, s.part_m[(int64_t)p * a.m + i]);
    float L = 0.f;
    for (int p = 0; p < a.nsplit; ++p) {
        const float mp = s.part_m[(int64_t)p * a.m + i];
        if (mp > -INFINITY) L += s.part_l[(int64_t)p * a.m + i] * expf(mp - M);
    }
    s.lse[i] = M + logf(L);
}

// ---------------------------------------------------------------------------------------------
// backward: dX = coef/tau * G . Y with G recomputed tile by tile
// ---------------------------------------------------------------------------------------------
template <int CPT, int NC>
__global__ void __launch_bounds__(kThreads) k_infonce_grad(GradArgs a) {
    constexpr int TJ = 16 * CPT;
    constexpr int GLD = kTI + 4;
    extern __shared__ __align__(16) float smem[];
    const GradSide s = a.s[blockIdx.z];
    const int d = (int)a.d, ld = d + 4;
    float* Xs = smem;
    float* Ys = Xs + kTI * ld;
    float* Gt = Ys + TJ * ld;   // [TJ][GLD]: Gt[j][i]
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int64_t row0 = (int64_t)blockIdx.x * kTI;
    load_rows(Xs, ld, s.x, row0, a.m, kTI, a.d);

    float lx[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int64_t grow = row0 + ty + 16 * r;
        lx[r] = grow < a.m ? s.lse_x[grow] : 0.f;
    }
    float out[kTI][NC];
#pragma unroll
    for (int i = 0; i < kTI; ++i)
#pragma unroll
        for (int q = 0; q < NC; ++q) out[i][q] = 0.f;

    const int64_t ntiles = (a.n + TJ - 1) / TJ;
    for (int64_t jt = 0; jt < ntiles; ++jt) {
        const int64_t col0 = jt * TJ;
        __syncthreads();   // previous tile's Ys / Gt fully consumed
        load_rows(Ys, ld, s.y, col0, a.n, TJ, a.d);
        __syncthreads();
        float acc[2][CPT];
        score_tile<CPT>(Xs, Ys, ld, d, ty, tx, acc);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int64_t gcol = col0 + tx + 16 * c;
            const float ly = gcol < a.n ? s.lse_y[gcol] : 0.f;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int64_t grow = row0 + ty + 16 * r;
                float g = 0.f;
                if (gcol < a.n && grow < a.m) {
                    const float v = acc[r][c] / a.temperature;
                    g = expf(v - lx[r]) + expf(v - ly);
                    if (gcol == a.row_offset + grow) g -= 2.0f;
                }
                Gt[(tx + 16 * c) * GLD + ty + 16 * r] = g;
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < TJ; ++j) {
            float y[NC];
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const int col = threadIdx.x + kThreads * q;
                y[q] = col < d ? Ys[j * ld + col] : 0.f;
            }
            const float4* g4 = reinterpret_cast<const float4*>(Gt + j * GLD);
#pragma unroll
            for (int i4 = 0; i4 < kTI / 4; ++i4) {
                const float4 g = g4[i4];
#pragma unroll
                for (int q = 0; q < NC; ++q) {
                    out[4 * i4 + 0][q] = fmaf(g.x, y[q], out[4 * i4 + 0][q]);
                    out[4 * i4 + 1][q] = fmaf(g.y, y[q], out[4 * i4 + 1][q]);
                    out[4 * i4 + 2][q] = fmaf(g.z, y[q], out[4 * i4 + 2][q]);
                    out[4 * i4 + 3][q] = fmaf(g.w, y[q], out[4 * i4 + 3][q]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kTI; ++i) {
        const int64_t grow = row0 + i;
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int col = threadIdx.x + kThreads * q;
            if (grow < a.m && col < d) s.dx[grow * a.d + col] = a.coef_over_tau * out[i][q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tail: cosine-term gradient + normalise backward, one warp per row
// ---------------------------------------------------------------------------------------------

// sharded step with in-kernel exchange: the gradient tail is the last reader of this step's peer-written buffers (stream order) -
// its last block tells every peer that they may be overwritten ("step done" credit of the next step's operand rows)
__device__ __forceinline__ void tail_signal_done(const TailArgs& a) {
    __syncthreads();                     // (nothing is published to the peers here: no per-thread system fence; p2p_signal_all fences once)
    if (threadIdx.x == 0 && p2p_last_block(a.peers, gridDim.x * gridDim.y)) p2p_signal_all(a.peers, kP2PKindDone);
}
// teachers phase of the split objective (TailSums): the last block to finish sums the cosine terms in a fixed order (thread-strided,
// shuffle tree, warps in order) and writes the four losses; the host-buffer step gets them straight into mapped pinned memory,
// sequence word last
__device__ __forceinline__ void tail_finish_losses(const TailArgs& a) {
    __shared__ float red[2][32];
    __shared__ int is_last;
    const TailSums& q = a.sums;
    __syncthreads();                     // this block's cosine terms are stored (thread 0's fence below is cumulative)
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(q.counter, 1u) == gridDim.x * gridDim.y - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const float con_sum = threadIdx.x == 0 ? __ldcg(q.con_sum) : 0.f;      // requested ahead of the sums
    float s1 = 0.f, s2 = 0.f;
    const int64_t m4 = a.m >> 2;
    if (q.cos_img) {
        const float4* c = reinterpret_cast<const float4*>(q.cos_img);
#pragma unroll 8
        for (int64_t i = threadIdx.x; i < m4; i += blockDim.x) { const float4 v = __ldcg(c + i); s1 += (1.0f - v.x) + (1.0f - v.y) + (1.0f - v.z) + (1.0f - v.w); }
        for (int64_t i = m4 * 4 + threadIdx.x; i < a.m; i += blockDim.x) s1 += 1.0f - __ldcg(q.cos_img + i);
    }
    if (q.cos_txt) {
        const float4* c = reinterpret_cast<const float4*>(q.cos_txt);
#pragma unroll 8
        for (int64_t i = threadIdx.x; i < m4; i += blockDim.x) { const float4 v = __ldcg(c + i); s2 += (1.0f - v.x) + (1.0f - v.y) + (1.0f - v.z) + (1.0f - v.w); }
        for (int64_t i = m4 * 4 + threadIdx.x; i < a.m; i += blockDim.x) s2 += 1.0f - __ldcg(q.cos_txt + i);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    float t1 = 0.f, t2 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t1 += red[0][w]; t2 += red[1][w]; }
    const float l_con = con_sum / (2.0f * (float)a.m);
    const float l_img = q.cos_img ? t1 / (float)a.m : 0.f;
    const float l_txt = q.cos_txt ? t2 / (float)a.m : 0.f;
    const float total = q.w_img * l_img + q.w_txt * l_txt + q.w_con * l_con;
    if (q.losses_host) *reinterpret_cast<float4*>(q.losses_host) = make_float4(total, l_img, l_txt, l_con);
    q.losses[0] = total; q.losses[1] = l_img; q.losses[2] = l_txt; q.losses[3] = l_con;
    *q.counter = 0u;
    if (q.losses_host) {
        __threadfence_system();          // (also orders this step's gradient stores before the word the host waits for)
        *reinterpret_cast<volatile unsigned int*>(q.losses_host + 4) = q.seq;
    }
}
template <typename T>
__global__ void __launch_bounds__(kThreads) k_loss_grad_tail(TailArgs a) {
    const TailSide s = a.s[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= a.m) {
        if (a.peers.world > 1) tail_signal_done(a);
        if (a.sums.counter) tail_finish_losses(a);
        return;
    }
    const int64_t d = a.d;
    const T* x = static_cast<const T*>(s.x) + row * d;
    const T* t = s.t ? static_cast<const T*>(s.t) + row * d : nullptr;
    const float* g_in = s.dx_hat ? s.dx_hat + row * d : nullptr;
    const float inv = s.inv_norm[row];
    float t_scale = 0.f;
    if (t) {
        float tt = 0.f;
        for (int64_t k = lane * 8; k < d; k += 256) {
            float tv[8];
            load8<T>(t + k, tv);
#pragma unroll
            for (int i = 0; i < 8; ++i) tt = fmaf(tv[i], tv[i], tt);
        }
        tt = warp_sum(tt);
        t_scale = -a.w_cos_over_B[blockIdx.y] / fmaxf(sqrtf(tt), kNormEps);
        if (s.cos_out) {          // the row's cosine term, exactly as k_cos_rows forms it
            float xt = 0.f;
            for (int64_t k = lane * 8; k < d; k += 256) {
                float xv[8], tv[8];
                load8<T>(x + k, xv);
                load8<T>(t + k, tv);
#pragma unroll
                for (int i = 0; i < 8; ++i) xt = fmaf(xv[i], tv[i], xt);
            }
            xt = warp_sum(xt);
            if (lane == 0) s.cos_out[row] = xt * inv / fmaxf(sqrtf(tt), kNormEps);
        }
    }
    // pass 1: dot = x_hat . g
    float dot = 0.f;
    for (int64_t k = lane * 8; k < d; k += 256) {
        float xv[8], g[8];
        load8<T>(x + k, xv);
        if (g_in) load8<float>(g_in + k, g);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = 0.f;
        }
        if (t) {
            float tv[8];
            load8<T>(t + k, tv);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = fmaf(t_scale, tv[i], g[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(xv[i] * inv, g[i], dot);
    }
    dot = warp_sum(dot);
    // F.normalize clamps the norm: below eps the denominator is a constant (no projection term)
    const bool clamped = inv >= 1.0f / kNormEps;
    if (clamped) dot = 0.f;
    T* out = static_cast<T*>(s.d_x) + row * d;
    const float scale = inv * a.grad_scale;
    for (int64_t k = lane * 8; k < d; k += 256) {
        float xv[8], g[8];
        load8<T>(x + k, xv);
        if (g_in) load8<float>(g_in + k, g);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = 0.f;
        }
        if (t) {
            float tv[8];
            load8<T>(t + k, tv);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = fmaf(t_scale, tv[i], g[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = (g[i] - xv[i] * inv * dot) * scale;
        store8<T>(out + k, g);
    }
    if (a.peers.world > 1) tail_signal_done(a);
    if (a.sums.counter) tail_finish_losses(a);
}

// Same for d <= 1024 and 16-bit I/O: all loads of the row up front (raw 16-byte words, see k_loss_prep_pair_nc), the projection
// and the store on registers.
template <typename T, int NC>
__global__ void __launch_bounds__(128) k_loss_grad_tail_nc(TailArgs a) {
    const TailSide s = a.s[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= a.m) {
        if (a.peers.world > 1) tail_signal_done(a);
        if (a.sums.counter) tail_finish_losses(a);
        return;
    }
    const int64_t d = a.d;
    const int nch = (int)(d >> 3);
    const bool has_t = s.t != nullptr, has_g = s.dx_hat != nullptr;
    uint4 xr[NC], tr[NC];
    float g[NC][8];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int ch = lane + 32 * k;
        xr[k] = make_uint4(0u, 0u, 0u, 0u); tr[k] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[k][i] = 0.f;
        if (ch < nch) {
            xr[k] = ldg_nc16(static_cast<const T*>(s.x) + row * d + ch * 8);
            if (has_t) tr[k] = ldg_nc16(static_cast<const T*>(s.t) + row * d + ch * 8);
            if (has_g) load8<float>(s.dx_hat + row * d + ch * 8, g[k]);
        }
    }
    const float inv = s.inv_norm[row];
    if (has_t) {
        float tt = 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (lane + 32 * k < nch) {
                float t[8];
                unpack8<T>(tr[k], t);
#pragma unroll
                for (int i = 0; i < 8; ++i) tt = fmaf(t[i], t[i], tt);
            }
        tt = warp_sum(tt);
        const float t_scale = -a.w_cos_over_B[blockIdx.y] / fmaxf(sqrtf(tt), kNormEps);
        if (s.cos_out) {          // the row's cosine term, exactly as k_cos_rows forms it (same chunk order per lane)
            float xt = 0.f;
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (lane + 32 * k < nch) {
                    float x[8], t[8];
                    unpack8<T>(xr[k], x);
                    unpack8<T>(tr[k], t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) xt = fmaf(x[i], t[i], xt);
                }
            xt = warp_sum(xt);
            if (lane == 0) s.cos_out[row] = xt * inv / fmaxf(sqrtf(tt), kNormEps);
        }
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (lane + 32 * k < nch) {
                float t[8];
                unpack8<T>(tr[k], t);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[k][i] = fmaf(t_scale, t[i], g[k][i]);
            }
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k)
        if (lane + 32 * k < nch) {
            float x[8];
            unpack8<T>(xr[k], x);
#pragma unroll
            for (int i = 0; i < 8; ++i) dot = fmaf(x[i] * inv, g[k][i], dot);
        }
    dot = warp_sum(dot);
    if (inv >= 1.0f / kNormEps) dot = 0.f;           // clamped norm: F.normalize has no projection term
    const float scale = inv * a.grad_scale;
    T* out = static_cast<T*>(s.d_x) + row * d;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nch) {
            float x[8], o[8];
            unpack8<T>(xr[k], x);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (g[k][i] - x[i] * inv * dot) * scale;
            store8<T>(out + ch * 8, o);
        }
    }
    if (a.peers.world > 1) tail_signal_done(a);
    if (a.sums.counter) tail_finish_losses(a);
}

// ---------------------------------------------------------------------------------------------
// loss sums: single CTA, fixed-order tree (deterministic)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_loss_sums(SumsArgs a) {
    __shared__ float red[3][1024];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int64_t i = threadIdx.x; i < a.m; i += 1024) {
        if (a.lse_u) s0 += a.lse_u[i] + a.lse_v[i] - 2.0f * a.diag[i];
        if (a.cos_img) s1 += 1.0f - a.cos_img[i];
        if (a.cos_txt) s2 += 1.0f - a.cos_txt[i];
    }
    red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
            red[2][threadIdx.x] += red[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (a.out3) { a.out3[0] = red[0][0]; a.out3[1] = red[1][0]; a.out3[2] = red[2][0]; }
        if (a.zero_counter) *a.zero_counter = 0u;
        if (a.losses) {   // single-GPU composite: finished losses {total, img, txt, con}
            const float l_con = red[0][0] / (2.0f * (float)a.m);
            const float l_img = a.cos_img ? red[1][0] / (float)a.m : 0.f;
            const float l_txt = a.cos_txt ? red[2][0] / (float)a.m : 0.f;
            a.losses[0] = a.w_img * l_img + a.w_txt * l_txt + a.w_con * l_con;
            a.losses[1] = l_img; a.losses[2] = l_txt; a.losses[3] = l_con;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int launch_loss_prep(const PrepArgs& a, int nsides, int dtype, int op_dtype, cudaStream_t st) {
    if (a.m == 0) return DCLIP_OK;
    dim3 grid((unsigned)ceil_div(a.m, kThreads / 32), nsides);
    if (op_dtype == DCLIP_F32) {
        DCLIP_DISPATCH_DTYPE(dtype, (k_loss_prep<T, float><<<grid, kThreads, 0, st>>>(a)));
    } else if (op_dtype == DCLIP_F16) {
        DCLIP_DISPATCH_DTYPE(dtype, (k_loss_prep<T, __half><<<grid, kThreads, 0, st>>>(a)));
    } else if (op_dtype == DCLIP_BF16) {
        DCLIP_DISPATCH_DTYPE(dtype, (k_loss_prep<T, __nv_bfloat16><<<grid, kThreads, 0, st>>>(a)));
    } else {
        return fail(DCLIP_ERR_INVALID, "loss_prep: bad op_dtype %d", op_dtype);
    }
    DCLIP_LAUNCH_OK("k_loss_prep");
    return DCLIP_OK;
}

template <typename T, typename OT>
static void launch_prep_pair_nc(const PrepArgs& a, cudaStream_t st) {
    const dim3 grid((unsigned)ceil_div(a.m, 4));
    switch ((a.d + 255) / 256) {
        case 1: k_loss_prep_pair_nc<T, OT, 1><<<grid, 128, 0, st>>>(a); break;
        case 2: k_loss_prep_pair_nc<T, OT, 2><<<grid, 128, 0, st>>>(a); break;
        case 3: k_loss_prep_pair_nc<T, OT, 3><<<grid, 128, 0, st>>>(a); break;
        default: k_loss_prep_pair_nc<T, OT, 4><<<grid, 128, 0, st>>>(a); break;
    }
}

bool loss_prep_pair_pushes(int64_t d, int dtype, int op_dtype) { return d <= 1024 && op_dtype == DCLIP_F16 && dtype != DCLIP_F32; }

int launch_loss_prep_pair(const PrepArgs& a, int dtype, int op_dtype, cudaStream_t st) {
    if (a.m == 0) return DCLIP_OK;
    if (a.peers.world > 1 && !loss_prep_pair_pushes(a.d, dtype, op_dtype))
        return fail(DCLIP_ERR_UNSUPPORTED, "loss_prep_pair: the in-kernel peer exchange needs 16-bit inputs and d <= 1024");
    if (a.d <= 1024 && op_dtype == DCLIP_F16 && dtype != DCLIP_F32) {        // the tensor tier's shapes: register-resident rows
        if (dtype == DCLIP_BF16) launch_prep_pair_nc<__nv_bfloat16, __half>(a, st);
        else launch_prep_pair_nc<__half, __half>(a, st);
        DCLIP_LAUNCH_OK("k_loss_prep_pair_nc");
        return DCLIP_OK;
    }
    dim3 grid((unsigned)ceil_div(a.m, kThreads / 32));
    if (op_dtype == DCLIP_F32) {
        DCLIP_DISPATCH_DTYPE(dtype, (k_loss_prep_pair<T, float><<<grid, kThreads, 0, st>>>(a)));
    } else if (op_dtype == DCLIP_F16) {
        DCLIP_DISPATCH_DTYPE(dtype, (k_loss_prep_pair<T, __half><<<grid, kThreads, 0, st>>>(a)));
    } else {
        return fail(DCLIP_ERR_INVALID, "loss_prep_pair: bad op_dtype %d", op_dtype);
    }
    DCLIP_LAUNCH_OK("k_loss_prep_pair");
    return DCLIP_OK;
}

int launch_cos_rows(const CosArgs& a, int dtype, cudaStream_t st) {
    if (a.m == 0) return DCLIP_OK;
    dim3 grid((unsigned)ceil_div(a.m, kThreads / 32), 2);
    DCLIP_DISPATCH_DTYPE(dtype, (k_cos_rows<T><<<grid, kThreads, 0, st>>>(a)));
    DCLIP_LAUNCH_OK("k_cos_rows");
    return DCLIP_OK;
}

int* tc_error_flag();   // loss_tc.cu

int launch_sharded_losses(const float* stat, int world, int64_t stride, int64_t off, int64_t B, int has_img, int has_txt, float w_img,
                          float w_txt, float w_con, float* losses, cudaStream_t st, const PeerCtx* peers) {
    PeerCtx pc{};
    if (peers) pc = *peers;
    k_sharded_losses<<<1, 32, 0, st>>>(stat, world, stride, off, (float)B, has_img, has_txt, w_img, w_txt, w_con, losses, pc, tc_error_flag());
    DCLIP_LAUNCH_OK("k_sharded_losses");
    return DCLIP_OK;
}

int launch_tc_finalize(const FinalizeArgs& a, cudaStream_t st) {
    k_tc_finalize<<<kFinCtas, kFinThreads, 0, st>>>(a);
    DCLIP_LAUNCH_OK("k_tc_finalize");
    return DCLIP_OK;
}

static int simt_cpt(int64_t d) {   // columns per thread of the score tile (TJ = 16*CPT)
    const size_t need2 = (size_t)(kTI + 32) * (d + 4) * 4 + 32 * (kTI + 4) * 4;
    const size_t need1 = (size_t)(kTI + 16) * (d + 4) * 4 + 16 * (kTI + 4) * 4;
    if (need2 <= 220 * 1024) return 2;
    if (need1 <= 220 * 1024) return 1;
    return 0;
}

int simt_rowlse_nsplit(int64_t m, int64_t n, int64_t d) {
    const int cpt = simt_cpt(d);
    if (cpt == 0) return 0;
    const int64_t ntiles = ceil_div(n, 16 * cpt);
    const int64_t rowtiles = ceil_div(m, kTI);
    int64_t want = ceil_div(2 * (int64_t)sm_count(), rowtiles);   // aim for >= 2 CTAs per SM overall
    if (want < 1) want = 1;
    if (want > ntiles) want = ntiles;
    if (want > 64) want = 64;
    return (int)want;
}

int launch_rowlse_simt(LseArgs a, int nsides, cudaStream_t st) {
    if (a.m == 0) return DCLIP_OK;
    const int cpt = simt_cpt(a.d);
    if (cpt == 0) return fail(DCLIP_ERR_UNSUPPORTED, "FP32 loss path supports d <= 1024 (got %lld)", (long long)a.d);
    const int TJ = 16 * cpt;
    const int64_t ntiles = ceil_div(a.n, TJ);
    a.tiles_per_split = (int)ceil_div(ntiles, a.nsplit);
    const size_t smem = (size_t)(kTI + TJ) * (a.d + 4) * 4;
    dim3 grid((unsigned)ceil_div(a.m, kTI), a.nsplit, nsides);
    if (cpt == 2) {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_rowlse_partial<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_rowlse_partial<2><<<grid, kThreads, smem, st>>>(a);
    } else {
        DCLIP_CUDA_OK(cudaFuncSetAttribute(k_rowlse_partial<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_rowlse_partial<1><<<grid, kThreads, smem, st>>>(a);
    }
    DCLIP_LAUNCH_OK("k_rowlse_partial");
    dim3 g2((unsigned)ceil_div(a.m, 256), nsides);
    k_lse_combine<<<g2, 256, 0, st>>>(a);
    DCLIP_LAUNCH_OK("k_lse_combine");
    return DCLIP_OK;
}

template <int CPT, int NC>
static int launch_grad_inst(const GradArgs& a, int nsides, cudaStream_t st) {
    constexpr int TJ = 16 * CPT;
    const size_t smem = ((size_t)(kTI + TJ) * (a.d + 4) + TJ * (kTI + 4)) * 4;
    DCLIP_CUDA_OK(cudaFuncSetAttribute(k_infonce_grad<CPT, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(a.m, kTI), 1, nsides);
    k_infonce_grad<CPT, NC><<<grid, kThreads, smem, st>>>(a);
    DCLIP_LAUNCH_OK("k_infonce_grad");
    return DCLIP_OK;
}

int launch_grad_simt(const GradArgs& a, int nsides, cudaStream_t st) {
    if (a.m == 0) return DCLIP_OK;
    const int cpt = simt_cpt(a.d);
    if (cpt == 0) return fail(DCLIP_ERR_UNSUPPORTED, "FP32 loss path supports d <= 1024 (got %lld)", (long long)a.d);
    const int nc = (int)ceil_div(a.d, kThreads);
    if (cpt == 2) {
        switch (nc) {
            case 1: return launch_grad_inst<2, 1>(a, nsides, st);
            case 2: return launch_grad_inst<2, 2>(a, nsides, st);
            case 3: return launch_grad_inst<2, 3>(a, nsides, st);
            case 4: return launch_grad_inst<2, 4>(a, nsides, st);
        }
    } else {
        switch (nc) {
            case 1: return launch_grad_inst<1, 1>(a, nsides, st);
            case 2: return launch_grad_inst<1, 2>(a, nsides, st);
            case 3: return launch_grad_inst<1, 3>(a, nsides, st);
            case 4: return launch_grad_inst<1, 4>(a, nsides, st);
        }
    }
    return fail(DCLIP_ERR_UNSUPPORTED, "FP32 loss path supports d <= 1024 (got %lld)", (long long)a.d);
}

template <typename T>
static void launch_tail_nc(const TailArgs& a, int nsides, cudaStream_t st) {
    const dim3 grid((unsigned)ceil_div(a.m, 4), nsides);
    switch ((a.d + 255) / 256) {
        case 1: k_loss_grad_tail_nc<T, 1><<<grid, 128, 0, st>>>(a); break;
        case 2: k_loss_grad_tail_nc<T, 2><<<grid, 128, 0, st>>>(a); break;
        case 3: k_loss_grad_tail_nc<T, 3><<<grid, 128, 0, st>>>(a); break;
        default: k_loss_grad_tail_nc<T, 4><<<grid, 128, 0, st>>>(a); break;
    }
}

int launch_grad_tail(const TailArgs& a, int nsides, int dtype, cudaStream_t st) {
    if (a.m == 0) return DCLIP_OK;
    if (a.d <= 1024 && dtype != DCLIP_F32) {
        if (dtype == DCLIP_BF16) launch_tail_nc<__nv_bfloat16>(a, nsides, st);
        else launch_tail_nc<__half>(a, nsides, st);
        DCLIP_LAUNCH_OK("k_loss_grad_tail_nc");
        return DCLIP_OK;
    }
    dim3 grid((unsigned)ceil_div(a.m, kThreads / 32), nsides);
    DCLIP_DISPATCH_DTYPE(dtype, (k_loss_grad_tail<T><<<grid, kThreads, 0, st>>>(a)));
    DCLIP_LAUNCH_OK("k_loss_grad_tail");
    return DCLIP_OK;
}

int launch_loss_sums(const SumsArgs& a, cudaStream_t st) {
    k_loss_sums<<<1, 1024, 0, st>>>(a);
    DCLIP_LAUNCH_OK("k_loss_sums");
    return DCLIP_OK;
}

}  // namespace dclip
