// C-ABI of the distillation objective: argument checking, workspace carving, path choice.
#include <cstdlib>

#include "common.cuh"
#include "loss_internal.h"

using namespace dclip;

namespace dclip { int* tc_error_flag(); }
namespace {
int* tc_error_flag_api() { return dclip::tc_error_flag(); }

#ifndef DCLIP_TENSOR_PATH_READY
#define DCLIP_TENSOR_PATH_READY 1   // validated on B200 (tests/test_loss_tensor_gpu.py)
#endif

// The ONE place DCLIP_PATH_AUTO is resolved (dclip_resolve_loss_path exports it to the host layer, which uses it for
// the single-GPU and the sharded entry points alike): the tcgen05 path for 16-bit inputs / fp16 operands when the
// shape is supported on an sm_100 device and the batch fills at least one 128-row tile, the exact FP32 path
// otherwise (fp32 inputs always: the 1e-5 tier; tiny batches: launch-bound either way, FP32 is exact).
int resolve_loss_path(int path, int dtype, int64_t B, int64_t d, float temperature) {
    if (path != DCLIP_PATH_AUTO) return path;
    if (dtype != DCLIP_F32 && B >= 128 && DCLIP_TENSOR_PATH_READY && tc_loss_shape_ok(d, temperature) && device_is_sm100()) return DCLIP_PATH_TENSOR;
    return DCLIP_PATH_FP32;
}
// building blocks: the operand dtype already says which tier produced the operands
int resolve_block_path(int path, int op_dtype) {
    if (path != DCLIP_PATH_AUTO) return path;
    return op_dtype == DCLIP_F16 ? DCLIP_PATH_TENSOR : DCLIP_PATH_FP32;
}

// C-ABI peer description -> device-side context (world <= 1 or NULL: no in-kernel exchange)
int make_peer_ctx(const dclip_peers* peers, PeerCtx* out) {
    *out = PeerCtx{};
    if (!peers || peers->world <= 1) return DCLIP_OK;
    DCLIP_REQUIRE(peers->bases && peers->world <= kP2PMaxWorld && peers->rank >= 0 && peers->rank < peers->world, "peers: bad world / rank");
    for (int r = 0; r < peers->world; ++r) {
        DCLIP_REQUIRE(peers->bases[r] != nullptr, "peers: null base for rank %d", r);
        out->base[r] = static_cast<char*>(peers->bases[r]);
    }
    out->world = peers->world; out->rank = peers->rank; out->flags_offset = peers->flags_offset;
    return DCLIP_OK;
}
bool inside(const PeerCtx& pc, const void* p) {      // p lies inside this rank's symmetric buffer (below the flag block)
    const char* c = static_cast<const char*>(p);
    return c >= pc.base[pc.rank] && c < pc.base[pc.rank] + pc.flags_offset;
}

int check_common(int64_t m, int64_t d) {
    DCLIP_REQUIRE(m >= 0 && d > 0, "bad shape m=%lld d=%lld", (long long)m, (long long)d);
    DCLIP_REQUIRE(d % 8 == 0, "d must be a multiple of 8 (got %lld)", (long long)d);
    return DCLIP_OK;
}

}  // namespace

namespace dclip {
// out = x * (*scale): the upstream gradient of the scalar loss (loss.backward(), training/train_contrastive_teacher.py:361) applied to
// the gradients the fused forward+backward call already produced; the factor stays on the device (no host sync, graph-capturable).
template <typename T>
__global__ void __launch_bounds__(256) k_scale_grad(const T* __restrict__ x, int64_t groups, const float* __restrict__ scale, T* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= groups) return;
    const float s = *scale;
    float v[8];
    load8<T>(x + g * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= s;
    store8<T>(out + g * 8, v);
}
}  // namespace dclip

extern "C" {

int dclip_loss_prep(const void* x, const void* teacher, int64_t m, int64_t d, int dtype, void* x_hat, int op_dtype,
                    float* inv_norm, float* cos, void* stream) {
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(x && x_hat && inv_norm, "loss_prep: null pointer");
    DCLIP_REQUIRE(aligned16(x) && aligned16(x_hat) && (!teacher || aligned16(teacher)), "loss_prep: pointers must be 16-byte aligned");
    PrepArgs a{};
    a.s[0] = PrepSide{x, teacher, x_hat, inv_norm, cos};
    a.m = m; a.d = d;
    return launch_loss_prep(a, 1, dtype, op_dtype, static_cast<cudaStream_t>(stream));
}

size_t dclip_infonce_rowlse_workspace_bytes(int64_t m, int64_t n, int64_t d, int path) {
    int ns = simt_rowlse_nsplit(m, n, d);
    if (path == DCLIP_PATH_TENSOR) ns = stream_nsplit(m, n, 1);
    if (ns < 1) ns = 1;
    return 2 * align_up((size_t)ns * m * sizeof(float), 256) + 256;
}

int dclip_infonce_rowlse(const void* x_hat, int64_t m, const void* y_hat, int64_t n, int64_t d, int op_dtype,
                         float temperature, int64_t row_offset, float* lse, float* diag, void* workspace,
                         size_t workspace_bytes, int path, void* stream) {
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(x_hat && y_hat && lse && n > 0, "rowlse: null pointer or empty n");
    DCLIP_REQUIRE(temperature > 0.f, "rowlse: temperature must be > 0");
    path = resolve_block_path(path, op_dtype);
    if (path == DCLIP_PATH_FP32) {
        DCLIP_REQUIRE(op_dtype == DCLIP_F32, "rowlse: FP32 path needs float32 operands");
        LseArgs a{};
        a.nsplit = simt_rowlse_nsplit(m, n, d);
        if (a.nsplit == 0) return fail(DCLIP_ERR_UNSUPPORTED, "FP32 loss path supports d <= 1024 (got %lld)", (long long)d);
        Workspace ws(workspace, workspace_bytes);
        float* pm = static_cast<float*>(ws.take((size_t)a.nsplit * m * sizeof(float)));
        float* pl = static_cast<float*>(ws.take((size_t)a.nsplit * m * sizeof(float)));
        if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "rowlse: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
        a.s[0] = LseSide{static_cast<const float*>(x_hat), static_cast<const float*>(y_hat), pm, pl, lse, diag};
        a.m = m; a.n = n; a.d = d; a.temperature = temperature; a.row_offset = row_offset;
        return launch_rowlse_simt(a, 1, static_cast<cudaStream_t>(stream));
    }
    if (path == DCLIP_PATH_TENSOR) {
        DCLIP_REQUIRE(op_dtype == DCLIP_F16, "rowlse: tensor path needs float16 operands (dclip_loss_prep with op_dtype = DCLIP_F16)");
        if (!tc_loss_shape_ok(d, temperature))
            return fail(DCLIP_ERR_UNSUPPORTED, "rowlse: tensor path needs d %% 128 == 0, 128 <= d <= 768 and temperature >= 0.025 (d=%lld tau=%g)", (long long)d, temperature);
        DCLIP_REQUIRE(aligned16(x_hat) && aligned16(y_hat), "rowlse: operands must be 16-byte aligned");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int ns = stream_nsplit(m, n, 1);
        Workspace ws(workspace, workspace_bytes);
        float* part = static_cast<float*>(ws.take((size_t)ns * m * sizeof(float)));
        if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "rowlse: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
        TcArgs a{};
        a.s[0] = TcSide{nullptr, nullptr, part};
        a.m = m; a.n = n; a.d = d; a.inv_tau = 1.0f / temperature; a.row_offset = row_offset;
        const void* xs[2] = {x_hat, x_hat};
        const void* ys[2] = {y_hat, y_hat};
        if (int e = launch_stream_tc(0, 1, xs, ys, a, ns, st)) return e;
        if (int e = launch_tc_lse_combine(part, ns, m, a.inv_tau, lse, st)) return e;
        if (diag) return launch_tc_diag(x_hat, y_hat, m, n, d, a.inv_tau, row_offset, diag, st);
        return DCLIP_OK;
    }
    return fail(DCLIP_ERR_UNSUPPORTED, "rowlse: path %d not available", path);
}

size_t dclip_infonce_grad_workspace_bytes(int64_t m, int64_t n, int64_t d, int path) {
    if (path == DCLIP_PATH_TENSOR) {
        const int ns = tc_nsplit(m, n, d, 1);
        return ns > 1 ? align_up((size_t)ns * m * d * sizeof(float), 256) + 256 : 256;
    }
    return 256;
}

int dclip_infonce_grad(const void* x_hat, int64_t m, const void* y_hat, int64_t n, int64_t d, int op_dtype,
                       float temperature, int64_t row_offset, const float* lse_x, const float* lse_y, float coef,
                       float* dx_hat, void* workspace, size_t workspace_bytes, int path, void* stream) {
    (void)workspace; (void)workspace_bytes;
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(x_hat && y_hat && lse_x && lse_y && dx_hat && n > 0, "infonce_grad: null pointer or empty n");
    DCLIP_REQUIRE(temperature > 0.f, "infonce_grad: temperature must be > 0");
    path = resolve_block_path(path, op_dtype);
    if (path == DCLIP_PATH_FP32) {
        DCLIP_REQUIRE(op_dtype == DCLIP_F32, "infonce_grad: FP32 path needs float32 operands");
        GradArgs a{};
        a.s[0] = GradSide{static_cast<const float*>(x_hat), static_cast<const float*>(y_hat), lse_x, lse_y, dx_hat};
        a.m = m; a.n = n; a.d = d; a.temperature = temperature; a.row_offset = row_offset;
        a.coef_over_tau = coef / temperature;
        return launch_grad_simt(a, 1, static_cast<cudaStream_t>(stream));
    }
    if (path == DCLIP_PATH_TENSOR) {
        DCLIP_REQUIRE(op_dtype == DCLIP_F16, "infonce_grad: tensor path needs float16 operands");
        if (!tc_loss_shape_ok(d, temperature))
            return fail(DCLIP_ERR_UNSUPPORTED, "infonce_grad: tensor path needs d %% 128 == 0, 128 <= d <= 768 and temperature >= 0.025 (d=%lld tau=%g)", (long long)d, temperature);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int ns = tc_nsplit(m, n, d, 1);
        float* out = dx_hat;
        if (ns > 1) {
            Workspace ws(workspace, workspace_bytes);
            out = static_cast<float*>(ws.take((size_t)ns * m * d * sizeof(float)));
            if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "infonce_grad: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
        }
        TcArgs a{};
        a.s[0] = TcSide{lse_x, lse_y, out};
        a.m = m; a.n = n; a.d = d; a.inv_tau = 1.0f / temperature; a.row_offset = row_offset;
        a.out_scale = coef / temperature * (1.0f / 8192.0f);
        const void* xs[2] = {x_hat, x_hat};
        const void* ys[2] = {y_hat, y_hat};
        if (int e = launch_infonce_bwd_tc(1, xs, ys, a, ns, st)) return e;
        if (ns > 1) return launch_tc_grad_combine(out, ns, m, d, dx_hat, st);
        return DCLIP_OK;
    }
    return fail(DCLIP_ERR_UNSUPPORTED, "infonce_grad: path %d not available", path);
}

size_t dclip_infonce_grad_pair_workspace_bytes(int64_t B, int64_t d, int path) {
    if (path == DCLIP_PATH_TENSOR) {
        const int ns = tc_nsplit(B, B, d, 2);
        return ns > 1 ? align_up((size_t)2 * ns * B * d * sizeof(float), 256) + 256 : 256;
    }
    return 256;
}

int dclip_infonce_grad_pair(const void* u_hat, const void* v_hat, int64_t B, int64_t d, int op_dtype, float temperature,
                            const float* lse_u, const float* lse_v, float coef, float* du_hat, float* dv_hat, void* workspace,
                            size_t workspace_bytes, int path, void* stream) {
    if (int e = check_common(B, d)) return e;
    DCLIP_REQUIRE(u_hat && v_hat && lse_u && lse_v && du_hat && dv_hat && B > 0, "infonce_grad_pair: null pointer or empty batch");
    DCLIP_REQUIRE(temperature > 0.f, "infonce_grad_pair: temperature must be > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    path = resolve_block_path(path, op_dtype);
    if (path == DCLIP_PATH_FP32) {
        DCLIP_REQUIRE(op_dtype == DCLIP_F32, "infonce_grad_pair: FP32 path needs float32 operands");
        GradArgs ga{};
        ga.s[0] = GradSide{(const float*)u_hat, (const float*)v_hat, lse_u, lse_v, du_hat};
        ga.s[1] = GradSide{(const float*)v_hat, (const float*)u_hat, lse_v, lse_u, dv_hat};
        ga.m = B; ga.n = B; ga.d = d; ga.temperature = temperature; ga.row_offset = 0;
        ga.coef_over_tau = coef / temperature;
        return launch_grad_simt(ga, 2, st);
    }
    DCLIP_REQUIRE(op_dtype == DCLIP_F16, "infonce_grad_pair: tensor path needs float16 operands");
    if (!tc_loss_shape_ok(d, temperature))
        return fail(DCLIP_ERR_UNSUPPORTED, "infonce_grad_pair: tensor path needs d %% 128 == 0, 128 <= d <= 768 and temperature >= 0.025 (d=%lld tau=%g)", (long long)d, temperature);
    const int ns = tc_nsplit(B, B, d, 2);
    float* du = du_hat;
    float* dv = dv_hat;
    if (ns > 1) {
        Workspace ws(workspace, workspace_bytes);
        du = static_cast<float*>(ws.take((size_t)2 * ns * B * d * sizeof(float)));
        if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "infonce_grad_pair: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
        dv = du + (size_t)ns * B * d;
    }
    TcArgs ga{};
    ga.s[0] = TcSide{lse_u, lse_v, du};
    ga.s[1] = TcSide{lse_v, lse_u, dv};
    ga.m = B; ga.n = B; ga.d = d; ga.inv_tau = 1.0f / temperature; ga.row_offset = 0;
    ga.out_scale = coef / temperature * (1.0f / 8192.0f);
    const void* xs[2] = {u_hat, v_hat};
    const void* ys[2] = {v_hat, u_hat};
    if (int e = launch_infonce_bwd_tc(2, xs, ys, ga, ns, st)) return e;
    if (ns > 1) {
        if (int e = launch_tc_grad_combine(du, ns, B, d, du_hat, st)) return e;
        if (int e = launch_tc_grad_combine(dv, ns, B, d, dv_hat, st)) return e;
    }
    return DCLIP_OK;
}

size_t dclip_infonce_coeffs_bytes(int64_t B) { return B > 0 ? tc_gmat_bytes(B) : 0; }
int64_t dclip_infonce_coeffs_ld(int64_t B) { return tc_gmat_ld(B); }
size_t dclip_infonce_coeffs_dot_bytes(int64_t B) { return B > 0 ? tc_gmat_dot_floats(B) * sizeof(float) : 0; }

int dclip_infonce_coeffs(const void* u_hat, const void* v_hat, int64_t B, int64_t d, float temperature, const float* lse_u,
                         const float* lse_v, void* coeffs, float* dots, void* stream) {
    if (int e = check_common(B, d)) return e;
    DCLIP_REQUIRE(u_hat && v_hat && lse_u && lse_v && coeffs && dots, "infonce_coeffs: null pointer");
    DCLIP_REQUIRE(B > 128, "infonce_coeffs: needs more than 128 rows (got %lld)", (long long)B);
    if (!tc_loss_shape_ok(d, temperature))
        return fail(DCLIP_ERR_UNSUPPORTED, "infonce_coeffs: tensor path needs d %% 128 == 0, 128 <= d <= 768 and temperature >= 0.025 (d=%lld tau=%g)", (long long)d, temperature);
    DCLIP_REQUIRE(aligned16(u_hat) && aligned16(v_hat) && aligned16(coeffs), "infonce_coeffs: pointers must be 16-byte aligned");
    return launch_gmat_coeffs(u_hat, v_hat, B, d, lse_u, lse_v, 1.0f / temperature, coeffs, dots, static_cast<cudaStream_t>(stream));
}

int dclip_infonce_grad_from_coeffs(const void* coeffs, const float* dots, const void* u_hat, const void* v_hat, int64_t B, int64_t d,
                                   float temperature, float coef, float* du_hat, float* dv_hat, const dclip_grad_tail* tail, void* stream) {
    if (int e = check_common(B, d)) return e;
    DCLIP_REQUIRE(coeffs && u_hat && v_hat && B > 128 && temperature > 0.f, "infonce_grad_from_coeffs: bad arguments");
    DCLIP_REQUIRE(tail || (du_hat && dv_hat), "infonce_grad_from_coeffs: give du_hat / dv_hat or a gradient tail");
    const float alpha = coef / temperature * (1.0f / 8192.0f);
    if (!tail) return launch_gmat_gemms(coeffs, dots, u_hat, v_hat, B, d, alpha, du_hat, dv_hat, nullptr, static_cast<cudaStream_t>(stream));
    DCLIP_REQUIRE(dots && tail->s_img && tail->s_txt && tail->inv_u && tail->inv_v && tail->d_s_img && tail->d_s_txt,
                  "infonce_grad_from_coeffs: incomplete gradient tail");
    DCLIP_REQUIRE((tail->t_img == nullptr) == (tail->t_txt == nullptr), "infonce_grad_from_coeffs: give both teachers or neither");
    DCLIP_REQUIRE(!tail->t_img || (tail->cos_img && tail->cos_txt && tail->tinv_img && tail->tinv_txt),
                  "infonce_grad_from_coeffs: teachers need their cosine terms and inverse norms");
    GmatTail gt{};
    gt.s[0] = GmatTailSide{tail->s_img, tail->t_img, tail->inv_u, tail->cos_img, tail->tinv_img, tail->d_s_img, tail->w_img_over_B};
    gt.s[1] = GmatTailSide{tail->s_txt, tail->t_txt, tail->inv_v, tail->cos_txt, tail->tinv_txt, tail->d_s_txt, tail->w_txt_over_B};
    gt.grad_scale = tail->grad_scale; gt.dtype = tail->dtype;
    return launch_gmat_gemms(coeffs, dots, u_hat, v_hat, B, d, alpha, nullptr, nullptr, &gt, static_cast<cudaStream_t>(stream));
}

int dclip_loss_grad_tail(const void* x, const void* teacher, const float* inv_norm, const float* dx_hat, int64_t m,
                         int64_t d, int dtype, float w_cos_over_B, float grad_scale, void* d_x, void* stream) {
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(x && inv_norm && d_x, "loss_grad_tail: null pointer");
    TailArgs a{};
    a.s[0] = TailSide{x, teacher, inv_norm, dx_hat, d_x, nullptr};
    a.m = m; a.d = d; a.w_cos_over_B[0] = w_cos_over_B; a.grad_scale = grad_scale;
    return launch_grad_tail(a, 1, dtype, static_cast<cudaStream_t>(stream));
}

int dclip_loss_partial_sums(const float* lse_u, const float* lse_v, const float* diag, const float* cos_img,
                            const float* cos_txt, int64_t m, float* out3, void* stream) {
    DCLIP_REQUIRE(out3 && m >= 0, "loss_partial_sums: null pointer");
    DCLIP_REQUIRE((lse_u && lse_v && diag) || (!lse_u && !lse_v && !diag), "loss_partial_sums: give lse_u, lse_v and diag together or none");
    SumsArgs a{lse_u, lse_v, diag, cos_img, cos_txt, m, out3, nullptr, 0.f, 0.f, 0.f};
    return launch_loss_sums(a, static_cast<cudaStream_t>(stream));
}

int dclip_tensor_path_ready(void) { return DCLIP_TENSOR_PATH_READY; }

int dclip_resolve_loss_path(int path, int dtype, int64_t B, int64_t d, float temperature) { return resolve_loss_path(path, dtype, B, d, temperature); }

int dclip_debug_pipeline_error(void* stream) { return tc_check_error_flag(static_cast<cudaStream_t>(stream)); }

int dclip_distill_loss_launch_count(int64_t m, int64_t n, int64_t d, int path, int sharded) {
    if (path == DCLIP_PATH_AUTO) path = DCLIP_PATH_FP32;   // callers pass the resolved path (dclip_resolve_loss_path)
    if (path == DCLIP_PATH_FP32 && !sharded && small_loss_shape_ok(m, d) && m == n) return 1;      // k_distill_small
    if (path == DCLIP_PATH_FP32) return sharded ? 11 : 6;   // prep, rowlse+combine, sums, grad, tail (x2 sides when sharded)
    // tensor path, single GPU: prep_pair, one-pass forward tiles, finalize, backward tiles (+ partial combines), tail
    if (!sharded && tc_gmat_ok(m, d)) return 5;   // prep_pair, forward tiles, finalize, G tiles, pair GEMM (dU, dV) with the fused tail
    if (!sharded) return 5 + (tc_nsplit(m, n, d, 2) > 1 ? 2 : 0);
    // fused sharded phases: prep, forward tiles, finalize, global losses, backward tiles (+ combine), tail
    // + the three peer-memory pushes (operands, statistics, step-done) that replace the NCCL all-gathers - unless the exchange
    // happens inside the kernels (sharded == 2: the `peers` argument of the three phases)
    return 6 + (tc_nsplit(m, n, d, 2) > 1 ? 1 : 0) + (sharded == 2 ? 0 : 3);
}

// ---- fused per-rank phases of the sharded objective (tensor path) -----------------------------
size_t dclip_sharded_loss_workspace_bytes(int64_t m, int64_t n, int64_t d) {
    const int nf = stream_nsplit(m, n, 2), nb = tc_nsplit(m, n, d, 2);
    size_t total = 2 * align_up((size_t)nf * m * 4, 256);                 // forward partial row sums (two sides)
    total += 2 * align_up((size_t)m * d * 4, 256);                        // du_hat, dv_hat
    if (nb > 1) total += align_up((size_t)2 * nb * m * d * 4, 256);       // backward partials
    return total + 1024;
}

int dclip_sharded_loss_prep(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t m, int64_t d, int dtype,
                            float temperature, void* u_all, void* v_all, int64_t row_offset, float* inv_u, float* inv_v, float* cos_img,
                            float* cos_txt, float* diag, const dclip_peers* peers, void* stream) {
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(s_img && s_txt && u_all && v_all && inv_u && inv_v && cos_img && cos_txt && diag, "sharded_loss_prep: null pointer");
    DCLIP_REQUIRE((t_img == nullptr) == (t_txt == nullptr), "sharded_loss_prep: give both teachers or neither");
    PrepArgs pa{};
    pa.s[0] = PrepSide{s_img, t_img, static_cast<char*>(u_all) + (size_t)row_offset * d * 2, inv_u, cos_img};
    pa.s[1] = PrepSide{s_txt, t_txt, static_cast<char*>(v_all) + (size_t)row_offset * d * 2, inv_v, cos_txt};
    pa.m = m; pa.d = d; pa.diag = diag; pa.inv_tau = 1.0f / temperature;
    if (int e = make_peer_ctx(peers, &pa.peers)) return e;
    if (pa.peers.world > 1) {
        DCLIP_REQUIRE(inside(pa.peers, u_all) && inside(pa.peers, v_all), "sharded_loss_prep: u_all / v_all must live in the symmetric buffer");
        pa.error_flag = tc_error_flag_api();
    }
    return launch_loss_prep_pair(pa, dtype, DCLIP_F16, static_cast<cudaStream_t>(stream));
}

int dclip_sharded_loss_fwd(const void* u_all, const void* v_all, int64_t m, int64_t n, int64_t d, float temperature, int64_t row_offset,
                           const float* cos_img, const float* cos_txt, const float* diag, float* stat_row, void* workspace,
                           size_t workspace_bytes, const dclip_peers* peers, void* stream) {
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(u_all && v_all && diag && stat_row && n >= m && m > 0, "sharded_loss_fwd: bad arguments");
    if (!tc_loss_shape_ok(d, temperature))
        return fail(DCLIP_ERR_UNSUPPORTED, "sharded_loss_fwd: tensor path needs d %% 128 == 0, 128 <= d <= 768 and temperature >= 0.025");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nf = stream_nsplit(m, n, 2);
    Workspace ws(workspace, workspace_bytes);
    float* pu = static_cast<float*>(ws.take((size_t)nf * m * 4));
    float* pv = static_cast<float*>(ws.take((size_t)nf * m * 4));
    if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "sharded_loss_fwd: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    const char* ul = static_cast<const char*>(u_all) + (size_t)row_offset * d * 2;
    const char* vl = static_cast<const char*>(v_all) + (size_t)row_offset * d * 2;
    TcArgs fa{};
    fa.s[0] = TcSide{nullptr, nullptr, pu};
    fa.s[1] = TcSide{nullptr, nullptr, pv};
    fa.m = m; fa.n = n; fa.d = d; fa.inv_tau = 1.0f / temperature; fa.row_offset = row_offset;
    if (int e = make_peer_ctx(peers, &fa.peers)) return e;
    if (fa.peers.world > 1) DCLIP_REQUIRE(inside(fa.peers, stat_row), "sharded_loss_fwd: stat_row must live in the symmetric buffer");
    const void* xs[2] = {ul, vl};
    const void* ys[2] = {v_all, u_all};
    if (int e = launch_stream_tc(0, 2, xs, ys, fa, nf, st)) return e;
    FinalizeArgs fz{pu, pv, nf, m, fa.inv_tau, stat_row, stat_row + m,
                    SumsArgs{stat_row, stat_row + m, diag, cos_img, cos_txt, m, stat_row + 2 * m, nullptr, 0.f, 0.f, 0.f}};
    fz.peers = fa.peers;
    return launch_tc_finalize(fz, st);
}

int dclip_sharded_loss_bwd(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, const void* u_all, const void* v_all,
                           int64_t m, int64_t n, int64_t d, int dtype, float temperature, int64_t row_offset, int rank, int world,
                           const float* stat_all, const float* inv_u, const float* inv_v, float w_img, float w_txt, float w_con,
                           float grad_scale, float* losses, void* d_s_img, void* d_s_txt, void* workspace, size_t workspace_bytes,
                           const dclip_peers* peers, void* stream) {
    if (int e = check_common(m, d)) return e;
    DCLIP_REQUIRE(s_img && s_txt && u_all && v_all && stat_all && inv_u && inv_v && losses, "sharded_loss_bwd: null pointer");
    DCLIP_REQUIRE(world >= 1 && rank >= 0 && rank < world && n == m * world, "sharded_loss_bwd: n must equal m * world");
    DCLIP_REQUIRE((d_s_img == nullptr) == (d_s_txt == nullptr), "sharded_loss_bwd: give both gradients or neither");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t stride = 2 * m + 4;
    PeerCtx pc;
    if (int e = make_peer_ctx(peers, &pc)) return e;
    DCLIP_REQUIRE(pc.world <= 1 || d_s_img, "sharded_loss_bwd: the in-kernel peer exchange needs the gradient pass (its tail signals \"step done\")");
    if (int e = launch_sharded_losses(stat_all, world, stride, 2 * m, n, t_img != nullptr, t_txt != nullptr, w_img, w_txt, w_con, losses, st, &pc)) return e;
    if (!d_s_img) return DCLIP_OK;
    const int nb = tc_nsplit(m, n, d, 2);
    Workspace ws(workspace, workspace_bytes);
    ws.take((size_t)stream_nsplit(m, n, 2) * m * 4); ws.take((size_t)stream_nsplit(m, n, 2) * m * 4);   // (forward partials: same carve)
    float* du_hat = static_cast<float*>(ws.take((size_t)m * d * 4));
    float* dv_hat = static_cast<float*>(ws.take((size_t)m * d * 4));
    float* part = nb > 1 ? static_cast<float*>(ws.take((size_t)2 * nb * m * d * 4)) : nullptr;
    if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "sharded_loss_bwd: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    const bool need_con = w_con != 0.f;
    if (need_con) {
        const float* mine = stat_all + (int64_t)rank * stride;
        TcArgs ga{};
        ga.s[0] = TcSide{mine, stat_all + m, nb > 1 ? part : du_hat, m, stride};                       // lse_y = lse_v of every rank
        ga.s[1] = TcSide{mine + m, stat_all, nb > 1 ? part + (size_t)nb * m * d : dv_hat, m, stride};  // lse_y = lse_u of every rank
        ga.m = m; ga.n = n; ga.d = d; ga.inv_tau = 1.0f / temperature; ga.row_offset = row_offset;
        ga.out_scale = w_con / (2.0f * (float)n) / temperature * (1.0f / 8192.0f);
        const void* xs[2] = {static_cast<const char*>(u_all) + (size_t)row_offset * d * 2, static_cast<const char*>(v_all) + (size_t)row_offset * d * 2};
        const void* ys[2] = {v_all, u_all};
        if (int e = launch_infonce_bwd_tc(2, xs, ys, ga, nb, st)) return e;
        if (nb > 1)
            if (int e = launch_tc_grad_combine2(part, nb, m, d, du_hat, dv_hat, st)) return e;
    }
    TailArgs ta{};
    ta.m = m; ta.d = d; ta.grad_scale = grad_scale;
    ta.s[0] = TailSide{s_img, t_img, inv_u, need_con ? du_hat : nullptr, d_s_img, nullptr};
    ta.s[1] = TailSide{s_txt, t_txt, inv_v, need_con ? dv_hat : nullptr, d_s_txt, nullptr};
    ta.w_cos_over_B[0] = w_img / (float)n;
    ta.w_cos_over_B[1] = w_txt / (float)n;
    ta.peers = pc;
    return launch_grad_tail(ta, 2, dtype, st);
}

// ---- single-GPU composite -------------------------------------------------------------------
// workspace: u_hat, v_hat [B,d] operands | inv_u, inv_v, cos_img, cos_txt, lse_u, lse_v, diag [B]
//            | partials 2 x 2 x [nsplit, B] | du_hat, dv_hat [B,d] fp32
size_t dclip_distill_loss_workspace_bytes(int64_t B, int64_t d, int path) {
    if (path == DCLIP_PATH_AUTO && tc_loss_shape_ok(d, 0.05f)) path = DCLIP_PATH_TENSOR;   // AUTO may resolve to either: size for both
    int ns = simt_rowlse_nsplit(B, B, d);
    if (ns < 1) ns = 1;
    size_t total = 0;
    total += 2 * align_up((size_t)B * d * 4, 256);          // operands (fp32 worst case)
    total += 8 * align_up((size_t)B * 4, 256);              // per-row vectors
    total += 4 * align_up((size_t)ns * B * 4, 256);         // partials
    total += 2 * align_up((size_t)B * d * 4, 256);          // dx_hat
    if (path == DCLIP_PATH_TENSOR) {
        const int tns = tc_nsplit(B, B, d, 2), fns = stream_nsplit(B, B, 2);
        total += 4 * align_up((size_t)fns * B * 4, 256);
        if (tns > 1) total += align_up((size_t)2 * tns * B * d * 4, 256);
        // one-pass forward: row partials [stream_nsplit(1 side)][B] + column partials [ceil(B / 128)][B]
        total += align_up((size_t)stream_nsplit(B, B, 1) * B * 4, 256) + align_up((size_t)ceil_div(B, 128) * B * 4, 256);
        if (tc_gmat_ok(B, d))     // fp16 gradient coefficients G + row / column sums of g * s + the teachers' inverse norms
            total += align_up(tc_gmat_bytes(B), 256) + align_up(tc_gmat_dot_floats(B) * 4, 256) + 2 * align_up((size_t)B * 4, 256);
    }
    return total + 256 + 1024;     // + the split objective's sums / block counter
}

// phase 0: the whole objective.  Phases 1 / 2 split it at the point where the TEACHER embeddings are first
// needed, so that a caller whose inputs arrive over PCIe can overlap the teachers' copy with the contrastive
// part (dclip_distill_loss_students / _teachers; both phases carve the same workspace):
//   phase 1 (students only): normalise, forward tiles, log-sum-exp, backward tiles -> du_hat, dv_hat, lse, diag
//   phase 2 (teachers):      cosine terms, the four losses, gradient tail (normalise backward + cosine gradient)
// The teachers phase of the host-buffer step is issued per side (TailCtl::sides): the image side's tail runs while the text teacher is
// still crossing PCIe, the text side's tail finishes the losses and stores them into host memory (TailSums).
struct TailCtl { int sides; float* losses_host; unsigned int seq; };      // sides: 0 both (one launch), 1 image side only, 2 text side + losses
static int distill_loss_impl(int phase, const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t B,
                             int64_t d, int dtype, float temperature, float w_img, float w_txt, float w_con,
                             float grad_scale, float* losses, void* d_s_img, void* d_s_txt, float* stats,
                             void* workspace, size_t workspace_bytes, int path, void* stream, TailCtl ctl = TailCtl{0, nullptr, 0u}) {
    if (int e = check_common(B, d)) return e;
    DCLIP_REQUIRE(B > 0, "distill_loss: empty batch");
    DCLIP_REQUIRE(s_img && s_txt && (losses || phase == 1), "distill_loss: null pointer");
    DCLIP_REQUIRE((t_img == nullptr) == (t_txt == nullptr), "distill_loss: give both teachers or neither");
    DCLIP_REQUIRE((d_s_img == nullptr) == (d_s_txt == nullptr), "distill_loss: give both gradients or neither");
    DCLIP_REQUIRE(temperature > 0.f, "distill_loss: temperature must be > 0");
    DCLIP_REQUIRE(aligned16(s_img) && aligned16(s_txt) && aligned16(workspace), "distill_loss: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    path = resolve_loss_path(path, dtype, B, d, temperature);
    const bool tensor = path == DCLIP_PATH_TENSOR;
    if (tensor && !tc_loss_shape_ok(d, temperature))
        return fail(DCLIP_ERR_UNSUPPORTED, "distill_loss: tensor path needs d %% 128 == 0, 128 <= d <= 768 and temperature >= 0.025 (d=%lld tau=%g)", (long long)d, temperature);
    if (!tensor && path != DCLIP_PATH_FP32) return fail(DCLIP_ERR_INVALID, "distill_loss: unknown path %d", path);

    // small batches (the teacher step's B = 32, the reference's own distillation batch): the whole objective in ONE
    // launch of one 8-CTA cluster instead of six launch-latency-bound kernels (DCLIP_LOSS_SMALL=0: generic path)
    if (phase == 0 && !tensor && small_loss_shape_ok(B, d) && (B * (ceil_div(ceil_div(d, 8), 8) * 8 + 4) * 16 + B * (B + 1) * 16) < 200 * 1024) {
        static const bool off = [] { const char* e = getenv("DCLIP_LOSS_SMALL"); return e && e[0] == '0'; }();
        if (!off)
            return launch_distill_small(s_img, s_txt, t_img, t_txt, B, d, dtype, temperature, w_img, w_txt, w_con, grad_scale, losses, d_s_img,
                                        d_s_txt, stats, st);
    }
    const int op_dtype = tensor ? DCLIP_F16 : DCLIP_F32;
    const size_t op_size = tensor ? 2 : 4;
    Workspace ws(workspace, workspace_bytes);
    void* u_hat = ws.take((size_t)B * d * op_size);
    void* v_hat = ws.take((size_t)B * d * op_size);
    float* vec[8];
    for (auto& v : vec) v = static_cast<float*>(ws.take((size_t)B * 4));
    float *inv_u = vec[0], *inv_v = vec[1], *cos_img = vec[2], *cos_txt = vec[3], *lse_u = vec[4], *lse_v = vec[5], *diag = vec[6];
    const int ns = tensor ? stream_nsplit(B, B, 2) : simt_rowlse_nsplit(B, B, d);   // forward column splits
    const int nsb = tensor ? tc_nsplit(B, B, d, 2) : 1;                                 // backward column splits
    if (ns == 0) return fail(DCLIP_ERR_UNSUPPORTED, "FP32 loss path supports d <= 1024 (got %lld)", (long long)d);
    float* part[4];
    for (auto& p : part) p = static_cast<float*>(ws.take((size_t)ns * B * 4));
    float* du_hat = static_cast<float*>(ws.take((size_t)B * d * 4));
    float* dv_hat = static_cast<float*>(ws.take((size_t)B * d * 4));
    float* du_part = du_hat;
    float* dv_part = dv_hat;
    if (tensor && nsb > 1 && (d_s_img || phase != 0) && !tc_gmat_ok(B, d)) {
        du_part = static_cast<float*>(ws.take((size_t)2 * nsb * B * d * 4));   // [side][split][B][d], sides contiguous
        dv_part = du_part + (size_t)nsb * B * d;
    }
    // tensor path forward: ONE pass over S gives the row sums (lse_u) and the column sums (lse_v)
    const int ns1 = tensor ? stream_nsplit(B, B, 1) : 0;
    const int nib = (int)ceil_div(B, 128);
    float* rowpart = tensor ? static_cast<float*>(ws.take((size_t)ns1 * B * 4)) : nullptr;
    float* colpart = tensor ? static_cast<float*>(ws.take((size_t)nib * B * 4)) : nullptr;
    const bool gmat = tensor && tc_gmat_ok(B, d);
    void* g_buf = gmat ? ws.take(tc_gmat_bytes(B)) : nullptr;
    float* dot_ws = gmat ? static_cast<float*>(ws.take(tc_gmat_dot_floats(B) * 4)) : nullptr;
    float* tinv_u = gmat ? static_cast<float*>(ws.take((size_t)B * 4)) : nullptr;
    float* tinv_v = gmat ? static_cast<float*>(ws.take((size_t)B * 4)) : nullptr;
    // split objective: raw contrastive sum of the students phase [0..2] and the teachers phase's block counter [4]
    float* split_sums = static_cast<float*>(ws.take(64));
    unsigned int* tail_counter = reinterpret_cast<unsigned int*>(split_sums + 4);
    // one-call objective: the GEMM epilogue finishes the gradients (no fp32 dX_hat round trip, no tail launch)
    const char* ft_env = getenv("DCLIP_GMAT_FUSED_TAIL");      // "0": fp32 dX_hat + the tail kernel (A/B measurements, tests)
    const bool fused_tail = gmat && phase == 0 && d_s_img && w_con != 0.f && !(ft_env && ft_env[0] == '0');
    if (!ws.ok()) return fail(DCLIP_ERR_WORKSPACE, "distill_loss: workspace too small (%zu < %zu)", workspace_bytes, ws.used);
    if (stats) { lse_u = stats; lse_v = stats + B; }

    const bool do_students = phase != 2, do_teachers = phase != 1;
    // phase 2 with gradients: the gradient tail reads the raw student / teacher rows anyway and leaves the cosine terms (TailSide::cos_out);
    // its last block finishes the losses (TailSums) - one pass over the 4 x B x d rows and ONE launch behind the teachers' copy
    const bool cos_in_tail = phase == 2 && d_s_img && t_img;
    if (phase == 2 && !cos_in_tail) {
        // cosine terms of the rows (the norms of the students are phase 1's), then the four losses
        if (t_img) {
            CosArgs ca{};
            ca.s[0] = CosSide{s_img, t_img, inv_u, cos_img};
            ca.s[1] = CosSide{s_txt, t_txt, inv_v, cos_txt};
            ca.m = B; ca.d = d;
            if (int e = launch_cos_rows(ca, dtype, st)) return e;
        }
        SumsArgs sa{lse_u, lse_v, diag, t_img ? cos_img : nullptr, t_txt ? cos_txt : nullptr, B, nullptr, losses, w_img, w_txt, w_con};
        if (int e = launch_loss_sums(sa, st)) return e;
    }
    PrepArgs pa{};
    pa.s[0] = PrepSide{s_img, phase == 1 ? nullptr : t_img, u_hat, inv_u, cos_img, tinv_u};
    pa.s[1] = PrepSide{s_txt, phase == 1 ? nullptr : t_txt, v_hat, inv_v, cos_txt, tinv_v};
    pa.m = B; pa.d = d;
    if (!do_students) {
    } else if (tensor) {
        pa.diag = diag; pa.inv_tau = 1.0f / temperature;
        if (int e = launch_loss_prep_pair(pa, dtype, op_dtype, st)) return e;
    } else {
        if (int e = launch_loss_prep(pa, 2, dtype, op_dtype, st)) return e;
    }

    const void* xs2[2] = {u_hat, v_hat};
    const void* ys2[2] = {v_hat, u_hat};
    if (!do_students) {
    } else if (tensor) {
        TcArgs fa{};
        fa.s[0] = TcSide{nullptr, nullptr, rowpart};
        fa.col_out = colpart;
        fa.m = B; fa.n = B; fa.d = d; fa.inv_tau = 1.0f / temperature; fa.row_offset = 0;
        if (int e = launch_stream_tc(1, 1, xs2, ys2, fa, ns1, st)) return e;
        FinalizeArgs fz{rowpart, colpart, ns1, B, fa.inv_tau, lse_u, lse_v,
                        SumsArgs{lse_u, lse_v, diag, t_img ? cos_img : nullptr, t_txt ? cos_txt : nullptr, B, phase == 1 ? split_sums : nullptr,
                                 phase == 0 ? losses : nullptr, w_img, w_txt, w_con, phase == 1 ? tail_counter : nullptr},
                        nib};
        if (int e = launch_tc_finalize(fz, st)) return e;
    } else {
        LseArgs la{};
        la.s[0] = LseSide{(const float*)u_hat, (const float*)v_hat, part[0], part[1], lse_u, diag};
        la.s[1] = LseSide{(const float*)v_hat, (const float*)u_hat, part[2], part[3], lse_v, nullptr};
        la.m = B; la.n = B; la.d = d; la.temperature = temperature; la.row_offset = 0; la.nsplit = ns;
        if (int e = launch_rowlse_simt(la, 2, st)) return e;
    }

    if (!tensor && phase == 0) {
        SumsArgs sa{lse_u, lse_v, diag, t_img ? cos_img : nullptr, t_txt ? cos_txt : nullptr, B, nullptr, losses, w_img, w_txt, w_con};
        if (int e = launch_loss_sums(sa, st)) return e;
    } else if (!tensor && phase == 1) {      // raw contrastive sum for the teachers phase (the tensor path's finalize kernel leaves it)
        SumsArgs sa{lse_u, lse_v, diag, nullptr, nullptr, B, split_sums, nullptr, 0.f, 0.f, 0.f, tail_counter};
        if (int e = launch_loss_sums(sa, st)) return e;
    }

    if (d_s_img || phase == 1) {
        const bool need_con = w_con != 0.f;
        if (!do_students) {
        } else if (need_con && gmat) {
            GmatTail gt{};
            gt.s[0] = GmatTailSide{s_img, t_img, inv_u, cos_img, tinv_u, d_s_img, w_img / (float)B};
            gt.s[1] = GmatTailSide{s_txt, t_txt, inv_v, cos_txt, tinv_v, d_s_txt, w_txt / (float)B};
            gt.grad_scale = grad_scale; gt.dtype = dtype;
            if (int e = launch_infonce_bwd_gmat(u_hat, v_hat, B, d, lse_u, lse_v, 1.0f / temperature,
                                                w_con / (2.0f * (float)B) / temperature * (1.0f / 8192.0f), g_buf, dot_ws, du_hat, dv_hat,
                                                fused_tail ? &gt : nullptr, st))
                return e;
            if (fused_tail) return DCLIP_OK;
        } else if (need_con && tensor) {
            TcArgs ga{};
            ga.s[0] = TcSide{lse_u, lse_v, du_part};
            ga.s[1] = TcSide{lse_v, lse_u, dv_part};
            ga.m = B; ga.n = B; ga.d = d; ga.inv_tau = 1.0f / temperature; ga.row_offset = 0;
            ga.out_scale = w_con / (2.0f * (float)B) / temperature * (1.0f / 8192.0f);
            if (int e = launch_infonce_bwd_tc(2, xs2, ys2, ga, nsb, st)) return e;
            if (nsb > 1) {
                if (int e = launch_tc_grad_combine(du_part, nsb, B, d, du_hat, st)) return e;
                if (int e = launch_tc_grad_combine(dv_part, nsb, B, d, dv_hat, st)) return e;
            }
        } else if (need_con) {
            GradArgs ga{};
            ga.s[0] = GradSide{(const float*)u_hat, (const float*)v_hat, lse_u, lse_v, du_hat};
            ga.s[1] = GradSide{(const float*)v_hat, (const float*)u_hat, lse_v, lse_u, dv_hat};
            ga.m = B; ga.n = B; ga.d = d; ga.temperature = temperature; ga.row_offset = 0;
            ga.coef_over_tau = w_con / (2.0f * (float)B) / temperature;
            if (int e = launch_grad_simt(ga, 2, st)) return e;
        }
        if (!do_teachers || !d_s_img) return DCLIP_OK;
        TailArgs ta{};
        ta.m = B; ta.d = d; ta.grad_scale = grad_scale;
        ta.s[0] = TailSide{s_img, t_img, inv_u, need_con ? du_hat : nullptr, d_s_img, cos_in_tail ? cos_img : nullptr};
        ta.s[1] = TailSide{s_txt, t_txt, inv_v, need_con ? dv_hat : nullptr, d_s_txt, cos_in_tail ? cos_txt : nullptr};
        ta.w_cos_over_B[0] = w_img / (float)B;
        ta.w_cos_over_B[1] = w_txt / (float)B;
        if (cos_in_tail && ctl.sides == 1) return launch_grad_tail(ta, 1, dtype, st);      // image side only; the text side's launch finishes the losses
        if (cos_in_tail) ta.sums = TailSums{tail_counter, split_sums, cos_img, cos_txt, losses, ctl.losses_host, ctl.seq, w_img, w_txt, w_con};
        if (cos_in_tail && ctl.sides == 2) {
            ta.s[0] = ta.s[1];
            ta.w_cos_over_B[0] = ta.w_cos_over_B[1];
            return launch_grad_tail(ta, 1, dtype, st);
        }
        if (int e = launch_grad_tail(ta, 2, dtype, st)) return e;
    }
    return DCLIP_OK;
}

int dclip_distill_loss_fwd_bwd(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t B,
                               int64_t d, int dtype, float temperature, float w_img, float w_txt, float w_con,
                               float grad_scale, float* losses, void* d_s_img, void* d_s_txt, float* stats,
                               void* workspace, size_t workspace_bytes, int path, void* stream) {
    return distill_loss_impl(0, s_img, s_txt, t_img, t_txt, B, d, dtype, temperature, w_img, w_txt, w_con, grad_scale, losses, d_s_img,
                             d_s_txt, stats, workspace, workspace_bytes, path, stream);
}

int dclip_distill_loss_students(const void* s_img, const void* s_txt, int64_t B, int64_t d, int dtype, float temperature, float w_con,
                                void* workspace, size_t workspace_bytes, int path, void* stream) {
    return distill_loss_impl(1, s_img, s_txt, nullptr, nullptr, B, d, dtype, temperature, 0.f, 0.f, w_con, 1.f, nullptr, nullptr, nullptr,
                             nullptr, workspace, workspace_bytes, path, stream);
}

int dclip_distill_loss_teachers(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t B, int64_t d,
                                int dtype, float temperature, float w_img, float w_txt, float w_con, float grad_scale, float* losses,
                                void* d_s_img, void* d_s_txt, void* workspace, size_t workspace_bytes, int path, void* stream) {
    return distill_loss_impl(2, s_img, s_txt, t_img, t_txt, B, d, dtype, temperature, w_img, w_txt, w_con, grad_scale, losses, d_s_img,
                             d_s_txt, nullptr, workspace, workspace_bytes, path, stream);
}

int dclip_host_distill_step(const void* h_s_img, const void* h_s_txt, const void* h_t_img, const void* h_t_txt, void* const* d_in,
                            int64_t B, int64_t d, int dtype, float temperature, float w_img, float w_txt, float w_con, float grad_scale,
                            float* d_losses, float* h_losses, void* d_s_img, void* d_s_txt, void* workspace, size_t workspace_bytes, int path,
                            void* compute_stream, void* copy_stream, void* ev_students, void* ev_teacher_img, void* ev_teachers, void* ev_done) {
    DCLIP_REQUIRE(h_s_img && h_s_txt && h_t_img && h_t_txt && d_in && d_in[0] && d_in[1] && d_in[2] && d_in[3] && d_losses && h_losses,
                  "host_distill_step: null pointer");
    DCLIP_REQUIRE(copy_stream && ev_students && ev_teacher_img && ev_teachers && ev_done && copy_stream != compute_stream,
                  "host_distill_step: needs a copy stream of its own and four events");
    cudaStream_t cs = static_cast<cudaStream_t>(compute_stream), xs = static_cast<cudaStream_t>(copy_stream);
    cudaEvent_t es = static_cast<cudaEvent_t>(ev_students), ei = static_cast<cudaEvent_t>(ev_teacher_img),
                et = static_cast<cudaEvent_t>(ev_teachers), ed = static_cast<cudaEvent_t>(ev_done);
    const size_t bytes = (size_t)B * d * dclip::dtype_size(dtype);
    // with gradients the teachers phase runs per side and its last block stores the losses straight into h_losses (mapped pinned
    // memory), sequence word last: the host waits on that word instead of a D2H copy + stream synchronisation
    const bool direct = d_s_img != nullptr;
    DCLIP_REQUIRE(!direct || dclip::aligned16(h_losses), "host_distill_step: h_losses must be 16-byte aligned");
    float* h_dev = nullptr;
    unsigned int seq = 0;
    if (direct) {
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_dev), h_losses, 0) != cudaSuccess || !h_dev) {
            (void)cudaGetLastError();
            return dclip::fail(DCLIP_ERR_INVALID, "host_distill_step: h_losses must be pinned host memory (cudaHostAlloc / cudaHostRegister), 8 floats");
        }
        seq = __atomic_load_n(reinterpret_cast<unsigned int*>(h_losses + 4), __ATOMIC_RELAXED) + 1u;
    }
    // the device input buffers are free again once the previous step's kernels are done
    DCLIP_CUDA_OK(cudaStreamWaitEvent(xs, ed, 0));
    DCLIP_CUDA_OK(cudaMemcpyAsync(d_in[0], h_s_img, bytes, cudaMemcpyHostToDevice, xs));
    DCLIP_CUDA_OK(cudaMemcpyAsync(d_in[1], h_s_txt, bytes, cudaMemcpyHostToDevice, xs));
    DCLIP_CUDA_OK(cudaEventRecord(es, xs));
    DCLIP_CUDA_OK(cudaMemcpyAsync(d_in[2], h_t_img, bytes, cudaMemcpyHostToDevice, xs));
    DCLIP_CUDA_OK(cudaEventRecord(ei, xs));
    DCLIP_CUDA_OK(cudaMemcpyAsync(d_in[3], h_t_txt, bytes, cudaMemcpyHostToDevice, xs));
    DCLIP_CUDA_OK(cudaEventRecord(et, xs));
    DCLIP_CUDA_OK(cudaStreamWaitEvent(cs, es, 0));
    if (int e = distill_loss_impl(1, d_in[0], d_in[1], nullptr, nullptr, B, d, dtype, temperature, 0.f, 0.f, w_con, 1.f, nullptr, nullptr, nullptr,
                                  nullptr, workspace, workspace_bytes, path, compute_stream)) return e;
    if (!direct) {
        DCLIP_CUDA_OK(cudaStreamWaitEvent(cs, et, 0));
        if (int e = distill_loss_impl(2, d_in[0], d_in[1], d_in[2], d_in[3], B, d, dtype, temperature, w_img, w_txt, w_con, grad_scale, d_losses,
                                      nullptr, nullptr, nullptr, workspace, workspace_bytes, path, compute_stream)) return e;
        DCLIP_CUDA_OK(cudaMemcpyAsync(h_losses, d_losses, 4 * sizeof(float), cudaMemcpyDeviceToHost, cs));
        DCLIP_CUDA_OK(cudaEventRecord(ed, cs));
        DCLIP_CUDA_OK(cudaStreamSynchronize(cs));
        return DCLIP_OK;
    }
    DCLIP_CUDA_OK(cudaStreamWaitEvent(cs, ei, 0));
    if (int e = distill_loss_impl(2, d_in[0], d_in[1], d_in[2], d_in[3], B, d, dtype, temperature, w_img, w_txt, w_con, grad_scale, d_losses,
                                  d_s_img, d_s_txt, nullptr, workspace, workspace_bytes, path, compute_stream, TailCtl{1, nullptr, 0u})) return e;
    DCLIP_CUDA_OK(cudaStreamWaitEvent(cs, et, 0));
    if (int e = distill_loss_impl(2, d_in[0], d_in[1], d_in[2], d_in[3], B, d, dtype, temperature, w_img, w_txt, w_con, grad_scale, d_losses,
                                  d_s_img, d_s_txt, nullptr, workspace, workspace_bytes, path, compute_stream, TailCtl{2, h_dev, seq})) return e;
    DCLIP_CUDA_OK(cudaEventRecord(ed, cs));
    // the one host wait of the step: the sequence word behind the losses (the stream is queried now and then so that a failed launch
    // or a device fault ends the wait with its error instead of hanging)
    const unsigned int* flag = reinterpret_cast<const unsigned int*>(h_losses + 4);
    for (unsigned int spins = 1; __atomic_load_n(flag, __ATOMIC_ACQUIRE) != seq; ++spins) {
        if ((spins & 0xfffu) == 0u) {
            const cudaError_t q = cudaStreamQuery(cs);
            if (q == cudaSuccess) {
                if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq) break;
                return dclip::fail(DCLIP_ERR_CUDA, "host_distill_step: the step finished without delivering its losses");
            }
            if (q != cudaErrorNotReady) return dclip::fail(DCLIP_ERR_CUDA, "host_distill_step: %s", cudaGetErrorString(q));
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    return DCLIP_OK;
}

int dclip_scale_grad(const void* x, int64_t n, int dtype, const float* scale, void* out, void* stream) {
    DCLIP_REQUIRE(x && scale && out && n > 0 && n % 8 == 0, "scale_grad: bad arguments (n must be a positive multiple of 8)");
    DCLIP_REQUIRE(dclip::aligned16(x) && dclip::aligned16(out), "scale_grad: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t groups = n / 8;
    DCLIP_DISPATCH_DTYPE(dtype, {
        dclip::k_scale_grad<T><<<(unsigned)dclip::ceil_div(groups, 256), 256, 0, st>>>(static_cast<const T*>(x), groups, scale, static_cast<T*>(out));
    });
    DCLIP_LAUNCH_OK("k_scale_grad");
    return DCLIP_OK;
}

}  // extern "C"
