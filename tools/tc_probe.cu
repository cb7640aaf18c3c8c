// tc_probe — standalone sm_100a probe for the hand-written tcgen05 / TMA primitives.
// Development tool (not part of the product): validates, on a real B200, the shared-memory
// descriptor encodings and canonical layouts that dclip_b200/csrc/tcgen05.cuh documents.
// Every wait is bounded, so a wrong guess prints FAIL instead of hanging the GPU.
//
//   tc_probe mma <variant>     single-tile MMA from host-built smem images + descriptors
//   tc_probe tma               TMA SWIZZLE_128B placement check
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../dclip_b200/csrc/tcgen05.cuh"

using namespace dclip::tc;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

// ------------------------------------------------------------------------------------------------
struct ProbeArgs {
    const uint8_t* a_img; int a_bytes;
    const uint8_t* b_img; int b_bytes;
    uint64_t a_hi, b_hi;          // descriptor high parts (LBO/SBO/version/swizzle)
    uint32_t a_off[16], b_off[16];  // per-MMA start byte offsets inside the images
    int n_mma;
    uint32_t idesc;
    int ncols;                    // N (accumulator columns)
    float* d_out;                 // [128][ncols]
    int* status;                  // 0 ok, 1 timeout
};

__global__ void __launch_bounds__(128) k_mma_probe(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + ((p.a_bytes + 1023) / 1024) * 1024;
    for (int i = threadIdx.x; i < p.a_bytes / 16; i += 128) reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(p.a_img)[i];
    for (int i = threadIdx.x; i < p.b_bytes / 16; i += 128) reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(p.b_img)[i];
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    uint32_t cols = 32;
    while ((int)cols < p.ncols) cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_base, cols);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.n_mma; ++i) {
            const uint64_t da = smem_desc(p.a_hi, smem_u32(a_s) + p.a_off[i]);
            const uint64_t db = smem_desc(p.b_hi, smem_u32(b_s) + p.b_off[i]);
            umma_f16(tmem, da, db, p.idesc, i > 0);
        }
        umma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0, 1u << 20);
    fence_after_thread_sync();
    if (!ok && threadIdx.x == 0) *p.status = 1;
    if (ok) {
        for (int c0 = 0; c0 < p.ncols; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) p.d_out[(size_t)threadIdx.x * p.ncols + c0 + j] = __uint_as_float(r[j]);
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, cols);
}

// host-side canonical layouts (see tcgen05.cuh)
static size_t off_kmajor_sw128(int r, int k, int rows) {     // element (r, k) of a [rows][K] K-major operand
    const int kb = k / 64, kk = k % 64;
    return (size_t)kb * rows * 128 + (size_t)(r / 8) * 1024 + (r % 8) * 128 + (((kk / 8) ^ (r % 8)) * 16) + (kk % 8) * 2;
}
static size_t off_mnmajor_sw128(int mn, int k, int K) {      // element (mn, k) of an MN-major operand
    const int g = mn / 64, m = mn % 64;
    return (size_t)g * (K / 8) * 1024 + (size_t)(k / 8) * 1024 + (k % 8) * 128 + (((m / 8) ^ (k % 8)) * 16) + (m % 8) * 2;
}

static int run_mma(int variant) {
    // variants: bit0: A MN-major, bit1: B MN-major, bit2: version bit off, bit3: LBO=16B for K-major (instead of 0)
    //           bit4: bf16 instead of fp16, bit5: N=64 (else 128)
    const bool a_mn = variant & 1, b_mn = variant & 2, no_ver = variant & 4, lbo16 = variant & 8, bf16 = variant & 16;
    const int M = 128, N = (variant & 32) ? 64 : 128, K = 64;
    std::vector<float> A(M * K), B(N * K);
    srand(123 + variant);
    for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;      // exactly representable in fp16/bf16
    for (auto& v : B) v = (float)((rand() % 17) - 8) / 8.0f;
    std::vector<uint8_t> a_img(M * K * 2, 0), b_img(N * K * 2, 0);
    auto put = [&](std::vector<uint8_t>& img, size_t off, float v) {
        uint16_t bits;
        if (bf16) { uint32_t u; memcpy(&u, &v, 4); bits = (uint16_t)(u >> 16); }
        else { __half h = __float2half(v); memcpy(&bits, &h, 2); }
        memcpy(&img[off], &bits, 2);
    };
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) put(a_img, a_mn ? off_mnmajor_sw128(m, k, K) : off_kmajor_sw128(m, k, M), A[m * K + k]);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) put(b_img, b_mn ? off_mnmajor_sw128(n, k, K) : off_kmajor_sw128(n, k, N), B[n * K + k]);

    ProbeArgs p{};
    uint8_t *da, *db;
    CK(cudaMalloc(&da, a_img.size())); CK(cudaMalloc(&db, b_img.size()));
    CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
    p.a_img = da; p.a_bytes = (int)a_img.size(); p.b_img = db; p.b_bytes = (int)b_img.size();
    auto hi = [&](bool mn, int rows_mn) -> uint64_t {
        uint32_t lbo, sbo = 1024;
        if (mn) lbo = (K / 8) * 1024;       // stride between 64-element M/N groups
        else lbo = lbo16 ? 16 : 0;
        (void)rows_mn;
        uint64_t h = make_smem_desc_hi(lbo, sbo, kSwizzle128B);
        if (no_ver) h &= ~(1ull << 46);
        return h;
    };
    p.a_hi = hi(a_mn, M); p.b_hi = hi(b_mn, N);
    p.n_mma = K / 16;
    for (int i = 0; i < p.n_mma; ++i) {
        p.a_off[i] = a_mn ? i * 2 * 1024 : i * 32;     // MN-major: 16 k = two 8-k groups; K-major: +32 B inside the atom
        p.b_off[i] = b_mn ? i * 2 * 1024 : i * 32;
    }
    p.idesc = make_idesc_f16(M, N, bf16 ? kFmtBF16 : kFmtF16, a_mn, b_mn);
    p.ncols = N;
    CK(cudaMalloc(&p.d_out, sizeof(float) * M * N));
    CK(cudaMemset(p.d_out, 0xFF, sizeof(float) * M * N));
    CK(cudaMalloc(&p.status, 4)); CK(cudaMemset(p.status, 0, 4));
    const size_t smem = ((a_img.size() + 1023) / 1024) * 1024 + b_img.size() + 1024;
    CK(cudaFuncSetAttribute(k_mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mma_probe<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mma variant %d: CUDA ERROR %s\n", variant, cudaGetErrorString(e)); return 2; }
    int status = 0;
    CK(cudaMemcpy(&status, p.status, 4, cudaMemcpyDeviceToHost));
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), p.d_out, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            double err = fabs(ref - D[m * N + n]);
            if (!(err < 1e-3)) ++bad;
            if (err > maxerr || std::isnan(err)) maxerr = err;
        }
    printf("mma variant %2d (A %s, B %s, ver %d, lbo16 %d, %s, N=%d): status=%d bad=%d/%d maxerr=%g -> %s\n", variant,
           a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", !no_ver, (int)lbo16, bf16 ? "bf16" : "fp16", N, status, bad, M * N, maxerr,
           (status == 0 && bad == 0) ? "PASS" : "FAIL");
    return (status == 0 && bad == 0) ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tma_probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, uint8_t* out, int bytes, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        tma_load_2d(smem, &tmap, &bar, c0, c1);
    }
    const bool ok = mbar_wait(&bar, 0, 1u << 20);
    if (!ok && threadIdx.x == 0) *status = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < bytes; i += 128) out[i] = smem[i];
}

static int run_tma() {
    const int R = 256, C = 192, box_rows = 128;
    std::vector<__half> h(R * C);
    for (int i = 0; i < R * C; ++i) h[i] = __float2half((float)(i % 2039) / 64.0f);
    __half* d;
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmap;
    if (!make_tmap_2d_sw128(&tmap, d, R, C, box_rows, 0)) { printf("tma: cuTensorMapEncodeTiled failed\n"); return 1; }
    const int bytes = box_rows * 128;
    uint8_t* out; int* status;
    CK(cudaMalloc(&out, bytes)); CK(cudaMalloc(&status, 4)); CK(cudaMemset(status, 0, 4));
    CK(cudaFuncSetAttribute(k_tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024));
    const int c0 = 64, c1 = 128;
    k_tma_probe<<<1, 128, bytes + 1024>>>(tmap, c0, c1, out, bytes, status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tma: CUDA ERROR %s\n", cudaGetErrorString(e)); return 2; }
    int st = 0; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> got(bytes);
    CK(cudaMemcpy(got.data(), out, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < box_rows; ++r)
        for (int k = 0; k < 64; ++k) {
            __half want = h[(size_t)(c1 + r) * C + c0 + k];
            __half have; memcpy(&have, &got[off_kmajor_sw128(r, k, box_rows)], 2);
            if (memcmp(&want, &have, 2) != 0) ++bad;
        }
    printf("tma SW128 placement: status=%d bad=%d/%d -> %s\n", st, bad, box_rows * 64, (st == 0 && bad == 0) ? "PASS" : "FAIL");
    return (st == 0 && bad == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "tma")) return run_tma();
    if (argc >= 3 && !strcmp(argv[1], "mma")) return run_mma(atoi(argv[2]));
    printf("usage: tc_probe tma | tc_probe mma <variant>\n");
    return 64;
}
