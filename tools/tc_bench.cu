// tc_bench — micro-benchmarks of the sm_100a primitives the tensor-path kernels are built from.
// Development tool: measures (1) tcgen05.mma issue/execute rate from resident shared memory for the
// tile shapes and operand majors we use, (2) TMA streaming rate through an mbarrier ring.
// Every wait is bounded.  Usage: tc_bench mma | tc_bench tma
#include <cuda_fp16.h>

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

// ---- (1) MMA rate: one thread issues `iters` x `per` MMAs on resident (zeroed) operands ---------
template <int commit_every>
__global__ void __launch_bounds__(128) k_mma_rate(int N, int a_mn, int b_mn, int iters, int per, int spread, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem = tmem_base;
    if (uniform_warp_idx() == 0 && elect_one()) {
        const uint32_t idesc = make_idesc_f16(128, N, kFmtF16, a_mn, b_mn);
        const uint64_t hiA = make_smem_desc_hi(a_mn ? 16384 : 0, 1024, kSwizzle128B);
        const uint64_t hiB = make_smem_desc_hi(b_mn ? 16384 : 0, 1024, kSwizzle128B);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint64_t dA = smem_desc(hiA, a0), dB = smem_desc(hiB, b0);
            const uint32_t stepA = spread ? (a_mn ? 128u : 2u) : 0u, stepB = b_mn ? 128u : 2u;
            for (int k8 = 0; k8 < per; k8 += 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                {
                    umma_f16(tmem + (k & 1) * 256, dA + (uint64_t)(k * stepA), dB + (uint64_t)((k & 3) * stepB), idesc, 1);
                    if constexpr (commit_every > 0) { if ((k + 1) % commit_every == 0) umma_commit(&bar2); }
                }
            }
            umma_commit(&bar);
            if (!mbar_wait(&bar, parity, 1u << 22)) { cycles[blockIdx.x] = -1; break; }
            parity ^= 1;
        }
        const long long t1 = clock64();
        if (cycles[blockIdx.x] != -1) cycles[blockIdx.x] = t1 - t0;
    }
    fence_before_thread_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

static void run_mma() {
    long long* cyc;
    CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(k_mma_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("%-6s %-4s %-4s %-7s %-6s | cycles/MMA (ideal M*N/256... 128*N*16/4096)\n", "N", "Amn", "Bmn", "spread", "grid");
    const int Ns[3] = {64, 128, 256};
    for (int grid : {1, 148})
        for (int ni = 0; ni < 3; ++ni)
            for (int maj = 0; maj < 4; ++maj)
                for (int spread = 0; spread < 2; ++spread) {
                    const int N = Ns[ni], a_mn = maj & 1, b_mn = maj >> 1;
                    if (b_mn && N > 64 && 0) continue;
                    CK(cudaMemset(cyc, 0, 148 * 8));
                    const int iters = 20, per = 64;
                    k_mma_rate<0><<<grid, 128, 200 * 1024>>>(N, a_mn, b_mn, iters, per, spread, cyc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("N=%d maj=%d: CUDA ERROR %s\n", N, maj, cudaGetErrorString(e)); exit(3); }
                    long long h[148];
                    CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
                    double avg = 0;
                    for (int i = 0; i < grid; ++i) avg += (double)h[i];
                    avg /= grid;
                    printf("%-6d %-4d %-4d %-7d %-6d | %8.1f   (ideal %d)%s\n", N, a_mn, b_mn, spread, grid, avg / (iters * per), N / 2,
                           h[0] < 0 ? "  TIMEOUT" : "");
                }
}

// ---- (1b) CTA-pair MMA rate: the leader issues `iters` x `per` cta_group::2 MMAs (M = 256) on resident operands ------
template <int commit_every, int commit_mode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k_mma2_rate(int N, int a_mn, int b_mn, int iters, int per,
                                                                             long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc_pair(&tmem_base, 512);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
    fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    fence_after_thread_sync();
    const uint32_t tmem = tmem_base;
    if (cluster_ctarank() == 0 && uniform_warp_idx() == 0 && elect_one()) {
        const uint32_t idesc = make_idesc_f16(256, N, kFmtF16, a_mn, b_mn);
        const uint64_t hiA = make_smem_desc_hi(a_mn ? 16384 : 0, 1024, kSwizzle128B);
        const uint64_t hiB = make_smem_desc_hi(b_mn ? 16384 : 0, 1024, kSwizzle128B);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint64_t dA = smem_desc(hiA, a0), dB = smem_desc(hiB, b0);
            const uint32_t stepA = a_mn ? 128u : 2u, stepB = b_mn ? 128u : 2u;
            for (int k8 = 0; k8 < per; k8 += 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    umma_f16_pair(tmem + (k & 1) * 256, dA + (uint64_t)(k * stepA), dB + (uint64_t)((k & 3) * stepB), idesc, 1);
                    if constexpr (commit_every > 0) if ((k + 1) % commit_every == 0) {
                        if (commit_mode == 1) umma_commit_pair(&bar2);             // multicast to both CTAs
                        else if (commit_mode == 2) umma_commit_pair(&bar2, 1);     // multicast form, leader only
                        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
                    }
                }
            }
            umma_commit_pair(&bar);
            if (!mbar_wait(&bar, parity, 1u << 22)) { cycles[blockIdx.x >> 1] = -1; break; }
            parity ^= 1;
        }
        const long long t1 = clock64();
        if (cycles[blockIdx.x >> 1] != -1) cycles[blockIdx.x >> 1] = t1 - t0;
    }
    fence_before_thread_sync();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) tmem_dealloc_pair(tmem, 512);
}

template <int CE>
static double one_mma1(int N, long long* cyc) {
    CK(cudaFuncSetAttribute(k_mma_rate<CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaMemset(cyc, 0, 148 * 8));
    k_mma_rate<CE><<<1, 128, 200 * 1024>>>(N, 0, 0, 20, 64, 1, cyc);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    return (double)h / (20 * 64);
}
template <int CE, int MODE>
static double one_mma2(int N, long long* cyc) {
    CK(cudaFuncSetAttribute((k_mma2_rate<CE, MODE>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaMemset(cyc, 0, 148 * 8));
    k_mma2_rate<CE, MODE><<<2, 128, 200 * 1024>>>(N, 0, 0, 20, 64, cyc);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    return (double)h / (20 * 64);
}

static void run_mma2() {
    long long* cyc;
    CK(cudaMalloc(&cyc, 148 * 8));
    printf("cycles per MMA with a tcgen05.commit every CE MMAs (CE = 0: only at the end of 64)\n");
    for (int N : {64, 128, 256})
        printf("single CTA  M=128 N=%-3d | CE=0 %6.1f  CE=8 %6.1f  CE=4 %6.1f  CE=2 %6.1f\n", N, one_mma1<0>(N, cyc), one_mma1<8>(N, cyc), one_mma1<4>(N, cyc), one_mma1<2>(N, cyc));
    for (int N : {128, 256}) {
        printf("CTA pair    M=256 N=%-3d | CE=0 %6.1f | multicast both: CE=8 %6.1f CE=4 %6.1f | leader only: CE=8 %6.1f CE=4 %6.1f | plain: CE=8 %6.1f CE=4 %6.1f\n", N,
               one_mma2<0, 1>(N, cyc), one_mma2<8, 1>(N, cyc), one_mma2<4, 1>(N, cyc), one_mma2<8, 2>(N, cyc), one_mma2<4, 2>(N, cyc),
               one_mma2<8, 3>(N, cyc), one_mma2<4, 3>(N, cyc));
        fflush(stdout);
    }
}

// ---- (2) TMA streaming rate ---------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_tma_rate(const __grid_constant__ CUtensorMap tm, int slots, int boxes_per_slot, int jobs, int rows_total,
                                                 int distinct, int mode, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16], empty[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int slot_bytes = boxes_per_slot * 16384;
    const int ntiles = rows_total / 128;
    if (uniform_warp_idx() == 0 && elect_one()) {           // producer
        uint32_t stage = 0, phase = 0;
        for (int j = 0; j < jobs; ++j) {
            if (!mbar_wait(&empty[stage], phase ^ 1, 1u << 22)) break;
            if (mode == 1) { mbar_arrive(&full[stage]); }     // handshake only, no TMA
            else {
                mbar_arrive_expect_tx(&full[stage], slot_bytes);
                const int tile = distinct ? (int)((blockIdx.x * 7 + j / 6) % ntiles) : (j / 6) % ntiles;
                for (int b = 0; b < boxes_per_slot; ++b)
                    tma_load_2d(smem + stage * slot_bytes + b * 16384, &tm, &full[stage], ((j % 6) * boxes_per_slot + b) * 64 % 768, tile * 128);
            }
            if (++stage == (uint32_t)slots) { stage = 0; phase ^= 1; }
        }
    } else if (uniform_warp_idx() == 1 && elect_one()) {   // consumer: wait + release
        uint32_t stage = 0, phase = 0;
        const long long t0 = clock64();
        bool ok = true;
        for (int j = 0; j < jobs && ok; ++j) {
            ok = mbar_wait(&full[stage], phase, 1u << 22);
            mbar_arrive(&empty[stage]);
            if (++stage == (uint32_t)slots) { stage = 0; phase ^= 1; }
        }
        cycles[blockIdx.x] = ok ? clock64() - t0 : -1;
    }
}

// multi-producer variant: warp w < nprod owns slot w (as the product kernels do); warp nprod consumes.
__global__ void __launch_bounds__(256) k_tma_rate_mp(const __grid_constant__ CUtensorMap tm, int nprod, int boxes_per_slot, int jobs, int rows_total,
                                                     long long* cycles, long long* latency) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16], empty[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < nprod; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int slot_bytes = boxes_per_slot * 16384;
    const int ntiles = rows_total / 128;
    const int w = uniform_warp_idx();
    if (w < nprod && elect_one()) {
        uint32_t phase = 0;
        for (int j = w; j < jobs; j += nprod, phase ^= 1) {
            if (!mbar_wait(&empty[w], phase ^ 1, 1u << 22)) break;
            mbar_arrive_expect_tx(&full[w], slot_bytes);
            const long long t0 = clock64();
            for (int b = 0; b < boxes_per_slot; ++b)
                tma_load_2d(smem + w * slot_bytes + b * 16384, &tm, &full[w], ((j % 6) * boxes_per_slot + b) * 64 % 768, ((j / 6) % ntiles) * 128);
            if (j == w && latency) {           // first job of this producer: measure issue -> data-landed latency
                mbar_wait(&full[w], 0, 1u << 22);
                latency[blockIdx.x * 8 + w] = clock64() - t0;
            }
        }
    } else if (w == nprod && elect_one()) {
        uint32_t stage = 0, phase = 0;
        const long long t0 = clock64();
        bool ok = true;
        for (int j = 0; j < jobs && ok; ++j) {
            ok = mbar_wait(&full[stage], phase, 1u << 22);
            mbar_arrive(&empty[stage]);
            if (++stage == (uint32_t)nprod) { stage = 0; phase ^= 1; }
        }
        cycles[blockIdx.x] = ok ? clock64() - t0 : -1;
    }
}

static void run_tma_mp() {
    const int rows = 4096, cols = 768;
    __half* d;
    CK(cudaMalloc(&d, (size_t)rows * cols * 2));
    CK(cudaMemset(d, 0, (size_t)rows * cols * 2));
    CUtensorMap tm;
    if (!make_tmap_2d_sw128(&tm, d, rows, cols, 128, 0)) { printf("tmap failed\n"); exit(2); }
    long long *cyc, *lat;
    CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaMalloc(&lat, 148 * 8 * 8));
    CK(cudaFuncSetAttribute(k_tma_rate_mp, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    for (int grid : {1, 148})
        for (int cfg = 0; cfg < 5; ++cfg) {
            const int nprod_[5] = {3, 6, 3, 6, 4};
            const int bps_[5] = {2, 1, 4, 2, 3};
            const int nprod = nprod_[cfg], bps = bps_[cfg];
            if (nprod * bps * 16384 > 200 * 1024) continue;
            const int jobs = 600;
            CK(cudaMemset(cyc, 0, 148 * 8));
            CK(cudaMemset(lat, 0, 148 * 64));
            k_tma_rate_mp<<<grid, 256, 220 * 1024>>>(tm, nprod, bps, jobs, rows, cyc, lat);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); exit(3); }
            long long h[148], hl[148 * 8];
            CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hl, lat, grid * 64, cudaMemcpyDeviceToHost));
            double avg = 0; bool to = false;
            for (int i = 0; i < grid; ++i) { avg += (double)h[i]; to |= h[i] < 0; }
            avg /= grid;
            double l = 0;
            for (int i = 0; i < grid; ++i) l += (double)hl[i * 8];
            l /= grid;
            const double bpc = (double)jobs * bps * 16384 / avg;
            printf("MP grid %-4d producers %-2d slotKB %-4d | %7.1f cycles/job  %7.1f B/clk/SM  %6.2f TB/s | first-load latency %6.0f cyc%s\n", grid, nprod,
                   bps * 16, avg / jobs, bpc, bpc * grid * clk_khz * 1e3 / 1e12, l, to ? "  TIMEOUT" : "");
            fflush(stdout);
        }
}

static void run_tma() {
    const int rows = 4096, cols = 768;
    __half* d;
    CK(cudaMalloc(&d, (size_t)rows * cols * 2));
    CK(cudaMemset(d, 0, (size_t)rows * cols * 2));
    CUtensorMap tm;
    if (!make_tmap_2d_sw128(&tm, d, rows, cols, 128, 0)) { printf("tmap failed\n"); exit(2); }
    long long* cyc;
    CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(k_tma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    printf("%-6s %-6s %-9s %-9s | bytes/clk/SM   aggregate TB/s (at measured SM clock)\n", "grid", "slots", "slotKB", "distinct");
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    for (int grid : {1, 148})
        for (int cfg = 0; cfg < 7; ++cfg) {
            // cfg: slots, boxes per slot, mode
            const int slots_[7] = {3, 6, 12, 3, 2, 6, 3};
            const int bps_[7] = {2, 1, 1, 4, 6, 2, 2};
            const int mode_[7] = {0, 0, 0, 0, 0, 0, 1};
            const int slots = slots_[cfg], bps = bps_[cfg], mode = mode_[cfg], distinct = 0;
            if (slots * bps * 16384 > 200 * 1024) continue;
            const int jobs = 600;
            CK(cudaMemset(cyc, 0, 148 * 8));
            k_tma_rate<<<grid, 64, 220 * 1024>>>(tm, slots, bps, jobs, rows, distinct, mode, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); exit(3); }
            long long h[148];
            CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
            double avg = 0; bool to = false;
            for (int i = 0; i < grid; ++i) { avg += (double)h[i]; to |= h[i] < 0; }
            avg /= grid;
            const double bpc = (double)jobs * bps * 16384 / avg;
            printf("grid %-4d slots %-3d slotKB %-4d mode %d | %7.1f cycles/job  %7.1f B/clk/SM  %6.2f TB/s%s\n", grid, slots, bps * 16, mode,
                   avg / jobs, bpc, bpc * grid * clk_khz * 1e3 / 1e12, to ? "  TIMEOUT" : "");
            fflush(stdout);
        }
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "mma")) { run_mma(); return 0; }
    if (argc >= 2 && !strcmp(argv[1], "mma2")) { run_mma2(); return 0; }
    if (argc >= 2 && !strcmp(argv[1], "tma")) { run_tma(); return 0; }
    if (argc >= 2 && !strcmp(argv[1], "tmamp")) { run_tma_mp(); return 0; }
    printf("usage: tc_bench mma | tma\n");
    return 64;
}
