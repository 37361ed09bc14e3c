// MN-major UMMA operands in the 32B / 64B / 128B swizzle modes, started at an arbitrary K row (pixel) of their tile, with
// the M = 128 rows of one MMA made of 128/C channel blocks that are `lbo_rows` pixels apart - what a weight-gradient GEMM
// over 16-, 32- and 64-channel activations needs to read several filter taps with one instruction.  Second part: how many
// cycles one such MMA costs as a function of N and of the swizzle mode (a chain of 256 MMAs on one accumulator).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I vsrlab_b200/csrc -I include -o tools/umma_mn_modes_test.bin tools/umma_mn_modes_test.cu
//
// X[96 pixels][C] bf16, pixel r at r * (2C) bytes, 16-byte chunks XOR-swizzled on address bits 7.. (TMA's pattern);
// Z[16 pixels][64] = identity (128B swizzle).  D[m][n] = X[n + s + (m / C) * lbo_rows][m % C] for n < 16.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace vsrb;

// layout field (bits 61..63): 2 = 128B, 4 = 64B, 6 = 32B swizzle; SBO = 8 K rows
__device__ __forceinline__ uint64_t mn_desc_mode(uint32_t saddr, uint32_t lbo_bytes, uint32_t pitch) {
    const uint32_t layout = pitch == 128 ? 2u : pitch == 64 ? 4u : 6u;
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((8u * pitch) >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t pitch) {
    const uint32_t mask = pitch == 128 ? 7u : pitch == 64 ? 3u : 1u;
    return off ^ (((off >> 7) & mask) << 4);
}

__global__ void __launch_bounds__(128) mn_modes_test(int C, int s, int lbo_rows, float* out) {
    extern __shared__ uint8_t raw_[];
    const uint32_t raw = smem_u32(raw_);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = raw_ + (base - raw);
    uint8_t* x_s = bp;                    // 96 rows x pitch
    uint8_t* z_s = bp + 12 * 1024;        // 16 rows x 128 B
    const uint32_t bar = base + 16 * 1024;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + 16 * 1024 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t pitch = 2u * C;
    for (int i = tid; i < 96 * C; i += 128) {
        const int r = i / C, c = i % C;
        *reinterpret_cast<__nv_bfloat16*>(x_s + swz((uint32_t)r * pitch + (uint32_t)c * 2u, pitch)) =
            __float2bfloat16_rn((float)((r * 3 + c * 5) % 31 - 15));
    }
    for (int i = tid; i < 16 * 64; i += 128) {
        const int k = i >> 6, n = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(z_s + swz((uint32_t)k * 128u + (uint32_t)n * 2u, 128)) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
        umma_bf16(tmem, mn_desc_mode(smem_u32(x_s) + (uint32_t)s * pitch, (uint32_t)lbo_rows * pitch, pitch),
                  mn_desc_mode(smem_u32(z_s), 0u, 128), idesc, 0u);
        umma_commit(bar);
    }
    __syncwarp();
    bool dead = false;
    int dbg = 0;
    mbar_wait(bar, 0, &dbg, 1, dead);
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + (tid & 31)) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// cycles per MMA: `reps` MMAs (M = 128, K = 16) on one accumulator, A in the given MN-major mode (or K-major SW128 when
// a_kmajor), B MN-major with pitch pb (N <= pb/2 per block; larger N re-reads the block through LBO = 0)
__global__ void __launch_bounds__(128) mn_cost_test(int pa, int pb, int N, int a_kmajor, int reps, long long* cycles, int start_rows = 0,
                                                    int cycle_rows = 0, int lbo_rows = 1, int n_acc = 1) {
    extern __shared__ uint8_t raw_[];
    const uint32_t raw = smem_u32(raw_);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = raw_ + (base - raw);
    const uint32_t bar = base + 48 * 1024;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + 48 * 1024 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(bp)[i] = 0;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    if (warp == 1 && elect_one()) {
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        if (!a_kmajor) idesc |= 1u << 15;
        uint64_t ad;
        if (a_kmajor) {     // K-major SW128: SBO = 1024 (8 rows), layout 2
            const uint32_t lo = ((base >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            ad = ((uint64_t)hi << 32) | lo;
        } else {
            ad = mn_desc_mode(base, (uint32_t)pa, (uint32_t)pa);
        }
        const uint64_t bd = mn_desc_mode(base + 24 * 1024, 0u, (uint32_t)pb);
        const long long t0 = clock64();
        if (start_rows || cycle_rows || lbo_rows != 1 || n_acc != 1) {
            // windows that start `start_rows` (+ 0..7 * cycle_rows) pixel rows into the tile, M blocks lbo_rows rows apart,
            // round-robin over n_acc accumulators; the eight descriptors are built before the clock starts
            uint64_t ds[8];
            for (int j = 0; j < 8; ++j)
                ds[j] = mn_desc_mode(base + (uint32_t)(start_rows + j * cycle_rows) * (uint32_t)pa, (uint32_t)(lbo_rows * pa), (uint32_t)pa);
            const long long t1 = clock64();
            for (int i = 0; i < reps; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) umma_bf16(tmem + (uint32_t)((j % n_acc) * N), ds[j], bd, idesc, (i + j) >= n_acc ? 1u : 0u);
            }
            umma_commit(bar);
            bool dead2 = false;
            int dbg2 = 0;
            mbar_wait(bar, 0, &dbg2, 1, dead2);
            *cycles = clock64() - t1;
            return;
        } else
        for (int i = 0; i < reps; ++i) umma_bf16(tmem, ad, bd, idesc, i ? 1u : 0u);
        umma_commit(bar);
        bool dead = false;
        int dbg = 0;
        mbar_wait(bar, 0, &dbg, 1, dead);
        *cycles = clock64() - t0;
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * sizeof(float));
    float* h = (float*)malloc(128 * 64 * sizeof(float));
    cudaFuncSetAttribute(mn_modes_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
    cudaFuncSetAttribute(mn_cost_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
    const int Cs[3] = {16, 32, 64};
    const int lbos[3] = {1, 2, 22};
    int total_bad = 0;
    for (int ci = 0; ci < 3; ++ci)
        for (int li = 0; li < 3; ++li)
            for (int s = 0; s < 12; ++s) {
                const int C = Cs[ci], lb = lbos[li];
                if (s + 15 + (128 / C - 1) * lb >= 96) continue;
                mn_modes_test<<<1, 128, 20 * 1024>>>(C, s, lb, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("C %d shift %d lbo %d: CUDA error %s\n", C, s, lb, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 16; ++n) {
                        const int r = n + s + (m / C) * lb, c = m % C;
                        bad += h[m * 64 + n] != (float)((r * 3 + c * 5) % 31 - 15);
                    }
                total_bad += bad;
                printf("C=%d (pitch %d B) start row %d, M blocks %d rows apart: %s (%d of 2048 wrong)\n", C, 2 * C, s, lb, bad ? "WRONG" : "exact", bad);
            }
    printf("correctness: %s\n", total_bad ? "FAILED" : "all exact");
    long long* dc;
    cudaMalloc(&dc, 8);
    const int reps = 256;
    const int Ns[5] = {16, 32, 64, 128, 256};
    for (int km = 0; km < 2; ++km)
        for (int pa = 32; pa <= 128; pa *= 2) {
            if (km && pa != 128) continue;
            for (int ni = 0; ni < 5; ++ni) {
                long long c = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    mn_cost_test<<<1, 128, 52 * 1024>>>(pa, 128, Ns[ni], km, reps, dc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("cost test: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
                }
                printf("A %s pitch %3d, B MN-major 128, M=128 N=%3d K=16: %.1f cycles per MMA\n", km ? "K-major " : "MN-major", pa, Ns[ni], (double)c / reps);
            }
        }
    // the weight-gradient kernel's access pattern: unaligned window starts, one-pixel M block distance, several accumulators
    for (int pa = 32; pa <= 128; pa *= 2)
        for (int variant = 0; variant < 6; ++variant) {
            const int start = (variant == 1 || variant == 3) ? 3 : 0, cyc = (variant == 2 || variant == 3) ? 11 : 0;
            const int lbo = variant == 4 ? 0 : 1, nacc = variant == 5 ? 4 : 1;
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                mn_cost_test<<<1, 128, 52 * 1024>>>(pa, 128, 32, 0, reps, dc, start, cyc, lbo, nacc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("cost test: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            }
            printf("A MN-major pitch %3d N=32: start row %d, cycling %2d rows, M blocks %d rows apart, %d accumulators: %.1f cycles per MMA\n", pa,
                   start, cyc, lbo, nacc, (double)c / reps);
        }
    return 0;
}
