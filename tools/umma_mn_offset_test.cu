// Does an MN-major, 128B-swizzled UMMA operand accept a start address that is NOT aligned to its 1024-byte swizzle atom
// (a window that starts s K-rows into the atom)?  This is what a weight-gradient GEMM (K = pixel index) needs to read the
// kx-shifted pixel windows of ONE activation tile instead of one halo copy per filter column.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I vsrlab_b200/csrc -I include -o tools/umma_mn_offset_test.bin tools/umma_mn_offset_test.cu
//
// X[40 pixels][64 ch] bf16 in the canonical MN-major SW128 layout (pixel r at r*128 B, 16-byte chunk j at j ^ (r & 7));
// Z[16 pixels][64] with Z[k][n] = (n == k); one K = 16 MMA, M = 128 as two 64-channel blocks LBO bytes apart:
//   D[m][n]      = sum_k X[k + s][m] * Z[k][n]      = X[n + s][m]            (n < 16)
//   D[64 + m][n] = X[n + s + lbo/128][m]
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace vsrb;

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(128) mn_offset_test(int s, int lbo_rows, float* out) {
    extern __shared__ uint8_t raw_[];
    const uint32_t raw = smem_u32(raw_);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = raw_ + (base - raw);
    uint8_t* x_s = bp;                    // 40 rows x 128 B
    uint8_t* z_s = bp + 8 * 1024;         // 16 rows x 128 B
    const uint32_t bar = base + 12 * 1024;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + 12 * 1024 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 40 * 64; i += 128) {
        const int r = i >> 6, c = i & 63;
        uint32_t off = (uint32_t)r * 128u + (uint32_t)c * 2u;
        off ^= ((off >> 7) & 7u) << 4;
        *reinterpret_cast<__nv_bfloat16*>(x_s + off) = __float2bfloat16_rn((float)((r * 3 + c * 5) % 31 - 15));
    }
    for (int i = tid; i < 16 * 64; i += 128) {
        const int k = i >> 6, n = i & 63;
        uint32_t off = (uint32_t)k * 128u + (uint32_t)n * 2u;
        off ^= ((off >> 7) & 7u) << 4;
        *reinterpret_cast<__nv_bfloat16*>(z_s + off) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
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
        umma_bf16(tmem, mn_desc(smem_u32(x_s) + (uint32_t)s * 128u, (uint32_t)lbo_rows * 128u), mn_desc(smem_u32(z_s), 0u), idesc, 0u);
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

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * sizeof(float));
    float* h = (float*)malloc(128 * 64 * sizeof(float));
    cudaFuncSetAttribute(mn_offset_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
    const int lbos[3] = {1, 2, 16};
    for (int li = 0; li < 3; ++li)
        for (int s = 0; s < 10; ++s) {
            const int lb = lbos[li];
            mn_offset_test<<<1, 128, 16 * 1024>>>(s, lb, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift %d lbo %d: CUDA error %s\n", s, lb, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 16; ++n) {
                    const int r = n + s + (m >= 64 ? lb : 0), c = m & 63;
                    bad += h[m * 64 + n] != (float)((r * 3 + c * 5) % 31 - 15);
                }
            printf("start row %d (+%d B), second M block %d rows further: %s (%d of 2048 wrong)\n", s, s * 128, lb, bad ? "WRONG" : "exact", bad);
        }
    return 0;
}
