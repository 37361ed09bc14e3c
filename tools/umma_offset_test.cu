// Does a K-major, 128B-swizzled UMMA operand accept a start address that is NOT aligned to its 1024-byte swizzle atom
// (a window that starts s rows into the atom), and what must the descriptor's base-offset field (bits 49-51) hold?
// This is the enabler for reading the kx-shifted pixel windows of a conv / weight gradient from ONE shared-memory tile.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I vsrlab_b200/csrc -I include -o tools/umma_offset_test.bin tools/umma_offset_test.cu
//   tools/umma_offset_test.bin
//
// A_full[136][64] bf16 sits in shared memory in the canonical K-major SW128 layout (row r at r*128 B, 16-byte chunk j at
// j ^ (r & 7)); B = 64x64 identity; D = A_window * B^T, so D[m][n] must equal A_full[m + s][n].
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace vsrb;

__global__ void __launch_bounds__(128) offset_test(int s, int base_off, float* out) {
    extern __shared__ uint8_t raw_[];
    const uint32_t raw = smem_u32(raw_);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = raw_ + (base - raw);
    uint8_t* a_s = bp;                    // 136 rows x 128 B  (17 KiB + slack)
    uint8_t* b_s = bp + 18 * 1024;        // 64 rows x 128 B
    const uint32_t bar = base + 27 * 1024;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + 27 * 1024 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 136 * 64; i += 128) {
        const int r = i >> 6, k = i & 63;
        uint32_t off = (uint32_t)r * 128u + (uint32_t)k * 2u;
        off ^= ((off >> 7) & 7u) << 4;
        *reinterpret_cast<__nv_bfloat16*>(a_s + off) = __float2bfloat16_rn((float)((r * 3 + k * 5) % 31 - 15));
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i >> 6, k = i & 63;
        uint32_t off = (uint32_t)n * 128u + (uint32_t)k * 2u;
        off ^= ((off >> 7) & 7u) << 4;
        *reinterpret_cast<__nv_bfloat16*>(b_s + off) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
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
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
        const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)base_off << 17) | (2u << 29);
        const uint32_t a0 = smem_u32(a_s) + (uint32_t)s * 128u, b0 = smem_u32(b_s);
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ((uint64_t)hi << 32) | ((((a0 >> 4) + k * 2) & 0x3FFFu) | (1u << 16));
            const uint64_t bd = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | ((((b0 >> 4) + k * 2) & 0x3FFFu) | (1u << 16));
            umma_bf16(tmem, ad, bd, idesc, k ? 1u : 0u);
        }
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
    cudaFuncSetAttribute(offset_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 30 * 1024);
    for (int s = 0; s < 8; ++s)
        for (int mode = 0; mode < 2; ++mode) {
            const int bo = mode ? s : 0;
            offset_test<<<1, 128, 30 * 1024>>>(s, bo, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", s, bo, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n)
                    bad += h[m * 64 + n] != (float)(((m + s) * 3 + n * 5) % 31 - 15);
            printf("start row %d (+%d B), base_offset field %d: %s (%d of 8192 wrong)\n", s, s * 128, bo, bad ? "WRONG" : "exact", bad);
        }
    return 0;
}
