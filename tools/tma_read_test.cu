// How fast can one CTA per SM stream global memory into shared memory through the TMA (cp.async.bulk), as a function of
// the bytes kept in flight?  Decides whether the DRAM-bound conv layers are limited by pipeline depth or by the engine.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I vsrlab_b200/csrc -I include -o tools/tma_read_test.bin tools/tma_read_test.cu
//   tools/tma_read_test.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace vsrb;

__global__ void __launch_bounds__(128) tma_stream(const uint8_t* src, size_t bytes_per_cta, int chunk, int slots, int ctas_per_sm_unused) {
    extern __shared__ uint8_t raw_[];
    const uint32_t raw = smem_u32(raw_);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base;                 // up to 32 barriers
    const uint32_t data = base + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint8_t* p = src + (size_t)blockIdx.x * bytes_per_cta;
    const int n = (int)(bytes_per_cta / chunk);
    int issued = 0, done = 0;
    bool dead = false;
    int dbg = 0;
    for (; issued < n && issued < slots; ++issued) {
        mbar_expect_tx(bars + 8 * issued, chunk);
        bulk_load(data + issued * chunk, p + (size_t)issued * chunk, chunk, bars + 8 * issued);
    }
    while (done < n) {
        const int s = done % slots;
        mbar_wait(bars + 8 * s, (done / slots) & 1, &dbg, 1, dead);
        ++done;
        if (issued < n) {
            mbar_expect_tx(bars + 8 * s, chunk);
            bulk_load(data + s * chunk, p + (size_t)issued * chunk, chunk, bars + 8 * s);
            ++issued;
        }
    }
}

// the ring-walk kernel's pattern: one barrier per slot, `parts` copies of chunk/parts bytes per slot issued by `parts` lanes
__global__ void __launch_bounds__(128) tma_stream_parts(const uint8_t* src, size_t bytes_per_cta, int chunk, int slots, int parts) {
    extern __shared__ uint8_t raw_[];
    const uint32_t raw = smem_u32(raw_);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base;
    const uint32_t data = base + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) mbar_init(bars + 8 * i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const uint8_t* p = src + (size_t)blockIdx.x * bytes_per_cta;
    const int n = (int)(bytes_per_cta / chunk), sub = chunk / parts;
    bool dead = false;
    int dbg = 0;
    for (int i = 0; i < n + slots; ++i) {
        const int s = i % slots;
        if (i >= slots) mbar_wait(bars + 8 * s, ((i - slots) / slots) & 1, &dbg, 1, dead);
        if (i < n) {
            if (lane == 0) mbar_expect_tx(bars + 8 * s, chunk);
            __syncwarp();
            if (lane < parts) bulk_load(data + s * chunk + lane * sub, p + (size_t)i * chunk + (size_t)lane * sub, sub, bars + 8 * s);
        }
        __syncwarp();
    }
}

int main() {
    const size_t total = (size_t)148 * 16 * 1024 * 1024;      // 2.3 GiB
    uint8_t* d;
    cudaMalloc(&d, total);
    cudaMemset(d, 1, total);
    cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    cudaFuncSetAttribute(tma_stream_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int chunks[] = {4096, 16384, 32768};
    for (int ci = 0; ci < 3; ++ci)
        for (int kb = 16; kb <= 192; kb *= 2) {
            const int chunk = chunks[ci];
            int slots = kb * 1024 / chunk;
            if (slots < 1 || slots > 32) continue;
            const int smem = 2048 + slots * chunk;
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                tma_stream<<<148, 128, smem>>>(d, total / 148, chunk, slots, 1);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = ms < best ? ms : best;
            }
            cudaError_t e = cudaGetLastError();
            printf("chunk %6d B x %2d slots = %3d KiB in flight per SM: %7.1f GB/s %s\n", chunk, slots, slots * chunk / 1024,
                   total / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    for (int parts = 1; parts <= 16; parts *= 2) {
        const int chunk = 16384, slots = 8, smem = 2048 + slots * chunk;
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            tma_stream_parts<<<148, 128, smem>>>(d, total / 148, chunk, slots, parts);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("16 KiB per barrier as %2d copies of %5d B, 8 slots: %7.1f GB/s\n", parts, chunk / parts, total / best / 1e6);
    }
    // two CTAs per SM (296 CTAs), 96 KiB in flight each
    for (int kb = 32; kb <= 96; kb += 32) {
        const int chunk = 16384, slots = kb * 1024 / chunk, smem = 2048 + slots * chunk;
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            tma_stream<<<296, 128, smem>>>(d, total / 296, chunk, slots, 2);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("2 CTAs/SM, chunk 16384 B x %2d slots each = %3d KiB in flight per SM: %7.1f GB/s\n", slots, 2 * kb, total / best / 1e6);
    }
    return 0;
}
