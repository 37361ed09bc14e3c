// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
//   D[pixel, cout] = sum over (segment, channel chunk, kx, ky) A[pixel shifted by (ky,kx), ck] * W[cout, ck]
//
// * A is never materialised as im2col.  For every (segment, chunk, kx) the producer issues ONE
//   TMA box load of (rows + kh - 1) x TW pixels x ck channels from the NHWC activation, shifted by
//   kx - kw/2 in x; out-of-image elements are zero-filled by TMA, which IS the conv's zero padding.
//   Inside the box, the A operand of filter row ky is the contiguous run of 128 pixels that starts
//   ky*TW pixels further down, so the kh filter rows reuse the same shared-memory bytes through
//   nothing but a different UMMA descriptor start address (always a whole number of 8-row swizzle
//   atoms because TW % 8 == 0).
// * W is pre-packed by vsrb_pack_conv_weight into the exact swizzled K-major shared-memory image
//   of each stage and copied with cp.async.bulk; when one (group, n_block) worth of weights fits,
//   it is loaded once per persistent CTA and stays resident.
// * One CTA per SM, persistent over output tiles.  Warp 0 = TMA producer, warp 1 = MMA issuer
//   (one elected lane issues tcgen05.mma, accumulators live in TMEM, double buffered), warp 2 =
//   TMEM allocator, warps 4..7 = epilogue (tcgen05.ld -> bias/act/residual/pixel-shuffle/skip ->
//   global).  smem full/empty and TMEM full/empty mbarrier rings connect the roles.
//
// Replaces: F.conv2d behind nn.Conv2d at reference conv.py:89-92,101-103; upsampling.py:10-12;
// basicvsr.py:75-82; realbasicvsr.py:28-29; spynet.py:16-21 (see include/vsrb200.h).
#include "common.cuh"

namespace vsrb {

static constexpr int kMaxSlots = 8;
static constexpr int kThreads = 256;
static constexpr int kSmemMax = 232448;   // 227 KiB opt-in maximum per CTA on sm_100
static constexpr int kCtrlBytes = 1024;

struct TcParams {
    CUtensorMap tmap[2];
    const uint8_t* w;     // packed weights (after the bias header)
    EpiParams epi;
    int n_seg;
    int seg_chunks[2], seg_ck[2], seg_rowbytes[2], seg_layout[2], seg_bstage[2], seg_abytes[2], seg_stages[2];
    int kh, kw;
    int H, W;
    int TW, rows_sub, MT, box_rows;
    int tiles_x, tiles_per_img;
    int imgs_per_group, groups;
    int n_tile, n_blocks;
    int stages_per_tile;
    int num_slots, slot_bytes;
    int resident;
    uint32_t wblock_bytes, wres_bytes;
    int acc_cols;
    int* dbg;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a mis-programmed pipeline raises the debug flag and lets the kernel run to
// completion with garbage instead of hanging the GPU.  `dead` is sticky per thread.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* dbg, int code, bool& dead) {
    if (dead) return;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        if (mbar_try_wait(bar, parity)) return;
        if ((it & 1023u) == 1023u) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000LL || *reinterpret_cast<volatile int*>(dbg) != 0) {
                atomicCAS(dbg, 0, code);
                dead = true;
                return;
            }
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, swizzled shared-memory operand descriptor (sm_100 "version 1"):
//  [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 8 rows |
//  [46,48) version=1 | [61,64) layout type (2 = 128B, 4 = 64B, 6 = 32B swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t rowbytes, uint32_t layout) {
    uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    uint32_t hi = ((rowbytes * 8u) >> 4) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | lo;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                // swizzle atoms need 1 KiB alignment
    uint8_t* base_ptr = smem_raw + (base - raw);
    // control block
    const uint32_t full0 = base, empty0 = base + 8 * kMaxSlots, tfull0 = base + 16 * kMaxSlots,
                   tempty0 = tfull0 + 16, wbar = tempty0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + 16 * kMaxSlots + 48);
    const uint32_t wres = base + kCtrlBytes;
    const uint32_t slots0 = wres + P.wres_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.y / P.n_blocks, qb = blockIdx.y - g * P.n_blocks;
    const int tiles_g = P.imgs_per_group * P.tiles_per_img;
    const int rows_tile = P.rows_sub * P.MT;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < P.num_slots; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, 4);
        }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint8_t* wsrc = P.w + ((size_t)g * P.n_blocks + qb) * P.wblock_bytes;
    bool dead = false;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            if (P.resident) {
                mbar_expect_tx(wbar, P.wblock_bytes);
                uint32_t off = 0;
                for (int s = 0; s < P.n_seg; ++s)
                    for (int i = 0; i < P.seg_stages[s]; ++i) {
                        bulk_load(wres + off, wsrc + off, P.seg_bstage[s], wbar);
                        off += P.seg_bstage[s];
                    }
            }
            int slot = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles_g; tile += gridDim.x) {
                const int li = tile / P.tiles_per_img;
                const int t = tile - li * P.tiles_per_img;
                const int img = g * P.imgs_per_group + li;
                const int ty = t / P.tiles_x, tx = t - ty * P.tiles_x;
                const int y0 = ty * rows_tile - P.kh / 2, x0 = tx * P.TW - P.kw / 2;
                uint32_t boff = 0;
                for (int s = 0; s < P.n_seg; ++s) {
                    for (int local = 0; local < P.seg_stages[s]; ++local) {
                        const int chunk = local / P.kw, kx = local - chunk * P.kw;
                        const uint32_t sa = slots0 + slot * P.slot_bytes;
                        mbar_wait(empty0 + 8 * slot, phase ^ 1, P.dbg, 1, dead);
                        mbar_expect_tx(full0 + 8 * slot, P.seg_abytes[s] + (P.resident ? 0 : P.seg_bstage[s]));
                        tma_load_4d(&P.tmap[s], full0 + 8 * slot, sa, chunk * P.seg_ck[s], x0 + kx, y0, img);
                        if (!P.resident) bulk_load(sa + P.seg_abytes[s], wsrc + boff, P.seg_bstage[s], full0 + 8 * slot);
                        boff += P.seg_bstage[s];
                        if (++slot == P.num_slots) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =================================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 @17, M>>4 @24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.n_tile >> 3) << 17) | (8u << 24);
            if (P.resident) mbar_wait(wbar, 0, P.dbg, 2, dead);
            int slot = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < tiles_g; tile += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, P.dbg, 3, dead);
                tc_fence_after();
                const uint32_t d0 = tmem_base + acc * P.acc_cols;
                uint32_t boff = 0;
                bool first = true;
                for (int s = 0; s < P.n_seg; ++s) {
                    const uint32_t rb = P.seg_rowbytes[s], lay = P.seg_layout[s];
                    const int ksteps = P.seg_ck[s] >> 4;
                    for (int local = 0; local < P.seg_stages[s]; ++local) {
                        const uint32_t sa = slots0 + slot * P.slot_bytes;
                        const uint32_t sb = P.resident ? (wres + boff) : (sa + P.seg_abytes[s]);
                        mbar_wait(full0 + 8 * slot, phase, P.dbg, 4, dead);
                        tc_fence_after();
                        for (int m = 0; m < P.MT; ++m) {
                            for (int ky = 0; ky < P.kh; ++ky) {
                                const uint32_t arow = sa + (uint32_t)((m * P.rows_sub + ky) * P.TW) * rb;
                                const uint32_t brow = sb + (uint32_t)(ky * P.n_tile) * rb;
                                for (int k = 0; k < ksteps; ++k) {
                                    const uint32_t accum = (first && ky == 0 && k == 0) ? 0u : 1u;
                                    umma_bf16(d0 + m * P.n_tile, make_desc(arow + k * 32, rb, lay),
                                              make_desc(brow + k * 32, rb, lay), idesc, accum);
                                }
                            }
                        }
                        umma_commit(empty0 + 8 * slot);   // frees the slot when these MMAs retire
                        first = false;
                        boff += P.seg_bstage[s];
                        if (++slot == P.num_slots) { slot = 0; phase ^= 1; }
                    }
                }
                umma_commit(tfull0 + 8 * acc);            // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue ===================================
        const int wq = warp - 4;                          // == warp % 4: the TMEM lane quarter this warp may read
        const int p = wq * 32 + lane;
        const int py = p / P.TW, px = p - py * P.TW;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles_g; tile += gridDim.x) {
            const int li = tile / P.tiles_per_img;
            const int t = tile - li * P.tiles_per_img;
            const int img = g * P.imgs_per_group + li;
            const int ty = t / P.tiles_x, tx = t - ty * P.tiles_x;
            mbar_wait(tfull0 + 8 * acc, acc_phase, P.dbg, 5, dead);
            tc_fence_after();
            for (int m = 0; m < P.MT; ++m) {
                const int y = ty * rows_tile + m * P.rows_sub + py, x = tx * P.TW + px;
                const bool valid = (y < P.H) && (x < P.W);
                const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * P.acc_cols + m * P.n_tile;
                for (int c0 = 0; c0 < P.n_tile; c0 += 16) {
                    float v[16];
                    tmem_ld16(t0 + c0, v);
                    if (valid) epi_store16<__nv_bfloat16>(P.epi, g, img, y, x, qb * P.n_tile + c0, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int g_sm_count = 0;
static bool g_attr_set = false;

int launch_conv_tc(const vsrb_conv_args* a, const ConvPlan& p, cudaStream_t stream) {
    EncodeTiledFn encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VSRB_E_NODEVICE;
    }
    if (!g_sm_count) {
        int dev = 0;
        VSRB_CUDA(cudaGetDevice(&dev));
        VSRB_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!g_attr_set) {
        VSRB_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        g_attr_set = true;
    }
    TcParams P;
    memset(&P, 0, sizeof(P));
    P.n_seg = p.n_seg; P.kh = p.kh; P.kw = p.kw; P.H = a->h; P.W = a->w;
    P.groups = p.groups; P.imgs_per_group = a->imgs_per_group;
    P.n_tile = p.n_tile; P.n_blocks = p.n_blocks; P.stages_per_tile = p.stages_per_tile;
    P.wblock_bytes = (uint32_t)p.wblock_bytes;
    P.w = reinterpret_cast<const uint8_t*>(a->packed) + p.bias_bytes;
    P.dbg = debug_flag();
    fill_epi(a, p, &P.epi);

    P.TW = a->w <= 8 ? 8 : 16;
    P.rows_sub = 128 / P.TW;
    const int units = p.groups * p.n_blocks;
    const int ctas_budget = (a->max_ctas > 0 ? a->max_ctas : g_sm_count);
    const int avail = kSmemMax - kCtrlBytes - 1024;   // control block + alignment slack
    int MT = (2 * 2 * p.n_tile <= 512) ? 2 : 1;
    if (MT == 2) {
        long tiles2 = (long)a->imgs_per_group * ceil_div(a->h, 2 * P.rows_sub) * ceil_div(a->w, P.TW) * units;
        if (a->h <= P.rows_sub || tiles2 < 2L * ctas_budget) MT = 1;
    }
    for (;; MT = 1) {
        P.MT = MT;
        P.box_rows = P.rows_sub * MT + p.kh - 1;
        int amax = 0, abmax = 0;
        for (int s = 0; s < p.n_seg; ++s) {
            P.seg_abytes[s] = P.box_rows * P.TW * p.seg[s].rowbytes;
            amax = amax > P.seg_abytes[s] ? amax : P.seg_abytes[s];
            int ab = P.seg_abytes[s] + p.b_stage_bytes[s];
            abmax = abmax > ab ? abmax : ab;
        }
        const int slot_res = (int)round_up(amax, 1024), slot_str = (int)round_up(abmax, 1024);
        const int wres = (int)round_up(p.wblock_bytes, 1024);
        if (P.box_rows <= 256 && wres + 3 * slot_res <= avail) {
            P.resident = 1; P.wres_bytes = wres; P.slot_bytes = slot_res;
            P.num_slots = (avail - wres) / slot_res;
        } else {
            P.resident = 0; P.wres_bytes = 0; P.slot_bytes = slot_str;
            P.num_slots = avail / slot_str;
        }
        if (P.num_slots > kMaxSlots) P.num_slots = kMaxSlots;
        if (P.num_slots >= 2 && P.box_rows <= 256) break;
        if (MT == 1) {
            set_error("conv tile does not fit shared memory (kh=%d n_tile=%d)", p.kh, p.n_tile);
            return VSRB_E_SMEM;
        }
    }
    P.acc_cols = P.MT * p.n_tile;
    P.tiles_x = ceil_div(a->w, P.TW);
    P.tiles_per_img = P.tiles_x * ceil_div(a->h, P.rows_sub * P.MT);
    for (int s = 0; s < p.n_seg; ++s) {
        const SegPlan& sp = p.seg[s];
        P.seg_chunks[s] = sp.chunks; P.seg_ck[s] = sp.ck; P.seg_rowbytes[s] = sp.rowbytes;
        P.seg_layout[s] = sp.layout; P.seg_bstage[s] = p.b_stage_bytes[s]; P.seg_stages[s] = sp.chunks * p.kw;
        cuuint64_t dims[4] = {(cuuint64_t)a->in_c[s], (cuuint64_t)a->w, (cuuint64_t)a->h, (cuuint64_t)a->batch};
        cuuint64_t strides[3] = {(cuuint64_t)a->in_c[s] * 2, (cuuint64_t)a->w * a->in_c[s] * 2,
                                 (cuuint64_t)a->h * a->w * a->in_c[s] * 2};
        cuuint32_t box[4] = {(cuuint32_t)sp.ck, (cuuint32_t)P.TW, (cuuint32_t)P.box_rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUtensorMapSwizzle sw = sp.ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (sp.ck == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        VSRB_CHECK_ARG(a->in_c[s] % 8 == 0, "segment %d: channel stride %d must be a multiple of 8", s, a->in_c[s]);
        VSRB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->in[s]) & 15) == 0, "segment %d: pointer not 16-byte aligned", s);
        CUresult r = encode(&P.tmap[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->in[s]), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed with %d (seg %d, c=%d w=%d h=%d b=%d box=%d,%d,%d)", (int)r, s,
                      a->in_c[s], a->w, a->h, a->batch, sp.ck, P.TW, P.box_rows);
            return VSRB_E_CUDA;
        }
    }
    const int tiles_g = a->imgs_per_group * P.tiles_per_img;
    int ctas_x = ctas_budget / units;
    if (ctas_x < 1) ctas_x = 1;
    if (ctas_x > tiles_g) ctas_x = tiles_g;
    const int smem = kCtrlBytes + 1024 + (int)P.wres_bytes + P.num_slots * P.slot_bytes;
    dim3 grid(ctas_x, units);
    conv_tc_kernel<<<grid, kThreads, smem, stream>>>(P);
    VSRB_LAUNCH_CHECK();
    return VSRB_OK;
}

}  // namespace vsrb
