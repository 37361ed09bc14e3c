// fp32 convolution on the FFMA pipes ("fp32 mode": full-precision operands and accumulate).
// Same geometry, packed-weight contract and epilogues as the tensor-core kernel; used when the
// caller needs the reference's fp32 numerics (max-abs <= 1e-4 end to end), which bf16 operands
// cannot deliver.
//
// Tile: 16 x 16 output pixels per 128-thread CTA, each thread 2 pixels (rows r and r+8) x NT
// output channels in registers.  Input channels are staged 8 at a time as channel-planar halo
// patches (conflict-free reads along x), weights one filter row at a time ([kw][8][NT], read as
// broadcast float4s).
#include "common.cuh"

namespace vsrb {

struct F32Params {
    const float* in[4];
    int in_c[4], seg_c[4];
    int n_seg;
    const float* w;   // [group][tap][cin_packed][cout_pad]
    EpiParams epi;
    int kh, kw, H, W, batch, imgs_per_group, cin_packed, cout_pad;
    int tiles_x, tiles_y;
};

static constexpr int kTH = 16, kTW = 16, kCC = 8;

template <int NT>
__global__ void __launch_bounds__(128) conv_f32_kernel(F32Params P) {
    extern __shared__ float sm[];
    const int PH = kTH + P.kh - 1, PW = kTW + P.kw - 1;
    float* patch = sm;                                  // [kCC][PH][PW]
    float* wrow = sm + kCC * PH * PW;                   // [kw][kCC][NT]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;             // ty in 0..7; pixels (ty, tx) and (ty+8, tx)
    int t = blockIdx.x;
    const int img = t / (P.tiles_x * P.tiles_y);
    t -= img * P.tiles_x * P.tiles_y;
    const int by = t / P.tiles_x, bx = t - by * P.tiles_x;
    const int y0 = by * kTH, x0 = bx * kTW;
    const int g = img / P.imgs_per_group;
    const int n0 = blockIdx.y * NT;

    float acc0[NT], acc1[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }

    int cbase = 0;   // position on the packed input-channel axis
    for (int s = 0; s < P.n_seg; ++s) {
        const float* inp = P.in[s] + (size_t)img * P.H * P.W * P.in_c[s];
        for (int c0 = 0; c0 < P.seg_c[s]; c0 += kCC) {
            const int cc = min(kCC, P.seg_c[s] - c0);
            __syncthreads();
            // halo patch, zero padded
            for (int i = tid; i < kCC * PH * PW; i += 128) {
                int c = i % kCC;
                int r = i / kCC;
                int px = r % PW, py = r / PW;
                int yy = y0 + py - P.kh / 2, xx = x0 + px - P.kw / 2;
                float v = 0.f;
                if (c < cc && yy >= 0 && yy < P.H && xx >= 0 && xx < P.W)
                    v = __ldg(inp + ((size_t)yy * P.W + xx) * P.in_c[s] + c0 + c);
                patch[(c * PH + py) * PW + px] = v;
            }
            for (int ky = 0; ky < P.kh; ++ky) {
                __syncthreads();
                for (int i = tid; i < P.kw * kCC * NT; i += 128) {
                    int n = i % NT;
                    int r = i / NT;
                    int c = r % kCC, kx = r / kCC;
                    float v = 0.f;
                    if (c < cc)
                        v = __ldg(P.w + (((size_t)g * P.kh * P.kw + ky * P.kw + kx) * P.cin_packed + cbase + c0 + c) * P.cout_pad +
                                  n0 + n);
                    wrow[(kx * kCC + c) * NT + n] = v;
                }
                __syncthreads();
                for (int kx = 0; kx < P.kw; ++kx) {
#pragma unroll
                    for (int c = 0; c < kCC; ++c) {
                        const float a0 = patch[(c * PH + ty + ky) * PW + tx + kx];
                        const float a1 = patch[(c * PH + ty + 8 + ky) * PW + tx + kx];
                        const float4* wp = reinterpret_cast<const float4*>(wrow + (kx * kCC + c) * NT);
#pragma unroll
                        for (int n = 0; n < NT / 4; ++n) {
                            float4 wv = wp[n];
                            acc0[4 * n] = fmaf(a0, wv.x, acc0[4 * n]);
                            acc0[4 * n + 1] = fmaf(a0, wv.y, acc0[4 * n + 1]);
                            acc0[4 * n + 2] = fmaf(a0, wv.z, acc0[4 * n + 2]);
                            acc0[4 * n + 3] = fmaf(a0, wv.w, acc0[4 * n + 3]);
                            acc1[4 * n] = fmaf(a1, wv.x, acc1[4 * n]);
                            acc1[4 * n + 1] = fmaf(a1, wv.y, acc1[4 * n + 1]);
                            acc1[4 * n + 2] = fmaf(a1, wv.z, acc1[4 * n + 2]);
                            acc1[4 * n + 3] = fmaf(a1, wv.w, acc1[4 * n + 3]);
                        }
                    }
                }
            }
        }
        cbase += P.seg_c[s];
    }

    const int x = x0 + tx;
    if (x < P.W) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int y = y0 + ty + half * 8;
            if (y >= P.H) continue;
#pragma unroll
            for (int c0 = 0; c0 < NT; c0 += 16) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = half ? acc1[c0 + i] : acc0[c0 + i];
                epi_bias16(P.epi, g, n0 + c0, v);
                epi_act16(P.epi, v);
                epi_store16<float>(P.epi, g, img, y, x, n0 + c0, v);
            }
        }
    }
}

int launch_conv_f32(const vsrb_conv_args* a, const ConvPlan& p, cudaStream_t stream) {
    F32Params P;
    memset(&P, 0, sizeof(P));
    P.n_seg = p.n_seg;
    for (int s = 0; s < p.n_seg; ++s) {
        P.in[s] = reinterpret_cast<const float*>(a->in[s]);
        P.in_c[s] = a->in_c[s];
        P.seg_c[s] = p.seg[s].c;
    }
    P.w = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(a->packed) + p.bias_bytes);
    fill_epi(a, p, &P.epi);
    P.kh = p.kh; P.kw = p.kw; P.H = a->h; P.W = a->w; P.batch = a->batch;
    P.imgs_per_group = a->imgs_per_group; P.cin_packed = p.cin_packed; P.cout_pad = p.cout_pad;
    P.tiles_x = ceil_div(a->w, kTW);
    P.tiles_y = ceil_div(a->h, kTH);
    if (a->epilogue == VSRB_EPI_NHWC) {
        VSRB_CHECK_ARG(a->out_c % 4 == 0 && (!a->residual || a->res_c % 4 == 0), "fp32 channel strides must be %% 4");
    }
    const int NT = p.cout_pad % 64 == 0 ? 64 : (p.cout_pad % 32 == 0 ? 32 : 16);
    const int PH = kTH + p.kh - 1, PW = kTW + p.kw - 1;
    const size_t smem = ((size_t)kCC * PH * PW + (size_t)p.kw * kCC * NT) * sizeof(float);
    dim3 grid(P.tiles_x * P.tiles_y * a->batch, p.cout_pad / NT);
    if (NT == 64) conv_f32_kernel<64><<<grid, 128, smem, stream>>>(P);
    else if (NT == 32) conv_f32_kernel<32><<<grid, 128, smem, stream>>>(P);
    else conv_f32_kernel<16><<<grid, 128, smem, stream>>>(P);
    VSRB_LAUNCH_CHECK();
    return VSRB_OK;
}

}  // namespace vsrb
