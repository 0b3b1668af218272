// 4x bilinear upscale of the dye + UQ32 round + RGB565 pack + byte swap
// (draw_routine arithmetic, ino:116-177).
//
// The image is TRANSPOSED relative to the sim grid: image rows run along the
// sim's fast axis i, image columns along the slow axis j (ino:165,174,183).
// A CTA therefore stages a 33x33-node block through shared memory: the dye is
// read coalesced along i, and each warp then owns 32 consecutive j so every
// image-row segment it writes is 256 contiguous bytes.
//
// The ramps are ACCUMULATED (c += dc, ino:137,151,160), not evaluated as
// c + k*dc; the adds are replayed in the reference's order.
#include "kernels.h"

namespace fs {

constexpr int UP_T = 32;            // cells per tile edge
constexpr int UP_PITCH = 33 * 3;    // words per staged row (odd => conflict-free column reads)

__global__ void __launch_bounds__(256)
upscale4_rgb565_kernel(uint16_t *__restrict__ out, const uint32_t *__restrict__ c, int dim_x,
                       int dim_y)
{
    __shared__ uint32_t s[33 * UP_PITCH];
    const int i0 = blockIdx.x * UP_T, j0 = blockIdx.y * UP_T;
    const int tid = threadIdx.y * 32 + threadIdx.x;

    // stage nodes (i0..i0+32, j0..j0+32), clipped: rows of 33*3 consecutive words
    const int ni = min(33, dim_x - i0), nj = min(33, dim_y - j0);
    for (int r = threadIdx.y; r < nj; r += 8) {
        const uint32_t *src = c + ((size_t)(j0 + r) * dim_x + i0) * 3;
        for (int w = threadIdx.x; w < ni * 3; w += 32) s[r * UP_PITCH + w] = __ldg(src + w);
    }
    __syncthreads();
    (void)tid;

    const int tj = threadIdx.x;  // cell column j = j0 + tj
    const int j = j0 + tj;
    if (j >= dim_y - 1) return;
    const size_t pitch = 4 * (size_t)(dim_y - 1);
    for (int ti = threadIdx.y; ti < UP_T; ti += 8) {
        const int i = i0 + ti;
        if (i >= dim_x - 1) break;
        uint32_t px[4][4][3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float c11 = __uint2float_rn(s[tj * UP_PITCH + ti * 3 + ch]);           // (i,   j)
            const float c12 = __uint2float_rn(s[(tj + 1) * UP_PITCH + ti * 3 + ch]);     // (i,   j+1)
            const float c21 = __uint2float_rn(s[tj * UP_PITCH + (ti + 1) * 3 + ch]);     // (i+1, j)
            const float c22 = __uint2float_rn(s[(tj + 1) * UP_PITCH + (ti + 1) * 3 + ch]);
            float left[4], right[4];
            float a = c11;
            const float da = __fmul_rn(__fsub_rn(c21, c11), 0.25f);  // ino:134
            float b = c12;
            const float db = __fmul_rn(__fsub_rn(c22, c12), 0.25f);  // ino:148
#pragma unroll
            for (int ii = 0; ii < 4; ii++) {
                left[ii] = a;
                a = __fadd_rn(a, da);
                right[ii] = b;
                b = __fadd_rn(b, db);
            }
#pragma unroll
            for (int ii = 0; ii < 4; ii++) {
                float r = left[ii];
                const float dr = __fmul_rn(__fsub_rn(right[ii], r), 0.25f);  // ino:157
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    px[ii][jj][ch] = __float2uint_rz(__fadd_rn(r, 0.5f));     // ino:168
                    r = __fadd_rn(r, dr);
                }
            }
        }
#pragma unroll
        for (int ii = 0; ii < 4; ii++) {
            uint32_t w[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                uint32_t v565 = ((px[ii][jj][0] & 0xF8000000u) >> 16) |
                                ((px[ii][jj][1] & 0xFC000000u) >> 21) |
                                ((px[ii][jj][2] & 0xF8000000u) >> 27);       // ino:170-172
                w[jj] = ((v565 & 0xFFu) << 8) | (v565 >> 8);                 // ino:173
            }
            uint2 q = make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
            *reinterpret_cast<uint2 *>(out + (4 * (size_t)i + ii) * pitch + 4 * (size_t)j) = q;
        }
    }
}

int launch_upscale4_rgb565(const Launch &L, uint16_t *out, const uint32_t *c, int dim_x, int dim_y)
{
    if (dim_x < 2 || dim_y < 2) return 0;
    dim3 block(32, 8), grid((dim_x - 1 + UP_T - 1) / UP_T, (dim_y - 1 + UP_T - 1) / UP_T);
    upscale4_rgb565_kernel<<<grid, block, 0, L.stream>>>(out, c, dim_x, dim_y);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int preload_upscale_kernels()
{
    FS_PRELOAD(upscale4_rgb565_kernel);
    return 0;
}

}  // namespace fs
