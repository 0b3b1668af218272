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
#include "upscale.cuh"

namespace fs {

constexpr int UP_T = 32;            // cells per tile edge
constexpr int UP_PITCH = 33 * 3;    // words per staged row (odd => conflict-free column reads)

__global__ void __launch_bounds__(256)
upscale4_rgb565_kernel(uint16_t *__restrict__ out, const uint32_t *__restrict__ c, int dim_x,
                       int dim_y, bool aligned8)
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
        uint32_t c11[3], c12[3], c21[3], c22[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            c11[ch] = s[tj * UP_PITCH + ti * 3 + ch];              // (i,   j)
            c12[ch] = s[(tj + 1) * UP_PITCH + ti * 3 + ch];        // (i,   j+1)
            c21[ch] = s[tj * UP_PITCH + (ti + 1) * 3 + ch];        // (i+1, j)
            c22[ch] = s[(tj + 1) * UP_PITCH + (ti + 1) * 3 + ch];  // (i+1, j+1)
        }
        upscale_cell_rgb565(out + 4 * (size_t)i * pitch + 4 * (size_t)j, pitch, aligned8, c11, c12, c21, c22);
    }
}

int launch_upscale4_rgb565(const Launch &L, uint16_t *out, const uint32_t *c, int dim_x, int dim_y)
{
    if (dim_x < 2 || dim_y < 2) return 0;
    dim3 block(32, 8), grid((dim_x - 1 + UP_T - 1) / UP_T, (dim_y - 1 + UP_T - 1) / UP_T);
    // 8-byte pixel-quad stores need an 8-byte aligned frame (a uint16_t* into a larger buffer may not be)
    const bool aligned8 = (uintptr_t)out % 8 == 0;
    upscale4_rgb565_kernel<<<grid, block, 0, L.stream>>>(out, c, dim_x, dim_y, aligned8);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int preload_upscale_kernels()
{
    FS_PRELOAD(upscale4_rgb565_kernel);
    return 0;
}

}  // namespace fs
