// One cell of the 4x bilinear upscale + UQ32 round + RGB565 pack + byte swap (draw_routine arithmetic,
// ino:116-177), shared by the stand-alone kernel and the dye-advect epilogue.
//
// The ramps are ACCUMULATED (c += dc, ino:137,151,160), not evaluated as c + k*dc; the adds are
// replayed in the reference's order.  The left column of a cell is recomputed from its own corners
// instead of being copied from the previous cell's right column (ino:141-143): same corner values, same
// operations, same bits.
#pragma once

#include "fs_common.cuh"

namespace fs {

// corners as UQ32 raw words: c11 = (i, j), c12 = (i, j+1), c21 = (i+1, j), c22 = (i+1, j+1), 3 channels each.
// dst = the cell's first pixel: image row 4*i, column 4*j; `pitch` in pixels; aligned8 = 8-byte stores allowed.
__device__ __forceinline__ void upscale_cell_rgb565(uint16_t *dst, size_t pitch, bool aligned8, const uint32_t (&c11)[3],
                                                    const uint32_t (&c12)[3], const uint32_t (&c21)[3],
                                                    const uint32_t (&c22)[3])
{
    uint32_t px[4][4][3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float f11 = __uint2float_rn(c11[ch]), f12 = __uint2float_rn(c12[ch]);   // ino:123-126
        const float f21 = __uint2float_rn(c21[ch]), f22 = __uint2float_rn(c22[ch]);
        float left[4], right[4];
        float a = f11;
        const float da = __fmul_rn(__fsub_rn(f21, f11), 0.25f);  // ino:134
        float b = f12;
        const float db = __fmul_rn(__fsub_rn(f22, f12), 0.25f);  // ino:148
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
            const uint32_t v565 = ((px[ii][jj][0] & 0xF8000000u) >> 16) | ((px[ii][jj][1] & 0xFC000000u) >> 21) |
                                  ((px[ii][jj][2] & 0xF8000000u) >> 27);   // ino:170-172
            w[jj] = ((v565 & 0xFFu) << 8) | (v565 >> 8);                   // ino:173
        }
        uint16_t *row = dst + (size_t)ii * pitch;
        if (aligned8) {
            *reinterpret_cast<uint2 *>(row) = make_uint2(w[0] | (w[1] << 16), w[2] | (w[3] << 16));
        } else {
#pragma unroll
            for (int jj = 0; jj < 4; jj++) row[jj] = (uint16_t)w[jj];
        }
    }
}

}  // namespace fs
