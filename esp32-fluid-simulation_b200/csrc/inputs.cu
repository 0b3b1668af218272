// The reference's own initial condition and input path on the device, so that
// ensembles can be seeded and driven without host round trips:
//   init_wheel_kernel + smooth_j_kernel + smooth_i_kernel   setup(), ino:196-241
//   touch_to_drags_kernel                                    touch_routine(), ino:63-96
// Pinned bit for bit against the sketch itself compiled on the host (oracle/ino_shim.cpp).
#include <cmath>

#include "kernels.h"

namespace fs {

// ---- setup(), ino:196-241 ------------------------------------------------------------------------
// Colour wheel (ino:204-219): angle = atan2f(-(i - center_i), j - center_j) in float, compared with
// -PI/3 and PI/3 in double.  The device evaluates atan2 in DOUBLE and compares with the midpoints of
// the two floats that surround each threshold — exactly the float comparison of a correctly rounded
// atan2f (checked on the host against glibc's atan2f for every offset up to +-8300 nodes: identical
// sectors, the closest angle is 1.3e-9 away from a midpoint).
__global__ void init_wheel_kernel(float2 *__restrict__ v, uint32_t *__restrict__ c, int batch, int dim_x, int dim_y,
                                  double mid_lo, double mid_hi, uint32_t full)
{
    const size_t n = (size_t)dim_x * dim_y;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * batch) return;
    const int node = (int)(k % n);
    const int j = node / dim_x, i = node - j * dim_x;
    const double angle = atan2((double)(float)(-(i - dim_x / 2)), (double)(float)(j - dim_y / 2));
    const int sector = angle < mid_lo ? 0 : (angle < mid_hi ? 1 : 2);   // red / green / blue (ino:211-217)
    v[k] = make_float2(0.0f, 0.0f);                                     // ino:197-201
    c[3 * k + 0] = sector == 0 ? full : 0u;
    c[3 * k + 1] = sector == 1 ? full : 0u;
    c[3 * k + 2] = sector == 2 ? full : 0u;
}

// 0.25f * a + 0.5f * b + 0.25f * c over Vector3<UQ32> operands (ino:227-228, 238-239): UQ32 -> float per
// operand, left-to-right float sum, one UQ32 conversion at the assignment
__device__ __forceinline__ uint32_t smooth121(uint32_t lo, uint32_t mid, uint32_t hi)
{
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(uq32_to_float(lo), 0.25f), __fmul_rn(uq32_to_float(mid), 0.5f)),
                              __fmul_rn(uq32_to_float(hi), 0.25f));
    return uq32_from_float(s);
}

// First pass (ino:220-230): IN PLACE along j for every i — node (i, j-1) is already smoothed when (i, j)
// is formed, so each (grid, i, channel) is one sequential chain; the chains are independent.
__global__ void smooth_j_kernel(uint32_t *__restrict__ c, int batch, int dim_x, int dim_y)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (grid, i, channel), channel fastest
    if (k >= (size_t)batch * dim_x * 3) return;
    const int g = (int)(k / ((size_t)dim_x * 3)), w = (int)(k - (size_t)g * dim_x * 3);   // w = 3*i + channel
    uint32_t *q = c + (size_t)g * dim_x * dim_y * 3 + w;
    const size_t row = (size_t)dim_x * 3;
    uint32_t centre = q[0], left = centre;                                // j == 0: left = center
    for (int j = 0; j < dim_y; j++) {
        const uint32_t right = j == dim_y - 1 ? centre : q[(size_t)(j + 1) * row];
        const uint32_t out = smooth121(left, centre, right);
        q[(size_t)j * row] = out;
        left = out;                                                       // the next node reads the SMOOTHED value
        centre = right;
    }
}

// Second pass (ino:231-241): in place along i for every j.
__global__ void smooth_i_kernel(uint32_t *__restrict__ c, int batch, int dim_x, int dim_y)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (grid, j, channel)
    if (k >= (size_t)batch * dim_y * 3) return;
    const int ch = (int)(k % 3);
    const size_t gj = k / 3;                                              // grid * dim_y + j
    uint32_t *q = c + gj * dim_x * 3 + ch;
    uint32_t centre = q[0], top = centre;
    for (int i = 0; i < dim_x; i++) {
        const uint32_t bot = i == dim_x - 1 ? centre : q[(size_t)(i + 1) * 3];
        const uint32_t out = smooth121(top, centre, bot);
        q[(size_t)i * 3] = out;
        top = out;
        centre = bot;
    }
}

int launch_init_color_wheel(const Launch &L, float2 *v, uint32_t *c, int batch, int dim_x, int dim_y)
{
    if (batch <= 0) return 0;
    // thresholds: the midpoints of the floats around -PI/3 and PI/3 (PI as in Arduino.h)
    const double PI_ = 3.1415926535897932384626433832795;
    auto mid = [](double t) {
        float a = (float)t;
        if ((double)a >= t) a = nextafterf(a, -INFINITY);
        const float b = nextafterf(a, INFINITY);
        return ((double)a + (double)b) / 2;
    };
    const uint32_t full = 0xFFFFFFFFu;   // UQ32(float(UINT32_MAX)) = UQ32(2^32): saturates (the defined deviation, uq32.h:13)
    const size_t n = (size_t)batch * dim_x * dim_y;
    init_wheel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(v, c, batch, dim_x, dim_y, mid(-PI_ / 3),
                                                                       mid(PI_ / 3), full);
    const size_t nj = (size_t)batch * dim_x * 3, ni = (size_t)batch * dim_y * 3;
    smooth_j_kernel<<<(unsigned)((nj + 127) / 128), 128, 0, L.stream>>>(c, batch, dim_x, dim_y);
    smooth_i_kernel<<<(unsigned)((ni + 127) / 128), 128, 0, L.stream>>>(c, batch, dim_x, dim_y);
    *L.launches += 3;
    return (int)cudaGetLastError();
}

// ---- touch_routine(), ino:63-96 --------------------------------------------------------------------
// Arduino map() (Arduino-ESP32 v3.3.1 WMath.cpp): long arithmetic, division truncating toward zero.
__device__ __forceinline__ long long arduino_map(long long x, long long in_min, long long in_max, long long out_min,
                                                 long long out_max)
{
    const long long run = in_max - in_min;
    if (run == 0) return -1;
    return (x - in_min) * (out_max - out_min) / run + out_min;
}

struct TouchArgs {
    fs_drag *drags;         // [batch][max_drags]
    int *counts;            // [batch]
    const int *samples;     // [batch][n_samples][3] = {touched, raw x, raw y}, one per polling period
    int n_samples, batch, max_drags, n_rows, n_cols;
    int min_x, max_x, min_y, max_y;
    float polling_ms;
};

// One warp per grid.  A sample produces a drag record iff it and its predecessor are touched (ino:81);
// records keep their order and the queue drops what does not fit (xQueueSend(..., 0), ino:85) — a
// ballot + popcount compaction, 32 samples per round.
__global__ void __launch_bounds__(128) touch_to_drags_kernel(const TouchArgs a)
{
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= a.batch) return;
    const int *s = a.samples + (size_t)g * a.n_samples * 3;
    fs_drag *out = a.drags + (size_t)g * a.max_drags;
    int filled = 0;
    for (int base = 0; base < a.n_samples; base += 32) {
        const int k = base + lane;
        bool emit = false;
        fs_drag rec = {};
        if (k < a.n_samples && k > 0 && s[3 * k] != 0 && s[3 * (k - 1)] != 0) {
            const int cx = (int)arduino_map(s[3 * k + 1], a.min_x, a.max_x, 0, a.n_cols);        // ino:77
            const int cy = (int)arduino_map(s[3 * k + 2], a.min_y, a.max_y, 0, a.n_rows);        // ino:78
            const int px = (int)arduino_map(s[3 * (k - 1) + 1], a.min_x, a.max_x, 0, a.n_cols);
            const int py = (int)arduino_map(s[3 * (k - 1) + 2], a.min_y, a.max_y, 0, a.n_rows);
            rec.cx = (uint16_t)cx;                                                               // Vector2<uint16_t>(Vector2<int>)
            rec.cy = (uint16_t)cy;
            rec.vx = __fdiv_rn(__fmul_rn((float)(cx - px), 1000.0f), a.polling_ms);              // ino:82-83
            rec.vy = __fdiv_rn(__fmul_rn((float)(cy - py), 1000.0f), a.polling_ms);
            emit = true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        const int at = filled + __popc(m & ((1u << lane) - 1));
        if (emit && at < a.max_drags) out[at] = rec;
        filled += __popc(m);
    }
    if (lane == 0) a.counts[g] = filled < a.max_drags ? filled : a.max_drags;
}

int launch_touch_to_drags(const Launch &L, fs_drag *drags, int *counts, const int *samples, int n_samples, int batch,
                          int max_drags, int n_rows, int n_cols, const int cal[4], int polling_ms)
{
    if (batch <= 0) return 0;
    TouchArgs a;
    a.drags = drags; a.counts = counts; a.samples = samples;
    a.n_samples = n_samples; a.batch = batch; a.max_drags = max_drags; a.n_rows = n_rows; a.n_cols = n_cols;
    a.min_x = cal[0]; a.max_x = cal[1]; a.min_y = cal[2]; a.max_y = cal[3];
    a.polling_ms = (float)polling_ms;
    touch_to_drags_kernel<<<(batch + 3) / 4, 128, 0, L.stream>>>(a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

}  // namespace fs
