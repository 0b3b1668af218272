// TEST INFRASTRUCTURE — not part of the product path.
//
// Compiles the reference SKETCH ITSELF (ESP32-fluid-simulation.ino), unmodified, from where it
// lies under /root/reference, against the stand-in headers in oracle/ino_stubs/, and exposes its
// four routines through extern "C" wrappers:
//   setup()          ino:194-246  initial conditions (colour wheel + in-place 1-2-1 smoothing)
//   loop()           ino:249-289  the sim step: call ORDER, drag overwrite, constants (K=10, w=1.96)
//   draw_routine()   ino:99-191   4x bilinear upscale + UQ32 round + RGB565 pack + byte swap
//   touch_routine()  ino:63-96    touch samples -> drag records (map(), finite difference, queue)
// so the restatements in oracle/fluid_oracle.c (oracle_step, oracle_upscale4_rgb565,
// oracle_init_color_wheel) and esp32-fluid-simulation_b200/synth.py (touch_drags) are pinned against
// COMPILED reference code instead of being restated twice (tests/test_oracle_vs_ref.py).
// Built by oracle/Makefile into oracle/_ref/libfluid_ref.so (same flags as ref_shim.cpp).
#include "ino_stubs/arduino_stub.h"

int g_tft_width = 240, g_tft_height = 320;       // CYD defaults: 61 x 81 nodes
uint16_t *g_frame = nullptr;
std::vector<TouchSample> g_touch_script;
size_t g_touch_pos = 0;

// vTaskDelay(POLLING_PERIOD / portTICK_PERIOD_MS) ends one poll of touch_routine (ino:93): advance the script
#define vTaskDelay(t) do { if ((t) != 0) g_touch_pos++; } while (0)

#include "ESP32-fluid-simulation.ino"

#undef vTaskDelay

namespace {
void set_dims(int dim_x, int dim_y)
{
    g_tft_width = (dim_x - 1) * SCALING;          // N_ROWS = dim_x (the fast axis, ino:37,253)
    g_tft_height = (dim_y - 1) * SCALING;         // N_COLS = dim_y
    delete[] velocity_field;
    delete[] color_field;
    velocity_field = new Vector2<float>[(size_t)dim_x * dim_y];
    color_field = new Vector3<UQ32>[(size_t)dim_x * dim_y];
}
}  // namespace

extern "C" {

// setup(), ino:194-246 -> the initial velocity and dye fields
void ref_ino_setup(float *v_out, uint32_t *c_out, int dim_x, int dim_y)
{
    set_dims(dim_x, dim_y);
    setup();
    memcpy(v_out, velocity_field, (size_t)dim_x * dim_y * sizeof(Vector2<float>));
    memcpy(c_out, color_field, (size_t)dim_x * dim_y * sizeof(Vector3<UQ32>));
}

// loop(), ino:249-289, with `n` drag records waiting in the queue.  K = 10, omega = 1.96, dt = DT
// are the sketch's literals.  Records are queued directly (not through xQueueSend, whose depth-10
// limit belongs to the touch task); out-of-range records would make ino:266-268 write out of bounds
// and must not be passed.
void ref_ino_loop(float *v, uint32_t *c, const void *drags, int n, int dim_x, int dim_y)
{
    set_dims(dim_x, dim_y);
    const size_t nn = (size_t)dim_x * dim_y;
    memcpy(velocity_field, v, nn * sizeof(Vector2<float>));
    memcpy(color_field, c, nn * sizeof(Vector3<UQ32>));
    drag_queue->q.clear();
    const unsigned char *b = (const unsigned char *)drags;
    for (int k = 0; k < n; k++) drag_queue->q.emplace_back(b + (size_t)k * sizeof(struct drag), b + (size_t)(k + 1) * sizeof(struct drag));
    xSemaphoreGive(color_consumed);               // the draw task has consumed the previous frame
    loop();
    memcpy(v, velocity_field, nn * sizeof(Vector2<float>));
    memcpy(c, color_field, nn * sizeof(Vector3<UQ32>));
}

// draw_routine(), ino:99-191: one frame.  out = (dim_x-1)*4 rows x (dim_y-1)*4 columns of RGB565.
void ref_ino_draw(uint16_t *out, const uint32_t *c, int dim_x, int dim_y)
{
    set_dims(dim_x, dim_y);
    memcpy(color_field, c, (size_t)dim_x * dim_y * sizeof(Vector3<UQ32>));
    g_frame = out;
    color_produced->count = 1;                    // one frame is ready
    try {
        draw_routine(nullptr);
    } catch (const StopTask &) {                  // blocked waiting for the NEXT frame: this one is done
    }
    g_frame = nullptr;
}

// touch_routine(), ino:63-96: `n` polls of the panel, samples = {touched, raw x, raw y}.  Returns
// the number of drag records the task queued (at most the queue depth, 10: ino:49,85), copied to out.
int ref_ino_touch(void *out, int max_out, const int *samples, int n, int dim_x, int dim_y)
{
    set_dims(dim_x, dim_y);                       // map() targets [0, N_COLS] x [0, N_ROWS] (ino:77-78)
    g_touch_script.clear();
    for (int k = 0; k < n; k++) g_touch_script.push_back(TouchSample{samples[3 * k], samples[3 * k + 1], samples[3 * k + 2]});
    g_touch_pos = 0;
    drag_queue->q.clear();
    try {
        touch_routine(nullptr);
    } catch (const StopTask &) {                  // script exhausted
    }
    int m = 0;
    struct drag msg;
    while (m < max_out && xQueueReceive(drag_queue, &msg, 0) == pdTRUE) memcpy((unsigned char *)out + (size_t)(m++) * sizeof(msg), &msg, sizeof(msg));
    return m;
}

int ref_ino_queue_depth(void) { return (int)drag_queue->cap; }

}  // extern "C"
