// Offset table (distortion_aware_ops.py:198-270) on the host (libm float, what the reference's eager scalar ops
// resolve to) and on the device (fp64-evaluate / fp32-round), plus the debug export of the sampling geometry.
#include <math.h>
#include <string.h>
#include <vector>

#include <atomic>

#include "da_geometry.cuh"

namespace sky {

static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{ 0 };
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct V3 {
    float x, y, z;
};

// tf.linalg.cross: three products-differences, every product and difference rounded on its own.
#pragma nv_exec_check_disable
template <class Ops>
__host__ __device__ inline V3 cross_tf(const V3 &a, const V3 &b)
{
    V3 o;
    o.x = Ops::sub(Ops::mul(a.y, b.z), Ops::mul(a.z, b.y));
    o.y = Ops::sub(Ops::mul(a.z, b.x), Ops::mul(a.x, b.z));
    o.z = Ops::sub(Ops::mul(a.x, b.y), Ops::mul(a.y, b.x));
    return o;
}

// Host arithmetic: plain fp32 operators (compiled with -ffp-contract=off) + libm float.
struct HostOps {
    static float add(float a, float b) { return a + b; }
    static float sub(float a, float b) { return a - b; }
    static float mul(float a, float b) { return a * b; }
    static float div(float a, float b) { return a / b; }
    static float tan_(float a) { return tanf(a); }
    static float cos_(float a) { return cosf(a); }
    static float sin_(float a) { return sinf(a); }
    static float atan2_(float a, float b) { return atan2f(a, b); }
    static float asin_(float a) { return asinf(a); }
};

#ifdef __CUDACC__
// Device arithmetic: IEEE fp32 intrinsics (never contracted) and transcendentals evaluated in fp64, rounded once.
struct DevOps {
    __device__ static float add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static float mul(float a, float b) { return __fmul_rn(a, b); }
    __device__ static float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static float tan_(float a) { return (float)tan((double)a); }
    __device__ static float cos_(float a) { return (float)cos((double)a); }
    __device__ static float sin_(float a) { return (float)sin((double)a); }
    __device__ static float atan2_(float a, float b) { return (float)atan2((double)a, (double)b); }
    __device__ static float asin_(float a) { return (float)asin((double)a); }
};
#endif

#define SKY_PI_D 3.141592653589793

// One row of the table.  row_out: [k*k][2].  Returns SKY_OK or SKY_ERR_UNDEFINED_COORDS.
#pragma nv_exec_check_disable
template <class Ops>
__host__ __device__ inline int offsets_row(int y, int h, int w, int k, int dilation, int skydome, float *row_out)
{
    const int n = k / 2, k2 = k * k, middle = n * (k + 1);                       // :201-202
    const float pi_f = (float)SKY_PI_D;
    const float unit_w = Ops::div((float)(2.0 * SKY_PI_D), (float)w);             // :204
    const float unit_h = Ops::div(pi_f, (float)(skydome ? 2 * h : h));           // :205
    const float rho = Ops::mul(Ops::tan_(unit_w), (float)dilation);              // :207
    const V3 v = { 0.f, 1.f, 0.f };                                              // :209
    const int xc = (int)(w * 0.5);                                               // :213
    const float theta = Ops::mul((float)((double)xc - 0.5 * (double)w), unit_w); // :220
    const float phi = skydome ? Ops::mul((float)(h - y), unit_h)                 // :221
                              : Ops::mul((float)((double)h * 0.5 - (double)y), unit_h);
    V3 p;                                                                        // :223-226
    p.x = Ops::mul(Ops::cos_(phi), Ops::cos_(theta));
    p.y = Ops::sin_(phi);
    p.z = Ops::mul(Ops::cos_(phi), Ops::sin_(theta));
    const V3 tx = cross_tf<Ops>(v, p);                                           // :228
    const V3 ty = cross_tf<Ops>(p, tx);                                          // :229

    float mid_y = 0.f, mid_x = 0.f;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 finds k[middle] (:261), pass 1 writes the differences
        int t = 0;
        for (int gy = n; gy >= -n; --gy)     // make_grid :192-194
            for (int gx = n; gx >= -n; --gx, ++t) {
                if (pass == 0 && t != middle) continue;
                const float fx = (float)gx, fy = (float)gy;
                V3 u;                        // :233-235
                u.x = Ops::add(p.x, Ops::mul(rho, Ops::add(Ops::mul(fx, tx.x), Ops::mul(fy, ty.x))));
                u.y = Ops::add(p.y, Ops::mul(rho, Ops::add(Ops::mul(fx, tx.y), Ops::mul(fy, ty.y))));
                u.z = Ops::add(p.z, Ops::mul(rho, Ops::add(Ops::mul(fx, tx.z), Ops::mul(fy, ty.z))));
                float theta_r;
                if (u.x > 0.f) theta_r = Ops::atan2_(u.z, u.x);                                   // :239-240
                else if (u.x < 0.f) {                                                             // :241-245
                    theta_r = (u.z >= 0.f) ? Ops::add(Ops::atan2_(u.z, u.x), pi_f) : Ops::sub(Ops::atan2_(u.z, u.x), pi_f);
                } else {                                                                          // :246-252
                    if (u.z > 0.f) theta_r = (float)(SKY_PI_D * 0.5);
                    else if (u.z < 0.f) theta_r = (float)(-SKY_PI_D * 0.5);
                    else return SKY_ERR_UNDEFINED_COORDS;
                }
                const float phi_r = Ops::asin_(u.y);                                              // :254
                const float x_r = Ops::mul(Ops::mul(Ops::add(Ops::div(theta_r, pi_f), 1.f), 0.5f), (float)w);   // :256
                float y_r;                                                                        // :257
                if (skydome) y_r = Ops::mul(Ops::sub(1.f, Ops::div(Ops::mul(2.f, phi_r), pi_f)), (float)h);
                else y_r = Ops::mul(Ops::sub(0.5f, Ops::div(phi_r, pi_f)), (float)h);
                if (pass == 0) { mid_y = y_r; mid_x = x_r; }
                else {
                    row_out[2 * t + 0] = Ops::sub(y_r, mid_y);
                    row_out[2 * t + 1] = Ops::sub(x_r, mid_x);
                }
            }
    }
    (void)k2;
    return SKY_OK;
}

__global__ void da_offsets_kernel(int h, int w, int k, int dilation, int skydome, float *out, int *status)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    float *row = out + (size_t)y * k * k * 2;
    int rc = offsets_row<DevOps>(y, h, w, k, dilation, skydome, row);
    if (rc == SKY_OK) {
        for (int t = 0; t < 2 * k * k; ++t)
            if (isnan(row[t])) rc = SKY_ERR_NAN_OFFSET;
    }
    // "undefined coordinates" is raised inside distortion() and therefore wins over a NaN table
    if (rc == SKY_ERR_UNDEFINED_COORDS) atomicExch(status, rc);
    else if (rc != SKY_OK) atomicCAS(status, SKY_OK, rc);
}

__global__ void da_sample_debug_kernel(int h, int w, int k, const float *__restrict__ offsets, int32_t *y0, int32_t *y1,
                                       int32_t *x0, int32_t *x1, float *w0, float *w1, float *w2, float *w3)
{
    const int k2 = k * k;
    const long total = (long)h * w * k2;
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, &ph0, &pht);
    pad_axis(w, k, &pw0, &pwt);
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int t = (int)(o % k2);
        const int j = (int)((o / k2) % w);
        const int i = (int)(o / ((long)k2 * w));
        const float yo = offsets[((size_t)i * k2 + t) * 2 + 0], xo = offsets[((size_t)i * k2 + t) * 2 + 1];
        const Sample s = da_sample(i, j, t / k, t % k, yo, xo, h + pht, w + pwt);
        y0[o] = s.y0; y1[o] = s.y1; x0[o] = s.x0; x1[o] = s.x1;
        w0[o] = s.w0; w1[o] = s.w1; w2[o] = s.w2; w3[o] = s.w3;
    }
}

static int check_geometry_args(int h, int w, int k, int dilation)
{
    SKY_REQUIRE(h > 0 && w > 0 && dilation > 0, SKY_ERR_INVALID, "h, w and dilation_rate must be positive (h=%d w=%d d=%d)", h, w, dilation);
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    // kernel_size 1 cannot be built by the reference either: tf.squeeze (:234) collapses the single tap and the tap
    // loop (:238-239) indexes a scalar.
    SKY_REQUIRE(k >= 3 && k <= 15, SKY_ERR_UNSUPPORTED, "kernel_size %d outside the supported odd range 3..15", k);
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

extern "C" int sky_version(void) { return 100; }
extern "C" const char *sky_last_error(void) { return sky::g_err; }
extern "C" long sky_launch_count(void) { return sky::g_launches.load(); }

extern "C" int sky_da_offsets_host(int h, int w, int k, int dilation, int skydome, float *out_host)
{
    int rc = check_geometry_args(h, w, k, dilation);
    if (rc != SKY_OK) return rc;
    SKY_REQUIRE(out_host != nullptr, SKY_ERR_INVALID, "out_host is NULL");
    for (int y = 0; y < h; ++y) {
        rc = offsets_row<HostOps>(y, h, w, k, dilation, skydome, out_host + (size_t)y * k * k * 2);
        SKY_REQUIRE(rc == SKY_OK, rc, "undefined coordinates");
    }
    for (size_t t = 0; t < (size_t)h * k * k * 2; ++t)
        SKY_REQUIRE(!isnan(out_host[t]), SKY_ERR_NAN_OFFSET,
                    "offset table contains NaN (kernel footprint leaves the sphere: asin of |y|>1 at h=%d w=%d k=%d dilation=%d)", h, w, k, dilation);
    return SKY_OK;
}

extern "C" int sky_da_offsets_device(int h, int w, int k, int dilation, int skydome, float *dev_out, int *dev_status, void *stream)
{
    int rc = check_geometry_args(h, w, k, dilation);
    if (rc != SKY_OK) return rc;
    SKY_REQUIRE(dev_out && dev_status, SKY_ERR_INVALID, "NULL device pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(dev_status, 0, sizeof(int), st));
    da_offsets_kernel<<<(h + 63) / 64, 64, 0, st>>>(h, w, k, dilation, skydome, dev_out, dev_status);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_da_sample_debug(int h, int w, int k, const float *dev_offsets, int32_t *y0, int32_t *y1, int32_t *x0,
                                   int32_t *x1, float *w0, float *w1, float *w2, float *w3, void *stream)
{
    int rc = check_geometry_args(h, w, k, 1);
    if (rc != SKY_OK) return rc;
    const long total = (long)h * w * k * k;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    da_sample_debug_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(h, w, k, dev_offsets, y0, y1, x0, x1, w0, w1, w2, w3);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
