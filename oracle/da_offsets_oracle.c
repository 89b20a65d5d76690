/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * CPU restatement, in plain C with glibc libm float functions, of the latitude-dependent sampling
 * geometry of the reference's distortion-aware convolution:
 *
 *   oracle_da_offsets   <- distortion_aware_ops.py:198-270  (conv2d.distortion; deconv2d.distortion 470-542 is a copy)
 *                          + make_grid 186-196
 *   oracle_da_sample    <- distortion_aware_ops.py:57-106    (conv2d.call: coordinates, clip, 360-degree wrap, corner
 *                          indices, bilinear weights) + _pad_input 125-150 + _get_conv_indices 152-168
 *
 * Parity status: the reference ships no golden vectors and TensorFlow is not installable here, so this file is
 * pinned against the reference's own *source* executed over a numpy/glibc TensorFlow stand-in
 * (tests/golden/make_golden.py + tests/golden/tf_shim.py -> tests/golden/*.npz), NOT against TensorFlow itself.
 * "parity unpinned" with respect to real TensorFlow numerics.
 *
 * fp32 semantics followed: every tf.* op is one separately rounded IEEE fp32 op
 * (no FMA contraction: build with -ffp-contract=off), python scalars are rounded to fp32 before they meet a tensor,
 * size-1 eager tensors take Eigen's scalar path => glibc tanf/cosf/sinf/atan2f/asinf.
 */
#include <math.h>
#include <stdint.h>

#define ORACLE_PI 3.141592653589793 /* np.math.pi, a python double (distortion_aware_ops.py:200) */

/* tf.linalg.cross(a, b): out0 = a1*b2 - a2*b1, out1 = a2*b0 - a0*b2, out2 = a0*b1 - a1*b0, each product and the
 * difference rounded separately. */
static void cross3(const float a[3], const float b[3], float out[3])
{
    float m0 = a[1] * b[2], m1 = a[2] * b[1];
    float m2 = a[2] * b[0], m3 = a[0] * b[2];
    float m4 = a[0] * b[1], m5 = a[1] * b[0];
    out[0] = m0 - m1;
    out[1] = m2 - m3;
    out[2] = m4 - m5;
}

/* Transcendental providers: glibc float (the oracle proper) or "evaluate in double, round once to float" (used only
 * to quantify how far a correctly-rounded evaluation sits from this glibc; see tests/test_oracle_offsets.py). */
typedef struct {
    float (*tan_)(float);
    float (*cos_)(float);
    float (*sin_)(float);
    float (*atan2_)(float, float);
    float (*asin_)(float);
} mathfn_t;

static float d_tan(float x) { return (float)tan((double)x); }
static float d_cos(float x) { return (float)cos((double)x); }
static float d_sin(float x) { return (float)sin((double)x); }
static float d_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
static float d_asin(float x) { return (float)asin((double)x); }

static const mathfn_t GLIBC_F32 = { tanf, cosf, sinf, atan2f, asinf };
static const mathfn_t VIA_F64 = { d_tan, d_cos, d_sin, d_atan2, d_asin };

/* out: [h][k*k][2] floats, last axis (y, x).  Returns 0, -1 for "undefined coordinates" (reference raises at :252),
 * -2 for an even kernel size (reference asserts at :188). */
static int offsets_impl(int h, int w, int k, int dilation, int skydome, float *out, const mathfn_t *mf)
{
    if (k % 2 != 1) return -2;
    const int n = k / 2;
    const int k2 = k * k;
    const int middle = n * (k + 1); /* :202 */
    const float pi_f = (float)ORACLE_PI;

    /* :204-205  tf.divide(python float, python int): both become fp32 tensors, then one fp32 division */
    const float unit_w = (float)(2.0 * ORACLE_PI) / (float)w;
    const float unit_h = pi_f / (float)(skydome ? h * 2 : h);
    /* :207 */
    const float rho = mf->tan_(unit_w) * (float)dilation;
    const float v[3] = { 0.f, 1.f, 0.f }; /* :209 */
    const int xc = (int)(w * 0.5);        /* :213 */

    for (int y = 0; y < h; ++y) {
        /* :220-221  python scalar (double) rounded to fp32, then fp32 multiply */
        const float theta = (float)((double)xc - 0.5 * (double)w) * unit_w;
        const float phi = skydome ? (float)(h - y) * unit_h : (float)((double)h * 0.5 - (double)y) * unit_h;

        /* :223-226 */
        float p_u[3];
        p_u[0] = mf->cos_(phi) * mf->cos_(theta);
        p_u[1] = mf->sin_(phi);
        p_u[2] = mf->cos_(phi) * mf->sin_(theta);

        float t_x[3], t_y[3];
        cross3(v, p_u, t_x);   /* :228 */
        cross3(p_u, t_x, t_y); /* :229 */

        float kk[49 * 49][2]; /* generous; k <= 49 */
        int t = 0;
        /* make_grid :192-194: y from r..-r outer, x from r..-r inner, entries [x, y] */
        for (int gy = n; gy >= -n; --gy) {
            for (int gx = n; gx >= -n; --gx, ++t) {
                float ur[3];
                for (int c = 0; c < 3; ++c) {
                    /* :233  rho * (r0 * t_x + r1 * t_y), python ints rounded to fp32 */
                    float a = (float)gx * t_x[c];
                    float b = (float)gy * t_y[c];
                    float s = a + b;
                    float r = rho * s;
                    ur[c] = p_u[c] + r; /* :235 */
                }
                float theta_r;
                if (ur[0] > 0.f) {                       /* :239-240 */
                    theta_r = mf->atan2_(ur[2], ur[0]);
                } else if (ur[0] < 0.f) {                /* :241-245 */
                    if (ur[2] >= 0.f) theta_r = mf->atan2_(ur[2], ur[0]) + pi_f;
                    else              theta_r = mf->atan2_(ur[2], ur[0]) - pi_f;
                } else {                                 /* :246-252 */
                    if (ur[2] > 0.f)      theta_r = (float)(ORACLE_PI * 0.5);
                    else if (ur[2] < 0.f) theta_r = (float)(-ORACLE_PI * 0.5);
                    else return -1;
                }
                float phi_r = mf->asin_(ur[1]);          /* :254 */
                /* :256  ((theta_r / pi + 1) * 0.5) * w */
                float q = theta_r / pi_f;
                q = q + 1.f;
                q = q * 0.5f;
                float x_r = q * (float)w;
                /* :257 */
                float y_r;
                if (skydome) {
                    float u = 2.f * phi_r;
                    u = u / pi_f;
                    u = 1.f - u;
                    y_r = u * (float)h;
                } else {
                    float u = phi_r / pi_f;
                    u = 0.5f - u;
                    y_r = u * (float)h;
                }
                kk[t][0] = y_r;
                kk[t][1] = x_r;
            }
        }
        for (t = 0; t < k2; ++t) { /* :261 */
            out[((long)y * k2 + t) * 2 + 0] = kk[t][0] - kk[middle][0];
            out[((long)y * k2 + t) * 2 + 1] = kk[t][1] - kk[middle][1];
        }
    }
    return 0;
}

int oracle_da_offsets(int h, int w, int k, int dilation, int skydome, float *out)
{
    return offsets_impl(h, w, k, dilation, skydome, out, &GLIBC_F32);
}

int oracle_da_offsets_via_f64(int h, int w, int k, int dilation, int skydome, float *out)
{
    return offsets_impl(h, w, k, dilation, skydome, out, &VIA_F64);
}

/* Padding decided by _pad_input (:125-150) for one axis: total pad k-1 (before = (k-1)/2) unless SAME and VALID
 * output sizes agree. */
static void pad_axis(int n, int k, int stride, int *before, int *total)
{
    int same_out = (n + stride - 1) / stride;
    int valid_out = (n - k + stride) / stride;
    if (same_out == valid_out) { *before = 0; *total = 0; }
    else { *total = k - 1; *before = (k - 1) / 2; }
}

/*
 * Sampling coordinates of conv2d.call (:63-106) for stride 1 on an h x w (unpadded) map.
 * offsets: [h][k2][2] from oracle_da_offsets.  Outputs are [h][w][k2] each:
 *   y0,y1,x0,x1  int32 corner coordinates in the PADDED frame (what tf.gather_nd receives, :94)
 *   w0..w3       fp32 bilinear weights (:103-106)
 * Returns 0, or -3 if a corner index falls outside the padded map (TF-CPU gather_nd would raise).
 */
int oracle_da_sample(int h, int w, int k, const float *offsets,
                     int32_t *y0o, int32_t *y1o, int32_t *x0o, int32_t *x1o,
                     float *w0o, float *w1o, float *w2o, float *w3o)
{
    const int k2 = k * k;
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, 1, &ph0, &pht);
    pad_axis(w, k, 1, &pw0, &pwt);
    const int in_h = h + pht, in_w = w + pwt;
    const float in_h_m1 = (float)(in_h - 1), in_w_f = (float)in_w, in_w_m1 = (float)(in_w - 1);
    int rc = 0;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j)
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) {
                    const int t = a * k + b;
                    const long o = ((long)i * w + j) * k2 + t;
                    /* :66-72  extract_patches gives y=i+a, x=j+b (padded frame); cast; add offset */
                    float y = (float)(i + a) + offsets[((long)i * k2 + t) * 2 + 0];
                    float x = (float)(j + b) + offsets[((long)i * k2 + t) * 2 + 1];
                    /* :73 clip_by_value(y, 0, in_h-1) == min(max(y, lo), hi) */
                    y = fminf(fmaxf(y, 0.f), in_h_m1);
                    /* :76-77 single wrap in the padded width */
                    if (x < 0.f) x = x + in_w_f;
                    if (x > in_w_m1) x = x - in_w_f;
                    /* :82-83 */
                    int32_t y0 = (int32_t)floorf(y), x0 = (int32_t)floorf(x);
                    int32_t y1 = y0 + 1, x1 = x0 + 1;
                    /* :86 */
                    y0 = y0 < 0 ? 0 : (y0 > in_h - 1 ? in_h - 1 : y0);
                    y1 = y1 < 0 ? 0 : (y1 > in_h - 1 ? in_h - 1 : y1);
                    /* :89-91 */
                    const int32_t x0_w = x0, x1_w = x1;
                    if (x0 < 0) x0 += in_w;
                    if (x1 < 0) x1 += in_w;
                    if (x0 > in_w - 1) x0 -= in_w;
                    if (x1 > in_w - 1) x1 -= in_w;
                    if (x0 < 0 || x0 >= in_w || x1 < 0 || x1 >= in_w) rc = -3;
                    /* :100-106 */
                    const float fy0 = (float)y0, fy1 = (float)y1, fx0 = (float)x0_w, fx1 = (float)x1_w;
                    const float dy1 = fy1 - y, dy0 = y - fy0, dx1 = fx1 - x, dx0 = x - fx0;
                    y0o[o] = y0; y1o[o] = y1; x0o[o] = x0; x1o[o] = x1;
                    w0o[o] = dy1 * dx1;
                    w1o[o] = dy1 * dx0;
                    w2o[o] = dy0 * dx1;
                    w3o[o] = dy0 * dx0;
                }
    return rc;
}
