// Sampling geometry of the distortion-aware convolution, restated once and shared by every kernel in the library
// (debug export, SIMT conv, tensor-core conv, backward kernels), so that the bit-exactness demonstrated through
// sky_da_sample_debug carries over to them.
//
// Follows distortion_aware_ops.py:63-106 (conv2d.call).  Every arithmetic step is an explicitly rounded fp32 intrinsic
// (__fadd_rn / __fsub_rn / __fmul_rn) so that nvcc can neither contract to FMA nor reassociate.
#pragma once
#include "sky_common.cuh"

namespace sky {

struct Sample {
    int y0, y1, x0, x1;    // corner coordinates in the reference's PADDED frame (:82-91)
    float w0, w1, w2, w3;  // bilinear weights (:103-106); corners (y0,x0) (y0,x1) (y1,x0) (y1,x1)
    float dy1, dy0, dx1, dx0;  // the factors: w0 = dy1*dx1, w1 = dy1*dx0, w2 = dy0*dx1, w3 = dy0*dx0
    bool wrapped;          // any 360-degree wrap applied to x (:76-77, :90-91)
};

// (i, j): output pixel; (a, b): tap row / column (tap = a*k + b, :152-168); in_h/in_w: padded map size.
__device__ __forceinline__ Sample da_sample(int i, int j, int a, int b, float y_off, float x_off, int in_h, int in_w)
{
    const float in_h_m1 = (float)(in_h - 1), in_w_f = (float)in_w, in_w_m1 = (float)(in_w - 1);
    float y = __fadd_rn((float)(i + a), y_off);  // :66-71
    float x = __fadd_rn((float)(j + b), x_off);  // :72
    y = fminf(fmaxf(y, 0.f), in_h_m1);           // :73
    if (x < 0.f) x = __fadd_rn(x, in_w_f);       // :76
    if (x > in_w_m1) x = __fsub_rn(x, in_w_f);   // :77
    int y0 = (int)floorf(y), x0 = (int)floorf(x);  // :82
    int y1 = y0 + 1, x1 = x0 + 1;                  // :83
    y0 = min(max(y0, 0), in_h - 1);                // :86
    y1 = min(max(y1, 0), in_h - 1);
    const int x0_w = x0, x1_w = x1;                // :89
    if (x0 < 0) x0 += in_w;                        // :90
    if (x1 < 0) x1 += in_w;
    if (x0 > in_w - 1) x0 -= in_w;                 // :91
    if (x1 > in_w - 1) x1 -= in_w;
    const bool wrapped = (x0 != x0_w) || (x1 != x1_w);
    const float dy1 = __fsub_rn((float)y1, y), dy0 = __fsub_rn(y, (float)y0);      // :103-106
    const float dx1 = __fsub_rn((float)x1_w, x), dx0 = __fsub_rn(x, (float)x0_w);
    Sample s;
    s.y0 = y0; s.y1 = y1; s.x0 = x0; s.x1 = x1;
    s.w0 = __fmul_rn(dy1, dx1);
    s.w1 = __fmul_rn(dy1, dx0);
    s.w2 = __fmul_rn(dy0, dx1);
    s.w3 = __fmul_rn(dy0, dx0);
    s.dy1 = dy1; s.dy0 = dy0; s.dx1 = dx1; s.dx0 = dx0;
    s.wrapped = wrapped;
    return s;
}

// Element offsets (into the UNPADDED NHWC tensor, channel 0) of the four corners of one sample, -1 where the corner
// lies in the zero halo of _pad_input (:125-150) or the index left the padded map (TF-GPU gather_nd yields 0 there).
struct CornerRef {
    int off[4];
    float w[4];
};

__device__ __forceinline__ CornerRef da_corners(const Sample &s, int b_img, int h, int w, int C, int ph0, int pw0)
{
    CornerRef r;
    const int ys[4] = { s.y0, s.y0, s.y1, s.y1 };
    const int xs[4] = { s.x0, s.x1, s.x0, s.x1 };
    r.w[0] = s.w0; r.w[1] = s.w1; r.w[2] = s.w2; r.w[3] = s.w3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int yy = ys[c] - ph0, xx = xs[c] - pw0;
        const bool ok = (yy >= 0) && (yy < h) && (xx >= 0) && (xx < w);
        r.off[c] = ok ? ((b_img * h + yy) * w + xx) * C : -1;
    }
    return r;
}

}  // namespace sky
