// Row-strip convolution: shared declarations of the kernel (strip_conv.cu) and its host-side planner (strip_plan.cu).
//
// The sampling offsets of the distortion-aware layer depend on (output row, tap) only (distortion_aware_ops.py:266-268), and the y
// offset of a tap depends on its kernel ROW only (t_x has no y component, :228-234).  For one output row i the layer is therefore
//
//     y[i, j, :] = sum over kernel rows a, integer column shifts s of   V_{i,a}[j + s, :] . Weff_{i,a,s}
//
// with V_{i,a} = dy1 * x[y0(i,a)] + dy0 * x[y1(i,a)] (the vertical half of the bilinear blend, :103-106, one blended input row per kernel
// row) and Weff_{i,a,s} = sum over the taps b of that kernel row whose left / right corner sits at shift s of dx1 / dx0 times the tap's
// [C, F] slice of the layer variable (the horizontal half, folded into the B operand).  A plain SAME convolution (ops.py:4-42) is the
// special case dy = (1, 0), dx = (1, 0).  So the A operand of the tensor core is never gathered per (pixel, tap): a "strip" (one
// blended input row plus a halo, TF32-rounded once) is written to shared memory, and every (a, s) term is ONE tcgen05.mma whose
// A descriptor points at row (s - u0) of that strip.
#pragma once
#include "da_conv.cuh"

namespace sky {

// One strip = the rows of the A operand that several shifted windows share.
//   kind 0: position p of the strip holds  wy0 * x[r0, col(p)] + wy1 * x[r1, col(p)],  col(p) = cm * (j0 + u0 + p) + c0  run through the
//           layer's column map (the reference's 360-degree wrap in the padded frame for distortion-aware layers, zero outside the map
//           for plain convolutions); r = -1 is a zero row.
//   kind 1: an exact tap — row m of the strip is output pixel m of the tile sampled with the reference's own per-pixel arithmetic
//           (da_sample) for tap r0 with offsets (wy0, wy1) = (y_off, x_off).  Used where a tap does not reduce to one shift for the
//           whole row (decided on the host by comparing with da_sample at every column).
struct StripDesc {
    int kind;
    int r0, r1;
    float wy0, wy1;
    int u0, cm, c0;
    int win_begin, win_end;      // this strip's windows in the window array
};
// One window = one K=32 slice of the contraction per channel chunk: A rows [start_row, start_row + 128) of the strip, B = weight tile
// wtile0 + cc * wcc_stride.
struct WinDesc {
    int start_row;
    int wtile0;
};
// One output row class: output row `out_row`, output columns oc0 + ocs * (0 .. ncols-1), and the strips that make it up.
struct RowPlan {
    int out_row, oc0;
    int strip_begin, strip_end;
};
// One term of an effective weight tile: coef * kernel[tap * C + c, f]
struct WeffTerm {
    int tap;
    float coef;
};

// Weight gradient over a forward plan (strip_wgrad.cu).  One MMA group = one tcgen05.mma chain whose A operand starts at strip row
// `start_row`; its 128 accumulator lanes are either the (up to) 128 channels of window win[0], or the channels of the wpg windows
// win[0 .. wpg-1] at consecutive column shifts (C == 64: two windows, C == 32: four — the operand is laid out so that the MN atoms of
// neighbouring column shifts follow each other).
struct WgGroup {
    int start_row;
    int win[4];                    // window index per 32-lane quarter, -1: the quarter is not stored
};
// One accumulation unit: the groups of one strip that share TMEM (the strip and the dy tile are produced once per pixel tile for all of them).
struct WgUnit {
    int row, strip;
    int group_begin, group_end;
};

struct StripPlan {                 // host view of a cached plan; the arrays live on the device
    const RowPlan *rows = nullptr;
    const StripDesc *strips = nullptr;
    const WinDesc *wins = nullptr;
    const int *term_begin = nullptr;      // [nwins + 1] (effective-weight plans only)
    const WeffTerm *terms = nullptr;
    int nrows = 0, nstrips = 0, nwins = 0, nterms = 0;
    int ncols = 0, ocs = 1;               // output columns per row class, output column stride
    int TW = 0, NB = 0, SR = 0;           // tile width (columns), panoramas per tile, strip rows (multiple of 8)
    int max_strips_row = 0, max_wins_row = 0, exact_strips = 0;
    int weff = 0;                         // 1: weight tiles are per-window effective weights (tile = window * CC + cc)
    const WgUnit *wg_units = nullptr;     // weight-gradient plans only
    const WgGroup *wg_groups = nullptr;
    int n_wg_units = 0, n_wg_groups = 0, max_groups_unit = 0, span_max = 0;
};

// plans are built once per layer geometry and cached for the life of the process (host tables + device copies)
// (device = false: host tables only — plan statistics / export without a GPU)
// transposed: the data-gradient plan (strips over dy rows, merged transposed effective weights)
int get_plan_da(const float *offsets_host, int h, int w, int k, const StripPlan **out, bool device = true, bool transposed = false);
int get_plan_plain(int h, int w, int k, int stride, int transposed, int out_h, int out_w, int tp_ph0, int tp_pw0, const StripPlan **out,
                   bool device = true);

// weight-gradient flavour of the forward plans: tiles of 8 columns x 8 panoramas (a column shift = 8 strip rows = two K atoms of the
// MN-major operand), `wpg` windows per MMA group (1, or 4 for 32-channel layers), at most `gmax` groups per unit
int get_plan_da_wgrad(const float *offsets_host, int h, int w, int k, int wpg, int gmax, const StripPlan **out, bool device = true);
int get_plan_plain_wgrad(int h, int w, int k, int stride, int out_h, int out_w, int wpg, int gmax, const StripPlan **out, bool device = true);

// strip kernel behind sky_conv2d_fwd / the data gradients (plain SAME conv, C % 32 == 0); SKY_ERR_UNSUPPORTED (no error text) when it
// does not apply
int launch_fwd_strip_plain(const FwdArgs &a);

// weight gradient on the strip formulation (strip_wgrad.cu); offsets_host != NULL: distortion-aware layer, else plain SAME conv of
// `stride`.  dw [k*k*C, F] is ADDED to (the caller zeroes it).  SKY_ERR_UNSUPPORTED (no error text) when it does not apply.
int launch_wgrad_strip(const float *x, const float *dy, const float *offsets_host, float *dw, int B, int h, int w, int C, int F, int k,
                       int stride, cudaStream_t stream);

// schedule of the strip weight gradient (strip_wgrad.cu): parts per accumulation, tiles per part, accumulations per wave, waves
void wgrad_schedule(int nuidx, int ntiles, int sms, bool big, int ucap, int *P_out, int *TP_out, int *U_out, int *nwaves_out);
// db[f] += column sums of the dense matrix dy [M][F] (bias gradients; db zeroed by the caller)
int launch_col_sum(const float *dy, float *db, int M, int F, cudaStream_t st);

#ifdef __CUDACC__
// The reference's treatment of a column index (distortion_aware_ops.py:76-77 on the float coordinate, :90-91 on the integer corners),
// restated on the integer part: `q` is a column in the PADDED frame before any wrap.  Returns the unpadded column, or -1 for a zero.
__device__ __forceinline__ int da_map_col(int q, int in_w, int pw0, int W)
{
    if (q < 0) q += in_w;                 // :76 (x < 0 -> x + in_w; the result is <= in_w - 1, so :77 does not fire after it)
    else if (q > in_w - 1) q -= in_w;     // :77
    if (q < 0) q += in_w;                 // :90
    if (q > in_w - 1) q -= in_w;          // :91
    const int c = q - pw0;
    return (c >= 0 && c < W) ? c : -1;
}
#endif

}  // namespace sky
