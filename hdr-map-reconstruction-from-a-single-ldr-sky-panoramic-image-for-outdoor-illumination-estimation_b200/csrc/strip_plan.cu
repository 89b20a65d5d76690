// Host-side planner of the row-strip convolution (strip_conv.cuh): turns a layer's sampling geometry into row classes, strips and
// windows, once per layer geometry, and keeps the device copies for the life of the process.
//
// Distortion-aware layers: every (output row, tap) is evaluated with the reference's own fp32 arithmetic (the host twin of da_sample,
// distortion_aware_ops.py:63-106) at EVERY output column.  A tap is folded into the strip formulation only if, at every column, the
// corner indices that arithmetic produces are exactly the ones the strip's integer column map yields for one shift, and the bilinear
// factors agree with the row's factors to fp32 rounding of the coordinate (the fraction of j + b + x_off depends on j only through the
// rounding of that sum).  Anything else — there is nothing else on the shapes of the path, but the check is what guarantees it —
// becomes an "exact tap" strip that the kernel samples per pixel with da_sample itself.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "strip_conv.cuh"

namespace sky {

namespace {

constexpr int SR_MAX = 192;     // strip rows the shared-memory budget of the kernel is sized for

struct HostSample {
    int y0, y1, x0, x1;
    float dy1, dy0, dx1, dx0;
};

// host twin of da_sample (da_geometry.cuh); compiled with -ffp-contract=off, every operation is one rounded fp32 operation
HostSample host_sample(int i, int j, int a, int b, float y_off, float x_off, int in_h, int in_w)
{
    const float in_h_m1 = (float)(in_h - 1), in_w_f = (float)in_w, in_w_m1 = (float)(in_w - 1);
    float y = (float)(i + a) + y_off;
    float x = (float)(j + b) + x_off;
    y = fminf(fmaxf(y, 0.f), in_h_m1);
    if (x < 0.f) x = x + in_w_f;
    if (x > in_w_m1) x = x - in_w_f;
    int y0 = (int)floorf(y), x0 = (int)floorf(x);
    int y1 = y0 + 1, x1 = x0 + 1;
    y0 = std::min(std::max(y0, 0), in_h - 1);
    y1 = std::min(std::max(y1, 0), in_h - 1);
    const int x0_w = x0, x1_w = x1;
    if (x0 < 0) x0 += in_w;
    if (x1 < 0) x1 += in_w;
    if (x0 > in_w - 1) x0 -= in_w;
    if (x1 > in_w - 1) x1 -= in_w;
    HostSample s;
    s.y0 = y0; s.y1 = y1; s.x0 = x0; s.x1 = x1;
    s.dy1 = (float)y1 - y; s.dy0 = y - (float)y0;
    s.dx1 = (float)x1_w - x; s.dx0 = x - (float)x0_w;
    return s;
}

// the kernel's integer column map (da_map_col in strip_conv.cu), same statements
int host_map_col(int q, int in_w, int pw0, int W)
{
    if (q < 0) q += in_w;
    else if (q > in_w - 1) q -= in_w;
    if (q < 0) q += in_w;
    if (q > in_w - 1) q -= in_w;
    const int c = q - pw0;
    return (c >= 0 && c < W) ? c : -1;
}

struct HostPlan {
    std::vector<RowPlan> rows;
    std::vector<StripDesc> strips;
    std::vector<WinDesc> wins;
    std::vector<int> term_begin;
    std::vector<WeffTerm> terms;
    std::vector<WgUnit> wg_units;
    std::vector<WgGroup> wg_groups;
    StripPlan dev;
    bool uploaded = false;
};

int pow2_floor(int v)
{
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// shifts (sorted, distinct) of one blended row -> strips whose span fits the strip buffer; appends strips + windows
struct ShiftWin {
    int shift;
    int wtile;                       // plain plans: the tap; effective-weight plans: filled in by the caller (window index)
    std::vector<WeffTerm> terms;
};

void emit_strips(HostPlan &hp, std::vector<ShiftWin> &sw, StripDesc proto, int NB, int cap, int *span_max, bool weff)
{
    std::sort(sw.begin(), sw.end(), [](const ShiftWin &l, const ShiftWin &r) { return l.shift < r.shift; });
    size_t q = 0;
    while (q < sw.size()) {
        const int start = sw[q].shift;
        StripDesc sd = proto;
        sd.u0 = start;
        sd.win_begin = (int)hp.wins.size();
        while (q < sw.size() && sw[q].shift - start <= cap) {
            WinDesc wd;
            wd.start_row = (sw[q].shift - start) * NB;
            wd.wtile0 = weff ? (int)hp.wins.size() : sw[q].wtile;
            if (weff) {
                hp.term_begin.push_back((int)hp.terms.size());
                hp.terms.insert(hp.terms.end(), sw[q].terms.begin(), sw[q].terms.end());
            }
            hp.wins.push_back(wd);
            *span_max = std::max(*span_max, sw[q].shift - start);
            ++q;
        }
        sd.win_end = (int)hp.wins.size();
        hp.strips.push_back(sd);
    }
}

template <class T>
int upload(const std::vector<T> &v, const T **out)
{
    *out = nullptr;
    if (v.empty()) return SKY_OK;
    T *d = nullptr;
    SKY_CHECK_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
    SKY_CHECK_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = d;
    return SKY_OK;
}

int finish_plan(HostPlan &hp, int TW, int NB, int span_max, int ncols, int ocs, int weff)
{
    StripPlan &d = hp.dev;
    d.nrows = (int)hp.rows.size(); d.nstrips = (int)hp.strips.size(); d.nwins = (int)hp.wins.size(); d.nterms = (int)hp.terms.size();
    d.ncols = ncols; d.ocs = ocs; d.TW = TW; d.NB = NB; d.weff = weff;
    d.span_max = span_max;
    d.SR = round_up((TW + span_max) * NB, 8);
    if (d.SR < BLOCK_M) d.SR = BLOCK_M;
    for (const RowPlan &r : hp.rows) {
        d.max_strips_row = std::max(d.max_strips_row, r.strip_end - r.strip_begin);
        int nw = 0;
        for (int s = r.strip_begin; s < r.strip_end; ++s) nw += hp.strips[s].win_end - hp.strips[s].win_begin;
        d.max_wins_row = std::max(d.max_wins_row, nw);
    }
    for (const StripDesc &s : hp.strips) d.exact_strips += s.kind == 1;
    if (weff) hp.term_begin.push_back((int)hp.terms.size());
    return SKY_OK;
}

// device copies are made on first use by a launch (cudaMalloc + synchronous copy: outside stream capture)
int ensure_device(HostPlan &hp)
{
    if (hp.uploaded) return SKY_OK;
    StripPlan &d = hp.dev;
    int rc;
    if ((rc = upload(hp.rows, &d.rows)) != SKY_OK) return rc;
    if ((rc = upload(hp.strips, &d.strips)) != SKY_OK) return rc;
    if ((rc = upload(hp.wins, &d.wins)) != SKY_OK) return rc;
    if (d.weff) {
        if ((rc = upload(hp.term_begin, &d.term_begin)) != SKY_OK) return rc;
        if ((rc = upload(hp.terms, &d.terms)) != SKY_OK) return rc;
    }
    if (!hp.wg_units.empty()) {
        if ((rc = upload(hp.wg_units, &d.wg_units)) != SKY_OK) return rc;
        if ((rc = upload(hp.wg_groups, &d.wg_groups)) != SKY_OK) return rc;
    }
    hp.uploaded = true;
    return SKY_OK;
}

// Weight-gradient units over a finished forward plan: the windows of a strip become MMA groups (one window each, or up to four windows
// at consecutive shifts for 32-channel layers), at most gmax groups (accumulators in TMEM) per unit.
void build_wgrad_units(HostPlan &hp, int NB, int wpg, int gmax)
{
    for (int r = 0; r < (int)hp.rows.size(); ++r)
        for (int si = hp.rows[r].strip_begin; si < hp.rows[r].strip_end; ++si) {
            const StripDesc &sd = hp.strips[si];
            std::vector<WgGroup> groups;
            for (int wi = sd.win_begin; wi < sd.win_end;) {
                WgGroup g;
                g.start_row = hp.wins[wi].start_row;
                g.win[0] = wi; g.win[1] = g.win[2] = g.win[3] = -1;
                int nx = wi + 1;
                if (wpg > 1)
                    for (; nx < sd.win_end; ++nx) {                 // windows are sorted by shift, shifts are distinct
                        const int q = (hp.wins[nx].start_row - g.start_row) / NB;
                        if (q > wpg - 1) break;
                        g.win[q] = nx;
                    }
                groups.push_back(g);
                wi = nx;
            }
            for (size_t g0 = 0; g0 < groups.size(); g0 += gmax) {
                WgUnit u;
                u.row = r; u.strip = si;
                u.group_begin = (int)hp.wg_groups.size();
                for (size_t g = g0; g < groups.size() && g < g0 + gmax; ++g) hp.wg_groups.push_back(groups[g]);
                u.group_end = (int)hp.wg_groups.size();
                hp.wg_units.push_back(u);
                hp.dev.max_groups_unit = std::max(hp.dev.max_groups_unit, u.group_end - u.group_begin);
            }
        }
    hp.dev.n_wg_units = (int)hp.wg_units.size();
    hp.dev.n_wg_groups = (int)hp.wg_groups.size();
}

// One kernel row of one output row: the vertical geometry (shared by its k taps) and the horizontal terms by column shift.
struct KernelRow {
    bool have_ref = false;
    HostSample ref{};
    std::map<int, std::vector<WeffTerm>> by_shift;      // unpadded, unwrapped column shift -> (tap, horizontal factor)
    std::vector<int> exact_taps;                         // taps that do not reduce to one shift for the whole row
};

// `relaxed`: a tap whose strict check fails is still folded when, at every column, the two-corner stencil of the reference arithmetic
// and the row's stencil agree as (column -> weight) maps to 2e-5 — the fraction of b + x_off sits within an ulp of an integer, so the
// reference's floor depends on j while the weights of the PHYSICAL columns do not (used by the data-gradient plan, whose TF32 operands
// are three orders of magnitude coarser).
int analyse_kernel_row(const float *off, int h, int w, int k, int i, int a, bool relaxed, KernelRow &kr)
{
    const int k2 = k * k;
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, &ph0, &pht);
    pad_axis(w, k, &pw0, &pwt);
    const int in_h = h + pht, in_w = w + pwt;
    const float tol = 8.f * (float)in_w * 1.1920929e-7f;      // a few ulp of the coordinate j + b + x_off
    for (int b = 0; b < k; ++b) {
        const int t = a * k + b;
        const float yo = off[((size_t)i * k2 + t) * 2 + 0], xo = off[((size_t)i * k2 + t) * 2 + 1];
        if (!(yo == yo) || !(xo == xo)) return SKY_ERR_UNSUPPORTED;        // NaN table: the caller's other path reports it
        const HostSample s0 = host_sample(i, 0, a, b, yo, xo, in_h, in_w);
        bool regular = true;
        if (!kr.have_ref) { kr.ref = s0; kr.have_ref = true; }
        else if (s0.y0 != kr.ref.y0 || s0.y1 != kr.ref.y1 || s0.dy1 != kr.ref.dy1 || s0.dy0 != kr.ref.dy0) regular = false;
        const double xs = (double)b + (double)xo;                            // x - j in the padded frame
        if (!(fabs(xs) < 1e6)) regular = false;
        int sq = 0;
        float dx1r = 1.f, dx0r = 0.f;
        if (regular) {
            const double fl = floor(xs);
            sq = (int)fl;
            dx0r = (float)(xs - fl);
            dx1r = (float)(1.0 - (xs - fl));
            bool strict = true, loose = true;
            for (int j = 0; j < w && (strict || (relaxed && loose)); ++j) {
                const HostSample s = host_sample(i, j, a, b, yo, xo, in_h, in_w);
                const int e0 = host_map_col(j + sq, in_w, pw0, w), e1 = host_map_col(j + sq + 1, in_w, pw0, w);
                const int g0 = (s.x0 - pw0 >= 0 && s.x0 - pw0 < w) ? s.x0 - pw0 : -1;
                const int g1 = (s.x1 - pw0 >= 0 && s.x1 - pw0 < w) ? s.x1 - pw0 : -1;
                if (fabsf(s.dx1 - dx1r) > tol || fabsf(s.dx0 - dx0r) > tol) strict = false;
                if (g0 != e0 && !(dx1r == 0.f && s.dx1 == 0.f)) strict = false;
                if (g1 != e1 && !(dx0r == 0.f && s.dx0 == 0.f)) strict = false;
                if (s.y0 != kr.ref.y0 || s.y1 != kr.ref.y1) { strict = false; loose = false; }
                // stencils as column -> weight maps (column -1 = zero halo: its weight is irrelevant)
                const int cols[4] = { g0, g1, e0, e1 };
                const float wts[4] = { s.dx1, s.dx0, -dx1r, -dx0r };
                for (int u = 0; u < 4 && loose; ++u) {
                    if (cols[u] < 0) continue;
                    float sum = 0.f;
                    for (int v = 0; v < 4; ++v)
                        if (cols[v] == cols[u]) sum += wts[v];
                    if (fabsf(sum) > 2e-5f) loose = false;
                }
            }
            regular = strict || (relaxed && loose);
        }
        if (!regular) { kr.exact_taps.push_back(t); continue; }
        const int su = sq - pw0;                                              // shift in unpadded, unwrapped columns
        if (dx1r != 0.f) kr.by_shift[su].push_back(WeffTerm{ t, dx1r });
        if (dx0r != 0.f) kr.by_shift[su + 1].push_back(WeffTerm{ t, dx0r });
    }
    return SKY_OK;
}

// force_tw > 0: tile geometry and strip span of the weight-gradient plans instead of the forward kernel's
int build_da(const float *off, int h, int w, int k, HostPlan &hp, int force_tw = 0, int force_nb = 0, int force_cap = 0)
{
    const int k2 = k * k;
    int ph0, pht;
    pad_axis(h, k, &ph0, &pht);
    const int TW = force_tw > 0 ? force_tw : pow2_floor(std::min(w, BLOCK_M)), NB = force_tw > 0 ? force_nb : BLOCK_M / TW;
    const int cap = force_tw > 0 ? force_cap : (SR_MAX - BLOCK_M) / NB;
    int span_max = 0;
    for (int i = 0; i < h; ++i) {
        RowPlan rp;
        rp.out_row = i; rp.oc0 = 0; rp.strip_begin = (int)hp.strips.size();
        std::vector<int> exact_taps;
        for (int a = 0; a < k; ++a) {
            KernelRow kr;
            const int rc = analyse_kernel_row(off, h, w, k, i, a, false, kr);
            if (rc != SKY_OK) return rc;
            exact_taps.insert(exact_taps.end(), kr.exact_taps.begin(), kr.exact_taps.end());
            if (!kr.have_ref || kr.by_shift.empty()) continue;
            const HostSample &ref = kr.ref;
            StripDesc proto{};
            proto.kind = 0;
            proto.r0 = (ref.y0 - ph0 >= 0 && ref.y0 - ph0 < h) ? ref.y0 - ph0 : -1;
            proto.r1 = (ref.y1 - ph0 >= 0 && ref.y1 - ph0 < h) ? ref.y1 - ph0 : -1;
            proto.wy0 = ref.dy1; proto.wy1 = ref.dy0;
            proto.cm = 1; proto.c0 = 0;
            const bool row0_dead = proto.r0 < 0 || proto.wy0 == 0.f, row1_dead = proto.r1 < 0 || proto.wy1 == 0.f;
            if (row0_dead && row1_dead) continue;                                     // the blended row is identically zero
            std::vector<ShiftWin> sw;
            for (auto &kv : kr.by_shift) sw.push_back(ShiftWin{ kv.first, 0, kv.second });
            emit_strips(hp, sw, proto, NB, cap, &span_max, true);
        }
        for (int t : exact_taps) {
            StripDesc sd{};
            sd.kind = 1; sd.r0 = t; sd.r1 = -1;
            sd.wy0 = off[((size_t)i * k2 + t) * 2 + 0]; sd.wy1 = off[((size_t)i * k2 + t) * 2 + 1];
            sd.u0 = 0; sd.cm = 1; sd.c0 = 0;
            sd.win_begin = (int)hp.wins.size();
            hp.term_begin.push_back((int)hp.terms.size());
            hp.terms.push_back(WeffTerm{ t, 1.f });
            hp.wins.push_back(WinDesc{ 0, (int)hp.wins.size() });
            sd.win_end = (int)hp.wins.size();
            hp.strips.push_back(sd);
        }
        rp.strip_end = (int)hp.strips.size();
        hp.rows.push_back(rp);
    }
    return finish_plan(hp, TW, NB, span_max, w, 1, 1);
}

// Data gradient of the distortion-aware layer as a row-strip convolution over dy (the transpose of the forward plan, in gather form):
// input row r of dx collects, from every forward (output row i, kernel row a) whose blended strip reads r with vertical weight wy, the
// terms  dy[i, c - s, :] . (wy * Weff_{i,a,s})^T.  Strips = plain dy rows, windows = (i, -s) with all coinciding terms merged into one
// effective weight, columns through the transposed column map (the unique dy column j with forward_map(j + s) = c).  No atomics.
int build_da_transposed(const float *off, int h, int w, int k, HostPlan &hp)
{
    int ph0, pht;
    pad_axis(h, k, &ph0, &pht);
    const int TW = pow2_floor(std::min(w, BLOCK_M)), NB = BLOCK_M / TW;
    const int cap = (SR_MAX - BLOCK_M) / NB;
    // per input row r: dy row i -> (transposed shift -> terms)
    std::vector<std::map<int, std::map<int, std::vector<WeffTerm>>>> acc(h);
    for (int i = 0; i < h; ++i)
        for (int a = 0; a < k; ++a) {
            KernelRow kr;
            const int rc = analyse_kernel_row(off, h, w, k, i, a, true, kr);
            if (rc != SKY_OK) return rc;
            if (!kr.exact_taps.empty()) return SKY_ERR_UNSUPPORTED;                   // the scatter kernel keeps such layers
            if (!kr.have_ref) continue;
            const int rr[2] = { kr.ref.y0 - ph0, kr.ref.y1 - ph0 };
            const float wy[2] = { kr.ref.dy1, kr.ref.dy0 };
            for (int c = 0; c < 2; ++c) {
                if (rr[c] < 0 || rr[c] >= h || wy[c] == 0.f) continue;
                for (auto &kv : kr.by_shift)
                    for (const WeffTerm &t : kv.second) acc[rr[c]][i][-kv.first].push_back(WeffTerm{ t.tap, t.coef * wy[c] });
            }
        }
    int span_max = 0;
    for (int r = 0; r < h; ++r) {
        RowPlan rp;
        rp.out_row = r; rp.oc0 = 0; rp.strip_begin = (int)hp.strips.size();
        for (auto &by_i : acc[r]) {
            StripDesc proto{};
            proto.kind = 0; proto.r0 = by_i.first; proto.r1 = -1; proto.wy0 = 1.f; proto.wy1 = 0.f; proto.cm = 1; proto.c0 = 0;
            std::vector<ShiftWin> sw;
            for (auto &kv : by_i.second) sw.push_back(ShiftWin{ kv.first, 0, kv.second });
            emit_strips(hp, sw, proto, NB, cap, &span_max, true);
        }
        rp.strip_end = (int)hp.strips.size();
        hp.rows.push_back(rp);
    }
    return finish_plan(hp, TW, NB, span_max, w, 1, 1);
}

// plain SAME convolution (tf.nn.conv2d, ops.py:41) of stride 1 / 2, or — transposed != 0 — the data gradient of one run as a forward
// pass over dy (same tap convention as the direct kernel: output pixel (i, j) reads tap (a, b) at ((i + a - ph0) / s, (j + b - pw0) / s)
// when both divisions are exact)
int build_plain(int h, int w, int k, int stride, int transposed, int OH, int OW, int tp_ph0, int tp_pw0, HostPlan &hp, int force_tw = 0,
                int force_nb = 0, int force_cap = 0)
{
    int ph0 = tp_ph0, pw0 = tp_pw0;
    if (!transposed) {
        const int th = (OH - 1) * stride + k - h, tw = (OW - 1) * stride + k - w;
        ph0 = (th > 0 ? th : 0) / 2; pw0 = (tw > 0 ? tw : 0) / 2;
    }
    const int col_classes = (transposed && stride == 2) ? 2 : 1;
    if (col_classes == 2 && (OW & 1)) return SKY_ERR_UNSUPPORTED;
    const int ncols = OW / col_classes;
    const int TW = force_tw > 0 ? force_tw : pow2_floor(std::min(ncols, BLOCK_M)), NB = force_tw > 0 ? force_nb : BLOCK_M / TW;
    const int cap = force_tw > 0 ? force_cap : (SR_MAX - BLOCK_M) / NB;
    int span_max = 0;
    for (int io = 0; io < OH; ++io)
        for (int pj = 0; pj < col_classes; ++pj) {
            RowPlan rp;
            rp.out_row = io; rp.oc0 = pj; rp.strip_begin = (int)hp.strips.size();
            for (int a = 0; a < k; ++a) {
                int r;
                if (!transposed) r = io * stride + a - ph0;
                else {
                    const int yy = io + a - ph0;
                    if (yy < 0 || (yy % stride) != 0) continue;
                    r = yy / stride;
                }
                if (r < 0 || r >= h) continue;
                const int nclass = (!transposed && stride == 2) ? 2 : 1;             // input column parity classes of a strided forward conv
                for (int par = 0; par < nclass; ++par) {
                    std::vector<ShiftWin> sw;
                    for (int b = 0; b < k; ++b) {
                        int shift;
                        if (!transposed) {
                            const int s = b - pw0;
                            if (stride == 1) shift = s;
                            else {
                                if (((s % 2) + 2) % 2 != par) continue;
                                shift = (s - par) / 2;
                            }
                        } else {
                            const int s = pj + b - pw0;
                            if (stride == 1) shift = s;
                            else {
                                if (s % 2 != 0) continue;
                                shift = s / 2;
                            }
                        }
                        sw.push_back(ShiftWin{ shift, a * k + b, {} });
                    }
                    if (sw.empty()) continue;
                    StripDesc proto{};
                    proto.kind = 0; proto.r0 = r; proto.r1 = -1; proto.wy0 = 1.f; proto.wy1 = 0.f;
                    proto.cm = (!transposed && stride == 2) ? 2 : 1; proto.c0 = par;
                    emit_strips(hp, sw, proto, NB, cap, &span_max, false);
                }
            }
            rp.strip_end = (int)hp.strips.size();
            hp.rows.push_back(rp);
        }
    return finish_plan(hp, TW, NB, span_max, ncols, col_classes, 0);
}

std::mutex g_mu;
std::map<std::tuple<int, int, int, int, unsigned long long>, HostPlan *> g_da;
std::map<std::tuple<int, int, int, int, int, int, int, int, int>, HostPlan *> g_plain;

unsigned long long fnv1a(const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    unsigned long long hsh = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { hsh ^= b[i]; hsh *= 1099511628211ull; }
    return hsh;
}

}  // namespace

static int export_plan(const HostPlan &hp, void *rows, void *strips, void *wins, int *term_begin, void *terms)
{
    if (rows) memcpy(rows, hp.rows.data(), hp.rows.size() * sizeof(RowPlan));
    if (strips) memcpy(strips, hp.strips.data(), hp.strips.size() * sizeof(StripDesc));
    if (wins) memcpy(wins, hp.wins.data(), hp.wins.size() * sizeof(WinDesc));
    if (term_begin) memcpy(term_begin, hp.term_begin.data(), hp.term_begin.size() * sizeof(int));
    if (terms) memcpy(terms, hp.terms.data(), hp.terms.size() * sizeof(WeffTerm));
    return SKY_OK;
}

int get_plan_da(const float *offsets_host, int h, int w, int k, const StripPlan **out, bool device, bool transposed)
{
    if (h <= 0 || w <= 0 || k < 3 || !(k & 1)) return SKY_ERR_UNSUPPORTED;
    const auto key = std::make_tuple(h, w, k, transposed ? 1 : 0, fnv1a(offsets_host, (size_t)h * k * k * 2 * sizeof(float)));
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_da.find(key);
    if (it == g_da.end()) {
        HostPlan *hp = new HostPlan();
        const int rc = transposed ? build_da_transposed(offsets_host, h, w, k, *hp) : build_da(offsets_host, h, w, k, *hp);
        if (rc != SKY_OK) { delete hp; return rc; }
        it = g_da.emplace(key, hp).first;
    }
    if (device) {
        const int rc = ensure_device(*it->second);
        if (rc != SKY_OK) return rc;
    }
    *out = &it->second->dev;
    return SKY_OK;
}

int get_plan_plain(int h, int w, int k, int stride, int transposed, int out_h, int out_w, int tp_ph0, int tp_pw0, const StripPlan **out,
                   bool device)
{
    const auto key = std::make_tuple(h, w, k, stride, transposed, out_h, out_w, tp_ph0, tp_pw0);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_plain.find(key);
    if (it == g_plain.end()) {
        HostPlan *hp = new HostPlan();
        const int rc = build_plain(h, w, k, stride, transposed, out_h, out_w, tp_ph0, tp_pw0, *hp);
        if (rc != SKY_OK) { delete hp; return rc; }
        it = g_plain.emplace(key, hp).first;
    }
    if (device) {
        const int rc = ensure_device(*it->second);
        if (rc != SKY_OK) return rc;
    }
    *out = &it->second->dev;
    return SKY_OK;
}

// weight-gradient plans share the caches: the `transposed` slot of the key carries 100 + 1000 * wpg + gmax
constexpr int WG_TW = 8, WG_NB = 8;
static int wg_cap(int wpg) { return wpg == 4 ? 8 : 4; }

int get_plan_da_wgrad(const float *offsets_host, int h, int w, int k, int wpg, int gmax, const StripPlan **out, bool device)
{
    if (h <= 0 || w <= 0 || k < 3 || !(k & 1) || gmax < 1) return SKY_ERR_UNSUPPORTED;
    const auto key = std::make_tuple(h, w, k, 100 + 1000 * wpg + gmax, fnv1a(offsets_host, (size_t)h * k * k * 2 * sizeof(float)));
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_da.find(key);
    if (it == g_da.end()) {
        HostPlan *hp = new HostPlan();
        const int rc = build_da(offsets_host, h, w, k, *hp, WG_TW, WG_NB, wg_cap(wpg));
        if (rc != SKY_OK) { delete hp; return rc; }
        build_wgrad_units(*hp, WG_NB, wpg, gmax);
        it = g_da.emplace(key, hp).first;
    }
    if (device) {
        const int rc = ensure_device(*it->second);
        if (rc != SKY_OK) return rc;
    }
    *out = &it->second->dev;
    return SKY_OK;
}

int get_plan_plain_wgrad(int h, int w, int k, int stride, int out_h, int out_w, int wpg, int gmax, const StripPlan **out, bool device)
{
    if (gmax < 1) return SKY_ERR_UNSUPPORTED;
    const auto key = std::make_tuple(h, w, k, stride, 100 + 1000 * wpg + gmax, out_h, out_w, 0, 0);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_plain.find(key);
    if (it == g_plain.end()) {
        HostPlan *hp = new HostPlan();
        const int rc = build_plain(h, w, k, stride, 0, out_h, out_w, 0, 0, *hp, WG_TW, WG_NB, wg_cap(wpg));
        if (rc != SKY_OK) { delete hp; return rc; }
        build_wgrad_units(*hp, WG_NB, wpg, gmax);
        it = g_plain.emplace(key, hp).first;
    }
    if (device) {
        const int rc = ensure_device(*it->second);
        if (rc != SKY_OK) return rc;
    }
    *out = &it->second->dev;
    return SKY_OK;
}

static void plan_info(const StripPlan &d, int *out8)
{
    out8[0] = d.nrows; out8[1] = d.nstrips; out8[2] = d.nwins; out8[3] = d.nterms; out8[4] = d.exact_strips;
    out8[5] = d.TW; out8[6] = d.NB; out8[7] = d.SR;
}

}  // namespace sky

using namespace sky;

// Plan export for tests and the design notes (host only, no device needed).  info: out8 = rows, strips, windows, terms, exact-tap strips,
// tile width, panoramas per tile, strip rows.  export: copies the host tables (RowPlan 16 B, StripDesc 40 B, WinDesc 8 B, WeffTerm 8 B).
extern "C" int sky_da_strip_plan_info(const float *offsets_host, int h, int w, int k, int transposed, int *out8)
{
    SKY_REQUIRE(offsets_host && out8, SKY_ERR_INVALID, "NULL pointer");
    const StripPlan *pl = nullptr;
    int rc = get_plan_da(offsets_host, h, w, k, &pl, false, transposed != 0);
    SKY_REQUIRE(rc == SKY_OK, rc, "no strip plan for h=%d w=%d k=%d", h, w, k);
    plan_info(*pl, out8);
    return SKY_OK;
}

extern "C" int sky_da_strip_plan_export(const float *offsets_host, int h, int w, int k, int transposed, void *rows, void *strips, void *wins,
                                        int *term_begin, void *terms)
{
    SKY_REQUIRE(offsets_host, SKY_ERR_INVALID, "NULL pointer");
    const StripPlan *pl = nullptr;
    int rc = get_plan_da(offsets_host, h, w, k, &pl, false, transposed != 0);
    SKY_REQUIRE(rc == SKY_OK, rc, "no strip plan for h=%d w=%d k=%d", h, w, k);
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_da)
        if (&kv.second->dev == pl) return export_plan(*kv.second, rows, strips, wins, term_begin, terms);
    return SKY_ERR_INVALID;
}

// Weight-gradient plans (strip_wgrad.cu): the forward plan with tiles of 8 columns x 8 panoramas, plus its accumulation units / MMA
// groups.  info: out12 = rows, strips, windows, terms, exact-tap strips, tile width, panoramas per tile, largest window shift of a strip,
// units, groups, most groups of a unit, 0.  export: the tables of sky_da_strip_plan_export + units {row, strip, group_begin, group_end}
// and groups {start_row, win[4]} (int32).
extern "C" int sky_da_strip_wgrad_plan_info(const float *offsets_host, int h, int w, int k, int wpg, int gmax, int *out12)
{
    SKY_REQUIRE(offsets_host && out12, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(wpg == 1 || wpg == 2 || wpg == 4, SKY_ERR_INVALID, "wpg must be 1, 2 or 4");
    const StripPlan *pl = nullptr;
    int rc = get_plan_da_wgrad(offsets_host, h, w, k, wpg, gmax, &pl, false);
    SKY_REQUIRE(rc == SKY_OK, rc, "no weight-gradient plan for h=%d w=%d k=%d", h, w, k);
    plan_info(*pl, out12);
    out12[7] = pl->span_max; out12[8] = pl->n_wg_units; out12[9] = pl->n_wg_groups; out12[10] = pl->max_groups_unit; out12[11] = 0;
    return SKY_OK;
}

extern "C" int sky_da_strip_wgrad_plan_export(const float *offsets_host, int h, int w, int k, int wpg, int gmax, void *rows, void *strips,
                                              void *wins, int *term_begin, void *terms, void *units, void *groups)
{
    SKY_REQUIRE(offsets_host, SKY_ERR_INVALID, "NULL pointer");
    const StripPlan *pl = nullptr;
    int rc = get_plan_da_wgrad(offsets_host, h, w, k, wpg, gmax, &pl, false);
    SKY_REQUIRE(rc == SKY_OK, rc, "no weight-gradient plan for h=%d w=%d k=%d", h, w, k);
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_da)
        if (&kv.second->dev == pl) {
            const HostPlan &hp = *kv.second;
            if (units) memcpy(units, hp.wg_units.data(), hp.wg_units.size() * sizeof(WgUnit));
            if (groups) memcpy(groups, hp.wg_groups.data(), hp.wg_groups.size() * sizeof(WgGroup));
            return export_plan(hp, rows, strips, wins, term_begin, terms);
        }
    return SKY_ERR_INVALID;
}

// the same for a plain SAME convolution (stride 1 / 2): out12 as above; export without terms (a window's tap is its wtile0)
extern "C" int sky_conv_strip_wgrad_plan_info(int h, int w, int k, int stride, int wpg, int gmax, int *out12)
{
    SKY_REQUIRE(out12 && (wpg == 1 || wpg == 2 || wpg == 4), SKY_ERR_INVALID, "bad arguments");
    const StripPlan *pl = nullptr;
    int rc = get_plan_plain_wgrad(h, w, k, stride, (h + stride - 1) / stride, (w + stride - 1) / stride, wpg, gmax, &pl, false);
    SKY_REQUIRE(rc == SKY_OK, rc, "no weight-gradient plan for this convolution");
    plan_info(*pl, out12);
    out12[7] = pl->span_max; out12[8] = pl->n_wg_units; out12[9] = pl->n_wg_groups; out12[10] = pl->max_groups_unit; out12[11] = 0;
    return SKY_OK;
}

extern "C" int sky_conv_strip_wgrad_plan_export(int h, int w, int k, int stride, int wpg, int gmax, void *rows, void *strips, void *wins,
                                                void *units, void *groups)
{
    const StripPlan *pl = nullptr;
    int rc = get_plan_plain_wgrad(h, w, k, stride, (h + stride - 1) / stride, (w + stride - 1) / stride, wpg, gmax, &pl, false);
    SKY_REQUIRE(rc == SKY_OK, rc, "no weight-gradient plan for this convolution");
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_plain)
        if (&kv.second->dev == pl) {
            const HostPlan &hp = *kv.second;
            if (units) memcpy(units, hp.wg_units.data(), hp.wg_units.size() * sizeof(WgUnit));
            if (groups) memcpy(groups, hp.wg_groups.data(), hp.wg_groups.size() * sizeof(WgGroup));
            return export_plan(hp, rows, strips, wins, nullptr, nullptr);
        }
    return SKY_ERR_INVALID;
}

extern "C" int sky_conv_strip_plan_info(int h, int w, int k, int stride, int transposed, int out_h, int out_w, int ph0, int pw0, int *out8)
{
    SKY_REQUIRE(out8, SKY_ERR_INVALID, "NULL pointer");
    const StripPlan *pl = nullptr;
    int rc = get_plan_plain(h, w, k, stride, transposed, out_h, out_w, ph0, pw0, &pl, false);
    SKY_REQUIRE(rc == SKY_OK, rc, "no strip plan for this convolution");
    plan_info(*pl, out8);
    return SKY_OK;
}

extern "C" int sky_conv_strip_plan_export(int h, int w, int k, int stride, int transposed, int out_h, int out_w, int ph0, int pw0, void *rows,
                                          void *strips, void *wins)
{
    const StripPlan *pl = nullptr;
    int rc = get_plan_plain(h, w, k, stride, transposed, out_h, out_w, ph0, pw0, &pl, false);
    SKY_REQUIRE(rc == SKY_OK, rc, "no strip plan for this convolution");
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &kv : g_plain)
        if (&kv.second->dev == pl) return export_plan(*kv.second, rows, strips, wins, nullptr, nullptr);
    return SKY_ERR_INVALID;
}
