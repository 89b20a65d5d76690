"""Row-strip plans (csrc/strip_plan.cu) checked on the CPU: the host tables the library builds are exported through the C ABI and the
strip formulation  y[i, j] = sum over strips, windows of  V[j + shift] . Weff  is evaluated in numpy (fp64) from them, then compared
with the oracle's materialised pad -> gather -> blend -> matmul dataflow (distortion_aware_ops.py:50-123) and with a direct SAME
convolution / its data gradient for the plain plans.  No device is needed: plans are built from the host copy of the offset table."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import da_oracle as O

ROW = np.dtype([("out_row", "i4"), ("oc0", "i4"), ("sb", "i4"), ("se", "i4")])
STRIP = np.dtype([("kind", "i4"), ("r0", "i4"), ("r1", "i4"), ("wy0", "f4"), ("wy1", "f4"), ("u0", "i4"), ("cm", "i4"), ("c0", "i4"),
                  ("wb", "i4"), ("we", "i4")])
WIN = np.dtype([("start_row", "i4"), ("wtile0", "i4")])
TERM = np.dtype([("tap", "i4"), ("coef", "f4")])


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def da_plan(pkg, off, h, w, k, transposed=0):
    lib, chk = pkg._lib.LIB, pkg._lib.check
    info = np.zeros(8, np.int32)
    chk(lib.sky_da_strip_plan_info(_vp(off), h, w, k, transposed, _vp(info)))
    nr, ns, nw, nt = (int(v) for v in info[:4])
    rows, strips, wins = np.zeros(nr, ROW), np.zeros(ns, STRIP), np.zeros(nw, WIN)
    tb, terms = np.zeros(nw + 1, np.int32), np.zeros(max(nt, 1), TERM)
    chk(lib.sky_da_strip_plan_export(_vp(off), h, w, k, transposed, _vp(rows), _vp(strips), _vp(wins), _vp(tb), _vp(terms)))
    return info, rows, strips, wins, tb, terms


def map_col(q, in_w, pw0, W):
    """da_map_col of strip_conv.cu."""
    if q < 0:
        q += in_w
    elif q > in_w - 1:
        q -= in_w
    if q < 0:
        q += in_w
    if q > in_w - 1:
        q -= in_w
    c = q - pw0
    return c if 0 <= c < W else -1


def emulate_da(x, kern, bias, off, k, plan, rnd=None):
    """The kernel's arithmetic from the exported plan: strips V = wy0 * x[r0] + wy1 * x[r1] (fp32, like the producer warps), effective
    weights sum coef * kernel[tap] (fp32, like strip_weff_pack_kernel), both passed through `rnd` (None: exact; tf32_emu.round_tf32: the
    TF32 operand rounding), contraction accumulated in fp64."""
    info, rows, strips, wins, tb, terms = plan
    B, h, w, C = x.shape
    F = kern.shape[1]
    NB = int(info[6])
    (ph0, _), pw = O.pad_amounts(h, k), O.pad_amounts(w, k)
    pw0, in_w = pw[0], w + sum(pw)
    f32 = np.float32
    rnd = rnd or (lambda a: a)
    x32, k32 = x.astype(f32), kern.astype(f32).reshape(k * k, C, F)
    y = np.zeros((B, h, w, F))
    samp = O.sample(h, w, k, off)
    jj = np.arange(w)
    for rp in rows:
        i = int(rp["out_row"])
        for sd in strips[rp["sb"]:rp["se"]]:
            v = None
            if sd["kind"] == 0:
                v0 = x32[:, sd["r0"]] if sd["r0"] >= 0 else np.zeros((B, w, C), f32)
                v1 = x32[:, sd["r1"]] if sd["r1"] >= 0 else np.zeros((B, w, C), f32)
                v = rnd((f32(sd["wy1"]) * v1 + (f32(sd["wy0"]) * v0).astype(f32)).astype(f32)).astype(np.float64)    # [B, w, C]
            for wi in range(sd["wb"], sd["we"]):
                weff = np.zeros((C, F), f32)
                for t in terms[tb[wi]:tb[wi + 1]]:
                    weff = (f32(t["coef"]) * k32[t["tap"]] + weff).astype(f32)
                weff = rnd(weff).astype(np.float64)
                if sd["kind"] == 1:
                    t = int(sd["r0"])           # exact tap: the oracle's own corners / weights per pixel
                    assert terms[tb[wi]]["tap"] == t and tb[wi + 1] - tb[wi] == 1
                    pix = np.zeros((B, w, C), f32)
                    for (yn, xn, wn) in (("y0", "x0", "w0"), ("y0", "x1", "w1"), ("y1", "x0", "w2"), ("y1", "x1", "w3")):
                        r, c = samp[yn][i, :, t] - ph0, samp[xn][i, :, t] - pw0
                        ok = (r >= 0) & (r < h) & (c >= 0) & (c < w)
                        pix[:, ok] += samp[wn][i, ok, t][None, :, None] * x32[:, r[ok], c[ok]]
                    y[:, i] += rnd(pix).astype(np.float64) @ weff
                    continue
                assert wins[wi]["start_row"] % NB == 0
                shift = int(sd["u0"]) + int(wins[wi]["start_row"]) // NB
                cols = np.array([map_col(int(sd["cm"]) * (j + shift) + int(sd["c0"]) + pw0, in_w, pw0, w) for j in jj])
                ok = cols >= 0
                y[:, i, ok] += v[:, cols[ok]] @ weff
    return y + bias.astype(np.float64)


@pytest.mark.parametrize("h,w,k", [(8, 32, 3), (16, 64, 3), (32, 128, 7), (4, 16, 3), (16, 64, 5)])
def test_da_strip_plan_reproduces_the_layer(pkg, h, w, k):
    rng = np.random.default_rng(h * 1000 + w + k)
    B, C, F = 2, 4, 5
    off = O.offsets(h, w, k)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = rng.standard_normal((k * k * C, F)).astype(np.float32)
    bias = rng.standard_normal(F).astype(np.float32)
    plan = da_plan(pkg, off, h, w, k)
    info = plan[0]
    assert info[0] == h and info[5] * info[6] == 128 and info[7] <= 192 and info[7] % 8 == 0
    want = O.conv2d_forward(x, kern, bias, k, acc_dtype=torch.float64).numpy()
    got = emulate_da(x, kern, bias, off, k, plan)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    # the only deviation: a row's horizontal factors are the exact fraction of b + x_off, the reference's are fp32 per pixel (j + b + x_off
    # rounded): a few ulp of the coordinate
    assert rel <= 2e-5, (rel, info)
    # every window starts inside the strip the kernel allocates
    _, rows, strips, wins, tb, terms = plan
    assert int(wins["start_row"].max()) + 128 <= int(info[7])


def test_da_strip_plan_statistics(pkg):
    """The trunk geometry (8x32, k=3): k strips per row except at the zenith, (2k-1) windows per strip at most."""
    off = O.offsets(8, 32, 3)
    info, rows, strips, wins, tb, terms = da_plan(pkg, off, 8, 32, 3)
    per_row = [int(r["se"] - r["sb"]) for r in rows]
    wins_per_row = [int(sum(s["we"] - s["wb"] for s in strips[r["sb"]:r["se"]])) for r in rows]
    print("strips per row", per_row, "windows per row", wins_per_row, "exact strips", int(info[4]), "SR", int(info[7]))
    assert max(wins_per_row[2:]) <= 3 * 5
    assert int(info[4]) <= 9        # exact taps, if any, only on the zenith row


def conv_plan(pkg, h, w, k, stride, transposed, oh, ow, ph0, pw0):
    lib, chk = pkg._lib.LIB, pkg._lib.check
    info = np.zeros(8, np.int32)
    chk(lib.sky_conv_strip_plan_info(h, w, k, stride, transposed, oh, ow, ph0, pw0, _vp(info)))
    rows, strips, wins = np.zeros(int(info[0]), ROW), np.zeros(int(info[1]), STRIP), np.zeros(int(info[2]), WIN)
    chk(lib.sky_conv_strip_plan_export(h, w, k, stride, transposed, oh, ow, ph0, pw0, _vp(rows), _vp(strips), _vp(wins)))
    return info, rows, strips, wins


def emulate_plain(x, taps, plan, OH, OW, ocs):
    """x [B,h,w,C]; taps [k*k, C, F] (the weight tile of tap t); returns [B,OH,OW,F]."""
    info, rows, strips, wins = plan
    B, h, w, C = x.shape
    NB = int(info[6])
    y = np.zeros((B, OH, OW, taps.shape[2]))
    ncols = OW // ocs
    for rp in rows:
        for sd in strips[rp["sb"]:rp["se"]]:
            assert sd["kind"] == 0 and sd["r1"] < 0 and sd["wy0"] == 1.0
            for wi in range(sd["wb"], sd["we"]):
                shift = int(sd["u0"]) + int(wins[wi]["start_row"]) // NB
                for jj in range(ncols):
                    c = int(sd["cm"]) * (jj + shift) + int(sd["c0"])
                    if 0 <= c < w:
                        y[:, rp["out_row"], rp["oc0"] + ocs * jj, :] += x[:, sd["r0"], c, :] @ taps[wins[wi]["wtile0"]]
    return y


@pytest.mark.parametrize("h,w,k,stride", [(8, 32, 3, 1), (16, 16, 3, 2), (8, 32, 4, 2), (4, 16, 4, 1), (32, 128, 7, 1), (5, 12, 3, 2)])
def test_plain_strip_plan_forward(pkg, h, w, k, stride):
    rng = np.random.default_rng(k * 100 + stride)
    B, C, F = 2, 3, 4
    x = rng.standard_normal((B, h, w, C))
    wt = rng.standard_normal((k, k, C, F))
    oh, ow = -(-h // stride), -(-w // stride)
    plan = conv_plan(pkg, h, w, k, stride, 0, oh, ow, 0, 0)
    got = emulate_plain(x, wt.reshape(k * k, C, F), plan, oh, ow, 1)
    th, tw = max((oh - 1) * stride + k - h, 0), max((ow - 1) * stride + k - w, 0)          # TensorFlow SAME: the smaller half in front
    xp = torch.nn.functional.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2))
    want = torch.nn.functional.conv2d(xp, torch.from_numpy(wt).permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1).numpy()
    assert np.allclose(got, want, atol=1e-10)


@pytest.mark.parametrize("h,w,k,stride", [(8, 32, 4, 2), (16, 16, 3, 2), (8, 16, 4, 1), (6, 12, 4, 2)])
def test_plain_strip_plan_data_gradient(pkg, h, w, k, stride):
    """dx of a SAME conv run as a forward pass over dy with the flipped, transposed kernel (the convention of sky_conv2d_bwd_data)."""
    rng = np.random.default_rng(k * 10 + stride)
    B, C, F = 2, 3, 4
    oh, ow = -(-h // stride), -(-w // stride)
    th, tw = max((oh - 1) * stride + k - h, 0), max((ow - 1) * stride + k - w, 0)
    ph0, pw0 = th // 2, tw // 2
    xt = torch.from_numpy(rng.standard_normal((B, h, w, C))).requires_grad_(True)
    wt = rng.standard_normal((k, k, C, F))
    xp = torch.nn.functional.pad(xt.permute(0, 3, 1, 2), (pw0, tw - pw0, ph0, th - ph0))
    y = torch.nn.functional.conv2d(xp, torch.from_numpy(wt).permute(3, 2, 0, 1), stride=stride).permute(0, 2, 3, 1)
    dy = rng.standard_normal(tuple(y.shape))
    y.backward(torch.from_numpy(dy))
    want = xt.grad.numpy()
    # tap (a, b) of the transposed pass multiplies dy by kernel[k-1-a, k-1-b]^T (sky_conv2d_transpose_weights)
    taps = np.stack([wt[k - 1 - a, k - 1 - b].T for a in range(k) for b in range(k)])      # [k*k, F, C]
    plan = conv_plan(pkg, oh, ow, k, stride, 1, h, w, k - 1 - ph0, k - 1 - pw0)
    got = emulate_plain(dy, taps, plan, h, w, 2 if stride == 2 else 1)
    assert np.allclose(got, want, atol=1e-10)


def map_col_t(q, in_w, W):
    """da_map_col_t of strip_conv.cu: the unique dy column j = q + m * in_w, |m| <= 2, inside the map."""
    for m in range(-2, 3):
        j = q + m * in_w
        if 0 <= j < W:
            return j
    return -1


def emulate_da_dgrad(dy, kern, C, off, k, plan, rnd=None):
    """dx from the transposed plan: strips are dy rows, windows carry the merged transposed effective weights."""
    info, rows, strips, wins, tb, terms = plan
    B, h, w, F = dy.shape
    NB = int(info[6])
    in_w = w + sum(O.pad_amounts(w, k))
    f32 = np.float32
    rnd = rnd or (lambda a: a)
    k32 = kern.astype(f32).reshape(k * k, C, F)
    dy_r = rnd(dy.astype(f32)).astype(np.float64)
    dx = np.zeros((B, h, w, C))
    for rp in rows:
        r = int(rp["out_row"])
        for sd in strips[rp["sb"]:rp["se"]]:
            assert sd["kind"] == 0 and sd["r1"] < 0 and sd["wy0"] == 1.0 and sd["cm"] == 1 and sd["c0"] == 0
            for wi in range(sd["wb"], sd["we"]):
                weff = np.zeros((C, F), f32)
                for t in terms[tb[wi]:tb[wi + 1]]:
                    weff = (f32(t["coef"]) * k32[t["tap"]] + weff).astype(f32)
                weff_t = rnd(weff).astype(np.float64).T                                   # [F, C]
                shift = int(sd["u0"]) + int(wins[wi]["start_row"]) // NB
                cols = np.array([map_col_t(c + shift, in_w, w) for c in range(w)])
                ok = cols >= 0
                dx[:, r, ok] += dy_r[:, sd["r0"], cols[ok]] @ weff_t
    return dx


@pytest.mark.parametrize("h,w,k", [(8, 32, 3), (16, 64, 3), (32, 128, 7), (4, 16, 3)])
def test_da_strip_plan_data_gradient(pkg, h, w, k):
    """The transposed plan against autograd through the oracle's materialised forward (what TF autodiff derives)."""
    rng = np.random.default_rng(h + w + k)
    B, C, F = 2, 4, 5
    off = O.offsets(h, w, k)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = rng.standard_normal((k * k * C, F)).astype(np.float32)
    bias = np.zeros(F, np.float32)
    dy = rng.standard_normal((B, h, w, F)).astype(np.float32)
    want = O.conv2d_backward(x, kern, bias, dy, k, acc_dtype=torch.float64)[0].numpy()
    plan = da_plan(pkg, off, h, w, k, transposed=1)
    info = plan[0]
    assert info[4] == 0 and info[7] <= 192
    got = emulate_da_dgrad(dy, kern, C, off, k, plan)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel <= 3e-5, (rel, info)
    _, rows, strips, wins, tb, terms = plan
    assert int(wins["start_row"].max()) + 128 <= int(info[7])
    print("transposed plan", (h, w, k), "windows per row", [int(sum(s["we"] - s["wb"] for s in strips[r["sb"]:r["se"]])) for r in rows][:12],
          "strips per row", [int(r["se"] - r["sb"]) for r in rows][:12], "SR", int(info[7]))


# --------------------------------------------------------------------------------------------------------------------------------
# weight-gradient plans (csrc/strip_wgrad.cu): units / MMA groups over the forward plan with tiles of 8 columns x 8 panoramas
# --------------------------------------------------------------------------------------------------------------------------------
UNIT = np.dtype([("row", "i4"), ("strip", "i4"), ("gb", "i4"), ("ge", "i4")])
GROUP = np.dtype([("start_row", "i4"), ("win", "i4", (4,))])


def da_wgrad_plan(pkg, off, h, w, k, wpg, gmax):
    lib, chk = pkg._lib.LIB, pkg._lib.check
    info = np.zeros(12, np.int32)
    chk(lib.sky_da_strip_wgrad_plan_info(_vp(off), h, w, k, wpg, gmax, _vp(info)))
    nr, ns, nw, nt = (int(v) for v in info[:4])
    rows, strips, wins = np.zeros(nr, ROW), np.zeros(ns, STRIP), np.zeros(nw, WIN)
    tb, terms = np.zeros(nw + 1, np.int32), np.zeros(max(nt, 1), TERM)
    units, groups = np.zeros(int(info[8]), UNIT), np.zeros(int(info[9]), GROUP)
    chk(lib.sky_da_strip_wgrad_plan_export(_vp(off), h, w, k, wpg, gmax, _vp(rows), _vp(strips), _vp(wins), _vp(tb), _vp(terms), _vp(units), _vp(groups)))
    return info, rows, strips, wins, tb, terms, units, groups


def emulate_da_wgrad(x, dy, off, k, plan):
    """What strip_wgrad_kernel computes, from the exported plan, in fp64: per unit and MMA group the accumulators
    G[c, f] = sum over panoramas, columns of strip[column + shift, c] * dy[column, f] (one per window of the group), added to the taps of the
    window with their coefficients."""
    info, rows, strips, wins, tb, terms, units, groups = plan
    B, h, w, C = x.shape
    F = dy.shape[-1]
    NB = int(info[6])
    (ph0, _), pw = O.pad_amounts(h, k), O.pad_amounts(w, k)
    pw0, in_w = pw[0], w + sum(pw)
    samp = O.sample(h, w, k, off)
    dk = np.zeros((k * k, C, F))
    x64, dy64 = x.astype(np.float64), dy.astype(np.float64)
    jj = np.arange(w)
    for u in units:
        rp, sd = rows[u["row"]], strips[u["strip"]]
        i = int(rp["out_row"])
        for g in groups[u["gb"]:u["ge"]]:
            for q, wi in enumerate(g["win"]):
                if wi < 0:
                    continue
                assert sd["wb"] <= wi < sd["we"]
                assert wins[wi]["start_row"] == g["start_row"] + q * NB          # quarter q of the MMA = the window one column shift further
                if sd["kind"] == 1:
                    t = int(sd["r0"])
                    pix = np.zeros((B, w, C))
                    for (yn, xn, wn) in (("y0", "x0", "w0"), ("y0", "x1", "w1"), ("y1", "x0", "w2"), ("y1", "x1", "w3")):
                        r, c = samp[yn][i, :, t] - ph0, samp[xn][i, :, t] - pw0
                        ok = (r >= 0) & (r < h) & (c >= 0) & (c < w)
                        pix[:, ok] += samp[wn][i, ok, t][None, :, None].astype(np.float64) * x64[:, r[ok], c[ok]]
                    G = np.einsum("bjc,bjf->cf", pix, dy64[:, i])
                else:
                    v0 = x64[:, sd["r0"]] if sd["r0"] >= 0 else np.zeros((B, w, C))
                    v1 = x64[:, sd["r1"]] if sd["r1"] >= 0 else np.zeros((B, w, C))
                    v = float(sd["wy0"]) * v0 + float(sd["wy1"]) * v1
                    shift = int(sd["u0"]) + int(wins[wi]["start_row"]) // NB
                    cols = np.array([map_col(int(sd["cm"]) * (j + shift) + int(sd["c0"]) + pw0, in_w, pw0, w) for j in jj])
                    ok = cols >= 0
                    G = np.einsum("bjc,bjf->cf", v[:, cols[ok]], dy64[:, i, ok])
                for t in terms[tb[wi]:tb[wi + 1]]:
                    dk[int(t["tap"])] += float(t["coef"]) * G
    return dk.reshape(k * k * C, F)


@pytest.mark.parametrize("h,w,k,wpg,gmax", [(8, 32, 3, 1, 4), (16, 64, 3, 1, 8), (32, 128, 7, 4, 16), (16, 64, 5, 4, 5), (4, 16, 3, 4, 16), (16, 64, 3, 2, 8), (8, 32, 5, 2, 4)])
def test_da_wgrad_plan_reproduces_the_weight_gradient(pkg, h, w, k, wpg, gmax):
    rng = np.random.default_rng(h + w + k + wpg)
    B, C, F = 3, 4, 5
    off = O.offsets(h, w, k)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    dy = rng.standard_normal((B, h, w, F)).astype(np.float32)
    kern = rng.standard_normal((k * k * C, F)).astype(np.float32)
    plan = da_wgrad_plan(pkg, off, h, w, k, wpg, gmax)
    info, rows, strips, wins, tb, terms, units, groups = plan
    # structure: tiles of 8 columns x 8 panoramas, every window in exactly one group, at most gmax accumulators per unit, window starts
    # on whole column shifts (8 strip rows = two K atoms of the MN-major operand), spans within the strip the kernel allocates
    assert int(info[5]) == 8 and int(info[6]) == 8
    seen = np.zeros(len(wins), np.int32)
    for u in units:
        assert 1 <= u["ge"] - u["gb"] <= gmax
        for g in groups[u["gb"]:u["ge"]]:
            assert g["start_row"] % 8 == 0 and g["win"][0] >= 0
            for wi in g["win"]:
                if wi >= 0:
                    seen[wi] += 1
            if wpg == 1:
                assert (g["win"][1:] < 0).all()
    assert (seen == 1).all()
    assert int(wins["start_row"].max()) // 8 <= int(info[7]) <= (8 if wpg == 4 else 4)
    if wpg > 1:
        assert all((g["win"][wpg:] < 0).all() for g in groups)
    want = O.conv2d_backward(x, kern, np.zeros(F, np.float32), dy, k, acc_dtype=torch.float64)[1].numpy()
    got = emulate_da_wgrad(x, dy, off, k, plan)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel <= 3e-5, (rel, info)


def test_wgrad_schedule_covers_every_tile_once(pkg):
    """The work split of the strip weight gradient (host logic of launch_wgrad_strip): CTA (ub, pb) of wave w owns tiles
    [pb * TP, pb * TP + TP) of accumulation w * U + ub — every (accumulation, tile) exactly once, never more CTAs than SMs, and the L2 cap."""
    lib, chk = pkg._lib.LIB, pkg._lib.check
    for (nuidx, ntiles, sms, big, ucap) in ((40, 16, 148, 0, 16), (340, 256, 148, 1, 16), (340, 256, 148, 1, 4), (1, 1, 148, 0, 16), (7, 3, 148, 0, 16),
                                            (500, 2, 148, 0, 16), (3, 1000, 148, 1, 16), (224, 64, 132, 0, 16)):
        out = np.zeros(4, np.int32)
        chk(lib.sky_wgrad_schedule_info(nuidx, ntiles, sms, big, ucap, _vp(out)))
        P, TP, U, waves = (int(v) for v in out)
        assert P >= 1 and TP >= 1 and U >= 1 and U * P <= sms
        assert P * TP >= ntiles and (P - 1) * TP < ntiles          # no empty part
        assert waves == -(-nuidx // U)
        if big and P < ntiles:
            assert U <= ucap
        seen = np.zeros((nuidx, ntiles), np.int32)
        for w in range(waves):
            for b in range(U * P):
                ub, pb = divmod(b, P)
                u = w * U + ub
                if u < nuidx:
                    seen[u, pb * TP:min(ntiles, pb * TP + TP)] += 1
        assert (seen == 1).all(), (nuidx, ntiles, P, TP, U)


@pytest.mark.parametrize("h,w,k,stride,wpg,gmax", [(8, 32, 3, 1, 1, 4), (16, 16, 3, 2, 4, 16), (8, 32, 4, 2, 2, 8), (4, 16, 4, 1, 1, 4), (16, 64, 7, 1, 4, 16),
                                                   (5, 12, 3, 2, 2, 8), (7, 9, 1, 1, 4, 16)])
def test_plain_wgrad_plan_reproduces_the_weight_gradient(pkg, h, w, k, stride, wpg, gmax):
    """Weight-gradient plan of a plain SAME convolution (tf.nn.conv2d, ops.py:41; stride 1 / 2, even kernels, 1x1) evaluated in numpy
    against autograd through the oracle's conv."""
    from oracle import model_oracle as M
    lib, chk = pkg._lib.LIB, pkg._lib.check
    info = np.zeros(12, np.int32)
    chk(lib.sky_conv_strip_wgrad_plan_info(h, w, k, stride, wpg, gmax, _vp(info)))
    rows, strips, wins = np.zeros(int(info[0]), ROW), np.zeros(int(info[1]), STRIP), np.zeros(int(info[2]), WIN)
    units, groups = np.zeros(int(info[8]), UNIT), np.zeros(int(info[9]), GROUP)
    chk(lib.sky_conv_strip_wgrad_plan_export(h, w, k, stride, wpg, gmax, _vp(rows), _vp(strips), _vp(wins), _vp(units), _vp(groups)))
    assert int(info[5]) == 8 and int(info[6]) == 8
    rng = np.random.default_rng(h * 100 + w + k + stride)
    B, C, F = 3, 4, 5
    x = rng.standard_normal((B, h, w, C))
    wt = torch.from_numpy(rng.standard_normal((k, k, C, F))).requires_grad_(True)
    y = M.conv2d_same(torch.from_numpy(x), wt, torch.zeros(F, dtype=torch.float64), stride=stride, acc_dtype=torch.float64)
    dy = rng.standard_normal(tuple(y.shape))
    y.backward(torch.from_numpy(dy))
    want = wt.grad.numpy().reshape(k * k, C, F)
    OW = dy.shape[2]
    dk = np.zeros((k * k, C, F))
    seen = np.zeros(len(wins), np.int32)
    for u in units:
        rp, sd = rows[u["row"]], strips[u["strip"]]
        assert 1 <= u["ge"] - u["gb"] <= gmax and sd["kind"] == 0 and sd["r1"] < 0 and sd["wy0"] == 1.0
        for g in groups[u["gb"]:u["ge"]]:
            for q, wi in enumerate(g["win"]):
                if wi < 0:
                    continue
                assert q < wpg and wins[wi]["start_row"] == g["start_row"] + 8 * q
                seen[wi] += 1
                shift = int(sd["u0"]) + int(wins[wi]["start_row"]) // 8
                for j in range(OW):
                    c = int(sd["cm"]) * (j + shift) + int(sd["c0"])
                    if 0 <= c < w:
                        dk[int(wins[wi]["wtile0"])] += np.einsum("bc,bf->cf", x[:, sd["r0"], c], dy[:, rp["out_row"], j])
    assert (seen == 1).all()
    assert np.linalg.norm(dk - want) / np.linalg.norm(want) < 1e-12
