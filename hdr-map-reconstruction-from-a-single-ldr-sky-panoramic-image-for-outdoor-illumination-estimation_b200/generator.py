"""Host-side mirror of the reference's generator.py: the same class names and constructor signatures
(``resBlock(filter_in, filter_out, k_h=3, k_w=3, strides=1, dilation_rate=1)``,
``resLayer(filters, filter_in, k_h, k_w, strides=1, dilation_rate=1)``, ``model(batch_size, im_height, im_width, da_kernel_size,
dilation_rate)``).  ``distortion_aware=True`` (default) wires every res-block conv the way the commented lines generator.py:14,18 do
(``distortion_aware_ops.conv2d``: the north-star path); ``distortion_aware=False`` gives the plain ``ops.conv2d`` wiring that is live
in the reference as committed (generator.py:13,17) — same variables ([k,k,C,F] kernels), so a reference SKY checkpoint maps onto it.

    resBlock.call  (generator.py:26-35):  identity(x) + IN(conv2(leaky_relu_0.1(IN(conv1(x)))))

Launches per res-block: conv1 (+ fused IN moments) -> IN apply + LeakyReLU -> conv2 (+ moments) -> IN apply + residual.
Training direction (train.train_step): ``model.train_forward`` keeps every intermediate, ``model.train_backward`` walks the graph in
reverse — instance-norm backward fused with the LeakyReLU masks, data gradients as gathers, pipelined tcgen05 weight gradients —
writing every gradient into the flat buffer of ``_flat.FlatVars``.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import LIB, check
from . import ops
from .distortion_aware_ops import _require_cuda, _stream, conv2d as da_conv2d, conv2d_backward, zero_
from .sunrad_net import sunRadNet


def _kernel_attr(conv):
    """Name of the kernel / bias attributes of the three conv flavours (reference variable names)."""
    if isinstance(conv, da_conv2d):
        return "kernel", "bias"
    return ("kernel" if isinstance(conv, ops.deconv2d) else "w"), "biases"


def conv_owner_list(conv):
    k, b = _kernel_attr(conv)
    return [(conv, k), (conv, b)]


def instnorm_backward(norm, z, stats, dy, act, slope, fv, extra=None):
    """tfa InstanceNormalization backward fused with the LeakyReLU mask (act = the activation that followed, or None)."""
    B, h, w, F = z.shape
    dz = torch.empty_like(z)
    sums = torch.empty(B, F, 2, dtype=torch.float64, device=z.device)
    check(LIB.sky_instnorm_bwd(z.data_ptr(), stats.data_ptr(), norm.gamma.data_ptr(), dy.data_ptr(), None if act is None else act.data_ptr(),
                               None if extra is None else extra.data_ptr(), sums.data_ptr(), dz.data_ptr(), fv.grad(norm, "gamma").data_ptr(),
                               fv.grad(norm, "beta").data_ptr(), B, h, w, F, norm.epsilon, float(slope), _stream()))
    return dz


def conv_backward(conv, x, dy, fv, need_dx=True, residual=None):
    """Weight / bias gradients into the flat buffer and (optionally) the data gradient (+ residual) of either conv flavour."""
    kname, bname = _kernel_attr(conv)
    dk, db = fv.grad(conv, kname), fv.grad(conv, bname)
    # fv.zeroed_each_step (set by train.Step, which zeroes the flat gradient buffer once before the backward pass): the weight-gradient
    # kernels add to the views instead of zeroing them first — two memset nodes less per conv in the step's graph
    acc = bool(getattr(fv, "zeroed_each_step", False))
    if isinstance(conv, da_conv2d):
        dx = None
        if need_dx and residual is not None:
            dx = residual.clone()
        return conv2d_backward(conv, x, dy, need_dx=need_dx, dx_out=dx, dk_out=dk, db_out=db, accumulate_dx=residual is not None,
                               accumulate_dw=acc)[0]
    conv.backward_filter(x, dy, dk, db, accumulate=acc and x.shape[-1] > 4)
    return conv.backward_data(x, dy, residual=residual) if need_dx else None


class InstanceNormalization:
    """tfa.layers.InstanceNormalization() defaults: axis=-1, epsilon=1e-3, center/scale with beta=0, gamma=1."""

    def __init__(self, epsilon=1e-3, device="cuda"):
        self.epsilon = float(epsilon)
        self.device = torch.device(device)
        self.gamma = None
        self.beta = None

    def build(self, input_shape):
        c = int(input_shape[-1])
        self.gamma = torch.ones(c, dtype=torch.float32, device=self.device)
        self.beta = torch.zeros(c, dtype=torch.float32, device=self.device)

    @property
    def trainable_variables(self):
        return [self.gamma, self.beta]

    def apply(self, x, stats, leaky_slope=None, residual=None):
        """Normalise `x` with the fp64 moments `stats` [B, C, 2] that the producing conv accumulated."""
        if self.gamma is None:
            self.build(tuple(x.shape))
        x = _require_cuda(x, "inputs")
        B, h, w, C = x.shape
        y = torch.empty_like(x)
        flags = _lib.EPI_NONE
        if leaky_slope is not None:
            flags |= _lib.EPI_LEAKY_RELU
        if residual is not None:
            residual = _require_cuda(residual, "residual")
            flags |= _lib.EPI_RESIDUAL
        check(LIB.sky_instnorm_apply(x.data_ptr(), stats.data_ptr(), self.gamma.data_ptr(), self.beta.data_ptr(),
                                     None if residual is None else residual.data_ptr(), y.data_ptr(), B, h, w, C,
                                     self.epsilon, flags, float(leaky_slope or 0.0), _stream()))
        return y


class resBlock:
    def __init__(self, filter_in, filter_out, k_h=3, k_w=3, strides=1, dilation_rate=1, *, distortion_aware=True, math_mode=None,
                 device="cuda"):
        if k_h != k_w:
            raise ValueError("the distortion-aware conv takes one kernel_size (generator.py:14)")

        def conv():
            if distortion_aware:                                                # generator.py:14,18
                return da_conv2d(filter_out, kernel_size=k_h, strides=strides, dilation_rate=dilation_rate, math_mode=math_mode,
                                 device=device)
            return ops.conv2d(filter_out, strides=strides, k_h=k_h, k_w=k_w, math_mode=math_mode, device=device)   # :13,17
        self.distortion_aware = bool(distortion_aware)
        self.filter_out = filter_out
        self.conv1 = conv()
        self.norm1 = InstanceNormalization(device=device)                       # :15
        self.conv2 = conv()
        self.norm2 = InstanceNormalization(device=device)                       # :19
        self.identity = None                                                    # :21-22 `lambda x: x`
        if filter_in != filter_out:                                             # :23-24 1x1 projection shortcut
            self.identity = ops.conv2d(filter_out, strides=1, k_h=1, k_w=1, math_mode=math_mode, device=device)
        self._stats = None

    def _moments(self, B, C, device):
        if self._stats is None or self._stats.shape[1] != B or self._stats.shape[2] != C:
            self._stats = torch.empty(2, B, C, 2, dtype=torch.float64, device=device)
        zero_(self._stats)
        return self._stats[0], self._stats[1]

    @property
    def trainable_variables(self):
        return (self.conv1.trainable_variables + self.norm1.trainable_variables + self.conv2.trainable_variables
                + self.norm2.trainable_variables + (self.identity.trainable_variables if self.identity is not None else []))

    def owner_list(self):
        """(object, attribute) of every trainable variable, in trainable_variables order."""
        out = []
        for conv, norm in ((self.conv1, self.norm1), (self.conv2, self.norm2)):
            out += conv_owner_list(conv) + [(norm, "gamma"), (norm, "beta")]
        if self.identity is not None:
            out += conv_owner_list(self.identity)
        return out

    def build(self, input_shape):
        if self.identity is not None and not self.identity.built:
            self.identity.build(tuple(input_shape))
        for conv, norm in ((self.conv1, self.norm1), (self.conv2, self.norm2)):
            if not conv.built:
                conv.build(tuple(input_shape))
            shape = tuple(input_shape[:3]) + (self.filter_out,)
            if norm.gamma is None:
                norm.build(shape)
            input_shape = shape

    def set_weights(self, w):
        """w: dict conv{1,2}_kernel / conv{1,2}_bias / norm{1,2}_gamma / norm{1,2}_beta (numpy or torch), the reference's
        variable names under each res-block; kernels as [k*k*C, F] (either wiring; the plain variable is the same memory as [k,k,C,F])."""
        for i, (conv, norm) in enumerate(((self.conv1, self.norm1), (self.conv2, self.norm2)), start=1):
            kname, bname = _kernel_attr(conv)
            kv = getattr(conv, kname)
            kv.copy_(torch.as_tensor(w[f"conv{i}_kernel"]).reshape(kv.shape))
            getattr(conv, bname).copy_(torch.as_tensor(w[f"conv{i}_bias"]))
            norm.gamma.copy_(torch.as_tensor(w[f"norm{i}_gamma"]))
            norm.beta.copy_(torch.as_tensor(w[f"norm{i}_beta"]))
        if self.identity is not None and "identity_kernel" in w:
            self.identity.w.copy_(torch.as_tensor(w["identity_kernel"]).reshape(self.identity.w.shape))
            self.identity.biases.copy_(torch.as_tensor(w["identity_bias"]))

    def call(self, x, save=False):
        x = _require_cuda(x, "inputs")
        if not self.conv1.built:
            self.build(tuple(x.shape))
        B = x.shape[0]
        s1, s2 = self._moments(B, self.filter_out, x.device)
        conv1 = self.conv1.call(x, stats=s1)                                    # :28
        actv1 = self.norm1.apply(conv1, s1, leaky_slope=0.1)                    # :29-30
        conv2 = self.conv2.call(actv1, stats=s2)                                # :32
        ident = x if self.identity is None else self.identity.call(x)           # :21-24
        out = self.norm2.apply(conv2, s2, residual=ident)                       # :33-35
        if save:
            self._saved = (x, conv1, actv1, conv2, s1, s2)
        return out

    __call__ = call

    def backward(self, dout, fv):
        """Gradient w.r.t. the block input; every variable gradient goes to fv (dgamma / dbeta accumulate: fv.flat_g zeroed per step)."""
        x, c1, a1, c2, s1, s2 = self._saved
        dc2 = instnorm_backward(self.norm2, c2, s2, dout, None, 0.1, fv)                         # IN2 (:33)
        da1 = conv_backward(self.conv2, a1, dc2, fv)                                             # conv2 (:32)
        dc1 = instnorm_backward(self.norm1, c1, s1, da1, a1, 0.1, fv)                            # lrelu (:30) + IN1 (:29)
        if self.identity is None:
            return conv_backward(self.conv1, x, dc1, fv, residual=dout)                          # conv1 (:28) + identity branch (:35)
        dx = conv_backward(self.identity, x, dout, fv)                                           # 1x1 projection (:24)
        return conv_backward(self.conv1, x, dc1, fv, residual=dx)


class resLayer:
    def __init__(self, filters, filter_in, k_h, k_w, strides=1, dilation_rate=1, *, distortion_aware=True, math_mode=None, device="cuda"):
        self.sequence = list()
        for f_in, f_out in zip([filter_in] + list(filters), filters):           # generator.py:43-44
            self.sequence.append(resBlock(f_in, f_out, k_h=k_h, k_w=k_w, strides=strides, dilation_rate=dilation_rate,
                                          distortion_aware=distortion_aware, math_mode=math_mode, device=device))

    @property
    def trainable_variables(self):
        return [v for unit in self.sequence for v in unit.trainable_variables]

    def owner_list(self):
        return [o for unit in self.sequence for o in unit.owner_list()]

    def build(self, input_shape):
        for unit in self.sequence:
            unit.build(input_shape)
            input_shape = tuple(input_shape[:3]) + (unit.filter_out,)

    def set_weights(self, blocks):
        for unit, w in zip(self.sequence, blocks):
            unit.set_weights(w)

    def call(self, x, save=False):
        for unit in self.sequence:                                              # :46-49
            x = unit(x, save=save)
        return x

    __call__ = call

    def backward(self, dout, fv):
        for unit in reversed(self.sequence):
            dout = unit.backward(dout, fv)
        return dout


class _NormAct:
    """conv -> InstanceNormalization -> leaky_relu(0.1), the pattern of generator.py:94-106 and :112-121, with the moments
    accumulated by the conv epilogue."""

    def __init__(self, conv, device):
        self.conv = conv
        self.norm = InstanceNormalization(device=device)
        self._stats = None

    def __call__(self, x, save=False):
        B = x.shape[0]
        F = self.conv.output_channels
        if self._stats is None or self._stats.shape[0] != B:
            self._stats = torch.empty(B, F, 2, dtype=torch.float64, device=x.device)
        zero_(self._stats)
        y = self.conv(x, stats=self._stats)
        a = self.norm.apply(y, self._stats, leaky_slope=0.1)
        if save:
            # the conv's own input: for ops.deconv2d that is the resized tensor (ops.py:122)
            xin = self.conv._last_resized if isinstance(self.conv, ops.deconv2d) else x
            self._saved = (x, xin, y, a)
        return a

    def owner_list(self):
        return conv_owner_list(self.conv) + [(self.norm, "gamma"), (self.norm, "beta")]

    def backward(self, da, fv, need_dx=True, dx_out=None, dx_accumulate=False):
        """da: gradient w.r.t. the stage output (after the LeakyReLU).  Returns the gradient w.r.t. the stage input; deconv stages can
        write it into dx_out, or (dx_accumulate) add it to what dx_out holds."""
        x, xin, z, a = self._saved
        dz = instnorm_backward(self.norm, z, self._stats, da, a, 0.1, fv)
        dxin = conv_backward(self.conv, xin, dz, fv, need_dx=need_dx)
        if not need_dx:
            return None
        if not isinstance(self.conv, ops.deconv2d):
            return dxin
        B, h, w, C = x.shape
        dx = dx_out if dx_out is not None else torch.empty_like(x)
        check(LIB.sky_resize_bilinear_bwd(dxin.data_ptr(), dx.data_ptr(), B, h, w, C, xin.shape[1], xin.shape[2],
                                          int(bool(dx_accumulate)), _stream()))                          # adjoint of ops.py:122
        return dx


class model:
    """generator.model (generator.py:51-175): encoder -> residual trunk -> sky decoder.  The sun branch
    (sun_decode / sun_rad_estimation / blending, :127-175) depends on sunpose_net, Grad-CAM and sunRadNet, which are not
    part of this round."""

    def __init__(self, batch_size=32, im_height=32, im_width=128, da_kernel_size=3, dilation_rate=1, *, distortion_aware=True,
                 math_mode=None, device="cuda"):
        self.fc_dim = int(im_height * im_width)
        self.im_height, self.im_width = im_height, im_width
        kw = dict(math_mode=math_mode, device=device)
        # sky encode (generator.py:60-67)
        self.conv1_d = ops.conv2d(output_channels=32, k_h=7, k_w=7, strides=1, **kw)
        self.conv2_d = ops.conv2d(output_channels=64, k_h=3, k_w=3, strides=2, **kw)
        self.conv3_d = ops.conv2d(output_channels=128, k_h=3, k_w=3, strides=2, **kw)
        self._enc = [_NormAct(c, device) for c in (self.conv1_d, self.conv2_d, self.conv3_d)]
        self.norm1_d, self.norm2_d, self.norm3_d = (e.norm for e in self._enc)
        self.res = resLayer((128,) * 6, 128, k_h=da_kernel_size, k_w=da_kernel_size, strides=1, dilation_rate=dilation_rate,
                            distortion_aware=distortion_aware, **kw)
        # sky_decode (generator.py:69-76)
        self.conv3_f = ops.deconv2d(output_channels=64, output_imshape=[int(im_height / 2), int(im_width / 2)], k_h=3, k_w=3,
                                    method='resize', **kw)
        self.conv2_f = ops.deconv2d(output_channels=32, output_imshape=[int(im_height), int(im_width)], k_h=3, k_w=3,
                                    method='resize', **kw)
        self._dec = [_NormAct(c, device) for c in (self.conv3_f, self.conv2_f)]
        self.norm3_f, self.norm2_f = (d.norm for d in self._dec)
        self.conv1_f = ops.conv2d(output_channels=3, k_h=7, k_w=7, strides=1, **kw)
        # sun_decode (generator.py:78-85)
        self.conv3_u = ops.deconv2d(output_channels=64, output_imshape=[int(im_height / 2), int(im_width / 2)], k_h=3, k_w=3,
                                    method='resize', **kw)
        self.conv2_u = ops.deconv2d(output_channels=32, output_imshape=[int(im_height), int(im_width)], k_h=3, k_w=3,
                                    method='resize', **kw)
        self._dec_u = [_NormAct(c, device) for c in (self.conv3_u, self.conv2_u)]
        self.norm3_u, self.norm2_u = (d.norm for d in self._dec_u)
        self.conv1_u = ops.conv2d(output_channels=3, k_h=7, k_w=7, strides=1, **kw)
        # enhanceSunRadiance (generator.py:88)
        self.sun = sunRadNet(**kw)
        self._gmax = None
        self._side = None

    def encode(self, x, training="training", save=False):
        for stage in self._enc:                       # generator.py:94-106
            x = stage(x, save=save)
        return self.res(x, save=save)                 # :108

    # ---- training direction (train.generator_in_step with training=True, train.py:239-349) ----------------------------------------
    def owner_list(self):
        """(object, attribute) of every trainable variable of the generator (order of the flat buffer)."""
        out = []
        for st in self._enc:
            out += st.owner_list()
        out += self.res.owner_list()
        for st in self._dec:
            out += st.owner_list()
        out += conv_owner_list(self.conv1_f)
        for st in self._dec_u:
            out += st.owner_list()
        out += conv_owner_list(self.conv1_u)
        for d in (self.sun.d1, self.sun.d2, self.sun.d3, self.sun.d4):
            out += [(d, "kernel")] + ([(d, "gamma"), (d, "beta")] if d.apply_norm else [])
        out += [(self.sun, "gb_kernel"), (self.sun, "gb_bias")]
        return out

    def decode_train(self, x, stages, last):
        """Decoder up to the raw output of its 7x7 conv (bias included, no activation): the tail kernel applies the rest."""
        for stage in stages:
            x = stage(x, save=True)
        return last(x), x

    def decode_backward(self, dc, stages, last, last_in, fv, dres, accumulate):
        """dc: gradient w.r.t. the raw 7x7 conv output; writes (accumulate=False) or adds the gradient w.r.t. the trunk output into dres."""
        g = conv_backward(last, last_in, dc, fv)                              # conv1_f / conv1_u (generator.py:120, 150)
        g = stages[1].backward(g, fv)                                         # conv2_* + IN + lrelu (+ resize adjoint)
        stages[0].backward(g, fv, dx_out=dres, dx_accumulate=accumulate)      # conv3_*

    def encode_backward(self, dres, fv):
        g = self.res.backward(dres, fv)
        g = self._enc[2].backward(g, fv)
        g = self._enc[1].backward(g, fv)
        self._enc[0].backward(g, fv, need_dx=False)

    def sky_decode(self, x, _input, training="training", log_decompress=False):
        for stage in self._dec:                       # generator.py:112-118
            x = stage(x)
        # :120-124  leaky_relu(conv1_f) + input -> relu, all in the conv epilogue; optionally followed by
        # tf_utils.hdr_logDecompression (inference.py:86)
        return self.conv1_f(x, leaky_slope=0.1, residual=_input, relu=True, log_decompress=log_decompress)

    def sky_inference(self, ldr):
        """inference.py:84-86: hdr_logDecompression(sky_decode(encode(ldr), ldr)) — the linear-radiance sky prediction."""
        return self.sky_decode(self.encode(ldr), ldr, log_decompress=True)

    def sun_decode(self, x, sun_cam1, sun_cam2, sun_cam3, sun_rad, training="training", *, blend_with=None, threshold=0.12,
                   log_decompress=False):
        """generator.py:127-156.  The CAM skip connections are commented out in the reference (:130-149), so the maps are unused
        here too.  `blend_with` = the sky prediction in the log domain: the alpha blend of inference.py:90-92,106-109 (and, with
        log_decompress, the final hdr_logDecompression of :110) then happen in conv1_u's epilogue."""
        for stage in self._dec_u:                     # :136-147
            x = stage(x)
        blend = None if blend_with is None else (blend_with, threshold)
        return self.conv1_u(x, leaky_slope=0.1, residual=sun_rad, relu=True, log_decompress=log_decompress, blend=blend)   # :150-156

    def sun_rad_estimation(self, jpeg_img_float, sun_cam1, sun_cam2, sun_cam3, sunpose_pred, training="training", *,
                           log_compress=False):
        """generator.py:158-169 -> (sun_rad_t [B,H,W,3], gamma, beta).  reduce_max (:160), both resizes (:161-162) and the concat
        (:164) are two small kernels in front of sunRadNet; log_compress=True returns hdr_logCompression(sun_rad_t)
        (inference.py:105) straight from the radiance kernel.  The maximum is over the tensor this process holds (its batch shard)."""
        ldr = _require_cuda(jpeg_img_float, "jpeg_img_float")
        B, H, W, _ = ldr.shape
        sm = _require_cuda(sunpose_pred, "sunpose_pred").reshape(B, H, W, 1)
        if self._gmax is None:
            self._gmax = torch.zeros(1, dtype=torch.float32, device=ldr.device)
        check(LIB.sky_max_nonneg(sm.data_ptr(), self._gmax.data_ptr(), sm.numel(), _stream()))            # :160
        plz = torch.empty((B, H, W, 8), dtype=torch.float32, device=ldr.device)                            # 6 channels + 2 of padding
        c1, c2, c3 = (_require_cuda(c, "sun_cam") for c in (sun_cam1, sun_cam2, sun_cam3))
        check(LIB.sky_sunrad_input(ldr.data_ptr(), c1.data_ptr(), c2.data_ptr(), c3.data_ptr(), plz.data_ptr(), B, H, W,
                                   c2.shape[1], c2.shape[2], c3.shape[1], c3.shape[2], 8, _stream()))     # :161-164
        out, gamma, beta = self.sun(sm, plz, training, x_max=self._gmax, log_compress_tiled=log_compress)  # :165
        if not log_compress:
            out = out.expand(B, H, W, 3).contiguous()                                                      # :167
        return out, gamma, beta

    def blending(self, sky_pred, sun_pred, training="training"):
        """generator.py:171-175: tf.add_n.  (generator_inference fuses it into conv1_u's epilogue instead.)"""
        return sky_pred + sun_pred

    def generator_inference(self, ldr, sun_model, threshold=0.12):
        """inference.generator_in_step (inference.py:81-112): LDR panorama -> linear HDR panorama."""
        from . import grad_cam
        # The sky branch (encode -> trunk -> sky_decode) and the sun branch up to the sun radiance (sun-position network -> Grad-CAM ->
        # sunRadNet) only share the input: the sun branch runs on a side stream (fork / join; also valid inside a CUDA-graph capture)
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            sunpose_cmf, (sunlayer1, sunlayer2, sunlayer3) = sun_model.sunposeEstimation(ldr, training=False)   # :87
            y_c = sun_model.class_score(sunpose_cmf)                                        # :98
            sun_cam1 = grad_cam.layer(y_c, sunlayer1)                                       # :100-102
            sun_cam2 = grad_cam.layer(y_c, sunlayer2)
            sun_cam3 = grad_cam.layer(y_c, sunlayer3)
            sun_rad_gamma, _, _ = self.sun_rad_estimation(ldr, sun_cam1, sun_cam2, sun_cam3, sunpose_cmf, training=False,
                                                          log_compress=True)                # :104-105
        res_out = self.encode(ldr, training=False)                                          # :83
        sky_pred_gamma = self.sky_decode(res_out, ldr, training=False)                      # :84 (:85 is folded into the blend)
        main.wait_stream(self._side)
        return self.sun_decode(res_out, sun_cam1, sun_cam2, sun_cam3, sun_rad_gamma, training=False, blend_with=sky_pred_gamma,
                               threshold=threshold, log_decompress=True)                    # :106-110

    def build(self, batch_size=None):
        """Create every variable by tracing shapes (the reference builds lazily on the first call)."""
        B = batch_size or 1
        H, W = self.im_height, self.im_width
        shapes = [(B, H, W, 3), (B, H, W, 32), (B, H // 2, W // 2, 64)]
        for st, shp in zip(self._enc, shapes):
            st.conv.build(shp)
            st.norm.build(shp[:3] + (st.conv.output_channels,))
        self.res.build((B, H // 4, W // 4, 128))
        for st, shp in zip(self._dec, [(B, H // 4, W // 4, 128), (B, H // 2, W // 2, 64)]):
            st.conv.build(shp)
            st.norm.build(shp[:3] + (st.conv.output_channels,))
        self.conv1_f.build((B, H, W, 32))
        for st, shp in zip(self._dec_u, [(B, H // 4, W // 4, 128), (B, H // 2, W // 2, 64)]):
            st.conv.build(shp)
            st.norm.build(shp[:3] + (st.conv.output_channels,))
        self.conv1_u.build((B, H, W, 32))
        self.sun.build(B, H, W)

    def set_weights(self, w):
        """w: dict keyed by the reference's attribute names: conv1_d/conv2_d/conv3_d/conv3_f/conv2_f/conv1_f ->
        (kernel [k,k,C,F], bias), norm*_d / norm*_f -> (gamma, beta), res -> list of res-block dicts."""
        for name in ("conv1_d", "conv2_d", "conv3_d", "conv1_f", "conv1_u"):
            if name not in w:
                continue
            layer = getattr(self, name)
            layer.w.copy_(torch.as_tensor(w[name][0]))
            layer.biases.copy_(torch.as_tensor(w[name][1]))
        for name in ("conv3_f", "conv2_f", "conv3_u", "conv2_u"):
            if name not in w:
                continue
            layer = getattr(self, name)
            layer.kernel.copy_(torch.as_tensor(w[name][0]))
            layer.biases.copy_(torch.as_tensor(w[name][1]))
        for name in ("norm1_d", "norm2_d", "norm3_d", "norm3_f", "norm2_f", "norm3_u", "norm2_u"):
            if name not in w:
                continue
            norm = getattr(self, name)
            norm.gamma.copy_(torch.as_tensor(w[name][0]))
            norm.beta.copy_(torch.as_tensor(w[name][1]))
        self.res.set_weights(w["res"])
        if "sun" in w:
            # sunRadNet: d1..d4 -> dict(kernel [4,4,C,F] (, gamma, beta, moving_mean, moving_variance)); gamma / beta -> (kernel [flat,1], bias [1])
            for name in ("d1", "d2", "d3", "d4"):
                d, src = getattr(self.sun, name), w["sun"][name]
                for key, val in src.items():
                    getattr(d, key).copy_(torch.as_tensor(val))
            for head in ("gamma", "beta"):
                getattr(self.sun, head + "_kernel").copy_(torch.as_tensor(w["sun"][head][0]))
                getattr(self.sun, head + "_bias").copy_(torch.as_tensor(w["sun"][head][1]))
