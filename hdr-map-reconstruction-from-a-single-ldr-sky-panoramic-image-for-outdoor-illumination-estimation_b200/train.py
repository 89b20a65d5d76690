"""Host-side mirror of the compute steps of the reference's train.py that are built so far — the forward (evaluation) direction:

    _preprocessing(hdr, crf, t, ...)        train.py:54-94    -> tf_utils.ldr_synth (one kernel)
    generator_in_step(args, training=False) train.py:239-349  generator + sun-position net + Grad-CAM + sunRadNet + blend, then every
                                                               loss term: KL, perceptual (VGG16 pool1-3), DoG, L1, LSGAN; total (:331)
    discriminator_in_step(args, training=False)  train.py:351-380
    test_step(ds, sunpose_gt)               train.py:417-442

The backward direction of the full step (two GradientTapes, RMSprop on G u Sun and on D, train.py:382-415) exists only for the
distortion-aware residual trunk (trunk_train.py); training=True raises here.
"""
from __future__ import annotations

import torch

from . import discriminator as _discriminator
from . import generator as _generator
from . import grad_cam, sunpose_net as _sunpose_net, tf_utils, vgg16 as _vgg16
from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream

IMSHAPE = (32, 128, 3)       # train.py:31
THRESHOLD = 0.12             # train.py:248


class Step:
    """Owns the models of train.py:183-199 (_gen, _sun, _dis, vgg, vgg2) and runs the step functions."""

    def __init__(self, batch_size=32, im_height=IMSHAPE[0], im_width=IMSHAPE[1], *, vgg_data_dict=None, distortion_aware_sunpose=True,
                 math_mode=None, device="cuda"):
        kw = dict(math_mode=math_mode, device=device)
        self._gen = _generator.model(batch_size=batch_size, im_height=im_height, im_width=im_width, **kw)
        self._sun = _sunpose_net.model(im_height=im_height, im_width=im_width, distortion_aware=distortion_aware_sunpose, **kw)
        self._dis = _discriminator.model(im_height=im_height, im_width=im_width, **kw)
        self.vgg = _vgg16.Vgg16(data_dict=vgg_data_dict, **kw)
        self.vgg2 = self.vgg                 # train.py:198-199 builds two instances from the same file: identical constants
        self._gen.build(batch_size)
        self._dis.build(batch_size, im_height, im_width)
        self._acc = torch.zeros(16, dtype=torch.float64, device=device)

    @staticmethod
    def _preprocessing(hdr, crf, t, sigma_s=None, sigma_c=None, noise_s=None, noise_c=None):
        return tf_utils.ldr_synth(hdr, t, crf, sigma_s, sigma_c, noise_s, noise_c, quantize=True)

    def generator_in_step(self, args, training=False):
        if training:
            raise NotImplementedError("the backward of the full generator step is not built (DESIGN.md section 7)")
        jpeg_img_float, hdr_t, sunpose_gt = (_require_cuda(a, "arg") for a in args)
        B, H, W, _ = jpeg_img_float.shape
        acc = self._acc.zero_()
        hdr_t_gamma = tf_utils.hdr_logCompression(hdr_t)                                                      # :246
        res_out = self._gen.encode(jpeg_img_float, training=False)                                            # :249
        sky_gamma = self._gen.sky_decode(res_out, jpeg_img_float, training=False)                             # :250
        sunpose_cmf, (sunlayer1, sunlayer2, sunlayer3) = self._sun.sunposeEstimation(jpeg_img_float, training=False)   # :253
        y_c = self._sun.class_score(sunpose_cmf, sunpose_gt)                                                  # :278-280
        sun_cam1 = grad_cam.layer(y_c, sunlayer1)                                                             # :282-284
        sun_cam2 = grad_cam.layer(y_c, sunlayer2)
        sun_cam3 = grad_cam.layer(y_c, sunlayer3)
        sun_rad_gamma, gamma, beta = self._gen.sun_rad_estimation(jpeg_img_float, sun_cam1, sun_cam2, sun_cam3, sunpose_cmf,
                                                                  training=False, log_compress=True)          # :286-287
        sun_gamma = self._gen.sun_decode(res_out, sun_cam1, sun_cam2, sun_cam3, sun_rad_gamma, training=False)   # :288
        y_final_gamma, y_final_lin = torch.empty_like(sky_gamma), torch.empty_like(sky_gamma)
        sky_pred_lin, sun_pred_lin = torch.empty_like(sky_gamma), torch.empty_like(sky_gamma)
        alpha = torch.empty((B, H, W, 1), dtype=torch.float32, device=sky_gamma.device)
        check(LIB.sky_blend_split(sky_gamma.data_ptr(), sun_gamma.data_ptr(), THRESHOLD, y_final_gamma.data_ptr(), y_final_lin.data_ptr(),
                                  sky_pred_lin.data_ptr(), sun_pred_lin.data_ptr(), alpha.data_ptr(), B * H * W, _stream()))   # :272-298
        disc_generated_output = self._dis([jpeg_img_float, y_final_lin], training=False)                      # :300
        sun_loss = tf_utils.kl_divergence(sunpose_gt, sunpose_cmf, acc[0:1])                                  # :303
        perceptual_loss = 0
        for i, (a, b) in enumerate(zip(self.vgg(y_final_gamma), self.vgg2(hdr_t_gamma))):                     # :306-312
            perceptual_loss = perceptual_loss + tf_utils.reduce_mean_abs_diff(a, b, acc[1 + i:2 + i])
        DoG_loss = tf_utils.DoG_l1(y_final_lin, hdr_t, acc[4:8])                                              # :315-321
        l1_loss = tf_utils.reduce_mean_abs_diff(y_final_lin, hdr_t, acc[8:9])                                 # :324
        d = disc_generated_output.contiguous()
        check(LIB.sky_loss_reduce(1, d.data_ptr(), None, d.numel(), acc[9:10].data_ptr(), _stream()))         # :327 gen_loss (:235)
        _gen_loss = acc[9:10] / d.numel()
        total_gen_loss = sun_loss + 1000. * DoG_loss + _gen_loss + 10. * l1_loss + 0.01 * perceptual_loss     # :330
        self.last_losses = dict(total=total_gen_loss, kl=sun_loss, perceptual=perceptual_loss, dog=DoG_loss, l1=l1_loss, gen=_gen_loss)
        alpha_c3 = alpha.expand(B, H, W, 3)
        sunpose_pred = sunpose_cmf.reshape(B, H, W, 1)
        return [total_gen_loss, y_final_gamma, sky_pred_lin, sun_pred_lin, gamma, beta, alpha_c3, sunpose_pred, sun_cam1, sun_cam2,
                sun_cam3, sun_rad_gamma, y_final_lin]

    def discriminator_in_step(self, args, training=False):
        if training:
            raise NotImplementedError("discriminator training (batch statistics + backward) is not built")
        jpeg_img_float, hdr_t, y_final_lin = args
        acc = torch.zeros(2, dtype=torch.float64, device=jpeg_img_float.device)
        real = self._dis([jpeg_img_float, hdr_t], training=False).contiguous()                                # :360
        fake = self._dis([jpeg_img_float, y_final_lin], training=False).contiguous()                          # :361
        check(LIB.sky_loss_reduce(1, real.data_ptr(), None, real.numel(), acc[0:1].data_ptr(), _stream()))    # real_loss (:236)
        check(LIB.sky_loss_reduce(2, fake.data_ptr(), None, fake.numel(), acc[1:2].data_ptr(), _stream()))    # generated_loss (:237)
        return 0.5 * (acc[1] / fake.numel() + acc[0] / real.numel())                                          # :368

    def test_step(self, ds, sunpose_gt):
        hdr_t, jpeg_img_float = ds                      # train.py:420-422 (rgb2bgr is a channel flip of the inputs; callers pass BGR)
        gen_pred = self.generator_in_step([jpeg_img_float, hdr_t, sunpose_gt], training=False)
        y_final_lin = gen_pred[-1]                      # == hdr_logDecompression(y_final_gamma) (:427)
        total_disc_loss = self.discriminator_in_step([jpeg_img_float, hdr_t, y_final_lin], training=False)
        return gen_pred, total_disc_loss
