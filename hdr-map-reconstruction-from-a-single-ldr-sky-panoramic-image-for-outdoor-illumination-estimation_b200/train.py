"""Host-side mirror of the compute steps of the reference's train.py:

    _preprocessing(hdr, crf, t, ...)        train.py:54-94    -> tf_utils.ldr_synth (one kernel)
    generator_in_step(args, training)       train.py:239-349  generator + sun-position net + Grad-CAM + sunRadNet + blend, then every
                                                               loss term: KL, perceptual (VGG16 pool1-3), DoG, L1, LSGAN; total (:331)
    discriminator_in_step(args, training)   train.py:351-380
    train_step(ds, sunpose_gt)              train.py:382-415  both tapes: d total_gen_loss / d (_gen + _sun variables), d total_disc_loss /
                                                               d (_dis variables), Keras RMSprop on each set
    test_step(ds, sunpose_gt)               train.py:417-442

``train_step`` is the reference's data flow with the autodiff written out: the forward pass keeps what the backward pass reads, the
loss adjoints (perceptual through the frozen VGG16, LSGAN through the discriminator in inference mode, DoG, L1) meet in one tail kernel
(``sky_train_tail_bwd``), and the gradients of every variable land in two flat buffers (generator + sun-position network, discriminator)
that data-parallel training all-reduces — the Dense bucket of the sun-position network as soon as it is complete, under the rest of the
backward pass — before two fused RMSprop launches.  Cross-sample couplings under batch sharding: BatchNormalization
statistics and tf.reduce_max(sunpose_pred) are per replica; loss adjoints are normalised by the GLOBAL batch so the summed gradient is
the global-batch mean.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import discriminator as _discriminator
from . import generator as _generator
from . import grad_cam, sunpose_net as _sunpose_net, tf_utils, vgg16 as _vgg16
from ._flat import FlatVars
from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream, zero_
from .train_sun import sunpose_backward, sunpose_owner_list
from .trunk_train import allreduce_flat_

IMSHAPE = (32, 128, 3)       # train.py:31
THRESHOLD = 0.12             # train.py:248
LEARNING_RATE = 1e-4         # train.py:34


class Step:
    """Owns the models of train.py:183-199 (_gen, _sun, _dis, vgg, vgg2), the two optimizers (:201-202) and runs the step functions."""

    def __init__(self, batch_size=32, im_height=IMSHAPE[0], im_width=IMSHAPE[1], *, vgg_data_dict=None, distortion_aware_sunpose=True,
                 distortion_aware=True, math_mode=None, device="cuda", lr=LEARNING_RATE, rho=0.9, eps=1e-7, concurrent=True):
        kw = dict(math_mode=math_mode, device=device)
        self.concurrent = bool(concurrent)   # train_step forks its independent parts onto two side streams
        self.H, self.W = im_height, im_width
        self._gen = _generator.model(batch_size=batch_size, im_height=im_height, im_width=im_width, distortion_aware=distortion_aware, **kw)
        self._sun = _sunpose_net.model(im_height=im_height, im_width=im_width, distortion_aware=distortion_aware_sunpose, **kw)
        self._dis = _discriminator.model(im_height=im_height, im_width=im_width, **kw)
        self.vgg = _vgg16.Vgg16(data_dict=vgg_data_dict, **kw)
        self.vgg2 = self.vgg                 # train.py:198-199 builds two instances from the same file: identical constants
        self._gen.build(batch_size)
        self._dis.build(batch_size, im_height, im_width)
        self._acc = torch.zeros(16, dtype=torch.float64, device=device)
        self.lr, self.rho, self.eps = lr, rho, eps
        self._marks = None
        self.fv_gen = self.fv_dis = None
        self._side = None
        self.last_losses = {}

    @staticmethod
    def _preprocessing(hdr, crf, t, sigma_s=None, sigma_c=None, noise_s=None, noise_c=None):
        """train._preprocessing (train.py:54-94); the JPEG round trip of :87-90 (a CPU codec) is omitted, SURVEY 8d."""
        return tf_utils.ldr_synth(hdr, t, crf, sigma_s, sigma_c, noise_s, noise_c, quantize=True)

    # ------------------------------------------------------------------------------------------------------------------------------
    # evaluation direction (training=False)
    # ------------------------------------------------------------------------------------------------------------------------------
    def generator_in_step(self, args, training=False):
        if training:
            raise ValueError("the training direction is train_step(): forward, both backward passes and the optimizers in one call")
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
            raise ValueError("the training direction is train_step(): forward, both backward passes and the optimizers in one call")
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

    # ------------------------------------------------------------------------------------------------------------------------------
    # training direction (train.py:382-415)
    # ------------------------------------------------------------------------------------------------------------------------------
    def init_training(self, batch_size):
        """Creates every lazily built variable and moves the two variable sets of train.py:402,405 into flat buffers with their
        RMSprop accumulators (train.py:201-202)."""
        if self.fv_gen is not None:
            return
        dev = self._gen.conv1_d.device
        x0 = torch.zeros((max(int(batch_size), 1), self.H, self.W, 3), dtype=torch.float32, device=dev)
        self._sun.sunposeEstimation(x0)
        conv_owners, fc_owners = sunpose_owner_list(self._sun)
        self.fv_gen = FlatVars(self._gen.owner_list() + conv_owners + fc_owners, dev)          # the Dense variables last: one early bucket
        self._fc_offset = self.fv_gen.offset(self._sun.fc1, "kernel")
        self.fv_dis = FlatVars(self._dis.owner_list(), dev)
        self.fv_gen.zeroed_each_step = True      # train_step zeroes flat_g[:fc_offset] before the backward pass: conv gradients accumulate
        self.fv_dis.zeroed_each_step = True      # and fv_dis.flat_g before the discriminator's
        self.ms_gen = torch.zeros_like(self.fv_gen.flat_w)
        self.ms_dis = torch.zeros_like(self.fv_dis.flat_w)

    def _sun_grads(self):
        fv, sun = self.fv_gen, self._gen.sun
        g = {name: {k: fv.grad(getattr(sun, name), k) for k in (("kernel", "gamma", "beta") if getattr(sun, name).apply_norm else ("kernel",))}
             for name in ("d1", "d2", "d3", "d4")}
        g["gb_kernel"], g["gb_bias"] = fv.grad(sun, "gb_kernel"), fv.grad(sun, "gb_bias")
        return g

    def train_step(self, ds, sunpose_gt, global_batch=None):
        """train.train_step (train.py:382-415) on this rank's shard -> [y_final_lin, sky_pred_lin, sun_pred_lin, alpha_c3, sunpose_pred,
        sun_cam1, sun_cam2, sun_cam3, sun_rad_lin(log domain x3), g_out, b_out]; losses in self.last_losses (fp64 device scalars)."""
        hdr_t, ldr = (_require_cuda(a, "ds") for a in ds)                                        # :385-387 (callers pass BGR)
        gt = _require_cuda(sunpose_gt, "sunpose_gt")
        B, H, W, _ = ldr.shape
        self.init_training(B)
        gen, sun, dis, fv, fvd = self._gen, self._sun, self._dis, self.fv_gen, self.fv_dis
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        Bg = int(global_batch) if global_batch is not None else B * world
        st = _stream()
        npix = B * H * W
        acc = zero_(self._acc)
        dev = ldr.device
        # Independent parts of the step run on side streams (fork / join with stream waits, which a CUDA-graph capture records as graph
        # edges): most launches of the step fill a fraction of the 148 SMs (the trunk convs at 32x128, B = 32: 64 tiles), so whatever does
        # not depend on them overlaps.  s1: the sun branch (forward) / the DoG term / the sun-position backward; s2: the adversarial term,
        # then the whole discriminator step.  `concurrent = False` keeps everything on the caller's stream.
        main = torch.cuda.current_stream()
        if self.concurrent and self._side is None:
            self._side = (torch.cuda.Stream(), torch.cuda.Stream())
        s1, s2 = self._side if self.concurrent else (main, main)

        marks = self._marks          # debug (tools/step_phases.py): events at the phase boundaries of every stream; None in production

        def mark(name, stream=None):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream if stream is not None else torch.cuda.current_stream())
                marks.append((name, ev))

        def fork(side):
            if side is not main:
                side.wait_stream(main)

        def join(side):
            if side is not main:
                main.wait_stream(side)

        # ================================ generator_in_step(training=True), forward (:239-331) ================================
        both = torch.empty((2 * B, H, W, 3), dtype=torch.float32, device=dev)                    # [y_final_gamma; hdr_t_gamma]: VGG16's batch
        y_gamma = both[:B]
        mark("start")
        fork(s1)
        with torch.cuda.stream(s1):
            sm, acts = sun.sunposeEstimation(ldr, training=True)                                 # :253
            mark("s1: sunpose forward done")
            sun.fc1.kernel_transposed()
            sun.fc2.kernel_transposed()
            y_c = sun.class_score(sm, gt)                                                        # :263-265 (outside the tape)
            cams = [grad_cam.layer(y_c, a) for a in acts]                                        # :267-269
            mark("s1: grad-cam done")
            sun_rad_gamma, _, _ = gen.sun_rad_estimation(ldr, cams[0], cams[1], cams[2], sm, training=True, log_compress=True)   # :286-287
            check(LIB.sky_kl_divergence(gt.data_ptr(), sm.data_ptr(), gt.numel(), acc[0:1].data_ptr(), _stream()))              # :303
            g_sm = torch.empty_like(sm)
            check(LIB.sky_kl_divergence_bwd(gt.data_ptr(), sm.data_ptr(), g_sm.data_ptr(), sm.numel(), 1.0 / Bg, 0, _stream()))
            mark("s1: sun branch forward done")
        check(LIB.sky_hdr_log_codec(hdr_t.data_ptr(), both[B:].data_ptr(), hdr_t.numel(), 0, st))                 # :246
        res_out = gen.encode(ldr, training=True, save=True)                                      # :249
        mark("main: encoder + trunk forward done")
        c_sky, sky_in = gen.decode_train(res_out, gen._dec, gen.conv1_f)                         # :250 up to conv1_f's raw output
        c_sun, sun_in = gen.decode_train(res_out, gen._dec_u, gen.conv1_u)                       # :288 up to conv1_u's raw output
        y_lin = torch.empty_like(c_sky)
        sky_lin, sun_lin = torch.empty_like(c_sky), torch.empty_like(c_sky)
        alpha = torch.empty((B, H, W, 1), dtype=torch.float32, device=dev)
        mark("main: decoders forward done")
        join(s1)
        check(LIB.sky_train_tail_fwd(c_sky.data_ptr(), c_sun.data_ptr(), ldr.data_ptr(), sun_rad_gamma.data_ptr(), hdr_t.data_ptr(), THRESHOLD,
                                     0.1, y_gamma.data_ptr(), y_lin.data_ptr(), sky_lin.data_ptr(), sun_lin.data_ptr(), alpha.data_ptr(),
                                     acc[8:9].data_ptr(), npix, st))                             # :251, 256-259, 289-298, 324
        # ---- loss terms with their adjoints ----
        fork(s1)
        fork(s2)
        with torch.cuda.stream(s1):
            # DoG (:315-321)
            base_y = torch.empty((B, 2 * H, 2 * W, 3), dtype=torch.float32, device=dev)
            base_t = torch.empty_like(base_y)
            check(LIB.sky_dog_base(y_lin.data_ptr(), base_y.data_ptr(), B, H, W, 3, _stream()))
            check(LIB.sky_dog_base(hdr_t.data_ptr(), base_t.data_ptr(), B, H, W, 3, _stream()))
            check(LIB.sky_dog_l1(base_y.data_ptr(), base_t.data_ptr(), B, 2 * H, 2 * W, 3, acc[4:8].data_ptr(), _stream()))
            dbase = torch.empty_like(base_y)
            check(LIB.sky_dog_l1_bwd(base_y.data_ptr(), base_t.data_ptr(), dbase.data_ptr(), B, 2 * H, 2 * W, 3,
                                     1000.0 / (base_y.numel() // B * Bg), _stream()))
            g_dog = torch.empty_like(y_lin)
            check(LIB.sky_dog_base_bwd(dbase.data_ptr(), g_dog.data_ptr(), B, H, W, 3, 0, _stream()))
            mark("s1: DoG done")
        with torch.cuda.stream(s2):
            # adversarial (:300, 327): the discriminator in inference mode on [ldr, y_final_lin], LSGAN gen_loss, data gradient back to y_final_lin
            cat16 = torch.empty((2 * B, H, W, 8), dtype=torch.float32, device=dev)
            dis.concat(ldr, hdr_t, out=cat16[:B])
            dis.concat(ldr, y_lin, out=cat16[B:])
            d_gen = dis.infer_forward(cat16[B:])
            hh, ww = d_gen.shape[1], d_gen.shape[2]
            n_d = B * (hh - 3) * (ww - 3)
            g_same = torch.empty_like(d_gen)
            check(LIB.sky_lsgan_bwd(d_gen.data_ptr(), g_same.data_ptr(), acc[9:10].data_ptr(), B, hh, ww, 1, hh - 2, 1, ww - 2, 1.0,
                                    1.0 / (n_d // B * Bg), _stream()))
            g_dis8 = dis.infer_backward_data(g_same)
            adv_done = torch.cuda.Event()
            adv_done.record(torch.cuda.current_stream())
            mark("s2: adversarial term done")
            # ============================ discriminator_in_step(training=True) + disc_tape.gradient (:351-380, 405) ============================
            # needs nothing of the generator's backward pass: it stays on this stream until the optimizers
            d_both = dis.train_forward(cat16, groups=2)                                          # :360-361, batch statistics per call
            g_both = torch.empty_like(d_both)
            check(LIB.sky_lsgan_bwd(d_both[:B].data_ptr(), g_both[:B].data_ptr(), acc[10:11].data_ptr(), B, hh, ww, 1, hh - 2, 1, ww - 2, 1.0,
                                    0.5 / (n_d // B * Bg), _stream()))                           # real_loss (:236)
            check(LIB.sky_lsgan_bwd(d_both[B:].data_ptr(), g_both[B:].data_ptr(), acc[11:12].data_ptr(), B, hh, ww, 1, hh - 2, 1, ww - 2, 0.0,
                                    0.5 / (n_d // B * Bg), _stream()))                           # generated_loss (:237)
            zero_(fvd.flat_g)
            dis.train_backward(g_both, fvd)
            mark("s2: discriminator step done")
        # perceptual (:306-312): VGG16 features of [prediction; target] in one batch, data gradient for the prediction half
        self.vgg.forward_saved(both)
        g_vgg4 = self.vgg.perceptual_backward(B, 0.01 * B / Bg, acc[1:4])
        mark("main: VGG forward + backward done")
        join(s1)
        if s2 is not main:
            main.wait_event(adv_done)
        # ---- the tail's adjoint: all four terms meet here (:331: total = kl + 1000 DoG + gen + 10 L1 + 0.01 perceptual) ----
        dc_sky, dc_sun, d_srg = torch.empty_like(c_sky), torch.empty_like(c_sky), torch.empty_like(c_sky)
        check(LIB.sky_train_tail_bwd(c_sky.data_ptr(), c_sun.data_ptr(), ldr.data_ptr(), sun_rad_gamma.data_ptr(), alpha.data_ptr(), y_lin.data_ptr(),
                                     hdr_t.data_ptr(), g_dog.data_ptr(), g_dis8.data_ptr(), g_vgg4.data_ptr(), 10.0 / (3 * npix // B * Bg), 255.0,
                                     0.1, dc_sky.data_ptr(), dc_sun.data_ptr(), d_srg.data_ptr(), npix, st))
        # ================================ gen_tape.gradient (:402) ================================
        zero_(fv.flat_g[:self._fc_offset])             # d gamma / d beta / atomically reduced kernels accumulate; the Dense gradients are overwritten
        work = []
        fork(s1)
        with torch.cuda.stream(s1):
            # sun radiance -> sun-position softmax (the max-normalisation of generator.py:160 included), joined with the KL adjoint
            gen.sun.train_backward(d_srg, self._sun_grads(), g_sm, accumulate_dsm=True, accumulate_dw=True)
            sunpose_backward(sun, g_sm, fv.grad, on_dense_done=lambda: work.append(self._start_tail_allreduce()), accumulate=True)
            mark("s1: sun branch backward done")
        mark("main: tail backward done")
        dres = torch.empty_like(res_out)
        gen.decode_backward(dc_sun, gen._dec_u, gen.conv1_u, sun_in, fv, dres, False)            # sun decoder (generator.py:127-156)
        gen.decode_backward(dc_sky, gen._dec, gen.conv1_f, sky_in, fv, dres, True)               # sky decoder (:110-125)
        mark("main: decoders backward done")
        gen.encode_backward(dres, fv)                                                            # trunk + encoder (:92-108)
        mark("main: trunk + encoder backward done")
        join(s1)
        join(s2)
        # ================================ optimizers (:403, 406) ================================
        self._finish_allreduce(work[0])
        self.apply_gradients()
        mark("main: optimizers done")
        kl = acc[0:1] / B
        perceptual = acc[1] / self.vgg._saved[3][:B].numel() + acc[2] / self.vgg._saved[6][:B].numel() + acc[3] / self.vgg._saved[10][:B].numel()
        dog = acc[4:8].sum() / base_y.numel()
        l1 = acc[8] / (3 * npix)
        gen_l = acc[9] / n_d
        total = kl + 1000.0 * dog + gen_l + 10.0 * l1 + 0.01 * perceptual
        disc = 0.5 * (acc[11] / n_d + acc[10] / n_d)
        self.last_losses = dict(total=total, kl=kl, perceptual=perceptual, dog=dog, l1=l1, gen=gen_l, disc=disc)
        gb = gen.sun._saved["gb"]
        return [y_lin, sky_lin, sun_lin, alpha.expand(B, H, W, 3), sm.reshape(B, H, W, 1), cams[0], cams[1], cams[2], sun_rad_gamma,
                torch.sigmoid(gb[:, 0]).max(), torch.sigmoid(gb[:, 1]).max()]                    # :411-415

    # ---- data-parallel gradient exchange: one all-reduce of the flat gradients per step, as two buckets -------------------------------
    def _start_tail_allreduce(self):
        """The Dense gradients of the sun-position network (99 % of the bytes) are the first ones the backward pass completes: their
        all-reduce starts here and runs under the rest of the backward pass."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.fv_gen.flat_g[self._fc_offset:], op=dist.ReduceOp.SUM, async_op=True)
        return None

    def _finish_allreduce(self, work):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            allreduce_flat_(self.fv_gen.flat_g[:self._fc_offset])
            allreduce_flat_(self.fv_dis.flat_g)
            if work is not None:
                work.wait()

    def apply_gradients(self):
        """Keras RMSprop (train.py:201-202, 403, 406) over the two flat buffers; the gradients are already global-batch means."""
        for fvx, ms in ((self.fv_gen, self.ms_gen), (self.fv_dis, self.ms_dis)):
            check(LIB.sky_rmsprop_step(fvx.flat_w.data_ptr(), ms.data_ptr(), fvx.flat_g.data_ptr(), fvx.flat_w.numel(), self.lr, self.rho, self.eps,
                                       1.0, _stream()))
            fvx.invalidate()
