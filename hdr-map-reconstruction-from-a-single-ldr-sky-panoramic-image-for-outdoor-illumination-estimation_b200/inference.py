"""Host-side mirror of the compute step of the reference's inference.py (``inference`` -> ``generator_in_step``,
inference.py:38-114): the generator (generator.model) and the sun-position network (sunpose_net.model) are built for the
panorama size, and one call maps a batch of LDR sky-dome panoramas to linear HDR radiance.  Checkpoint restore, cv2 image I/O
and the .hdr writer around it (inference.py:48-79, 131-157) are outside the hot path."""
from __future__ import annotations

from . import generator as _generator
from . import sunpose_net as _sunpose_net

IMSHAPE = (32, 128, 3)      # inference.py:34
THRESHOLD = 0.12            # inference.py:36


def build_models(batch_size=32, im_height=IMSHAPE[0], im_width=IMSHAPE[1], *, distortion_aware_sunpose=True, math_mode=None,
                 device="cuda"):
    """inference.py:44-46: _gen = generator.model(...), _sun = sunpose_net.model(...)."""
    gen = _generator.model(batch_size=batch_size, im_height=im_height, im_width=im_width, math_mode=math_mode, device=device)
    sun = _sunpose_net.model(im_height=im_height, im_width=im_width, distortion_aware=distortion_aware_sunpose,
                             math_mode=math_mode, device=device)
    gen.build(batch_size)
    return gen, sun


def generator_in_step(gen, sun, ldr, training=False):
    """inference.py:81-112."""
    if training:
        raise NotImplementedError("inference.generator_in_step runs with training=False (inference.py:114)")
    return gen.generator_inference(ldr, sun, threshold=THRESHOLD)
