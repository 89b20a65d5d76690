"""Host-side mirror of the reference's grad_cam.py: ``layer(y_c, A_k)`` (grad_cam.py:29-45).

    grad = tf.gradients(y_c, A_k)[0]; weights = reduce_mean(grad, axis=(1, 2));
    cam = relu(einsum('bc,bwhc->bwh', weights, A_k))[..., None]

TensorFlow differentiates symbolically; here `y_c` is the `ClassScore` handle returned by
``sunpose_net.model.class_score(softmax)``, which runs the backward sweep of the sun-position network once and serves the
gradient with respect to each of its three activation maps (sunpose_net.py:72).  The channel mean and the weighted channel sum
are two small kernels behind ``sky_gradcam``.
"""
from __future__ import annotations

import torch

from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream


def layer(y_c, A_k):
    A_k = _require_cuda(A_k, "A_k")
    grad = y_c.gradient(A_k)                                              # grad_cam.py:31
    B, h, w, C = A_k.shape
    wsum = torch.empty((B, C), dtype=torch.float32, device=A_k.device)    # :34 (sum; the 1/(h w) is applied by the second kernel)
    cam = torch.empty((B, h, w, 1), dtype=torch.float32, device=A_k.device)
    check(LIB.sky_gradcam(grad.data_ptr(), A_k.data_ptr(), wsum.data_ptr(), cam.data_ptr(), B, h, w, C, _stream()))   # :35-43
    return cam
