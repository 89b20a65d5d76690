"""Reference-side glue for libskydome_tf_ops.so (skydome_ops.cc): the body a maintainer puts into the reference's
distortion_aware_ops.conv2d, leaving `__init__`, `build` (add_weight kernel / bias, :30-44) and `distortion` (:198-270) untouched.
NOT importable in this repository's image (no TensorFlow); shipped as the binding SURVEY 8(b) asks for."""
import tensorflow as tf

_sky = tf.load_op_library("libskydome_tf_ops.so")


class DaConv2DLayerMixin:
    """Drop into `class conv2d(Layer)` of the reference: replaces `call` (distortion_aware_ops.py:50-123)."""

    def call(self, inputs):
        k, C = self.kernel_size, inputs.shape[-1]
        offsets = self.offset[0, :, 0]                              # [h, k*k, 2]: the table is replicated over w (:266-268)

        @tf.custom_gradient
        def _op(x, kernel, bias):
            packed = _sky.da_pack_weights(kernel, channels=C, kernel_size=k)
            y = _sky.da_conv2d(x, offsets, offsets, packed, bias, filters=self.filters, kernel_size=k)

            def grad(dy):
                dx = _sky.da_conv2d_grad_input(dy, offsets, kernel, channels=C, kernel_size=k)
                dk, db = _sky.da_conv2d_grad_filter(x, dy, offsets, kernel_size=k)
                return dx, dk, db
            return y, grad

        return _op(inputs, self.kernel, self.bias)
