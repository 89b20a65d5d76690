// TensorFlow custom-op wrapper over the C ABI of include/skydome_b200.h.
//
// NOT BUILT IN THIS REPOSITORY: the image has no TensorFlow headers.  A maintainer with TensorFlow >= 2.4 builds it with
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC skydome_ops.cc -o libskydome_tf_ops.so $TF_CFLAGS $TF_LFLAGS \
//       -I../../include -L<package dir> -lskydome_b200 -DGOOGLE_CUDA=1 -I/usr/local/cuda/include
//
// and loads it from the reference's distortion_aware_ops.py (see skydome_ops.py next to this file).  One op per entry point; the
// three below are the ones conv2d.call / its autodiff need (distortion_aware_ops.py:50-123).  Offsets are passed twice: as a device
// tensor (what the kernels read) and, through `offsets_host`, as a host-memory input (what sizes the shared-memory band).
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

#include "skydome_b200.h"

namespace tf = tensorflow;

namespace {

void *CudaStream(tf::OpKernelContext *ctx)
{
    // TF owns one compute stream per GPU device; every sky_* call is asynchronous on the stream it is given
    return ctx->op_device_context()->stream()->platform_specific_handle().stream;
}

#define SKY_TF_CHECK(ctx, call)                                                       \
    do {                                                                              \
        const int _rc = (call);                                                       \
        OP_REQUIRES(ctx, _rc == SKY_OK, tf::errors::Internal(#call ": ", sky_last_error())); \
    } while (0)

}  // namespace

// ---- DaPackWeights: kernel [k*k*C, F] -> packed uint8 image for the tensor-core kernels (re-run when the variable changes) ----------
REGISTER_OP("DaPackWeights")
    .Input("kernel: float")
    .Attr("channels: int")
    .Attr("kernel_size: int")
    .Output("packed: uint8")
    .SetShapeFn([](tf::shape_inference::InferenceContext *c) {
        c->set_output(0, c->Vector(c->UnknownDim()));
        return tf::Status();
    });

class DaPackWeightsOp : public tf::OpKernel {
 public:
    explicit DaPackWeightsOp(tf::OpKernelConstruction *c) : OpKernel(c)
    {
        OP_REQUIRES_OK(c, c->GetAttr("channels", &channels_));
        OP_REQUIRES_OK(c, c->GetAttr("kernel_size", &k_));
    }
    void Compute(tf::OpKernelContext *ctx) override
    {
        const tf::Tensor &kernel = ctx->input(0);
        const int F = static_cast<int>(kernel.dim_size(1));
        const size_t bytes = sky_da_packed_weight_bytes(channels_, F, k_, SKY_MATH_TF32);
        tf::Tensor *packed = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({static_cast<tf::int64>(bytes)}), &packed));
        SKY_TF_CHECK(ctx, sky_da_pack_weights(kernel.flat<float>().data(), packed->flat<tf::uint8>().data(), channels_, F, k_,
                                              SKY_MATH_TF32, CudaStream(ctx)));
    }

 private:
    int channels_, k_;
};
REGISTER_KERNEL_BUILDER(Name("DaPackWeights").Device(tf::DEVICE_GPU), DaPackWeightsOp);

// ---- DaConv2D: conv2d.call (distortion_aware_ops.py:50-123) -------------------------------------------------------------------------
REGISTER_OP("DaConv2D")
    .Input("x: float")              // [B, h, w, C]
    .Input("offsets: float")        // [h, k*k, 2] on the device: self.offset[0, :, 0] of the reference layer
    .Input("offsets_host: float")   // the same table in host memory
    .Input("packed: uint8")
    .Input("bias: float")
    .Attr("filters: int")
    .Attr("kernel_size: int")
    .Output("y: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext *c) {
        tf::shape_inference::ShapeHandle x;
        TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 4, &x));
        int filters;
        TF_RETURN_IF_ERROR(c->GetAttr("filters", &filters));
        c->set_output(0, c->MakeShape({c->Dim(x, 0), c->Dim(x, 1), c->Dim(x, 2), filters}));
        return tf::Status();
    });

class DaConv2DOp : public tf::OpKernel {
 public:
    explicit DaConv2DOp(tf::OpKernelConstruction *c) : OpKernel(c)
    {
        OP_REQUIRES_OK(c, c->GetAttr("filters", &filters_));
        OP_REQUIRES_OK(c, c->GetAttr("kernel_size", &k_));
    }
    void Compute(tf::OpKernelContext *ctx) override
    {
        const tf::Tensor &x = ctx->input(0), &off = ctx->input(1), &off_host = ctx->input(2), &packed = ctx->input(3), &bias = ctx->input(4);
        const int B = x.dim_size(0), h = x.dim_size(1), w = x.dim_size(2), C = x.dim_size(3);
        tf::Tensor *y = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, h, w, filters_}), &y));
        SKY_TF_CHECK(ctx, sky_da_conv2d_fwd(x.flat<float>().data(), off.flat<float>().data(), off_host.flat<float>().data(),
                                            packed.flat<tf::uint8>().data(), bias.flat<float>().data(), y->flat<float>().data(),
                                            /*residual=*/nullptr, /*stats=*/nullptr, B, h, w, C, filters_, k_, SKY_EPI_NONE, 0.f,
                                            SKY_MATH_TF32, CudaStream(ctx)));
    }

 private:
    int filters_, k_;
};
REGISTER_KERNEL_BUILDER(Name("DaConv2D").Device(tf::DEVICE_GPU).HostMemory("offsets_host"), DaConv2DOp);

// ---- DaConv2DGradInput / DaConv2DGradFilter: what TF autodiff derives from :50-123 (offsets are constants) --------------------------
REGISTER_OP("DaConv2DGradInput")
    .Input("dy: float").Input("offsets: float").Input("kernel: float")
    .Attr("channels: int").Attr("kernel_size: int")
    .Output("dx: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext *c) {
        tf::shape_inference::ShapeHandle dy;
        TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 4, &dy));
        int channels;
        TF_RETURN_IF_ERROR(c->GetAttr("channels", &channels));
        c->set_output(0, c->MakeShape({c->Dim(dy, 0), c->Dim(dy, 1), c->Dim(dy, 2), channels}));
        return tf::Status();
    });

class DaConv2DGradInputOp : public tf::OpKernel {
 public:
    explicit DaConv2DGradInputOp(tf::OpKernelConstruction *c) : OpKernel(c)
    {
        OP_REQUIRES_OK(c, c->GetAttr("channels", &channels_));
        OP_REQUIRES_OK(c, c->GetAttr("kernel_size", &k_));
    }
    void Compute(tf::OpKernelContext *ctx) override
    {
        const tf::Tensor &dy = ctx->input(0), &off = ctx->input(1), &kernel = ctx->input(2);
        const int B = dy.dim_size(0), h = dy.dim_size(1), w = dy.dim_size(2), F = dy.dim_size(3);
        tf::Tensor *dx = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, h, w, channels_}), &dx));
        SKY_TF_CHECK(ctx, sky_da_conv2d_bwd_data(dy.flat<float>().data(), off.flat<float>().data(), kernel.flat<float>().data(),
                                                 dx->flat<float>().data(), B, h, w, channels_, F, k_, /*accumulate=*/0, CudaStream(ctx)));
    }

 private:
    int channels_, k_;
};
REGISTER_KERNEL_BUILDER(Name("DaConv2DGradInput").Device(tf::DEVICE_GPU), DaConv2DGradInputOp);

REGISTER_OP("DaConv2DGradFilter")
    .Input("x: float").Input("dy: float").Input("offsets: float")
    .Attr("kernel_size: int")
    .Output("dkernel: float").Output("dbias: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext *c) {
        c->set_output(0, c->Matrix(c->UnknownDim(), c->UnknownDim()));
        c->set_output(1, c->Vector(c->UnknownDim()));
        return tf::Status();
    });

class DaConv2DGradFilterOp : public tf::OpKernel {
 public:
    explicit DaConv2DGradFilterOp(tf::OpKernelConstruction *c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("kernel_size", &k_)); }
    void Compute(tf::OpKernelContext *ctx) override
    {
        const tf::Tensor &x = ctx->input(0), &dy = ctx->input(1), &off = ctx->input(2);
        const int B = x.dim_size(0), h = x.dim_size(1), w = x.dim_size(2), C = x.dim_size(3), F = dy.dim_size(3);
        tf::Tensor *dk = nullptr, *db = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({k_ * k_ * C, F}), &dk));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({F}), &db));
        SKY_TF_CHECK(ctx, sky_da_conv2d_bwd_filter(x.flat<float>().data(), dy.flat<float>().data(), off.flat<float>().data(),
                                                   dk->flat<float>().data(), db->flat<float>().data(), B, h, w, C, F, k_, CudaStream(ctx)));
    }

 private:
    int k_;
};
REGISTER_KERNEL_BUILDER(Name("DaConv2DGradFilter").Device(tf::DEVICE_GPU), DaConv2DGradFilterOp);
