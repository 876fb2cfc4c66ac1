// TensorFlow custom-op shim over libdeepsphere_b200.so (include/deepsphere_b200.h) — the binding a maintainer of the
// reference adds next to src/deepsphere/gnn_layers.py so that Chebyshev.call / Monomial.call (gnn_layers.py:113-159,
// :281-309), HealpyPool (healpy_layers.py:48-63) and the pseudo-convolutions run on the sm_100a kernels:
//
//   g++ -std=c++17 -shared -fPIC deepsphere_tf_ops.cc -o deepsphere_tf_ops.so \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags() + tf.sysconfig.get_link_flags()))') \
//       -I ../../include -L ../lib -ldeepsphere_b200 -DGOOGLE_CUDA=1
//   _ds = tf.load_op_library("deepsphere_tf_ops.so")          (python side: deepsphere_tf.py in this directory)
//
// TensorFlow is not installable in the image this repository is developed in: this file is source only;
// tests/test_tf_shim_cpu.py checks that every C-ABI call below matches a declaration of the header (names and argument
// counts) and skips the load test when TensorFlow is absent.
//
// Conventions: tensors stay in the reference's own layout [B, M, F] fp32 on the GPU; the plan handle (created once in
// Chebyshev.__init__ through ds_plan_create_coo from the COO triple the reference already builds, gnn_layers.py:68-72)
// travels as an int64 attr; the op takes its CUDA stream from the kernel context, so it is ordered like any other TF op.
#include <algorithm>
#include <cstdint>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "deepsphere_b200.h"

namespace tf = tensorflow;

namespace {

void* StreamOf(tf::OpKernelContext* ctx) { return static_cast<void*>(ctx->eigen_gpu_device().stream()); }

struct ConvAttrs {
  int64_t plan = 0;
  int recursion = 0, k = 1, act = 0, mode = 0;
  explicit ConvAttrs(tf::OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("plan", &plan));
    OP_REQUIRES_OK(c, c->GetAttr("recursion", &recursion));
    OP_REQUIRES_OK(c, c->GetAttr("k", &k));
    OP_REQUIRES_OK(c, c->GetAttr("act", &act));
    OP_REQUIRES_OK(c, c->GetAttr("mode", &mode));
  }
  const ds_plan_t* handle() const { return reinterpret_cast<const ds_plan_t*>(plan); }
};

const float* OptionalPtr(const tf::Tensor& t) { return t.NumElements() > 0 ? t.flat<float>().data() : nullptr; }

}  // namespace

// ---- y, basis = graph_conv_forward(x, kernel, bias) ------------------------------------------------------------------
REGISTER_OP("DsGraphConvForward")
    .Input("x: float")        // [B, M, Fin]
    .Input("kernel: float")   // [K*Fin, Fout], row order f*K + k (gnn_layers.py:145-147)
    .Input("bias: float")     // [Fout] or empty
    .Attr("plan: int")
    .Attr("recursion: int")   // DS_RECURSION_CHEBYSHEV / DS_RECURSION_MONOMIAL
    .Attr("k: int")
    .Attr("act: int")
    .Attr("mode: int")
    .Output("y: float")       // [B, M, Fout]
    .Output("basis: float")   // [K-1, B, M, Fin] when the path materialises it, else empty (fused lattice kernel)
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      tf::shape_inference::ShapeHandle x, w;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 3, &x));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 2, &w));
      c->set_output(0, c->MakeShape({c->Dim(x, 0), c->Dim(x, 1), c->Dim(w, 1)}));
      c->set_output(1, c->UnknownShape());
      return tf::OkStatus();
    });

class DsGraphConvForwardOp : public tf::OpKernel {
 public:
  explicit DsGraphConvForwardOp(tf::OpKernelConstruction* c) : OpKernel(c), a_(c) {}
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    const tf::Tensor& w = ctx->input(1);
    const int64_t B = x.dim_size(0), M = x.dim_size(1), Fin = x.dim_size(2), Fout = w.dim_size(1);
    OP_REQUIRES(ctx, w.dim_size(0) == a_.k * Fin, tf::errors::InvalidArgument("kernel must be [K*Fin, Fout]"));
    tf::Tensor *y = nullptr, *basis = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {B, M, Fout}, &y));
    const bool writes = ds_graph_conv_forward_writes_basis(a_.handle(), a_.k, B, Fin, Fout, a_.mode) != 0;
    const int64_t n_basis = writes ? ds_graph_conv_basis_elems(M, B, Fin, a_.k) : 0;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {n_basis}, &basis));
    const int rc = ds_graph_conv_forward(a_.handle(), a_.recursion, a_.k, B, Fin, Fout, x.flat<float>().data(),
                                         w.flat<float>().data(), OptionalPtr(ctx->input(2)), a_.act,
                                         y->flat<float>().data(), n_basis ? basis->flat<float>().data() : nullptr,
                                         a_.mode, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(ds_last_error()));
  }

 private:
  ConvAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("DsGraphConvForward").Device(tf::DEVICE_GPU), DsGraphConvForwardOp);

// ---- dx, dkernel, dbias = graph_conv_backward(x, kernel, y, dy, basis) ------------------------------------------------
REGISTER_OP("DsGraphConvBackward")
    .Input("x: float")
    .Input("kernel: float")
    .Input("y: float")       // forward output (needed when act != linear), may be empty otherwise
    .Input("dy: float")
    .Input("basis: float")   // what the forward returned (may be empty: recomputed / fused path)
    .Attr("plan: int")
    .Attr("recursion: int")
    .Attr("k: int")
    .Attr("act: int")
    .Attr("mode: int")
    .Attr("has_bias: bool")
    .Output("dx: float")
    .Output("dkernel: float")
    .Output("dbias: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(1));
      c->set_output(2, c->UnknownShape());
      return tf::OkStatus();
    });

class DsGraphConvBackwardOp : public tf::OpKernel {
 public:
  explicit DsGraphConvBackwardOp(tf::OpKernelConstruction* c) : OpKernel(c), a_(c) {
    OP_REQUIRES_OK(c, c->GetAttr("has_bias", &has_bias_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    const tf::Tensor& w = ctx->input(1);
    const tf::Tensor& dy = ctx->input(3);
    const tf::Tensor& basis = ctx->input(4);
    const int64_t B = x.dim_size(0), M = x.dim_size(1), Fin = x.dim_size(2), Fout = w.dim_size(1);
    tf::Tensor *dx = nullptr, *dw = nullptr, *db = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, w.shape(), &dw));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, {has_bias_ ? Fout : 0}, &db));
    const int64_t n_ws = ds_graph_conv_backward_workspace_elems(M, B, Fin, Fout, a_.k, basis.NumElements() > 0, a_.act);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, {n_ws}, &ws));
    const int rc = ds_graph_conv_backward(a_.handle(), a_.recursion, a_.k, B, Fin, Fout, x.flat<float>().data(),
                                          w.flat<float>().data(), OptionalPtr(ctx->input(2)), dy.flat<float>().data(),
                                          a_.act, OptionalPtr(basis), dx->flat<float>().data(), dw->flat<float>().data(),
                                          has_bias_ ? db->flat<float>().data() : nullptr, ws.flat<float>().data(),
                                          a_.mode, StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(ds_last_error()));
  }

 private:
  ConvAttrs a_;
  bool has_bias_ = false;
};
REGISTER_KERNEL_BUILDER(Name("DsGraphConvBackward").Device(tf::DEVICE_GPU), DsGraphConvBackwardOp);

// ---- HealpyPool (healpy_layers.py:48-63) ------------------------------------------------------------------------------
REGISTER_OP("DsPoolForward").Input("x: float").Attr("p: int").Attr("pool_type: int").Output("y: float");
REGISTER_OP("DsPoolBackward").Input("x: float").Input("dy: float").Attr("p: int").Attr("pool_type: int").Output("dx: float");

class DsPoolForwardOp : public tf::OpKernel {
 public:
  explicit DsPoolForwardOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("p", &p_));
    OP_REQUIRES_OK(c, c->GetAttr("pool_type", &type_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    const int64_t B = x.dim_size(0), M = x.dim_size(1), F = x.dim_size(2);
    tf::Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, {B, M >> (2 * p_), F}, &y));
    const int rc = ds_pool_forward(B, M, F, p_, type_, x.flat<float>().data(), y->flat<float>().data(), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(ds_last_error()));
  }

 private:
  int p_ = 1, type_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("DsPoolForward").Device(tf::DEVICE_GPU), DsPoolForwardOp);

class DsPoolBackwardOp : public tf::OpKernel {
 public:
  explicit DsPoolBackwardOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("p", &p_));
    OP_REQUIRES_OK(c, c->GetAttr("pool_type", &type_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& x = ctx->input(0);
    const tf::Tensor& dy = ctx->input(1);
    const int64_t B = x.dim_size(0), M = x.dim_size(1), F = x.dim_size(2);
    tf::Tensor* dx = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    const int rc = ds_pool_backward(B, M, F, p_, type_, x.flat<float>().data(), dy.flat<float>().data(),
                                    dx->flat<float>().data(), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(ds_last_error()));
  }

 private:
  int p_ = 1, type_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("DsPoolBackward").Device(tf::DEVICE_GPU), DsPoolBackwardOp);

// ---- BatchNormalization(center=False, scale=False) + bias + activation (gnn_layers.py:53,152-159), single device ------
REGISTER_OP("DsBnBiasActForward")
    .Input("z: float").Input("bias: float").Input("moving_mean: Ref(float)").Input("moving_var: Ref(float)")
    .Attr("act: int").Attr("training: bool").Attr("eps: float").Attr("momentum: float")
    .Output("y: float").Output("mean_rstd: float");

class DsBnBiasActForwardOp : public tf::OpKernel {
 public:
  explicit DsBnBiasActForwardOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("act", &act_));
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_));
    OP_REQUIRES_OK(c, c->GetAttr("eps", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("momentum", &momentum_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& z = ctx->input(0);
    const int64_t B = z.dim_size(0), M = z.dim_size(1), F = z.dim_size(2);
    tf::Tensor mm = ctx->mutable_input(2, true), mv = ctx->mutable_input(3, true);
    tf::Tensor *y = nullptr, *mr = nullptr, sums, ws, scratch;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, z.shape(), &y));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, {2 * F}, &mr));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_DOUBLE, {2 * F}, &sums));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_DOUBLE, {std::max<int64_t>(ds_bn_workspace_doubles(B, M, F), 1)}, &ws));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, {2 * F}, &scratch));
    int rc = 0;
    if (training_)
      rc = ds_bn_stats(B, M, F, 0, M, z.flat<float>().data(), sums.flat<double>().data(), ws.flat<double>().data(),
                       StreamOf(ctx));
    // (multi-GPU: all-reduce `sums` and the row count here — tf.distribute / NCCL — before the second call)
    if (rc == 0)
      rc = ds_bn_bias_act_forward(B, M, F, z.flat<float>().data(), sums.flat<double>().data(), (double)(B * M), nullptr,
                                  eps_, momentum_, training_ ? 1 : 0, mm.flat<float>().data(), mv.flat<float>().data(),
                                  OptionalPtr(ctx->input(1)), act_, mr->flat<float>().data(),
                                  scratch.flat<float>().data(), y->flat<float>().data(), StreamOf(ctx));
    OP_REQUIRES(ctx, rc == 0, tf::errors::Internal(ds_last_error()));
  }

 private:
  int act_ = 0;
  bool training_ = false;
  float eps_ = 1e-5f, momentum_ = 0.9f;
};
REGISTER_KERNEL_BUILDER(Name("DsBnBiasActForward").Device(tf::DEVICE_GPU), DsBnBiasActForwardOp);
