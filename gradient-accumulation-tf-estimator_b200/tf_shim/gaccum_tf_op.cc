// gaccum_tf_op.cc -- the `tf.load_op_library` custom op that puts libgaccum.so behind the
// reference's unchanged `create_optimizer(...)` (reference optimization.py:25-104).
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: TensorFlow (headers and runtime) is absent here
// (DESIGN.md "Boundary").  Build where TensorFlow >= 2.4 with tf.compat.v1 exists:
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC gaccum_tf_op.cc -o libgaccum_tf.so $TF_CFLAGS $TF_LFLAGS \
//       -I../../include -L../csrc -lgaccum -Wl,-rpath,'$ORIGIN/../csrc' -DGOOGLE_CUDA=1
//
// The op is a thin adapter: it reads raw device pointers and the device stream out of the
// OpKernelContext and calls the C ABI (include/gaccum.h).  It never allocates, never
// synchronises, and launches exactly one kernel (`gaccum_step`).
//
//   GaccumStep(params: N * Ref(float), grads: N * float, accum: Ref(float), m: Ref(float),
//              v: Ref(float), global_step: int64 [host], lr: float [host],
//              beta_powers: float[2] [host])
//     attrs: N, accum_n, variant, beta1, beta2, epsilon, weight_decay_rate, clip_norm,
//            decay_mask: list(bool)
//
// State (accum/m/v) are ordinary TF variables created by tf_shim/optimization.py, so the
// Estimator's Saver checkpoints them like the reference's per-variable accumulators
// (optimization.py:78, 137-148) -- as three packed tensors instead of 3T.
#include <vector>

#include "gaccum.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

namespace tf = tensorflow;

REGISTER_OP("GaccumStep")
    .Input("params: Ref(N * float)")
    .Input("grads: N * float")
    .Input("accum: Ref(float)")
    .Input("m: Ref(float)")
    .Input("v: Ref(float)")
    .Input("global_step: int64")
    .Input("lr: float")
    .Input("beta_powers: float")
    .Attr("N: int >= 1")
    .Attr("accum_n: int >= 1")
    .Attr("variant: int = 0")
    .Attr("beta1: float = 0.9")
    .Attr("beta2: float = 0.999")
    .Attr("epsilon: float = 1e-6")
    .Attr("weight_decay_rate: float = 0.01")
    .Attr("clip_norm: float = 1.0")
    .Attr("decay_mask: list(bool)")
    .SetIsStateful()
    .SetShapeFn(tf::shape_inference::NoOutputs)
    .Doc("One micro-step of the gradient-accumulation train_op (reference optimization.py:76-104).");

class GaccumStepOp : public tf::OpKernel {
 public:
  explicit GaccumStepOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    float b1, b2, eps, wd, clip;
    OP_REQUIRES_OK(c, c->GetAttr("N", &n_));
    OP_REQUIRES_OK(c, c->GetAttr("accum_n", &accum_n_));
    OP_REQUIRES_OK(c, c->GetAttr("variant", &hp_.variant));
    OP_REQUIRES_OK(c, c->GetAttr("beta1", &b1));
    OP_REQUIRES_OK(c, c->GetAttr("beta2", &b2));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps));
    OP_REQUIRES_OK(c, c->GetAttr("weight_decay_rate", &wd));
    OP_REQUIRES_OK(c, c->GetAttr("clip_norm", &clip));
    OP_REQUIRES_OK(c, c->GetAttr("decay_mask", &decay_));
    // attrs are fp32 in the GraphDef; the Python side passes the exact doubles the reference uses
    // through string attrs if bit-parity of (1.0 - beta) matters -- see tf_shim/optimization.py.
    hp_.reserved = 0;
    hp_.beta1 = b1 == 0.9f ? 0.9 : b1;
    hp_.beta2 = b2 == 0.999f ? 0.999 : b2;
    hp_.epsilon = eps == 1e-6f ? 1e-6 : (eps == 1e-8f ? 1e-8 : eps);
    hp_.weight_decay_rate = wd == 0.01f ? 0.01 : wd;
    hp_.clip_norm = clip;
  }

  ~GaccumStepOp() override { gaccum_plan_destroy(plan_); }

  void Compute(tf::OpKernelContext* c) override {
    // Ref inputs: take the variables' buffers without copying (lock_held = true is fine: the
    // reference relies on control dependencies for ordering, optimization.py:82,86).
    std::vector<float*> params(n_);
    std::vector<const float*> grads(n_);
    std::vector<int64_t> numels(n_);
    for (int i = 0; i < n_; ++i) {
      tf::Tensor p = c->mutable_input(i, /*lock_held=*/true);
      const tf::Tensor& g = c->input(n_ + i);
      OP_REQUIRES(c, p.NumElements() == g.NumElements(),
                  tf::errors::InvalidArgument("grad ", i, " does not match its variable"));
      params[i] = p.flat<float>().data();
      grads[i] = g.flat<float>().data();
      numels[i] = p.NumElements();
    }
    tf::Tensor accum = c->mutable_input(2 * n_ + 0, true);
    tf::Tensor m = c->mutable_input(2 * n_ + 1, true);
    tf::Tensor v = c->mutable_input(2 * n_ + 2, true);
    const int64_t step = c->input(2 * n_ + 3).scalar<int64_t>()();   // HostMemory
    const float lr = c->input(2 * n_ + 4).scalar<float>()();          // HostMemory
    auto bp = c->input(2 * n_ + 5).flat<float>();                     // HostMemory

    if (plan_ == nullptr) {   // first run: shapes are static in a TF1 graph
      std::vector<uint8_t> decay(decay_.begin(), decay_.end());
      int dev = 0;
      cudaGetDevice(&dev);
      OP_REQUIRES(c, gaccum_plan_create(&plan_, n_, numels.data(), decay.data(), &hp_, dev) == 0,
                  tf::errors::Internal("gaccum_plan_create: ", gaccum_last_error()));
      OP_REQUIRES(c, accum.NumElements() >= gaccum_padded_size(plan_),
                  tf::errors::InvalidArgument("accum slab too small"));
    }
    gaccum_step_args a{};
    a.global_step = step;
    a.accum_n = accum_n_;
    a.lr = lr;
    a.beta1_power = bp(0);
    a.beta2_power = bp(1);
    // TF's compute stream for this device: the kernel is stream-ordered after the backward pass
    // that produced `grads` and before whatever reads the variables next.
    auto* stream = c->op_device_context()->stream();
    gaccum_stream_t cu_stream = stream->platform_specific_handle().stream;   // CUstream
    const int rc = gaccum_step(plan_, grads.data(), params.data(), accum.flat<float>().data(),
                               m.flat<float>().data(), v.flat<float>().data(), &a, cu_stream);
    OP_REQUIRES(c, rc == 0, tf::errors::Internal("gaccum_step: ", gaccum_last_error()));
  }

 private:
  int n_ = 0, accum_n_ = 1;
  gaccum_hparams hp_{};
  std::vector<bool> decay_;
  gaccum_plan* plan_ = nullptr;
};

REGISTER_KERNEL_BUILDER(Name("GaccumStep")
                            .Device(tf::DEVICE_GPU)
                            .HostMemory("global_step")
                            .HostMemory("lr")
                            .HostMemory("beta_powers"),
                        GaccumStepOp);
