// tf_op_driver.cc -- runs tf_shim/gaccum_tf_op.cc (compiled UNMODIFIED against tests/tf_mock's headers) the way a
// TensorFlow executor would: looks the op and its GPU kernel up in the registry the adapter's REGISTER_OP /
// REGISTER_KERNEL_BUILDER statements filled, constructs the kernel from an attribute bag, builds the input list IN THE ORDER
// THE ADAPTER'S OWN OpDef DECLARES (so a Compute() that indexes its inputs differently from its registration fails here),
// places inputs in device memory unless the kernel registration says HostMemory, and calls Compute() once per micro-step
// on a non-default stream.  TEST INFRASTRUCTURE; see tests/tf_mock/tensorflow/core/framework/op_kernel.h for what the mock
// can and cannot prove.
//
//   tf_op_driver --registry                      print what the adapter registered (CPU)
//   tf_op_driver --errors                        constructor / Compute error paths that need no GPU (CPU)
//   tf_op_driver <GaccumStep|GaccumStepV2> <in.bin> <out.bin>      a trajectory on cuda:0; formats of tests/abi_consumer.cc
#include <cuda_runtime_api.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "gaccum.h"
#include "tensorflow/core/framework/op_kernel.h"

namespace tf = tensorflow;

#define CHECK_CUDA(x)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
  } while (0)

template <typename T>
static bool rd(FILE* f, T* v, size_t n = 1) { return std::fread(v, sizeof(T), n, f) == n; }

// "params: Ref(N * float)" -> {name, is_list, is_ref, is_resource, dtype}
struct InputSpec { std::string name, type; bool list = false, ref = false, resource = false; };
static InputSpec parse_input(const std::string& s) {
  InputSpec r;
  const size_t colon = s.find(':');
  r.name = s.substr(0, colon);
  std::string t = s.substr(colon + 1);
  t.erase(std::remove(t.begin(), t.end(), ' '), t.end());
  if (t.rfind("Ref(", 0) == 0) { r.ref = true; t = t.substr(4, t.size() - 5); }
  if (t.rfind("N*", 0) == 0) { r.list = true; t = t.substr(2); }
  r.resource = t == "resource";
  r.type = t;
  return r;
}

static void set_common_attrs(tf::OpKernelConstruction* c, int T, int accum_n, int variant, const char* b1, const char* b2, const char* eps,
                             const char* wd, const char* clip, const std::vector<bool>& mask) {
  c->attrs["N"].i = T;
  c->attrs["accum_n"].i = accum_n;
  c->attrs["variant"].i = variant;
  c->attrs["beta1"].s = b1;
  c->attrs["beta2"].s = b2;
  c->attrs["epsilon"].s = eps;
  c->attrs["weight_decay_rate"].s = wd;
  c->attrs["clip_norm"].s = clip;
  c->attrs["decay_mask"].bl = mask;
}

static int print_registry() {
  auto& reg = tf::MockRegistry::Get();
  for (auto& [name, def] : reg.ops) {
    std::printf("op %s stateful=%d\n", name.c_str(), (int)def.stateful);
    for (auto& i : def.inputs) std::printf("  input %s\n", i.c_str());
    for (auto& a : def.attrs) std::printf("  attr %s\n", a.c_str());
    auto k = reg.kernels.find(name);
    if (k == reg.kernels.end()) { std::printf("  NO KERNEL\n"); continue; }
    std::printf("  kernel device=%s host_memory=", k->second.first.device.c_str());
    for (auto& h : k->second.first.host_memory) std::printf("%s,", h.c_str());
    std::printf("\n");
  }
  return 0;
}

// Everything below builds contexts the same way; the struct owns the memory the Tensors view.
struct Harness {
  const tf::OpDef* def = nullptr;
  const tf::KernelDef* kdef = nullptr;
  std::unique_ptr<tf::OpKernel> kernel;
  std::vector<InputSpec> specs;
  tf::Stream stream{nullptr};
  tf::DeviceContext dctx{&stream};
  tf::DeviceBase device;
  std::vector<std::unique_ptr<tf::Var>> vars;      // V2: params[0..T), accum, m, v

  bool host_memory(const std::string& n) const { return std::find(kdef->host_memory.begin(), kdef->host_memory.end(), n) != kdef->host_memory.end(); }

  tf::Status create(const std::string& op, tf::OpKernelConstruction* c) {
    auto& reg = tf::MockRegistry::Get();
    auto o = reg.ops.find(op);
    auto k = reg.kernels.find(op);
    if (o == reg.ops.end() || k == reg.kernels.end()) return tf::errors::InvalidArgument("op ", op, " not registered");
    def = &o->second;
    kdef = &k->second.first;
    for (auto& i : def->inputs) specs.push_back(parse_input(i));
    kernel.reset(k->second.second(c));
    return c->status;
  }
};

static int error_paths() {
  // (1) decay_mask with the wrong number of entries is refused by the constructor
  {
    tf::OpKernelConstruction c;
    set_common_attrs(&c, 3, 4, 0, "0.9", "0.999", "1e-06", "0.01", "1.0", {true, false});
    Harness h;
    tf::Status s = h.create("GaccumStep", &c);
    if (s.ok() || s.message().find("decay_mask") == std::string::npos) { std::fprintf(stderr, "(1) expected a decay_mask error, got '%s'\n", s.ToString().c_str()); return 1; }
  }
  // (2) a missing attr is reported, not defaulted
  {
    tf::OpKernelConstruction c;
    set_common_attrs(&c, 1, 4, 0, "0.9", "0.999", "1e-06", "0.01", "1.0", {true});
    c.attrs.erase("epsilon");
    Harness h;
    if (h.create("GaccumStepV2", &c).ok()) { std::fprintf(stderr, "(2) missing attr accepted\n"); return 1; }
  }
  // (3) a gradient whose size differs from its variable's is an InvalidArgument before anything is launched
  // (4) without a CUDA device Compute fails with the library's message (no CPU fallback) -- only checked when there is none
  {
    tf::OpKernelConstruction c;
    set_common_attrs(&c, 1, 4, 0, "0.9", "0.999", "1e-06", "0.01", "1.0", {true});
    Harness h;
    if (!h.create("GaccumStep", &c).ok()) { std::fprintf(stderr, "(3) construction failed: %s\n", c.status.ToString().c_str()); return 1; }
    float* fake = reinterpret_cast<float*>(uintptr_t(0x10000));
    int64_t step = 0; float lr = 1e-3f; float bp[2] = {0.9f, 0.999f};
    auto make = [&](int64_t grad_numel) {
      tf::OpKernelContext ctx;
      ctx.dev_ctx = &h.dctx; ctx.dev = &h.device;
      for (auto& sp : h.specs) {
        tf::Tensor t;
        if (sp.name == "global_step") t = tf::Tensor(tf::DT_INT64, &step, 1);
        else if (sp.name == "lr") t = tf::Tensor(tf::DT_FLOAT, &lr, 1);
        else if (sp.name == "beta_powers") t = tf::Tensor(tf::DT_FLOAT, bp, 2);
        else if (sp.name == "grads") t = tf::Tensor(tf::DT_FLOAT, fake, grad_numel);
        else if (sp.name == "params") t = tf::Tensor(tf::DT_FLOAT, fake, 64);
        else t = tf::Tensor(tf::DT_FLOAT, fake, 64);
        ctx.inputs.push_back(t); ctx.resources.push_back(nullptr); ctx.is_ref.push_back(sp.ref);
      }
      return ctx;
    };
    tf::OpKernelContext bad = make(63);
    h.kernel->Compute(&bad);
    if (bad.status().ok() || bad.status().message().find("does not match") == std::string::npos) { std::fprintf(stderr, "(3) got '%s'\n", bad.status().ToString().c_str()); return 1; }
    if (gaccum_device_count() < 1) {
      tf::OpKernelContext ctx = make(64);
      h.kernel->Compute(&ctx);
      if (ctx.status().ok() || ctx.status().message().find("gaccum_plan_create") == std::string::npos) { std::fprintf(stderr, "(4) got '%s'\n", ctx.status().ToString().c_str()); return 1; }
      std::printf("no device: %s\n", ctx.status().message().c_str());
    }
  }
  std::printf("error paths ok\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 2 && std::string(argv[1]) == "--registry") return print_registry();
  if (argc == 2 && std::string(argv[1]) == "--errors") return error_paths();
  if (argc != 4) { std::fprintf(stderr, "usage: %s --registry | --errors | <op> <in.bin> <out.bin>\n", argv[0]); return 64; }
  const std::string op = argv[1];
  FILE* in = std::fopen(argv[2], "rb");
  if (!in) { std::perror(argv[2]); return 64; }
  int32_t T, N, steps, variant;
  double init_lr, clip_norm;
  int64_t train_steps, warmup_steps;
  if (!rd(in, &T) || !rd(in, &N) || !rd(in, &steps) || !rd(in, &variant) || !rd(in, &init_lr) || !rd(in, &train_steps) ||
      !rd(in, &warmup_steps) || !rd(in, &clip_norm)) return 64;
  std::vector<int64_t> numel(T);
  std::vector<std::string> names(T);
  for (int t = 0; t < T; ++t) {
    int32_t len;
    if (!rd(in, &numel[t]) || !rd(in, &len)) return 64;
    names[t].resize(len);
    if (len && !rd(in, &names[t][0], (size_t)len)) return 64;
  }
  if (gaccum_device_count() < 1) { std::fprintf(stderr, "no CUDA device\n"); return 77; }
  CHECK_CUDA(cudaSetDevice(0));

  // ---- graph-construction time (what tf_shim/optimization.py does in Python): mask, slab sizes, attrs as repr() strings ----
  std::vector<const char*> cnames(T);
  for (int t = 0; t < T; ++t) cnames[t] = names[t].c_str();
  const char* excl[] = {"LayerNorm", "layer_norm", "bias"};
  std::vector<uint8_t> decay(T, 0);
  if (variant == GACCUM_ADAM_WEIGHT_DECAY && gaccum_decay_mask(T, cnames.data(), 0.01, excl, 3, decay.data()) != 0) return 3;
  gaccum_hparams hp_layout = {variant, 0, 0.9, 0.999, 1e-6, 0.01, 1.0};
  gaccum_plan* layout = nullptr;
  if (gaccum_plan_create(&layout, T, numel.data(), decay.data(), &hp_layout, -1) != 0) { std::fprintf(stderr, "%s\n", gaccum_last_error()); return 3; }
  const int64_t padded = gaccum_padded_size(layout);
  std::vector<int64_t> off(T);
  gaccum_offsets(layout, off.data());
  gaccum_plan_destroy(layout);

  tf::OpKernelConstruction cons;
  char clip_s[64];
  std::snprintf(clip_s, sizeof clip_s, "%.17g", clip_norm);
  if (variant == GACCUM_ADAM_WEIGHT_DECAY) set_common_attrs(&cons, T, N, 0, "0.9", "0.999", "1e-06", "0.01", clip_s, std::vector<bool>(decay.begin(), decay.end()));
  else set_common_attrs(&cons, T, N, 1, "0.9", "0.999", "1e-08", "0.0", "0.0", std::vector<bool>(T, false));
  Harness h;
  tf::Status cs = h.create(op, &cons);
  if (!cs.ok()) { std::fprintf(stderr, "construction: %s\n", cs.ToString().c_str()); return 3; }

  // ---- the executor's memory: variables, gradients, scalars (host or device as the kernel registration says) ----------
  cudaStream_t cu_stream;
  CHECK_CUDA(cudaStreamCreateWithFlags(&cu_stream, cudaStreamNonBlocking));
  h.stream = tf::Stream(cu_stream);
  h.device.info.gpu_id = 0;
  const size_t slab_bytes = (size_t)(padded > 0 ? padded : 32) * sizeof(float);
  float* slabs[3];
  for (auto& s : slabs) { CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&s), slab_bytes)); CHECK_CUDA(cudaMemset(s, 0, slab_bytes)); }
  std::vector<float*> params(T), grads(T);
  std::vector<float> host;
  for (int t = 0; t < T; ++t) {
    const size_t bytes = (size_t)(numel[t] > 0 ? numel[t] : 1) * sizeof(float);
    CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&params[t]), bytes)); CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&grads[t]), bytes));
    host.resize((size_t)numel[t]);
    if (numel[t] && !rd(in, host.data(), (size_t)numel[t])) return 64;
    CHECK_CUDA(cudaMemcpy(params[t], host.data(), (size_t)numel[t] * sizeof(float), cudaMemcpyHostToDevice));
  }
  const bool v2 = h.specs[0].resource;
  if (v2) {
    for (int t = 0; t < T; ++t) { h.vars.emplace_back(new tf::Var); *h.vars.back()->tensor() = tf::Tensor(tf::DT_FLOAT, params[t], numel[t]); }
    for (int k = 0; k < 3; ++k) { h.vars.emplace_back(new tf::Var); *h.vars.back()->tensor() = tf::Tensor(tf::DT_FLOAT, slabs[k], padded); }
  }
  // scalars: host memory only if the kernel registration declared it; otherwise the executor would hand device memory
  struct Scalars { int64_t step; float lr; float bp[2]; } hs{0, 0.f, {0.9f, 0.999f}};
  Scalars* ds = nullptr;
  CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&ds), sizeof(Scalars)));
  auto scalar_ptr = [&](const std::string& name, size_t offset) -> void* {
    return h.host_memory(name) ? reinterpret_cast<char*>(&hs) + offset : reinterpret_cast<char*>(ds) + offset;
  };
  for (const char* n : {"global_step", "lr", "beta_powers"})
    if (!h.host_memory(n)) std::fprintf(stderr, "warning: %s is not HostMemory: Compute will be handed a device pointer\n", n);

  FILE* out = std::fopen(argv[3], "wb");
  if (!out) { std::perror(argv[3]); return 64; }
  for (int s = 0; s < steps; ++s) {
    for (int t = 0; t < T; ++t) {
      host.resize((size_t)numel[t]);
      if (numel[t] && !rd(in, host.data(), (size_t)numel[t])) return 64;
      CHECK_CUDA(cudaMemcpyAsync(grads[t], host.data(), (size_t)numel[t] * sizeof(float), cudaMemcpyHostToDevice, cu_stream));
      CHECK_CUDA(cudaStreamSynchronize(cu_stream));
    }
    hs.lr = variant == GACCUM_ADAM_WEIGHT_DECAY ? gaccum_learning_rate(init_lr, train_steps, warmup_steps, hs.step) : (float)init_lr;
    CHECK_CUDA(cudaMemcpy(ds, &hs, sizeof hs, cudaMemcpyHostToDevice));
    // ---- one executor step: inputs in the OpDef's order ----
    tf::OpKernelContext ctx;
    ctx.dev_ctx = &h.dctx; ctx.dev = &h.device;
    int slab_i = 0;
    for (auto& sp : h.specs) {
      auto push = [&](tf::Tensor t, tf::Var* var) { ctx.inputs.push_back(t); ctx.resources.push_back(var); ctx.is_ref.push_back(sp.ref); };
      if (sp.name == "params") {
        for (int t = 0; t < T; ++t) sp.resource ? push(tf::Tensor(tf::DT_RESOURCE, nullptr, 1), h.vars[t].get()) : push(tf::Tensor(tf::DT_FLOAT, params[t], numel[t]), nullptr);
      } else if (sp.name == "accum" || sp.name == "m" || sp.name == "v") {
        const int k = sp.name == "accum" ? 0 : sp.name == "m" ? 1 : 2;
        ++slab_i;
        sp.resource ? push(tf::Tensor(tf::DT_RESOURCE, nullptr, 1), h.vars[T + k].get()) : push(tf::Tensor(tf::DT_FLOAT, slabs[k], padded), nullptr);
      } else if (sp.name == "grads") {
        for (int t = 0; t < T; ++t) push(tf::Tensor(tf::DT_FLOAT, grads[t], numel[t]), nullptr);
      } else if (sp.name == "global_step") {
        push(tf::Tensor(tf::DT_INT64, scalar_ptr("global_step", offsetof(Scalars, step)), 1), nullptr);
      } else if (sp.name == "lr") {
        push(tf::Tensor(tf::DT_FLOAT, scalar_ptr("lr", offsetof(Scalars, lr)), 1), nullptr);
      } else if (sp.name == "beta_powers") {
        push(tf::Tensor(tf::DT_FLOAT, scalar_ptr("beta_powers", offsetof(Scalars, bp)), 2), nullptr);
      } else {
        std::fprintf(stderr, "unknown input %s\n", sp.name.c_str());
        return 5;
      }
    }
    if (slab_i != 3) { std::fprintf(stderr, "op def lacks accum/m/v\n"); return 5; }
    h.kernel->Compute(&ctx);
    if (!ctx.status().ok()) { std::fprintf(stderr, "Compute: %s\n", ctx.status().ToString().c_str()); return 6; }
    const bool applied = gaccum_is_apply_step(hs.step, N) != 0;
    if (applied && variant == GACCUM_ADAM) { hs.bp[0] *= 0.9f; hs.bp[1] *= 0.999f; }   // TF1 Adam's _finish (non-slot variables)
    const float hdr[4] = {hs.lr, 0.f, 0.f, applied ? 1.f : 0.f};
    ++hs.step;                                                                          // optimization.py:102-103 (the Python side's assign)
    CHECK_CUDA(cudaStreamSynchronize(cu_stream));
    std::fwrite(hdr, sizeof(float), 4, out);
    for (int t = 0; t < T; ++t) {
      const size_t n = (size_t)numel[t];
      host.resize(n);
      const float* src[4] = {params[t], slabs[0] + off[t], slabs[1] + off[t], slabs[2] + off[t]};
      for (int k = 0; k < 4; ++k) {
        if (n) CHECK_CUDA(cudaMemcpy(host.data(), src[k], n * sizeof(float), cudaMemcpyDeviceToHost));
        std::fwrite(host.data(), sizeof(float), n, out);
      }
    }
  }
  std::fclose(out);
  std::fclose(in);
  h.kernel.reset();                                   // ~GaccumStepOpT destroys the plan
  std::printf("tf_op_driver ok: %s, %d tensors, %d micro-steps, N=%d, variant %d\n", op.c_str(), T, steps, N, variant);
  return 0;
}
