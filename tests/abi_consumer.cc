// abi_consumer.cc -- a non-Python consumer of include/gaccum.h, built with plain g++ (no nvcc, no torch).
//
// It drives libgaccum.so the way the TensorFlow custom op of INTEGRATION.md (`GaccumStepOp::Compute`) would:
// device tensors it allocated itself with cudaMalloc (scattered, one allocation per variable), packed
// accum / adam_m / adam_v slabs sized by gaccum_padded_size(), a non-default non-blocking stream, one
// gaccum_step() per micro-step with no synchronisation in between, host-side scalars (learning rate, apply
// predicate, decay mask) from the library's own helpers.  tests/test_abi_consumer.py feeds it a trajectory
// exported from a golden fixture (produced by the reference's optimization.py, tests/golden/) and compares
// what it writes back.
//
//   abi_consumer --host-only                 header + link check, host-side helpers, layout-only plan
//   abi_consumer <in.bin> <out.bin>          run a trajectory on cuda:0
//
// in.bin (little endian):  int32 T, N, steps, variant; double init_lr; int64 num_train_steps, num_warmup_steps;
//                          double clip_norm; per tensor: int64 numel, int32 name_len, name bytes;
//                          T initial parameter arrays (fp32); steps x T gradient arrays (fp32)
// out.bin:                 per step: float lr, clip_scale, global_norm, applied; T x (param, accum, m, v) arrays
#include <cuda_runtime_api.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gaccum.h"

#define CHECK_CUDA(x)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define CHECK_G(x)                                                                          \
  do {                                                                                      \
    int rc_ = (x);                                                                          \
    if (rc_ != GACCUM_OK) { std::fprintf(stderr, "%s -> %d: %s\n", #x, rc_, gaccum_last_error()); return 3; } \
  } while (0)

template <typename T>
static bool rd(FILE* f, T* v, size_t n = 1) { return std::fread(v, sizeof(T), n, f) == n; }

static int host_only() {
  if (gaccum_version() != GACCUM_VERSION) { std::fprintf(stderr, "header %d vs library %d\n", GACCUM_VERSION, gaccum_version()); return 1; }
  // optimization.py:29-54 at a few points; :91 predicate; :179-194 mask
  const float lr0 = gaccum_learning_rate(2e-5, 207900, 20790, 0);
  const float lr1 = gaccum_learning_rate(2e-5, 207900, 20790, 20790);
  if (lr0 != 0.0f || !(lr1 > 1.79e-5f && lr1 < 1.81e-5f)) { std::fprintf(stderr, "schedule: %g %g\n", lr0, lr1); return 1; }
  if (!gaccum_is_apply_step(0, 4) || gaccum_is_apply_step(3, 4) || !gaccum_is_apply_step(8, 4)) return 1;
  const char* names[] = {"bert/encoder/layer_0/output/dense/kernel:0", "bert/encoder/layer_0/output/dense/bias:0",
                         "bert/embeddings/LayerNorm/gamma:0"};
  const char* excl[] = {"LayerNorm", "layer_norm", "bias"};
  uint8_t mask[3] = {9, 9, 9};
  if (gaccum_decay_mask(3, names, 0.01, excl, 3, mask) != GACCUM_OK || mask[0] != 1 || mask[1] != 0 || mask[2] != 0) return 1;
  // layout-only plan: offsets are multiples of 32 elements; compute must be refused (no CPU fallback)
  const int64_t numels[3] = {100, 7, 2048};
  gaccum_hparams hp = {GACCUM_ADAM_WEIGHT_DECAY, 0, 0.9, 0.999, 1e-6, 0.01, 1.0};
  gaccum_plan* plan = nullptr;
  if (gaccum_plan_create(&plan, 3, numels, mask, &hp, -1) != GACCUM_OK) { std::fprintf(stderr, "%s\n", gaccum_last_error()); return 1; }
  int64_t off[3];
  if (gaccum_offsets(plan, off) != GACCUM_OK || off[0] != 0 || off[1] != 128 || off[2] != 160 || gaccum_padded_size(plan) != 160 + 2048) return 1;
  if (gaccum_num_elements(plan) != 2155 || gaccum_algorithmic_bytes(plan, 1) != 36 * 2155 || gaccum_algorithmic_bytes(plan, 0) != 12 * 2155) return 1;
  gaccum_step_args a = {0, 4, 0, 1e-3f, 0.9f, 0.999f, 0.0f};
  float* fake = reinterpret_cast<float*>(uintptr_t(0x1000));
  const float* g[3] = {fake, fake, fake};
  float* p[3] = {fake, fake, fake};
  if (gaccum_step(plan, g, p, fake, fake, fake, &a, nullptr) != GACCUM_ENODEVICE) { std::fprintf(stderr, "layout-only plan computed?!\n"); return 1; }
  if (std::strstr(gaccum_last_error(), "no CPU fallback") == nullptr) return 1;
  gaccum_plan_destroy(plan);
  std::printf("host-only ok (libgaccum %d)\n", gaccum_version());
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 2 && std::string(argv[1]) == "--host-only") return host_only();
  if (argc != 3) { std::fprintf(stderr, "usage: %s --host-only | <in.bin> <out.bin>\n", argv[0]); return 64; }
  FILE* in = std::fopen(argv[1], "rb");
  if (!in) { std::perror(argv[1]); return 64; }
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
  if (gaccum_device_count() < 1) { std::fprintf(stderr, "no CUDA device: libgaccum has no CPU fallback\n"); return 77; }
  CHECK_CUDA(cudaSetDevice(0));

  // ---- what the op's constructor does: decay mask from variable names, plan, slabs --------------------------------
  std::vector<const char*> cnames(T);
  for (int t = 0; t < T; ++t) cnames[t] = names[t].c_str();
  const char* excl[] = {"LayerNorm", "layer_norm", "bias"};
  std::vector<uint8_t> decay(T, 0);
  gaccum_hparams hp;
  if (variant == GACCUM_ADAM_WEIGHT_DECAY) {
    hp = {GACCUM_ADAM_WEIGHT_DECAY, 0, 0.9, 0.999, 1e-6, 0.01, clip_norm};              // optimization.py:59-65, 84
    CHECK_G(gaccum_decay_mask(T, cnames.data(), hp.weight_decay_rate, excl, 3, decay.data()));
  } else {
    hp = {GACCUM_ADAM, 0, 0.9, 0.999, 1e-8, 0.0, 0.0};                                   // tf.train.AdamOptimizer defaults
  }
  gaccum_plan* plan = nullptr;
  CHECK_G(gaccum_plan_create(&plan, T, numel.data(), decay.data(), &hp, 0));
  const int64_t padded = gaccum_padded_size(plan);
  std::vector<int64_t> off(T);
  CHECK_G(gaccum_offsets(plan, off.data()));
  float *accum = nullptr, *m = nullptr, *v = nullptr;
  const size_t slab_bytes = (size_t)(padded > 0 ? padded : 32) * sizeof(float);
  CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&accum), slab_bytes)); CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&m), slab_bytes)); CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&v), slab_bytes));
  CHECK_CUDA(cudaMemset(accum, 0, slab_bytes)); CHECK_CUDA(cudaMemset(m, 0, slab_bytes)); CHECK_CUDA(cudaMemset(v, 0, slab_bytes));
  // ---- the framework's tensors: one allocation per variable / per gradient (scattered) ---------------------------
  std::vector<float*> params(T), grads(T);
  std::vector<float> host;
  for (int t = 0; t < T; ++t) {
    const size_t bytes = (size_t)(numel[t] > 0 ? numel[t] : 1) * sizeof(float);
    CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&params[t]), bytes)); CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&grads[t]), bytes));
    host.resize((size_t)numel[t]);
    if (numel[t] && !rd(in, host.data(), (size_t)numel[t])) return 64;
    CHECK_CUDA(cudaMemcpy(params[t], host.data(), (size_t)numel[t] * sizeof(float), cudaMemcpyHostToDevice));
  }
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));        // TF's compute stream is not the default stream
  gaccum_stats* stats_host = nullptr;
  CHECK_CUDA(cudaMallocHost(reinterpret_cast<void**>(&stats_host), sizeof(gaccum_stats)));
  FILE* out = std::fopen(argv[2], "wb");
  if (!out) { std::perror(argv[2]); return 64; }

  // an error path must not poison the plan: NULL gradient table -> GACCUM_EINVAL with a message
  gaccum_step_args bad = {0, N, 0, 0.f, 0.9f, 0.999f, 0.f};
  if (gaccum_step(plan, nullptr, params.data(), accum, m, v, &bad, stream) != GACCUM_EINVAL || !*gaccum_last_error()) {
    std::fprintf(stderr, "expected GACCUM_EINVAL for a NULL gradient table\n");
    return 4;
  }

  // ---- what Compute() does, once per session.run(train_op) ----------------------------------------------------------
  float b1p = 0.9f, b2p = 0.999f;                                               // TF1 Adam non-slot variables
  int64_t global_step = 0;
  std::vector<const float*> gptr(T);
  for (int s = 0; s < steps; ++s) {
    for (int t = 0; t < T; ++t) {
      host.resize((size_t)numel[t]);
      if (numel[t] && !rd(in, host.data(), (size_t)numel[t])) return 64;
      CHECK_CUDA(cudaMemcpyAsync(grads[t], host.data(), (size_t)numel[t] * sizeof(float), cudaMemcpyHostToDevice, stream));
      CHECK_CUDA(cudaStreamSynchronize(stream));                                 // `host` is reused: the test harness, not the op, syncs here
      gptr[t] = grads[t];
    }
    gaccum_step_args a;
    a.global_step = global_step;                                                // pre-increment value (optimization.py:77)
    a.accum_n = N;
    a.reserved = 0;
    a.lr = variant == GACCUM_ADAM_WEIGHT_DECAY ? gaccum_learning_rate(init_lr, train_steps, warmup_steps, global_step) : (float)init_lr;
    a.beta1_power = b1p; a.beta2_power = b2p; a.reserved2 = 0.f;
    CHECK_G(gaccum_step(plan, gptr.data(), params.data(), accum, m, v, &a, stream));     // asynchronous, one kernel
    CHECK_G(gaccum_read_stats(plan, stats_host, stream));
    if (gaccum_is_apply_step(global_step, N) && variant == GACCUM_ADAM) { b1p *= 0.9f; b2p *= 0.999f; }
    ++global_step;                                                              // optimization.py:102-103
    // ---- test harness: dump the state after this micro-step ----
    CHECK_CUDA(cudaStreamSynchronize(stream));
    const float hdr[4] = {stats_host->lr, stats_host->clip_scale, stats_host->global_norm, stats_host->applied};
    std::fwrite(hdr, sizeof(float), 4, out);
    for (int t = 0; t < T; ++t) {
      const size_t n = (size_t)numel[t];
      host.resize(n);
      const float* src[4] = {params[t], accum + off[t], m + off[t], v + off[t]};
      for (int k = 0; k < 4; ++k) {
        if (n) CHECK_CUDA(cudaMemcpy(host.data(), src[k], n * sizeof(float), cudaMemcpyDeviceToHost));
        std::fwrite(host.data(), sizeof(float), n, out);
      }
    }
  }
  std::fclose(out);
  std::fclose(in);
  CHECK_G(gaccum_plan_destroy(plan));
  for (int t = 0; t < T; ++t) { cudaFree(params[t]); cudaFree(grads[t]); }
  cudaFree(accum); cudaFree(m); cudaFree(v); cudaFreeHost(stats_host); cudaStreamDestroy(stream);
  std::printf("abi_consumer ok: %d tensors, %d micro-steps, N=%d, variant %d\n", T, steps, N, variant);
  return 0;
}
