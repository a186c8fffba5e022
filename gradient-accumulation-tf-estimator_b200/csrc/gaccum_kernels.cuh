// gaccum_kernels.cuh -- sm_100a kernels of the gradient-accumulation train_op.
//
// All kernels grid-stride over a static tile table (<= 2048 elements of one tensor per tile), so the
// work split -- and therefore every reduction -- is deterministic.
//
//   accumulate_kernel    a += G                                  optimization.py:81,93    (12 B/elem)
//                        one tile per CTA (hardware block scheduler)
//   apply_kernel         the apply branch WITHOUT clipping (plain Adam of the example scripts, or
//                        clip_norm <= 0): a' = a + G; n = a'/N; Adam; a = 0 in a single pass (36 B/elem).
//                        Its CLIP=true instantiation is the first two-pass version, kept behind
//                        GACCUM_TUNE for A/B runs.
//   apply_clip2_kernel   the apply branch WITH tf.clip_by_global_norm                  (36 B/elem)
//                        optimization.py:80-88, 128-177.  The global norm of ALL tensors is needed before
//                        ANY element can be updated, so it is one cooperative launch with two passes
//                        around a grid barrier:
//                          pass 1  stream G and a, a' = a + G, reduce sum((a'/N)^2)
//                                  (thread fp32 per tile -> fp64 running sum -> warp shuffle -> shared
//                                  memory -> one fp64 partial per CTA); a' is PARKED ON CHIP: oldest
//                                  tiles in shared memory, next in Tensor Memory (tcgen05.st), the rest
//                                  written back in place tagged L2::evict_last
//                          barrier every CTA adds the per-CTA partials in the same order
//                                  (bit-identical gn and clip scale everywhere)
//                          pass 2  tiles in REVERSE order (youngest a' lines are still in L2, the oldest
//                                  never left the SM): clip, AdamWeightDecay/Adam, write p, m, v, a = 0
// Arithmetic uses round-to-nearest intrinsics (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn) so nvcc
// cannot contract mul+add into FMA: the reference graph is un-fused, one rounding per TF op, and we
// reproduce it bit for bit (the kernels are HBM-bound, the extra flops are free).
// Measurements and the experiments behind each choice: profiles/r01_tune_sweep.md.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gaccum {
namespace cg = cooperative_groups;

constexpr int kThreads = 256;                    // 8 warps per CTA
constexpr int kUnroll = 2;                       // 128-bit vectors per thread per stream per tile
constexpr int kTile = kThreads * 4 * kUnroll;    // 2048 elements = 8 KB per stream
constexpr int kSlabAlign = 32;                   // tensors start at multiples of 32 elements (128 B)

// One unit of work: <= kTile consecutive elements of ONE tensor (so pointers and the decay flag
// are tile-uniform).  16 bytes, read with a single LDG.128.
struct __align__(16) TileDesc {
  uint32_t tensor_flags;   // bits 0..30 tensor index, bit 31 = apply weight decay (optimization.py:166)
  uint32_t len;            // elements in this tile (1..kTile)
  uint32_t toff;           // element offset inside the tensor (multiple of kTile)
  uint32_t soff32;         // element offset inside the slabs, in units of 32 elements
};

// Scattered inputs arrive as a pointer table that lives in the kernel-parameter (constant)
// space: no device-side table to keep in sync, launches stay re-entrant and graph-capturable.
// CAP = 0 is the packed layout (grads/params are slabs with the same offsets as accum).
template <int CAP>
struct PtrTable {
  const float* g[CAP];
  float* p[CAP];
};
template <>
struct PtrTable<0> {
  const float* g;
  float* p;
};

struct Scalars {
  float nf;      // fp32(N)                                   optimization.py:83
  float inv_nf;  // 1/N when N is a power of two (exact), else 0 -> the kernels divide
  float lr;      // learning rate of this micro-step          optimization.py:29-54
  float b1, b2;  // fp32(beta)                                optimization.py:151,153
  float omb1;    // A: fp32(1.0 - beta1) from double (:152);  B: 1.0f - fp32(beta1)
  float omb2;
  float eps;     //                                           optimization.py:157
  float wd;      //                                           optimization.py:167
  float clip;    //                                           optimization.py:84
  float alpha;   // B only: lr*sqrt(1-b2^t)/(1-b1^t)          TF1 ApplyAdam
};

template <int CAP>
struct KernelParams {
  const TileDesc* tiles;
  int32_t num_tiles;
  float* accum;
  float* m;
  float* v;
  double* partials;   // one per CTA (apply with clip)
  float* tile_sumsq;  // one per tile (dynamic apply): sum((a'/N)^2) of that tile
  uint32_t* tickets;  // [0] pass-1 ticket counter, [1] pass-2 ticket counter (zero between launches)
  float* stats;       // gaccum_stats
  uint32_t tune;      // kTune* bits (cache-policy experiments; uniform branches)
  int32_t stash_tiles;  // apply_clip2_kernel: tiles of a' each CTA keeps in shared memory between the passes
  int32_t tmem_tiles;   // ... and in Tensor Memory (0 or kTmemTiles)
  unsigned long long* debug;  // GACCUM_EXPERIMENTS: 4 timestamps (ns) per CTA, else nullptr
  Scalars sc;
  PtrTable<CAP> tab;
};

constexpr uint32_t kTuneKeepA = 1u;      // pass 1: a / a' lines get L2 evict_last priority
constexpr uint32_t kTuneStreamState = 2u;  // pass 2: p, m, v (and the spent a') move with evict-first
constexpr uint32_t kTuneAccTiles = 4u;   // accumulate: one tile per CTA (hardware scheduler) instead of persistent
constexpr uint32_t kTuneStaticApply = 8u;  // (retired: the warp-ticket dynamic apply was slower and was removed; bit kept for sweep numbering)
constexpr uint32_t kTuneSkipPass1 = 16u;   // TIMING EXPERIMENTS ONLY (results are wrong): skip the norm pass
constexpr uint32_t kTuneSkipPass2 = 32u;   // TIMING EXPERIMENTS ONLY: skip the update pass
constexpr uint32_t kTuneOwnBarrier = 128u;  // clip-apply v2: ordinary launch + atomic grid barrier (no cooperative launch)
constexpr uint32_t kTuneApplyV1 = 256u;     // clip-apply: use the first two-pass kernel (no on-chip stash)
constexpr uint32_t kTuneTmemStash = 512u;   // clip-apply v2: also park a' tiles in Tensor Memory (tcgen05.st / tcgen05.ld)
constexpr uint32_t kTunePrefetch = 1024u;   // clip-apply v2: L2 software prefetch of the next tiles in pass 1
constexpr int kPrefetchDistance = 1;        // iterations ahead
constexpr uint32_t kAccAssign = 1u << 30;    // accumulate_kernel stores G instead of adding it (host-session gather of small tensors)
constexpr uint32_t kTuneSkipZero = 64u;    // TIMING EXPERIMENTS ONLY (dp kernel): skip zeroing non-owned tiles

// ---------------------------------------------------------------------------------------------
// memory helpers: G is read exactly once -> streaming (evict-first) loads; zeroing the
// accumulator is a streaming store.  a' must survive in L2 from pass 1 to pass 2, so it can be
// tagged evict_last while everything that is touched once is tagged evict-first.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
#ifdef GACCUM_G_NC
  float4 v;   // G is never written by these kernels: read-only path, no L1 allocation
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_policy(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_policy(float4* p, const float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

template <int CAP>
__device__ __forceinline__ const float* grad_ptr(const PtrTable<CAP>& tab, const TileDesc& d) {
  if constexpr (CAP == 0) {
    return tab.g ? tab.g + (size_t)d.soff32 * kSlabAlign : nullptr;
  } else {
    const float* b = tab.g[d.tensor_flags & 0x7fffffffu];
    return b ? b + d.toff : nullptr;
  }
}
template <int CAP>
__device__ __forceinline__ float* param_ptr(const PtrTable<CAP>& tab, const TileDesc& d) {
  if constexpr (CAP == 0) {
    return tab.p + (size_t)d.soff32 * kSlabAlign;
  } else {
    return tab.p[d.tensor_flags & 0x7fffffffu] + d.toff;
  }
}
__device__ __forceinline__ bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// per-element math, one rounding per reference op
// ---------------------------------------------------------------------------------------------
// optimization.py:83  (1.0*a)/N  -- the multiply by 1.0 is exact.  When N is a power of two, 1/N is
// exact and a * (1/N) is the same correctly rounded real number as a / N (bit-identical, subnormals
// included), which saves a ~12-instruction IEEE division per element; otherwise divide.
__device__ __forceinline__ float normalize(float a, float nf, float inv_nf) {
  return inv_nf != 0.f ? __fmul_rn(a, inv_nf) : __fdiv_rn(a, nf);
}

template <int VARIANT>
__device__ __forceinline__ void adam_elem(float c, float& p, float& m, float& v, bool decay,
                                          const Scalars& sc) {
  if constexpr (VARIANT == 0) {
    // optimization.py:151-171
    const float m2 = __fadd_rn(__fmul_rn(sc.b1, m), __fmul_rn(sc.omb1, c));
    const float v2 = __fadd_rn(__fmul_rn(sc.b2, v), __fmul_rn(sc.omb2, __fmul_rn(c, c)));
    float u = __fdiv_rn(m2, __fadd_rn(__fsqrt_rn(v2), sc.eps));
    if (decay) u = __fadd_rn(u, __fmul_rn(sc.wd, p));
    p = __fsub_rn(p, __fmul_rn(sc.lr, u));
    m = m2;
    v = v2;
  } else {
    // TF1 ApplyAdam: m += (g-m)(1-b1); v += (g*g-v)(1-b2); var -= (m*alpha)/(sqrt(v)+eps)
    const float m2 = __fadd_rn(m, __fmul_rn(__fsub_rn(c, m), sc.omb1));
    const float v2 = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(c, c), v), sc.omb2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m2, sc.alpha), __fadd_rn(__fsqrt_rn(v2), sc.eps)));
    m = m2;
    v = v2;
  }
}

// tf.clip_by_global_norm (TF 1.15): scale = clip * min(1/gn, 1/clip) + (gn - gn)
__device__ __forceinline__ float clip_scale(float gn, float clip) {
  const float inv = __fdiv_rn(1.0f, gn);
  const float invc = __fdiv_rn(1.0f, clip);
  float mn = inv < invc ? inv : invc;
  if (inv != inv) mn = inv;
  return __fadd_rn(__fmul_rn(clip, mn), __fsub_rn(gn, gn));
}

// ---------------------------------------------------------------------------------------------
// accumulate: a += G                                                   optimization.py:81,93
// ---------------------------------------------------------------------------------------------
template <int CAP>
__device__ __forceinline__ void accumulate_tile(const TileDesc d, const KernelParams<CAP>& prm) {
  const float* __restrict__ g = grad_ptr(prm.tab, d);
  if (g == nullptr) return;   // optimization.py:132 -- tensors without a gradient are skipped
  float* __restrict__ a = prm.accum + (size_t)d.soff32 * kSlabAlign;
  const uint32_t len = d.len, tid = threadIdx.x;
  if (prm.tune & kAccAssign) {              // gather: a = G (G may live in pinned host memory)
    for (uint32_t i = tid; i < len; i += kThreads) a[i] = ld_stream(g + i);
    return;
  }
  if (aligned16(g)) {
    const uint32_t nvec = len >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* a4 = reinterpret_cast<float4*>(a);
    float4 vg[kUnroll], va[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) { vg[u] = ld_stream(g4 + i); va[u] = a4[i]; }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        va[u].x = __fadd_rn(va[u].x, vg[u].x); va[u].y = __fadd_rn(va[u].y, vg[u].y);
        va[u].z = __fadd_rn(va[u].z, vg[u].z); va[u].w = __fadd_rn(va[u].w, vg[u].w);
        a4[i] = va[u];
      }
    }
    const uint32_t i = (nvec << 2) + tid;      // < 4 trailing elements
    if (i < len) a[i] = __fadd_rn(a[i], ld_stream(g + i));
  } else {
    for (uint32_t i = tid; i < len; i += kThreads) a[i] = __fadd_rn(a[i], ld_stream(g + i));
  }
}

template <int CAP>
__global__ void __launch_bounds__(kThreads)
accumulate_kernel(const __grid_constant__ KernelParams<CAP> prm) {
  int t = blockIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0 && !(prm.tune & kAccAssign)) {   // the gather pass is not a step
    prm.stats[0] = 0.f; prm.stats[1] = prm.sc.lr; prm.stats[2] = 0.f; prm.stats[3] = 1.f;
  }
  if (t >= prm.num_tiles) return;
  TileDesc d = prm.tiles[t];
  while (true) {
    const int tn = t + gridDim.x;
    TileDesc dn;
    if (tn < prm.num_tiles) dn = prm.tiles[tn];   // prefetch the next descriptor
    accumulate_tile(d, prm);
    if (tn >= prm.num_tiles) break;
    t = tn; d = dn;
  }
}

// ---------------------------------------------------------------------------------------------
// apply, pass 1 (clip only): a' = a + G written back, returns acc + sum((a'/N)^2) over the tile
// ---------------------------------------------------------------------------------------------
template <bool HAS_G, int CAP>
__device__ __forceinline__ float norm_tile(const TileDesc d, const KernelParams<CAP>& prm, float acc) {
  const float* __restrict__ g = nullptr;
  if constexpr (HAS_G) g = grad_ptr(prm.tab, d);
  float* __restrict__ a = prm.accum + (size_t)d.soff32 * kSlabAlign;
  const uint32_t len = d.len, tid = threadIdx.x;
  const float nf = prm.sc.nf, inv_nf = prm.sc.inv_nf;
  const bool keep = (prm.tune & kTuneKeepA) != 0;
  const uint64_t pol = policy_evict_last();
  if (g == nullptr || aligned16(g)) {
    const uint32_t nvec = len >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* a4 = reinterpret_cast<float4*>(a);
    float4 vg[kUnroll], va[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) { va[u] = keep ? ld_policy(a4 + i, pol) : a4[i]; if (g) vg[u] = ld_stream(g4 + i); }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        if (g) {
          va[u].x = __fadd_rn(va[u].x, vg[u].x); va[u].y = __fadd_rn(va[u].y, vg[u].y);
          va[u].z = __fadd_rn(va[u].z, vg[u].z); va[u].w = __fadd_rn(va[u].w, vg[u].w);
          if (keep) st_policy(a4 + i, va[u], pol); else a4[i] = va[u];
        }
        const float nx = normalize(va[u].x, nf, inv_nf), ny = normalize(va[u].y, nf, inv_nf),
                    nz = normalize(va[u].z, nf, inv_nf), nw = normalize(va[u].w, nf, inv_nf);
        acc = fmaf(nx, nx, acc); acc = fmaf(ny, ny, acc); acc = fmaf(nz, nz, acc); acc = fmaf(nw, nw, acc);
      }
    }
    const uint32_t i = (nvec << 2) + tid;
    if (i < len) {
      float x = a[i];
      if (g) { x = __fadd_rn(x, ld_stream(g + i)); a[i] = x; }
      const float n = normalize(x, nf, inv_nf);
      acc = fmaf(n, n, acc);
    }
  } else {
    for (uint32_t i = tid; i < len; i += kThreads) {
      const float x = __fadd_rn(a[i], ld_stream(g + i));
      a[i] = x;
      const float n = normalize(x, nf, inv_nf);
      acc = fmaf(n, n, acc);
    }
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// apply, update pass.  LOAD_G: single-pass mode (no clip) adds G here; after pass 1 it is false.
// ---------------------------------------------------------------------------------------------
template <int VARIANT, bool CLIP, bool LOAD_G, int CAP>
__device__ __forceinline__ void update_tile(const TileDesc d, const KernelParams<CAP>& prm, const float s) {
  const float* __restrict__ g = nullptr;
  if constexpr (LOAD_G) g = grad_ptr(prm.tab, d);
  const size_t soff = (size_t)d.soff32 * kSlabAlign;
  float* __restrict__ a = prm.accum + soff;
  float* __restrict__ m = prm.m + soff;
  float* __restrict__ v = prm.v + soff;
  float* __restrict__ p = param_ptr(prm.tab, d);
  const bool decay = (d.tensor_flags >> 31) != 0;
  const uint32_t len = d.len, tid = threadIdx.x;
  const Scalars& sc = prm.sc;

  auto elem = [&](float ax, float gx, float& px, float& mx, float& vx) {
    if (LOAD_G) ax = __fadd_rn(ax, gx);               // optimization.py:81 (gx = 0 never used: see callers)
    float c = normalize(ax, sc.nf, sc.inv_nf);                    // :83
    if (CLIP) c = __fmul_rn(c, s);                     // :84
    adam_elem<VARIANT>(c, px, mx, vx, decay, sc);      // :85
  };

  if (aligned16(p) && (g == nullptr || aligned16(g))) {
    const uint32_t nvec = len >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* a4 = reinterpret_cast<float4*>(a);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* p4 = reinterpret_cast<float4*>(p);
    float4 va[kUnroll], vg[kUnroll], vp[kUnroll], vm[kUnroll], vv[kUnroll];
    const bool strm = (prm.tune & kTuneStreamState) != 0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        if (strm) { va[u] = __ldcs(a4 + i); vp[u] = __ldcs(p4 + i); vm[u] = __ldcs(m4 + i); vv[u] = __ldcs(v4 + i); }
        else { va[u] = a4[i]; vp[u] = p4[i]; vm[u] = m4[i]; vv[u] = v4[i]; }
        if (LOAD_G && g) vg[u] = ld_stream(g4 + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        if (LOAD_G && g) {
          elem(va[u].x, vg[u].x, vp[u].x, vm[u].x, vv[u].x); elem(va[u].y, vg[u].y, vp[u].y, vm[u].y, vv[u].y);
          elem(va[u].z, vg[u].z, vp[u].z, vm[u].z, vv[u].z); elem(va[u].w, vg[u].w, vp[u].w, vm[u].w, vv[u].w);
        } else {
          // no gradient for this tile: a + 0 would turn -0 into +0 only; skip the add entirely
          auto e0 = [&](float ax, float& px, float& mx, float& vx) {
            float c = normalize(ax, sc.nf, sc.inv_nf);
            if (CLIP) c = __fmul_rn(c, s);
            adam_elem<VARIANT>(c, px, mx, vx, decay, sc);
          };
          e0(va[u].x, vp[u].x, vm[u].x, vv[u].x); e0(va[u].y, vp[u].y, vm[u].y, vv[u].y);
          e0(va[u].z, vp[u].z, vm[u].z, vv[u].z); e0(va[u].w, vp[u].w, vm[u].w, vv[u].w);
        }
        if (strm) { __stcs(p4 + i, vp[u]); __stcs(m4 + i, vm[u]); __stcs(v4 + i, vv[u]); }
        else { p4[i] = vp[u]; m4[i] = vm[u]; v4[i] = vv[u]; }
        __stcs(a4 + i, make_float4(0.f, 0.f, 0.f, 0.f));   // optimization.py:86-87
      }
    }
    const uint32_t i = (nvec << 2) + tid;
    if (i < len) {
      float ax = a[i], px = p[i], mx = m[i], vx = v[i];
      if (LOAD_G && g) ax = __fadd_rn(ax, ld_stream(g + i));
      float c = normalize(ax, sc.nf, sc.inv_nf);
      if (CLIP) c = __fmul_rn(c, s);
      adam_elem<VARIANT>(c, px, mx, vx, decay, sc);
      p[i] = px; m[i] = mx; v[i] = vx; a[i] = 0.f;
    }
  } else {
    for (uint32_t i = tid; i < len; i += kThreads) {
      float ax = a[i], px = p[i], mx = m[i], vx = v[i];
      if (LOAD_G && g) ax = __fadd_rn(ax, ld_stream(g + i));
      float c = normalize(ax, sc.nf, sc.inv_nf);
      if (CLIP) c = __fmul_rn(c, s);
      adam_elem<VARIANT>(c, px, mx, vx, decay, sc);
      p[i] = px; m[i] = mx; v[i] = vx; a[i] = 0.f;
    }
  }
}

// Deterministic CTA reduction of one double per thread -> total in thread 0.  Threads add each
// tile's 8-element fp32 partial into an fp64 running sum, so the norm of a 335 M-element model is
// good to ~1e-7 relative even for adversarial (constant) data.
__device__ __forceinline__ double block_reduce_to_double(double x, double* smem /* blockDim.x/32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = x;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (int)blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) tot += smem[w];
  }
  return tot;
}

template <int VARIANT, bool CLIP, bool HAS_G, int CAP>
__global__ void __launch_bounds__(kThreads)
apply_kernel(const __grid_constant__ KernelParams<CAP> prm) {
  __shared__ double red[kThreads / 32];
  __shared__ float s_bcast[2];
  const int nt = prm.num_tiles;
  float s = 1.0f, gn = 0.0f;

  if constexpr (CLIP) {
    // ---- pass 1: a' = a + G (written back), sum of squares of a'/N -----------------------
    double acc = 0.0;
    int t = blockIdx.x;
    if (t < nt && !(prm.tune & kTuneSkipPass1)) {
      TileDesc d = prm.tiles[t];
      while (true) {
        const int tn = t + gridDim.x;
        TileDesc dn;
        if (tn < nt) dn = prm.tiles[tn];
        acc += (double)norm_tile<HAS_G>(d, prm, 0.f);
        if (tn >= nt) break;
        t = tn; d = dn;
      }
    }
    const double part = block_reduce_to_double(acc, red);
    if (threadIdx.x == 0) prm.partials[blockIdx.x] = part;
    cg::this_grid().sync();
    // ---- every CTA combines the per-CTA partials in the same fixed order -----------------
    if (threadIdx.x < 32) {
      double tot = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) tot += __ldcg(prm.partials + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (threadIdx.x == 0) {
        // tf.linalg.global_norm: sqrt(2 * sum_i l2_loss(n_i)) == sqrt(sum n^2), fp32
        const float g_norm = __fsqrt_rn((float)tot);
        s_bcast[0] = clip_scale(g_norm, prm.sc.clip);
        s_bcast[1] = g_norm;
      }
    }
    __syncthreads();
    s = s_bcast[0]; gn = s_bcast[1];
    // ---- pass 2, reverse order: most recently written a' lines first (L2 hits) ----------
    if (blockIdx.x < nt && !(prm.tune & kTuneSkipPass2)) {
      int t2 = blockIdx.x + ((nt - 1 - blockIdx.x) / gridDim.x) * gridDim.x;   // my last tile
      TileDesc d = prm.tiles[t2];
      while (true) {
        const int tn = t2 - (int)gridDim.x;
        TileDesc dn;
        if (tn >= 0) dn = prm.tiles[tn];
        update_tile<VARIANT, true, false>(d, prm, s);
        if (tn < 0) break;
        t2 = tn; d = dn;
      }
    }
  } else {
    // ---- no clipping: one pass, 36 B/elem ------------------------------------------------
    int t = blockIdx.x;
    if (t < nt) {
      TileDesc d = prm.tiles[t];
      while (true) {
        const int tn = t + gridDim.x;
        TileDesc dn;
        if (tn < nt) dn = prm.tiles[tn];
        update_tile<VARIANT, false, HAS_G>(d, prm, 1.0f);
        if (tn >= nt) break;
        t = tn; d = dn;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    prm.stats[0] = 1.f; prm.stats[1] = prm.sc.lr; prm.stats[2] = gn; prm.stats[3] = s;
  }
}


// =============================================================================================
// apply with clipping, v2: the same static two-pass schedule as apply_kernel, plus
//   * an ON-CHIP STASH: the first `stash_tiles` tiles a CTA reduces in pass 1 never leave the SM --
//     a' goes to shared memory instead of HBM/L2 and pass 2 (which walks the CTA's tiles in reverse,
//     i.e. reaches them last, when L2 would long have evicted them) consumes it from there.  3 CTAs x 9 tiles x 8 KB x 148 SMs = 32 MB of the 115 MB a'
//     slab at BERT-Small; what is left competes for far fewer L2 lines.
//   * optionally an ordinary launch with an atomic grid barrier instead of a cooperative launch
//     (kTuneOwnBarrier): the grid never exceeds the co-resident capacity, so the barrier cannot
//     deadlock on an otherwise idle device, and ~6 us of cooperative-launch overhead go away.
// Thread t of a CTA reads back exactly the shared-memory words it wrote, so the stash needs no
// synchronisation and is bank-conflict free (consecutive lanes, consecutive 16-byte words).
// =============================================================================================
// ---------------------------------------------------------------------------------------------
// Tensor Memory as a scratchpad.  TMEM (256 KB per SM, 512 columns x 128 lanes x 32 bit) normally
// holds tcgen05.mma accumulators; this kernel has no MMA, so it is idle silicon -- 37 MB across the
// chip, more than the shared-memory stash.  A kernel that touches TMEM is limited to ONE CTA per
// SM by the driver, so the TMEM variant runs 768-thread CTAs made of three 256-thread groups that
// behave exactly like the three co-resident CTAs of the plain variant (virtual block id =
// blockIdx * 3 + group).  The CTA allocates all 512 columns; every warp parks a' values in the 32
// lanes it may address (lane quadrant = warp % 4; the 6 warps sharing a quadrant take 80 columns each):
// tcgen05.st 32x32b.x8 writes the thread's 8 words of a tile to 8 consecutive columns of its own
// lane, tcgen05.ld reads them back in pass 2.  A thread only ever reads what it wrote itself.
// ---------------------------------------------------------------------------------------------
constexpr int kGroups3 = 3;                       // 768-thread CTA = three 256-thread groups, one CTA per SM
constexpr int kTmemCols = 512;                    // a TMEM-using kernel gets one CTA per SM: take all columns
constexpr int kTmemColsPerWarp = 80;              // 6 warps share a lane quadrant: 6 x 80 = 480 <= 512
constexpr int kTmemTiles = kTmemColsPerWarp / 8;  // 10 tiles per group
constexpr uint32_t kNoTmem = 0xffffffffu;

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"((uint32_t)kTmemCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)kTmemCols) : "memory");
}
__device__ __forceinline__ void tmem_store8(uint32_t taddr, const float4& a, const float4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
                 "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)),
                 "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_load8(uint32_t taddr, float4& a, float4& b) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  a = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
  b = make_float4(__uint_as_float(r4), __uint_as_float(r5), __uint_as_float(r6), __uint_as_float(r7));
}
// TMEM address of this warp's slot for stashed tile `slot` (0..kTmemTiles-1): lane quadrant = warp % 4,
// column block = warp / 4 (0..5 across the three groups)
__device__ __forceinline__ uint32_t tmem_slot_addr(uint32_t base, int slot) {
  const uint32_t warp = threadIdx.x >> 5;
  return base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * kTmemColsPerWarp + (uint32_t)slot * 8u;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void atomic_grid_barrier(uint32_t* ctr, uint32_t target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_gpu(ctr) < target) { __nanosleep(32); }
  }
  __syncthreads();
}

// Pass 1 is split into "issue the loads" and "finish", so that the loads of several tiles can be in
// flight before the first one is consumed: with one tile at a time a thread has only 4 LDG.128
// outstanding (49 KB per SM), and the measured per-CTA timeline shows pass 1 latency-bound at
// 4.4 TB/s; with kPass1Tiles tiles the SM keeps >= 100 KB in flight and the pass becomes HBM-bound.
constexpr int kPass1Tiles = 2;
struct NormRegs {
  float4 va[kUnroll], vg[kUnroll];
  const float* g;
  bool vec;          // false: unaligned gradient pointer -> scalar path, nothing was loaded
};

// Software prefetch into L2 of a tile the group will reduce one iteration later: prefetch.global.L2
// needs no destination registers, so it adds bytes in flight without adding register pressure (the
// thing that kept the 24-warp pass 1 latency-bound).  Threads 0..63 cover the 64 lines of `a`,
// threads 64..127 those of G.
template <bool HAS_G, int CAP>
__device__ __forceinline__ void norm_prefetch(const TileDesc& d, const KernelParams<CAP>& prm) {
  const uint32_t tid = threadIdx.x & (kThreads - 1);
  const uint32_t line = tid & 63u;
  if (line * 32u >= d.len) return;
  const float* ptr = nullptr;
  if (tid < 64) ptr = prm.accum + (size_t)d.soff32 * kSlabAlign;
  else if (tid < 128) { if constexpr (HAS_G) ptr = grad_ptr(prm.tab, d); }
  if (ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + line * 32u));
}

template <bool HAS_G, int CAP>
__device__ __forceinline__ void norm_issue(const TileDesc& d, const KernelParams<CAP>& prm, const bool on_chip,
                                           const uint64_t pol, NormRegs& r) {
  r.g = nullptr;
  if constexpr (HAS_G) r.g = grad_ptr(prm.tab, d);
  r.vec = (r.g == nullptr || aligned16(r.g));
  if (!r.vec) return;
  const float4* g4 = reinterpret_cast<const float4*>(r.g);
  const float4* a4 = reinterpret_cast<const float4*>(prm.accum + (size_t)d.soff32 * kSlabAlign);
  const uint32_t nvec = d.len >> 2, tid = threadIdx.x & (kThreads - 1);   // thread index inside the 256-thread group
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t i = u * kThreads + tid;
    if (i < nvec) { r.va[u] = on_chip ? __ldcs(a4 + i) : ld_policy(a4 + i, pol); if (r.g) r.vg[u] = ld_stream(g4 + i); }
  }
}

template <bool HAS_G, int CAP, bool USE_TMEM>
__device__ __forceinline__ float norm_finish(const TileDesc& d, const KernelParams<CAP>& prm, NormRegs& r,
                                             float4* __restrict__ stash, const uint32_t tmem, const uint64_t pol) {
  static_assert(kUnroll == 2, "the TMEM stash moves exactly two float4 per thread per tile");
  const float* __restrict__ g = r.g;
  float* __restrict__ a = prm.accum + (size_t)d.soff32 * kSlabAlign;
  const uint32_t len = d.len, tid = threadIdx.x & (kThreads - 1);
  const float nf = prm.sc.nf, inv_nf = prm.sc.inv_nf;
  float acc = 0.f;
  if (r.vec) {
    const uint32_t nvec = len >> 2;
    float4* a4 = reinterpret_cast<float4*>(a);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        float4& x = r.va[u];
        if (g) {
          x.x = __fadd_rn(x.x, r.vg[u].x); x.y = __fadd_rn(x.y, r.vg[u].y);
          x.z = __fadd_rn(x.z, r.vg[u].z); x.w = __fadd_rn(x.w, r.vg[u].w);
        }
        if (stash) stash[i] = x;
        else if (tmem != kNoTmem) {}
        else if (g) st_policy(a4 + i, x, pol);
        const float nx = normalize(x.x, nf, inv_nf), ny = normalize(x.y, nf, inv_nf), nz = normalize(x.z, nf, inv_nf), nw = normalize(x.w, nf, inv_nf);
        acc = fmaf(nx, nx, acc); acc = fmaf(ny, ny, acc); acc = fmaf(nz, nz, acc); acc = fmaf(nw, nw, acc);
      }
    }
    if constexpr (USE_TMEM) {
      if (tmem != kNoTmem) tmem_store8(tmem, r.va[0], r.va[1]);   // full tile: every lane of every warp is here
    }
    const uint32_t i = (nvec << 2) + tid;      // < 4 tail elements always travel through global memory
    if (i < len) {
      float x = a[i];
      if (g) { x = __fadd_rn(x, ld_stream(g + i)); a[i] = x; }
      const float n = normalize(x, nf, inv_nf);
      acc = fmaf(n, n, acc);
    }
  } else {
    for (uint32_t i = tid; i < len; i += kThreads) {
      const float x = __fadd_rn(a[i], ld_stream(g + i));
      a[i] = x;
      const float n = normalize(x, nf, inv_nf);
      acc = fmaf(n, n, acc);
    }
  }
  return acc;
}

template <int VARIANT, int CAP, bool USE_TMEM>
__device__ __forceinline__ void update_tile2(const TileDesc d, const KernelParams<CAP>& prm, const float s,
                                             const float4* __restrict__ stash, const uint32_t tmem) {
  const size_t soff = (size_t)d.soff32 * kSlabAlign;
  float* __restrict__ a = prm.accum + soff;
  float* __restrict__ m = prm.m + soff;
  float* __restrict__ v = prm.v + soff;
  float* __restrict__ p = param_ptr(prm.tab, d);
  const bool decay = (d.tensor_flags >> 31) != 0;
  const uint32_t len = d.len, tid = threadIdx.x & (kThreads - 1);
  const Scalars& sc = prm.sc;
  auto elem = [&](float ax, float& px, float& mx, float& vx) {
    const float c = __fmul_rn(normalize(ax, sc.nf, sc.inv_nf), s);     // optimization.py:83-84
    adam_elem<VARIANT>(c, px, mx, vx, decay, sc);           // :85
  };
  if (aligned16(p)) {
    const uint32_t nvec = len >> 2;
    float4* a4 = reinterpret_cast<float4*>(a);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* p4 = reinterpret_cast<float4*>(p);
    float4 va[kUnroll], vp[kUnroll], vm[kUnroll], vv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        vp[u] = __ldcs(p4 + i); vm[u] = __ldcs(m4 + i); vv[u] = __ldcs(v4 + i);
        if (stash) va[u] = stash[i];
        else if (tmem == kNoTmem) va[u] = __ldcs(a4 + i);
      }
    }
    if constexpr (USE_TMEM) {
      if (tmem != kNoTmem) tmem_load8(tmem, va[0], va[1]);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        elem(va[u].x, vp[u].x, vm[u].x, vv[u].x); elem(va[u].y, vp[u].y, vm[u].y, vv[u].y);
        elem(va[u].z, vp[u].z, vm[u].z, vv[u].z); elem(va[u].w, vp[u].w, vm[u].w, vv[u].w);
        __stcs(p4 + i, vp[u]); __stcs(m4 + i, vm[u]); __stcs(v4 + i, vv[u]);
        __stcs(a4 + i, make_float4(0.f, 0.f, 0.f, 0.f));   // optimization.py:86-87
      }
    }
    const uint32_t i = (nvec << 2) + tid;
    if (i < len) {
      float px = p[i], mx = m[i], vx = v[i];
      elem(a[i], px, mx, vx);
      p[i] = px; m[i] = mx; v[i] = vx; a[i] = 0.f;
    }
  } else {
    for (uint32_t i = tid; i < len; i += kThreads) {
      float px = p[i], mx = m[i], vx = v[i];
      elem(a[i], px, mx, vx);
      p[i] = px; m[i] = mx; v[i] = vx; a[i] = 0.f;
    }
  }
}

// a tile may be stashed only if BOTH passes will take the vector path for it
template <bool HAS_G, int CAP>
__device__ __forceinline__ bool stashable(const TileDesc& d, const KernelParams<CAP>& prm) {
  bool ok = aligned16(param_ptr(prm.tab, d));
  if constexpr (HAS_G) { const float* g = grad_ptr(prm.tab, d); ok = ok && (g == nullptr || aligned16(g)); }
  return ok;
}

template <int VARIANT, bool HAS_G, int CAP, bool USE_TMEM>
__global__ void __launch_bounds__(kThreads * (USE_TMEM ? kGroups3 : 1))
apply_clip2_kernel(const __grid_constant__ KernelParams<CAP> prm) {
  constexpr int GROUPS = USE_TMEM ? kGroups3 : 1;
  extern __shared__ float4 stash_all[];                 // GROUPS x stash_tiles x (kTile/4) float4
  __shared__ double red[kThreads * GROUPS / 32];
  __shared__ float s_bcast[2];
  const int grp = (int)threadIdx.x / kThreads;          // 256-thread group = virtual CTA
  const int nt = prm.num_tiles, G = (int)gridDim.x * GROUPS, b = (int)blockIdx.x * GROUPS + grp;
  float4* const stash_mem = stash_all + (size_t)grp * prm.stash_tiles * (kTile / 4);
  const int my_count = b < nt ? (nt - 1 - b) / G + 1 : 0;    // tiles b, b+G, ... of this CTA
  // Stash the OLDEST tiles of pass 1 (k < stash_tiles): pass 2 runs in reverse, so the youngest a'
  // lines are still in L2 when they are needed, while the oldest would have been evicted long before
  // pass 2 reaches them -- shared memory and L2 cover complementary ends of the sequence.
  const int n_stashed = min(my_count, prm.stash_tiles);
  // Tensor Memory takes the next-oldest tiles (full, vector-path tiles only: tcgen05.st/ld are
  // warp-collective, every lane must carry data)
  __shared__ uint32_t s_tmem_base;
  const int n_tmem = USE_TMEM ? prm.tmem_tiles : 0;
  if (USE_TMEM && n_tmem > 0) {
    if (threadIdx.x < 32) tmem_alloc(&s_tmem_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t tmem_base = (USE_TMEM && n_tmem > 0) ? s_tmem_base : 0u;
  auto tmem_for = [&](int k, const TileDesc& d) -> uint32_t {
    if constexpr (!USE_TMEM) return kNoTmem;
    const int slot = k - n_stashed;
    if (slot < 0 || slot >= n_tmem || d.len != (uint32_t)kTile || !stashable<HAS_G>(d, prm)) return kNoTmem;
    return tmem_slot_addr(tmem_base, slot);
  };
  const uint64_t pol = policy_evict_last();
  const bool own_barrier = (prm.tune & kTuneOwnBarrier) != 0;

  auto stamp = [&](int which) {
    if (prm.debug && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      prm.debug[blockIdx.x * 4 + which] = t;
    }
  };
  stamp(0);
  // ---- pass 1: kPass1Tiles tiles per iteration, all loads issued before the first is consumed ----
  double acc = 0.0;
  for (int k0 = 0; k0 < my_count; k0 += kPass1Tiles) {
    TileDesc d[kPass1Tiles];
    NormRegs r[kPass1Tiles];
    float4* st[kPass1Tiles];
    uint32_t tm[kPass1Tiles];
#pragma unroll
    for (int j = 0; j < kPass1Tiles; ++j) {
      const int k = k0 + j;
      if (k < my_count) {
        d[j] = prm.tiles[b + k * G];
        st[j] = (k < n_stashed && stashable<HAS_G>(d[j], prm)) ? stash_mem + (size_t)k * (kTile / 4) : nullptr;
        tm[j] = st[j] ? kNoTmem : tmem_for(k, d[j]);
        norm_issue<HAS_G>(d[j], prm, st[j] != nullptr || tm[j] != kNoTmem, pol, r[j]);
      }
    }
    if (prm.tune & kTunePrefetch) {
#pragma unroll
      for (int j = 0; j < kPass1Tiles; ++j) {
        const int k = k0 + kPass1Tiles * kPrefetchDistance + j;
        if (k < my_count) norm_prefetch<HAS_G>(prm.tiles[b + k * G], prm);
      }
    }
#pragma unroll
    for (int j = 0; j < kPass1Tiles; ++j)
      if (k0 + j < my_count)      // tile partials are added in tile order: the sum does not depend on kPass1Tiles
        acc += (double)norm_finish<HAS_G, CAP, USE_TMEM>(d[j], prm, r[j], st[j], tm[j], pol);
  }
  const double part = block_reduce_to_double(acc, red);
  if (threadIdx.x == 0) prm.partials[blockIdx.x] = part;
  stamp(1);
  if (own_barrier) atomic_grid_barrier(prm.tickets + 3, gridDim.x); else cg::this_grid().sync();
  stamp(2);
  // ---- every CTA combines the per-CTA partials in the same fixed order ---------------------------
  if (threadIdx.x < 32) {
    double tot = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) tot += __ldcg(prm.partials + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (threadIdx.x == 0) {
      const float g_norm = __fsqrt_rn((float)tot);        // tf.linalg.global_norm
      s_bcast[0] = clip_scale(g_norm, prm.sc.clip);
      s_bcast[1] = g_norm;
    }
  }
  __syncthreads();
  const float s = s_bcast[0], gn = s_bcast[1];
  // ---- pass 2, reverse: stash first, then the most recently written L2 lines -------------------------
  if (my_count > 0) {
    TileDesc d = prm.tiles[b + (my_count - 1) * G];
    for (int k = my_count - 1; k >= 0; --k) {
      TileDesc dn;
      if (k > 0) dn = prm.tiles[b + (k - 1) * G];
      const float4* st = (k < n_stashed && stashable<HAS_G>(d, prm)) ? stash_mem + (size_t)k * (kTile / 4) : nullptr;
      update_tile2<VARIANT, CAP, USE_TMEM>(d, prm, s, st, st ? kNoTmem : tmem_for(k, d));
      d = dn;
    }
  }
  __syncthreads();
  stamp(3);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    prm.stats[0] = 1.f; prm.stats[1] = prm.sc.lr; prm.stats[2] = gn; prm.stats[3] = s;
  }
  if (USE_TMEM && n_tmem > 0) {
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base);
  }
  if (own_barrier) {          // the last CTA to leave re-arms the barrier counter for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(prm.tickets + 2, 1u) == gridDim.x - 1) { prm.tickets[3] = 0; prm.tickets[2] = 0; __threadfence(); }
    }
  }
}

}  // namespace gaccum
