// gaccum_kernels.cuh -- sm_100a kernels of the gradient-accumulation train_op.
//
// All kernels walk a static tile table (<= 2048 elements of one tensor per tile), so the work split --
// and therefore every reduction -- is deterministic.
//
//   accumulate_kernel    a += G                                  optimization.py:81,93    (12 B/elem)
//                        one tile per CTA (hardware block scheduler)
//   apply_kernel         the apply branch WITHOUT clipping (plain Adam of the example scripts, or
//                        clip_norm <= 0): a' = a + G; n = a'/N; Adam; a = 0 in a single pass (36 B/elem).
//   apply_clip_kernel    the apply branch WITH tf.clip_by_global_norm                  (36 B/elem)
//                        optimization.py:80-88, 128-177.  The global norm of ALL tensors is needed before
//                        ANY element can be updated, so it is one cooperative launch, one CTA per SM, with
//                        two passes around a grid barrier:
//                          pass 1  PRODUCER WARPS stream G into shared-memory tile slots with TMA bulk
//                                  copies (cp.async.bulk + mbarrier complete_tx; descriptors and L2
//                                  prefetches of a run ahead), three 256-thread CONSUMER GROUPS load a,
//                                  take G out of the slots, a' = a + G, reduce sum((a'/N)^2)
//                                  (thread fp32 per tile -> fp64 running sum -> warp shuffle -> shared
//                                  memory -> one fp64 partial per CTA); a' is PARKED ON CHIP: the first
//                                  tiles in Tensor Memory (tcgen05.st), the last ones in the very slots
//                                  their G landed in (the ring becomes the stash), the rest written back
//                                  in place tagged L2::evict_last
//                          barrier every CTA adds the per-CTA partials in the same order
//                                  (bit-identical gn and clip scale everywhere)
//                          pass 2  L2-resident tiles youngest first, then the slots, then Tensor Memory:
//                                  clip, AdamWeightDecay/Adam, write p, m, v, a = 0
// Arithmetic uses round-to-nearest intrinsics (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn) so nvcc
// cannot contract mul+add into FMA: the reference graph is un-fused, one rounding per TF op, and we
// reproduce it bit for bit (the kernels are HBM-bound, the extra flops are free).
// Measurements and the experiments behind each choice: profiles/r01_tune_sweep.md, profiles/r02_*.md.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gaccum {
namespace cg = cooperative_groups;

constexpr int kThreads = 256;                    // 8 warps per CTA / per consumer group
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
  float* stats;       // gaccum_stats
  uint32_t flags;     // kFlag* bits (all of them produce correct results)
  int32_t stash_tiles;  // apply_clip_kernel: tiles of a' each consumer group keeps in shared memory between the passes
  int32_t tmem_tiles;   // ... and in Tensor Memory (0 or kTmemTiles)
#ifdef GACCUM_EXPERIMENTS
  unsigned long long* debug;  // 16 words per CTA: 4 timestamps (ns) + wait-cycle counters (tools/cta_timeline.py)
#endif
  Scalars sc;
  PtrTable<CAP> tab;
};

constexpr uint32_t kFlagAssign = 1u;        // accumulate_kernel stores G instead of adding it (host-session gather of small tensors)
constexpr uint32_t kFlagNoL2Prefetch = 2u;  // apply_clip_kernel: producer does not run bulk L2 prefetches ahead (A/B measurement)
constexpr uint32_t kFlagNoCrossPrefetch = 4u;  // apply_clip_kernel: no L2 prefetch of pass 2's first p/m/v tiles before the barrier

// ---------------------------------------------------------------------------------------------
// memory helpers: G is read exactly once -> streaming (evict-first) loads; zeroing the
// accumulator is a streaming store.  a' must survive in L2 from pass 1 to pass 2, so it is
// tagged evict_last while everything that is touched once is tagged evict-first.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
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
  if (prm.flags & kFlagAssign) {              // gather: a = G (G may live in pinned host memory)
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
  if (blockIdx.x == 0 && threadIdx.x == 0 && !(prm.flags & kFlagAssign)) {   // the gather pass is not a step
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
// apply without clipping: one pass.  HAS_G = false applies the accumulators as they are.
// ---------------------------------------------------------------------------------------------
template <int VARIANT, bool LOAD_G, int CAP>
__device__ __forceinline__ void update_tile(const TileDesc d, const KernelParams<CAP>& prm) {
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

  // no gradient for this tile: a + 0 would turn -0 into +0 only; the add is skipped entirely
  auto elem = [&](float ax, float gx, bool has_g, float& px, float& mx, float& vx) {
    if (has_g) ax = __fadd_rn(ax, gx);                            // optimization.py:81
    const float c = normalize(ax, sc.nf, sc.inv_nf);              // :83
    adam_elem<VARIANT>(c, px, mx, vx, decay, sc);                 // :85
  };

  if (aligned16(p) && (g == nullptr || aligned16(g))) {
    const uint32_t nvec = len >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* a4 = reinterpret_cast<float4*>(a);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* p4 = reinterpret_cast<float4*>(p);
    float4 va[kUnroll], vg[kUnroll], vp[kUnroll], vm[kUnroll], vv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        va[u] = __ldcs(a4 + i); vp[u] = __ldcs(p4 + i); vm[u] = __ldcs(m4 + i); vv[u] = __ldcs(v4 + i);
        if (LOAD_G && g) vg[u] = ld_stream(g4 + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        const bool hg = LOAD_G && g;
        elem(va[u].x, vg[u].x, hg, vp[u].x, vm[u].x, vv[u].x); elem(va[u].y, vg[u].y, hg, vp[u].y, vm[u].y, vv[u].y);
        elem(va[u].z, vg[u].z, hg, vp[u].z, vm[u].z, vv[u].z); elem(va[u].w, vg[u].w, hg, vp[u].w, vm[u].w, vv[u].w);
        __stcs(p4 + i, vp[u]); __stcs(m4 + i, vm[u]); __stcs(v4 + i, vv[u]);
        __stcs(a4 + i, make_float4(0.f, 0.f, 0.f, 0.f));   // optimization.py:86-87
      }
    }
    const uint32_t i = (nvec << 2) + tid;
    if (i < len) {
      float px = p[i], mx = m[i], vx = v[i];
      const bool hg = LOAD_G && g;
      elem(a[i], hg ? ld_stream(g + i) : 0.f, hg, px, mx, vx);
      p[i] = px; m[i] = mx; v[i] = vx; a[i] = 0.f;
    }
  } else {
    for (uint32_t i = tid; i < len; i += kThreads) {
      float px = p[i], mx = m[i], vx = v[i];
      const bool hg = LOAD_G && g;
      elem(a[i], hg ? ld_stream(g + i) : 0.f, hg, px, mx, vx);
      p[i] = px; m[i] = mx; v[i] = vx; a[i] = 0.f;
    }
  }
}

template <int VARIANT, bool HAS_G, int CAP>
__global__ void __launch_bounds__(kThreads)
apply_kernel(const __grid_constant__ KernelParams<CAP> prm) {
  const int nt = prm.num_tiles;
  int t = blockIdx.x;
  if (t < nt) {
    TileDesc d = prm.tiles[t];
    while (true) {
      const int tn = t + gridDim.x;
      TileDesc dn;
      if (tn < nt) dn = prm.tiles[tn];
      update_tile<VARIANT, HAS_G>(d, prm);
      if (tn >= nt) break;
      t = tn; d = dn;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    prm.stats[0] = 1.f; prm.stats[1] = prm.sc.lr; prm.stats[2] = 0.f; prm.stats[3] = 1.f;
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
    const int nw = ((int)blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) tot += smem[w];
  }
  return tot;
}

// =============================================================================================
// apply with clipping: ONE cooperative launch, one CTA per SM =
//     3 consumer groups x 256 threads  +  3 producer warps (one per group)          (864 threads)
//
// Every consumer group behaves like an independent 256-thread CTA with virtual block id
// b = blockIdx * 3 + group and owns tiles b, b + 3*grid, b + 6*grid, ...  (its sequence j = 0..C-1).
//
// Pass 1 is latency-bound when the consumers issue all of their own loads: an LDG-fed loop can keep only
// as many bytes in flight as it has registers to land them in, and it stalls once per iteration on a
// dependent descriptor load (measured in round 1: 4.6 TB/s, two tiles = 32 registers per thread in
// flight).  Shared memory that a TMA ring would need is the same shared memory the a' stash needs (L2 can
// hold only ~50 MB of a' across the two passes, and above 196 KB of shared memory the SM's L1 shrinks to
// 28 KB, which starves pass 2) -- so here THE RING IS THE STASH:
//   * each group has M tile slots (8 KB).  Its producer warp fetches tile descriptors 32 at a time (one per
//     lane, shuffled out) and streams the GRADIENT tile j into slot j % M with one TMA bulk copy
//     (cp.async.bulk shared <- global, completing on the slot's `full` mbarrier by byte count), as soon as
//     the consumers have released the slot (`empty` mbarrier, one arrive per warp).  Up to M x 8 KB of G
//     per group are in flight, no registers involved;
//   * the consumers load the ACCUMULATOR tile themselves, two tiles ahead (16 registers), helped by a
//     bulk L2 prefetch the producer issues a few tiles ahead; they wait on `full`, read G out of the slot
//     (thread t reads word t: conflict free), a' = a + G, reduce sum((a'/N)^2);
//   * where a' goes depends on j: the first n_tm tiles -> Tensor Memory; the middle ones -> back in place
//     in global memory tagged L2::evict_last; the LAST M tiles -> IN PLACE OVER G IN THEIR SLOT, which is
//     never recycled again: at the end of pass 1 the ring has become the stash, no byte of shared memory
//     was ever only a staging buffer.
// The producers never wait for the grid barrier: when their last copy is issued they prefetch the first
// p/m/v tile of pass 2 into L2, so HBM keeps streaming while the CTAs rendezvous.
// Pass 2 takes the L2-resident tiles first, youngest first, then the slots, then Tensor Memory.  Thread t of
// a group reads back exactly the words it wrote, so neither stash needs synchronisation.
//
// Tiles whose gradient pointer is not 16-byte aligned (views into a flat buffer) or that are shorter than
// one float4 cannot be moved by bulk copies: producer and consumers evaluate the same predicate
// (bulk_vecs); for such tiles the producer only completes the slot's barrier and the consumers load the
// tile with scalar LDGs.  Every tile goes through the same full/empty protocol, so phases never skew.
// =============================================================================================
// ---------------------------------------------------------------------------------------------
// Tensor Memory as a scratchpad.  TMEM (256 KB per SM, 512 columns x 128 lanes x 32 bit) normally
// holds tcgen05.mma accumulators; this kernel has no MMA, so it is idle silicon -- 37 MB across the
// chip, more than the shared-memory stash.  The CTA allocates all 512 columns (one CTA per SM by
// construction); every consumer warp parks a' values in the 32 lanes it may address
// (lane quadrant = warp % 4; the 6 warps sharing a quadrant take 80 columns each):
// tcgen05.st 32x32b.x8 writes the thread's 8 words of a tile to 8 consecutive columns of its own
// lane, tcgen05.ld reads them back in pass 2.  A thread only ever reads what it wrote itself.
// ---------------------------------------------------------------------------------------------
constexpr int kGroups = 3;                        // consumer groups per CTA
constexpr int kConsumerThreads = kThreads * kGroups;        // 768
constexpr int kClipThreads = kConsumerThreads + 32 * kGroups;   // + one producer warp per group = 864
constexpr int kTmemCols = 512;                    // one CTA per SM: take all columns
constexpr int kTmemColsPerWarp = 80;              // 6 warps share a lane quadrant: 6 x 80 = 480 <= 512
constexpr int kTmemTiles = kTmemColsPerWarp / 8;  // 10 tiles per group
constexpr uint32_t kNoTmem = 0xffffffffu;
constexpr int kMaxSlots = 8;                      // tile slots per group (3 x 8 x 8 KB = 192 KB: the largest pool that keeps L1 at 60 KB)
#ifdef GACCUM_A_VIA_TMA                           // measurement variant: the accumulator tile travels by TMA too (slot = G | a)
constexpr int kSlotVecs = 2 * (kTile / 4);
#else
constexpr int kSlotVecs = kTile / 4;
#endif
#ifndef GACCUM_A_PREFETCH_TILES
#define GACCUM_A_PREFETCH_TILES 3
#endif
constexpr int kAPrefetch = GACCUM_A_PREFETCH_TILES;   // bulk L2 prefetch distance for the accumulator stream, in tiles per group

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"((uint32_t)kTmemCols) : "memory");
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

// Number of float4 vectors of this tile that take the vector path (0: scalar fallback).  The producer and
// the consumers MUST agree on this.
__device__ __forceinline__ uint32_t bulk_vecs(const TileDesc& d, const float* g) {
  return (g == nullptr || aligned16(g)) ? (d.len >> 2) : 0u;
}

// a tile may be stashed on chip only if BOTH passes will take the vector path for it
template <bool HAS_G, int CAP>
__device__ __forceinline__ bool stashable(const TileDesc& d, const KernelParams<CAP>& prm) {
  bool ok = aligned16(param_ptr(prm.tab, d));
  if constexpr (HAS_G) { const float* g = grad_ptr(prm.tab, d); ok = ok && (g == nullptr || aligned16(g)); }
  return ok;
}

// how a group's C tiles are split between Tensor Memory, L2 and the slots
struct TileClasses {
  int count;      // C
  int first_st;   // tiles j >= first_st stay in slot j % M
  int n_tm;       // tiles j < n_tm go to Tensor Memory
};
__device__ __forceinline__ TileClasses classify(int count, int slots, int tmem_tiles) {
  TileClasses c;
  c.count = count;
  c.first_st = count - min(slots, count);
  c.n_tm = min(tmem_tiles, c.first_st);
  return c;
}

struct ARegs { float4 v[kUnroll]; };

// ---- pass 1, consumer side: issue the accumulator loads of one tile (consumed two tiles later) ----
template <bool HAS_G, int CAP>
__device__ __forceinline__ void norm_issue_a(const TileDesc& d, const KernelParams<CAP>& prm, const bool to_l2,
                                             const uint64_t pol, ARegs& r) {
  const float* g = nullptr;
  if constexpr (HAS_G) g = grad_ptr(prm.tab, d);
  const uint32_t nvec = bulk_vecs(d, g), tid = threadIdx.x & (kThreads - 1);
  const float4* a4 = reinterpret_cast<const float4*>(prm.accum + (size_t)d.soff32 * kSlabAlign);
#ifndef GACCUM_A_VIA_TMA
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t i = u * kThreads + tid;
    if (i < nvec) r.v[u] = to_l2 ? ld_policy(a4 + i, pol) : __ldcs(a4 + i);   // the line a' returns to keeps evict_last
  }
#endif
}

// ---- pass 1, consumer side: finish one tile.  slot: this tile's slot (G lands there); in_slot: a' stays in it ----
// release: ring phase, the slot is handed back to the producer once G is in registers
template <bool HAS_G, int CAP>
__device__ __forceinline__ float norm_finish(const TileDesc& d, const KernelParams<CAP>& prm, ARegs& r, float4* slot,
                                             uint64_t* full, uint64_t* empty, const uint32_t parity, const bool release,
                                             const bool in_slot, const uint32_t tmem, const uint64_t pol,
                                             long long& dbg_wait /* experiments build: cycles spent waiting for G */) {
  const float* __restrict__ g = nullptr;
  if constexpr (HAS_G) g = grad_ptr(prm.tab, d);
  float* __restrict__ a = prm.accum + (size_t)d.soff32 * kSlabAlign;
  const uint32_t len = d.len, tid = threadIdx.x & (kThreads - 1);
  const float nf = prm.sc.nf, inv_nf = prm.sc.inv_nf;
  const uint32_t nvec = bulk_vecs(d, g);
  float acc = 0.f;
#ifdef GACCUM_EXPERIMENTS
  const long long t_w0 = clock64();
#endif
  mbar_wait(full, parity);                      // G has landed (or the producer had nothing to copy)
#ifdef GACCUM_EXPERIMENTS
  dbg_wait += clock64() - t_w0;
#endif
  if (nvec > 0) {
    float4* a4 = reinterpret_cast<float4*>(a);
    float4 gg[kUnroll];
    if (g) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t i = u * kThreads + tid;
        if (i < nvec) gg[u] = slot[i];
      }
    }
#ifdef GACCUM_A_VIA_TMA
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) r.v[u] = slot[kTile / 4 + i];
    }
#endif
    if (release) {                              // ring phase: hand the slot back as soon as G is in registers
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(empty);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = u * kThreads + tid;
      if (i < nvec) {
        float4& x = r.v[u];
        if (g) {
          x.x = __fadd_rn(x.x, gg[u].x); x.y = __fadd_rn(x.y, gg[u].y);
          x.z = __fadd_rn(x.z, gg[u].z); x.w = __fadd_rn(x.w, gg[u].w);
        }
        if (in_slot) slot[i] = x;               // in place over G: the slot is now stash
        else if (tmem != kNoTmem) {}
        else if (g) st_policy(a4 + i, x, pol);
        const float nx = normalize(x.x, nf, inv_nf), ny = normalize(x.y, nf, inv_nf),
                    nz = normalize(x.z, nf, inv_nf), nw = normalize(x.w, nf, inv_nf);
        acc = fmaf(nx, nx, acc); acc = fmaf(ny, ny, acc); acc = fmaf(nz, nz, acc); acc = fmaf(nw, nw, acc);
      }
    }
    if (tmem != kNoTmem) tmem_store8(tmem, r.v[0], r.v[1]);   // full tile: every lane of every warp carries data
    const uint32_t i = (nvec << 2) + tid;      // < 4 tail elements always travel through global memory
    if (i < len) {
      float xs = a[i];
      if (g) { xs = __fadd_rn(xs, ld_stream(g + i)); a[i] = xs; }
      const float n = normalize(xs, nf, inv_nf);
      acc = fmaf(n, n, acc);
    }
  } else {
    // unaligned gradient view, or a tile shorter than one float4: nothing was copied, nothing is stashed
    if (release) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(empty);
    }
    for (uint32_t i = tid; i < len; i += kThreads) {
      float xs = a[i];
      if (g) { xs = __fadd_rn(xs, ld_stream(g + i)); a[i] = xs; }
      const float n = normalize(xs, nf, inv_nf);
      acc = fmaf(n, n, acc);
    }
  }
  return acc;
}

// ---- pass 2: one tile ------------------------------------------------------------------------------
template <int VARIANT, int CAP>
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
    if (tmem != kNoTmem) tmem_load8(tmem, va[0], va[1]);
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

// Dynamic shared memory of apply_clip_kernel: 3 groups x slots x 8 KB tile slots.
template <int VARIANT, bool HAS_G, int CAP>
__global__ void __launch_bounds__(kClipThreads, 1)
apply_clip_kernel(const __grid_constant__ KernelParams<CAP> prm) {
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  __shared__ double red[kClipThreads / 32];
  __shared__ float s_bcast[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(8) uint64_t s_full[kGroups][kMaxSlots];
  __shared__ __align__(8) uint64_t s_empty[kGroups][kMaxSlots];

  const int warp = (int)threadIdx.x >> 5;
  const bool is_producer = warp >= kConsumerThreads / 32;
  const int grp = is_producer ? warp - kConsumerThreads / 32 : (int)threadIdx.x / kThreads;   // group served / group id
  const int nt = prm.num_tiles, G = (int)gridDim.x * kGroups;
  const int M = prm.stash_tiles;                                   // slots per group (1..kMaxSlots)
  const int b = (int)blockIdx.x * kGroups + grp;                   // virtual block id
  const TileClasses tc = classify(b < nt ? (nt - 1 - b) / G + 1 : 0, M, prm.tmem_tiles);
  float4* const slots = reinterpret_cast<float4*>(smem_dyn) + (size_t)grp * M * kSlotVecs;
  const uint64_t pol_last = policy_evict_last();

#ifdef GACCUM_EXPERIMENTS
  auto stamp = [&](int which) {
    if (prm.debug && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      prm.debug[blockIdx.x * 16 + which] = t;
    }
  };
#else
  auto stamp = [](int) {};
#endif
  stamp(0);

  // ---- set-up: mbarriers, Tensor Memory ---------------------------------------------------------------
  if (threadIdx.x == 0) {
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
      for (int s = 0; s < M; ++s) { mbar_init(&s_full[g][s], 1); mbar_init(&s_empty[g][s], kThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (prm.tmem_tiles > 0 && warp == 1) tmem_alloc(&s_tmem_base);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = prm.tmem_tiles > 0 ? s_tmem_base : 0u;

  double acc = 0.0;
  if (is_producer) {
    // =========================== producer warp of group `grp` ===========================
    const int lane = (int)threadIdx.x & 31;
    const uint64_t pol_first = policy_evict_first();
    const bool l2_prefetch = !(prm.flags & kFlagNoL2Prefetch);
    auto fetch = [&](int j0, TileDesc& d) -> bool {      // this lane's descriptor of the batch starting at j0
      const int j = j0 + lane;
      if (j >= tc.count) return false;
      d = prm.tiles[b + j * G];
      return true;
    };
    auto prefetch_a = [&](const TileDesc& d) {
      const float* g = nullptr;
      if constexpr (HAS_G) g = grad_ptr(prm.tab, d);
      const uint32_t nv = bulk_vecs(d, g);
      if (nv > 0) bulk_prefetch_l2(prm.accum + (size_t)d.soff32 * kSlabAlign, nv * 16u);
    };
    TileDesc dc, dn;
    bool vc = fetch(0, dc), vn = fetch(32, dn);
    if (l2_prefetch && vc && lane < kAPrefetch) prefetch_a(dc);      // the first tiles of the group
    int slot = 0;
    uint32_t use = 0;                                                  // how many times slot 0.. have been used: j / M
#ifdef GACCUM_EXPERIMENTS
    long long dbg_empty = 0;
    const long long dbg_p0 = clock64();
#endif
    for (int j0 = 0; j0 < tc.count; j0 += 32) {
      const int nb = min(32, tc.count - j0);
      for (int l = 0; l < nb; ++l) {
        TileDesc d;
        d.tensor_flags = __shfl_sync(0xffffffffu, dc.tensor_flags, l);
        d.len = __shfl_sync(0xffffffffu, dc.len, l);
        d.toff = __shfl_sync(0xffffffffu, dc.toff, l);
        d.soff32 = __shfl_sync(0xffffffffu, dc.soff32, l);
        // accumulator stream: the lane that holds tile j + kAPrefetch pulls it into L2
        if (l2_prefetch) {
          const int lp = l + kAPrefetch;
          if (lp < 32) { if (lane == lp && vc) prefetch_a(dc); }
          else { if (lane == lp - 32 && vn) prefetch_a(dn); }
        }
        if (lane == 0) {
          const float* g = nullptr;
          if constexpr (HAS_G) g = grad_ptr(prm.tab, d);
          const uint32_t nvec = bulk_vecs(d, g);
#ifdef GACCUM_EXPERIMENTS
          const long long t_e0 = clock64();
#endif
          mbar_wait(&s_empty[grp][slot], (use & 1u) ^ 1u);           // the consumers have released the slot
#ifdef GACCUM_EXPERIMENTS
          dbg_empty += clock64() - t_e0;
#endif
#ifdef GACCUM_A_VIA_TMA
          if (nvec > 0) {
            mbar_arrive_expect_tx(&s_full[grp][slot], nvec * 16u * (g ? 2u : 1u));
            if (g) bulk_g2s(slots + (size_t)slot * kSlotVecs, g, nvec * 16u, &s_full[grp][slot], pol_first);
            bulk_g2s(slots + (size_t)slot * kSlotVecs + kTile / 4, prm.accum + (size_t)d.soff32 * kSlabAlign, nvec * 16u, &s_full[grp][slot], pol_first);
          } else {
            mbar_arrive(&s_full[grp][slot]);
          }
#else
          if (g != nullptr && nvec > 0) {
            mbar_arrive_expect_tx(&s_full[grp][slot], nvec * 16u);
            bulk_g2s(slots + (size_t)slot * kSlotVecs, g, nvec * 16u, &s_full[grp][slot], pol_first);
          } else {
            mbar_arrive(&s_full[grp][slot]);                         // nothing to copy: complete the phase
          }
#endif
        }
        if (++slot == M) { slot = 0; ++use; }
      }
      dc = dn; vc = vn;
      vn = fetch(j0 + 64, dn);
    }
#ifdef GACCUM_EXPERIMENTS
    if (prm.debug && lane == 0) {
      prm.debug[blockIdx.x * 16 + 7 + grp] = (unsigned long long)dbg_empty;               // cycles the producer waited for a free slot
      prm.debug[blockIdx.x * 16 + 13 + grp] = (unsigned long long)(clock64() - dbg_p0);   // cycles until the last copy was issued
    }
#endif
    // pass 1 is fully issued: pull the first pass-2 tile's p, m, v into L2 while the CTAs rendezvous
    if (!(prm.flags & kFlagNoCrossPrefetch) && lane == 0 && tc.count > 0) {
      const int j = tc.first_st > tc.n_tm ? tc.first_st - 1 : tc.first_st;
      const TileDesc d = prm.tiles[b + j * G];
      const float* p = param_ptr(prm.tab, d);
      const uint32_t nv = d.len >> 2;
      if (nv > 0) {
        const size_t soff = (size_t)d.soff32 * kSlabAlign;
        if (aligned16(p)) bulk_prefetch_l2(p, nv * 16u);
        bulk_prefetch_l2(prm.m + soff, nv * 16u);
        bulk_prefetch_l2(prm.v + soff, nv * 16u);
      }
    }
  } else if (tc.count > 0) {
    // =========================== consumer group: pass 1 ===========================
    const int C = tc.count;
    auto desc = [&](int j) { return prm.tiles[b + j * G]; };
    auto is_l2 = [&](int j) { return j >= tc.n_tm && j < tc.first_st; };
    TileDesc d0 = desc(0), d1 = d0;
    if (C > 1) d1 = desc(1);
    ARegs r0, r1;
    norm_issue_a<HAS_G>(d0, prm, is_l2(0), pol_last, r0);
    if (C > 1) norm_issue_a<HAS_G>(d1, prm, is_l2(1), pol_last, r1);
    int slot = 0;
    uint32_t use = 0;
    long long dbg_wait = 0;
#ifdef GACCUM_EXPERIMENTS
    const long long dbg_t0 = clock64();
#endif
    auto finish = [&](int j, const TileDesc& d, ARegs& r) -> float {
      const bool in_slot = j >= tc.first_st && stashable<HAS_G>(d, prm);
      const bool ring = j < tc.first_st;
      uint32_t tm = kNoTmem;
      if (j < tc.n_tm && d.len == (uint32_t)kTile && stashable<HAS_G>(d, prm)) tm = tmem_slot_addr(tmem_base, j);
      // slots of the final M tiles are never handed back: `ring` decides the empty-arrive, `in_slot` the destination
      const float x = norm_finish<HAS_G>(d, prm, r, slots + (size_t)slot * kSlotVecs, &s_full[grp][slot], &s_empty[grp][slot],
                                         use & 1u, ring, in_slot, tm, pol_last, dbg_wait);
      if (++slot == M) { slot = 0; ++use; }
      return x;
    };
    for (int j = 0; j < C; j += 2) {
      TileDesc d2 = d0, d3 = d1;
      if (j + 2 < C) d2 = desc(j + 2);
      if (j + 3 < C) d3 = desc(j + 3);
      acc += (double)finish(j, d0, r0);
      if (j + 2 < C) norm_issue_a<HAS_G>(d2, prm, is_l2(j + 2), pol_last, r0);
      if (j + 1 < C) {
        acc += (double)finish(j + 1, d1, r1);
        if (j + 3 < C) norm_issue_a<HAS_G>(d3, prm, is_l2(j + 3), pol_last, r1);
      }
      d0 = d2; d1 = d3;
    }
#ifdef GACCUM_EXPERIMENTS
    if (prm.debug && (threadIdx.x & (kThreads - 1)) == 0) {
      prm.debug[blockIdx.x * 16 + 4 + grp] = (unsigned long long)dbg_wait;                 // cycles waiting for G
      prm.debug[blockIdx.x * 16 + 10 + grp] = (unsigned long long)(clock64() - dbg_t0);   // cycles of the group's pass 1
    }
#endif
  }
  const double part = block_reduce_to_double(acc, red);
  if (threadIdx.x == 0) prm.partials[blockIdx.x] = part;
  stamp(1);
  cg::this_grid().sync();
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
  // ---- pass 2: L2-resident tiles youngest first, then the slots, then Tensor Memory ----------------
  if (!is_producer && tc.count > 0) {
    auto run = [&](int j) {
      const TileDesc d = prm.tiles[b + j * G];
      const bool ok = stashable<HAS_G>(d, prm);
      const float4* st = (j >= tc.first_st && ok) ? slots + (size_t)(j % M) * kSlotVecs : nullptr;
      uint32_t tm = kNoTmem;
      if (j < tc.n_tm && d.len == (uint32_t)kTile && ok) tm = tmem_slot_addr(tmem_base, j);
      update_tile2<VARIANT>(d, prm, s, st, tm);
    };
    for (int j = tc.first_st - 1; j >= tc.n_tm; --j) run(j);
    for (int j = tc.first_st; j < tc.count; ++j) run(j);
    for (int j = 0; j < tc.n_tm; ++j) run(j);
  }
  __syncthreads();
  stamp(3);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    prm.stats[0] = 1.f; prm.stats[1] = prm.sc.lr; prm.stats[2] = gn; prm.stats[3] = s;
  }
  if (prm.tmem_tiles > 0 && warp == 1) tmem_dealloc(tmem_base);
}

}  // namespace gaccum
