// gaccum_dp.cuh -- the data-parallel apply step as ONE kernel over NVLink peer memory.
//
// Replaces what reference distributedExample/04 does with MultiWorkerMirroredStrategy:
// per-variable all-reduces on EVERY micro-step (04:55,58,70) followed by a replicated
// apply (04:59-66).  Here every rank accumulates locally for the whole window and the exchange
// happens once, inside the apply kernel, tile by tile:
//
//   flags 0    (system-scope release/acquire words in symmetric memory, one block signals, every block
//              waits): every rank's accumulator is final
//   pass 1     rank r owns the tiles [tile_lo, tile_hi).  For each owned tile it LOADS THE TILE
//              FROM EVERY RANK'S ACCUMULATOR over NVLink (peer pointers, fixed rank order
//              0..W-1 => deterministic sum), writes the reduced a' into its own slab and
//              reduces sum((a'/N)^2)                                  == reduce-scatter + norm
//   flags 1    per-rank partial norms are exchanged through the control blocks; every rank
//              adds the W partials in rank order => bit-identical gn and clip scale everywhere
//   pass 2     clip + AdamWeightDecay/Adam on the owned tiles (m, v are only ever touched by
//              their owner: ZeRO-1 style), and the new parameters are STORED INTO EVERY RANK'S
//              PARAMETER SLAB over NVLink                               == all-gather
//              meanwhile all non-owned tiles of the local accumulator are zeroed (:86-87)
//   flags 2    all peers' parameter stores have landed before this kernel completes
//   (a flag wait longer than 60 s means a dead peer: the kernel traps instead of hanging the GPU)
//
// NVLink bytes per rank and direction: 2 * (W-1)/W * 4P (the all-reduce lower bound); HBM
// bytes of the update shrink to 1/W.  Parameters must live in one packed, peer-mapped slab
// (plan offsets); the host side (PyTorch symmetric memory, NVSHMEM, cuMem IPC, ...) only
// supplies the W base pointers.
#pragma once

#include "gaccum_kernels.cuh"

namespace gaccum {

constexpr int kMaxRanks = 8;
// control block layout (uint32 words), one block per rank, zero-initialised once:
//   [phase * kMaxRanks + src]  epoch flags, phase 0..2
//   byte 128: double norm_partial[kMaxRanks]
constexpr int kCtrlFlagWords = 3 * kMaxRanks;
constexpr int kCtrlNormByteOffset = 128;
constexpr int kCtrlBytes = 256;
constexpr unsigned long long kDpTimeoutNs = 60ull * 1000 * 1000 * 1000;   // a flag wait longer than 60 s is a dead peer

struct DpParams {
  const TileDesc* tiles;
  int32_t num_tiles, tile_lo, tile_hi;
  float* m;
  float* v;
  double* partials;
  float* stats;
  float* bcast;          // plan-owned: {scale, gn}
  uint32_t tune;
  Scalars sc;
  int32_t rank, world;
  uint32_t epoch;
  float* accum[kMaxRanks];
  float* param[kMaxRanks];
  uint32_t* ctrl[kMaxRanks];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Peer accumulator data is read exactly once per kernel, after flag round 0 (every block polls its
// own rank's control block and issues one acquire fence at system scope), L1 is invalidated at every
// launch and peer lines bypass the local L2: a plain LDG.128 is coherent here and lets the compiler
// keep all W x kUnroll x tiles-per-iteration loads in flight (NVLink latency is ~2 us; ordering them
// would serialise it).
__device__ __forceinline__ float4 ld_peer(const float4* p) { return *p; }
__device__ __forceinline__ float ld_peer(const float* p) { return *p; }

// Cross-GPU flags.  Signalling is done by ONE block (thread t < W writes rank t's control block with
// release semantics at system scope); waiting is done by EVERY block on its own rank's control
// block (local memory), so no grid-wide barrier is needed to fan the news out.
__device__ __forceinline__ void dp_signal(const DpParams& prm, int phase) {
  const int t = threadIdx.x;
  if (t < prm.world) st_release_sys(prm.ctrl[t] + phase * kMaxRanks + prm.rank, prm.epoch);
}
__device__ __forceinline__ void dp_wait(const DpParams& prm, int phase) {
  const int t = threadIdx.x;
  if (t < prm.world) {
    const uint32_t* mine = prm.ctrl[prm.rank] + phase * kMaxRanks + t;
    // poll with relaxed loads (no fence per probe), acquire once when the flag has flipped
    uint32_t seen;
    unsigned long long t_start = 0;
    uint32_t spins = 0;
    do {
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      // A peer that never arrives (crashed rank, mismatched epoch) must not hang the GPU for ever:
      // after kDpTimeoutNs the kernel traps, which surfaces as a sticky CUDA error on the host.
      if (seen != prm.epoch && (++spins & 0x3fffu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t_start == 0) t_start = now;
        else if (now - t_start > kDpTimeoutNs) __trap();
      }
    } while (seen != prm.epoch);
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncthreads();
}

// pass 1 on TPI owned tiles at once: a' = sum_w a_w (rank order), stored locally; returns
// acc + sum((a'/N)^2).  TPI x kUnroll x W 128-bit loads are in flight per thread: NVLink reads
// have ~2-3 us latency, so ~16 outstanding vectors per thread are needed to fill the links.
template <int TPI>
__device__ __forceinline__ float dp_reduce_tiles(const TileDesc (&d)[TPI], const int ntile, const DpParams& prm,
                                                 float acc, const uint64_t pol) {
  const uint32_t tid = threadIdx.x;
  const float nf = prm.sc.nf, inv_nf = prm.sc.inv_nf;
  const int W = prm.world;
  float4 part[TPI][kUnroll][kMaxRanks / (TPI > 1 ? (TPI > 2 ? 4 : 2) : 1)];
  constexpr int WCAP = kMaxRanks / (TPI > 1 ? (TPI > 2 ? 4 : 2) : 1);   // TPI=4 -> W<=2, TPI=2 -> W<=4, TPI=1 -> W<=8
#pragma unroll
  for (int j = 0; j < TPI; ++j) {
    if (j < ntile) {
      const size_t soff = (size_t)d[j].soff32 * kSlabAlign;
      const uint32_t nvec = d[j].len >> 2;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t i = u * kThreads + tid;
        if (i < nvec) {
#pragma unroll
          for (int w = 0; w < WCAP; ++w)
            if (w < W) part[j][u][w] = ld_peer(reinterpret_cast<const float4*>(prm.accum[w] + soff) + i);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < TPI; ++j) {
    if (j < ntile) {
      const size_t soff = (size_t)d[j].soff32 * kSlabAlign;
      const uint32_t len = d[j].len, nvec = len >> 2;
      float4* mine = reinterpret_cast<float4*>(prm.accum[prm.rank] + soff);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const uint32_t i = u * kThreads + tid;
        if (i < nvec) {
          float4 s = part[j][u][0];
#pragma unroll
          for (int w = 1; w < WCAP; ++w)
            if (w < W) {
              s.x = __fadd_rn(s.x, part[j][u][w].x); s.y = __fadd_rn(s.y, part[j][u][w].y);
              s.z = __fadd_rn(s.z, part[j][u][w].z); s.w = __fadd_rn(s.w, part[j][u][w].w);
            }
          st_policy(mine + i, s, pol);
          const float nx = normalize(s.x, nf, inv_nf), ny = normalize(s.y, nf, inv_nf), nz = normalize(s.z, nf, inv_nf), nw = normalize(s.w, nf, inv_nf);
          acc = fmaf(nx, nx, acc); acc = fmaf(ny, ny, acc); acc = fmaf(nz, nz, acc); acc = fmaf(nw, nw, acc);
        }
      }
      const uint32_t i = (nvec << 2) + tid;
      if (i < len) {
        float s = ld_peer(prm.accum[0] + soff + i);
        for (int w = 1; w < W; ++w) s = __fadd_rn(s, ld_peer(prm.accum[w] + soff + i));
        prm.accum[prm.rank][soff + i] = s;
        const float n = normalize(s, nf, inv_nf);
        acc = fmaf(n, n, acc);
      }
    }
  }
  return acc;
}

template <int TPI>
__device__ __forceinline__ double dp_pass1(const DpParams& prm, const uint64_t pol) {
  double acc = 0.0;
  const int lo = prm.tile_lo, hi = prm.tile_hi, G = (int)gridDim.x;
  for (int t0 = lo + (int)blockIdx.x; t0 < hi; t0 += TPI * G) {
    TileDesc d[TPI];
    int n = 0;
#pragma unroll
    for (int j = 0; j < TPI; ++j)
      if (t0 + j * G < hi) { d[j] = prm.tiles[t0 + j * G]; n = j + 1; }
    acc += (double)dp_reduce_tiles<TPI>(d, n, prm, 0.f, pol);
  }
  return acc;
}

// pass 2 on one owned tile: update from the local reduced a', broadcast p' to every rank
template <int VARIANT>
__device__ __forceinline__ void dp_update_tile(const TileDesc d, const DpParams& prm, const float s) {
  const size_t soff = (size_t)d.soff32 * kSlabAlign;
  float* __restrict__ a = prm.accum[prm.rank] + soff;
  float* __restrict__ m = prm.m + soff;
  float* __restrict__ v = prm.v + soff;
  const float* __restrict__ p = prm.param[prm.rank] + soff;
  const bool decay = (d.tensor_flags >> 31) != 0;
  const uint32_t len = d.len, tid = threadIdx.x, nvec = len >> 2;
  const Scalars& sc = prm.sc;
  const int W = prm.world;
  auto elem = [&](float ax, float& px, float& mx, float& vx) {
    const float c = sc.clip > 0.f ? __fmul_rn(normalize(ax, sc.nf, sc.inv_nf), s) : normalize(ax, sc.nf, sc.inv_nf);
    adam_elem<VARIANT>(c, px, mx, vx, decay, sc);
  };
  float4* a4 = reinterpret_cast<float4*>(a);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* p4 = reinterpret_cast<const float4*>(p);
  float4 va[kUnroll], vp[kUnroll], vm[kUnroll], vv[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t i = u * kThreads + tid;
    if (i < nvec) { va[u] = __ldcs(a4 + i); vp[u] = __ldcs(p4 + i); vm[u] = __ldcs(m4 + i); vv[u] = __ldcs(v4 + i); }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t i = u * kThreads + tid;
    if (i < nvec) {
      elem(va[u].x, vp[u].x, vm[u].x, vv[u].x); elem(va[u].y, vp[u].y, vm[u].y, vv[u].y);
      elem(va[u].z, vp[u].z, vm[u].z, vv[u].z); elem(va[u].w, vp[u].w, vm[u].w, vv[u].w);
      __stcs(m4 + i, vm[u]); __stcs(v4 + i, vv[u]);
      __stcs(a4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
#pragma unroll
      for (int w = 0; w < kMaxRanks; ++w)      // all-gather: peer stores over NVLink (and the local copy)
        if (w < W) *(reinterpret_cast<float4*>(prm.param[w] + soff) + i) = vp[u];
    }
  }
  const uint32_t i = (nvec << 2) + tid;
  if (i < len) {
    float px = p[i], mx = m[i], vx = v[i];
    elem(a[i], px, mx, vx);
    m[i] = mx; v[i] = vx; a[i] = 0.f;
    for (int w = 0; w < W; ++w) prm.param[w][soff + i] = px;
  }
}

__device__ __forceinline__ void dp_zero_tile(const TileDesc d, const DpParams& prm) {
  float* a = prm.accum[prm.rank] + (size_t)d.soff32 * kSlabAlign;
  const uint32_t len = d.len, tid = threadIdx.x, nvec = len >> 2;
  float4* a4 = reinterpret_cast<float4*>(a);
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const uint32_t i = u * kThreads + tid;
    if (i < nvec) __stcs(a4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  const uint32_t i = (nvec << 2) + tid;
  if (i < len) a[i] = 0.f;
}

template <int VARIANT>
__global__ void __launch_bounds__(kThreads)
dp_apply_kernel(const __grid_constant__ DpParams prm) {
  __shared__ double red[kThreads / 32];
  __shared__ float s_bcast[2];
  cg::grid_group grid = cg::this_grid();
  const uint64_t pol = policy_evict_last();
  const int lo = prm.tile_lo, hi = prm.tile_hi, nt = prm.num_tiles;

  // ---- flag 0: my accumulators are final (stream order: the local accumulate ran before this
  //      kernel); every block waits until that is true of every rank ------------------------------
  if (blockIdx.x == 0) dp_signal(prm, 0);
  dp_wait(prm, 0);

  // ---- pass 1: reduce-scatter over peer loads + norm partial -------------------------------------
  double acc = 0.0;
  if (!(prm.tune & kTuneSkipPass1)) {
    if (prm.world <= 2) acc = dp_pass1<4>(prm, pol);
    else if (prm.world <= 4) acc = dp_pass1<2>(prm, pol);
    else acc = dp_pass1<1>(prm, pol);
  }
  const double part = block_reduce_to_double(acc, red);
  if (threadIdx.x == 0) prm.partials[blockIdx.x] = part;
  grid.sync();                     // gpu-scope: a' (local, read only by this rank) and the partials are visible

  // ---- flag 1: this rank's partial norm goes to every rank (itself included) ---------------------
  if (blockIdx.x == 0) {
    __shared__ double s_tot;
    if (threadIdx.x < 32) {
      double tot = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) tot += __ldcg(prm.partials + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (threadIdx.x == 0) s_tot = tot;
    }
    __syncthreads();
    if ((int)threadIdx.x < prm.world) {
      double* slot = reinterpret_cast<double*>(reinterpret_cast<char*>(prm.ctrl[threadIdx.x]) + kCtrlNormByteOffset) + prm.rank;
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot), "d"(s_tot) : "memory");
    }
    dp_signal(prm, 1);             // same thread: the release store orders the partial before the flag
  }
  // every block: wait for all W partials (also proves every rank finished READING my accumulators),
  // then add them in rank order -> bit-identical gn and s on every block of every rank
  dp_wait(prm, 1);
  if (threadIdx.x == 0) {
    const double* slots = reinterpret_cast<const double*>(reinterpret_cast<const char*>(prm.ctrl[prm.rank]) + kCtrlNormByteOffset);
    double tot = 0.0;
    for (int w = 0; w < prm.world; ++w) {
      double x;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(slots + w) : "memory");
      tot += x;
    }
    const float g_norm = prm.sc.clip > 0.f ? __fsqrt_rn((float)tot) : 0.f;
    s_bcast[0] = prm.sc.clip > 0.f ? clip_scale(g_norm, prm.sc.clip) : 1.f;
    s_bcast[1] = g_norm;
    if (blockIdx.x == 0) {
      prm.stats[0] = 1.f; prm.stats[1] = prm.sc.lr; prm.stats[2] = g_norm; prm.stats[3] = s_bcast[0];
    }
  }
  __syncthreads();
  const float s = s_bcast[0];

  // ---- pass 2: sharded update + all-gather by peer stores; zero everything I do not own ---------
  if (!(prm.tune & kTuneSkipPass2))
    for (int t = hi - 1 - (int)blockIdx.x; t >= lo; t -= (int)gridDim.x) dp_update_tile<VARIANT>(prm.tiles[t], prm, s);
  if (!(prm.tune & kTuneSkipZero))
    for (int t = (int)blockIdx.x; t < nt; t += (int)gridDim.x)
      if (t < lo || t >= hi) dp_zero_tile(prm.tiles[t], prm);
  __threadfence_system();
  grid.sync();

  // ---- flag 2: every peer's parameter stores into my slab are complete before the kernel ends ----
  if (blockIdx.x == 0) { dp_signal(prm, 2); dp_wait(prm, 2); }
}

}  // namespace gaccum
