// ring_kernel.cuh — sm_100a fast path of the risk-bounded safe-halfspace computation for fp32 samples, 4096 < N <= 10 240
// (BASELINE config 4: N = 10 000).  Same arithmetic contract and outputs as halfspace_kernel.cuh;
// what changes is the pipeline around the two sweeps:
//
//   * ONE CTA per SM: 8 sweep warps + 4 helper warps, and a shared-memory RING of 23 x 8 KB chunks (2.3 halfspaces at
//     N = 10 000) instead of one slot per CTA.  The producer warp streams halfspace after halfspace into the ring with
//     cp.async.bulk (TMA), chunk by chunk, as sweep B releases them: the copies of halfspaces b+1 and b+2 are in flight or
//     landed while b is worked on, HBM never waits for a slot.
//   * Software pipelining: the sweep warps run  sweep A(b+1) -> sweep B(b) -> exact phase(b).  Everything scalar that
//     sits between sweep A and sweep B of one halfspace — canonical direction (director warp), window placement (placer
//     warp) — runs while the sweep warps are busy with the neighbouring halfspace, and the select / CVaR / offsets of b
//     (finisher warp) run during b+1.  The sweep warps never wait for a helper; they wait for data only.
//   * sweep B classifies with one predicate per sample: "surely above the window" is a 1.0 / 0.0 factor (FSET) folded
//     into full-rate FFMAs, "inside the fp32 uncertainty band" is |q| <= half with q measured from the band centre; kept
//     samples go straight to a per-thread list (no mask extraction pass).
//   * No team-wide barrier per halfspace: hand-offs are bar.arrive towards helpers that block in bar.sync.
//
// Roles:  producer  TMA copies into the ring                                   (1 lane)
//         director  canonical mean -> h = unit(mean - ego), mean halfspace     core/geometry.py:35-53, core/halfspaces.py:70-106
//         placer    fp32 direction with a rigorous bound + statistical window around the predicted kc-th largest loss
//         finisher  exact kc-th largest loss among the window candidates, CVaR, offsets   core/risk_metrics.py:84-338
//                   (two warps, one per parity: a finisher has two halfspace periods for its serial chain)
// Anything unusual (window miss, list overflow, non-finite or degenerate data) sets the halfspace's redo flag; the ABI then
// runs the exact general path on the flagged halfspaces (streaming_kernel with redo flags).  The window and the fp32
// bounds decide how fast the exact threshold is found, never the result.
#pragma once

#include "halfspace_kernel.cuh"

namespace drcvar {

#ifdef DRCVAR_PROFILE_PHASES
#define RG_PH_DECL long long rg_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long rg_last = clock64();
#define RG_PH_MARK(k) { const long long rg_now = clock64(); rg_t[k] += rg_now - rg_last; rg_last = rg_now; }
#else
#define RG_PH_DECL
#define RG_PH_MARK(k)
#endif

constexpr int kRgWarps = 8;                                   // sweep warps (the team): the canonical 256-thread lane map
constexpr int kRgTeam = kRgWarps * 32;                        // 256
constexpr int kRgHelpers = 5;                                 // producer, director, placer, two finishers (even / odd halfspaces)
constexpr int kRgThreads = kRgTeam + 32 * kRgHelpers;
constexpr int kRgProducerWarp = 8, kRgDirectorWarp = 9, kRgPlacerWarp = 10, kRgFinisherWarp = 11;   // finishers: 11, 12
constexpr int kRgRowBytes = kRgTeam * 16;                     // 4 KB: one 16-byte vector per sweep thread
constexpr int kRgChunkBytes = 2 * kRgRowBytes;                // 8 KB: TMA / release granularity
constexpr int kRgRing = 23;                                   // chunks in the ring
constexpr int kRgMaxChunks = 10;                              // per halfspace: N <= 10 240
constexpr int kRgMaxN = kRgMaxChunks * kRgChunkBytes / 8;
constexpr int kRgMinN = 4097;                                 // >= 5 chunks per halfspace: at most 5 halfspaces in the ring
constexpr int kRgFullBars = 8;                                // one "landed" barrier per halfspace in flight
constexpr int kRgListStride = 9;                              // float2 entries per thread (odd stride: conflict-free columns)
constexpr int kRgListCap = 9;
constexpr int kRgListGuard = 4 * kRgMaxChunks;                // a thread that keeps everything overruns its neighbours, never the array
constexpr int kRgDenseCap = 80;                               // kept samples per sweep warp after compaction
constexpr int kRgCandCap = 80;                                // window candidates per sweep warp
constexpr int kRgWarpCand = kRgCandCap + 32;                  // + per-lane (sum dx, sum dy) of the "surely above" set

// Named barriers: a helper that waits for the team blocks in hardware (bar.sync); the team only arrives.
constexpr int kRgBarA = 2;        // +par: team (256, arrive) -> placer + director (64, sync): red[par] is complete
constexpr int kRgBarFull = 4;     // +par: team (256, arrive) -> finisher (32, sync): candidates of the halfspace are complete
constexpr int kRgBarPlaced = 6;   // +par: placer (32, arrive) -> team (256, sync): window / thresholds are in ctl
constexpr int kRgBarH = 8;        // +par: director (32, arrive) -> team (256, sync): canonical h / mean / flags are in ctl

struct RgBars {
  unsigned long long full[kRgFullBars];   // all chunks of halfspace (it % 8) have landed
  unsigned long long empty[kRgRing];      // the 8 sweep warps are done with the chunk
  unsigned long long fdone[3];            // finisher -> everyone: halfspace (it % 3) is finished, its buffers are free
};

struct RgRed {                           // per parity: what sweep A hands to the placer and the director
  double T[8][2];                        // canonical group totals (x, y) of the 8 groups of 32 tree slots
  float q[kRgWarps][4];                  // per warp: sum dx^2, sum dy^2, sum dx dy, sum |d|^2 bound
  float first[2];
  float pad[2];
};

struct RgFin {                           // per parity: team -> finisher
  int wc[kRgWarps];                      // samples surely / exactly above the window
  int nc[kRgWarps];                      // window candidates
  int ovf[kRgWarps];
  int pad[8];
};

__host__ __device__ inline size_t rg_smem_bytes() {
  return static_cast<size_t>(kRgRing) * kRgChunkBytes                                    // ring
         + sizeof(float2) * (kRgTeam * kRgListStride + kRgListGuard)                     // per-thread lists of kept samples
         + sizeof(float2) * kRgWarps * kRgDenseCap                                       // the same, compacted per warp
         + sizeof(double) * 2 * kRgWarps * kRgWarpCand                                   // cand [2][warps][kRgWarpCand]
         + sizeof(unsigned) * 2 * kHistBuckets                                           // hist [2][256]
         + sizeof(double) * 2 * kResolveMax                                              // small [2]
         + 2 * sizeof(RgRed) + 2 * sizeof(RgFin) + 3 * sizeof(Ctl) + sizeof(RgBars);
}

__device__ __forceinline__ float4 rg_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// sweep B, one sample: q = h_a . (xi - first) - mid in fp32 (mid = centre of the fp32 uncertainty band around the window).
//   up   = q < -half         -> shifted coordinate sums + count, by multiplication with m = 1.0 / 0.0 (FSET + full-rate
//                               FFMA / FADD: no predicate, so the samples of a group of rows pipeline freely)
//   keep = |q| <= half       -> raw copy into this thread's list (exact fp64 loss later); the only predicate
__device__ __forceinline__ void rg_classify(float q, float neg_half, float half, float dx, float dy, float x, float y,
                                            float& ax, float& ay, float& cnt, uint32_t& lp) {
  asm volatile(
      "{\n\t.reg .pred k;\n\t.reg .f32 m, aq;\n\t"
      "set.lt.f32.f32 m, %4, %5;\n\t"
      "abs.f32 aq, %4;\n\t"
      "setp.le.f32 k, aq, %6;\n\t"
      "fma.rn.f32 %0, m, %7, %0;\n\t"
      "fma.rn.f32 %1, m, %8, %1;\n\t"
      "add.f32 %2, %2, m;\n\t"
      "@k st.shared.v2.f32 [%3], {%9, %10};\n\t"
      "@k add.u32 %3, %3, 8;\n\t}"
      : "+f"(ax), "+f"(ay), "+f"(cnt), "+r"(lp)
      : "f"(q), "f"(neg_half), "f"(half), "f"(dx), "f"(dy), "f"(x), "f"(y)
      : "memory");
}

__global__ void __launch_bounds__(kRgThreads, 1) halfspace_ring_kernel(const KernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  float2* list_base = reinterpret_cast<float2*>(smem_raw + kRgRing * kRgChunkBytes);
  float2* dense_base = list_base + kRgTeam * kRgListStride + kRgListGuard;
  double* cand_base = reinterpret_cast<double*>(dense_base + kRgWarps * kRgDenseCap);
  unsigned* hist_base = reinterpret_cast<unsigned*>(cand_base + 2 * kRgWarps * kRgWarpCand);
  double* small_base = reinterpret_cast<double*>(hist_base + 2 * kHistBuckets);
  RgRed* red_base = reinterpret_cast<RgRed*>(small_base + 2 * kResolveMax);
  RgFin* fin_base = reinterpret_cast<RgFin*>(red_base + 2);
  Ctl* ctl_base = reinterpret_cast<Ctl*>(fin_base + 2);
  RgBars* bars = reinterpret_cast<RgBars*>(ctl_base + 3);

  if (tid == 0) {
    for (int j = 0; j < kRgFullBars; ++j) mbar_init(&bars->full[j], 1);
    for (int j = 0; j < kRgRing; ++j) mbar_init(&bars->empty[j], kRgWarps);
    for (int j = 0; j < 3; ++j) mbar_init(&bars->fdone[j], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 2 * kHistBuckets; i += kRgThreads) hist_base[i] = 0;
  if (tid < 3) {
    ctl_base[tid].small_n = 0;
    ctl_base[tid].z_learned = 0;
    ctl_base[tid].z_missrun = 0;
    ctl_base[tid].z_lo_use = a.z_mid_f - a.z_half_f;
    ctl_base[tid].z_hi_use = a.z_mid_f + a.z_half_f;
    ctl_base[tid].z_est = 0.f;
  }
  __syncthreads();

  const int n_vec = N >> 1;                                   // 16-byte vectors of a halfspace (N is even)
  const int rows_full = n_vec / kRgTeam;                      // rows in which every sweep thread has a vector
  const int rows_all = (n_vec + kRgTeam - 1) / kRgTeam;
  const int nch = (rows_all + 1) >> 1;                        // chunks per halfspace
  const uint32_t copy_bytes = static_cast<uint32_t>(N) * 8u;

  // ============================================================================================ producer warp
  if (warp == kRgProducerWarp) {
    if (lane == 0) {
      int it = 0;
      int slot = 0;        // ring chunk of the next copy
      int lap = 0;         // times the ring has been filled completely
      for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++it) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b) * a.stride_b * 4;
        unsigned long long* fb = &bars->full[it & (kRgFullBars - 1)];
        mbar_expect_tx(fb, copy_bytes);
        for (int c = 0; c < nch; ++c) {
          if (lap > 0) mbar_wait(&bars->empty[slot], (lap - 1) & 1);   // the team has swept the chunk's previous content twice
          const uint32_t off = static_cast<uint32_t>(c) * kRgChunkBytes;
          const uint32_t n = copy_bytes - off < kRgChunkBytes ? copy_bytes - off : kRgChunkBytes;
          bulk_g2s(smem_raw + static_cast<uint32_t>(slot) * kRgChunkBytes, src + off, n, fb);
          if (++slot == kRgRing) {
            slot = 0;
            ++lap;
          }
        }
      }
    }
    return;
  }

  // ============================================================================================ director warp
  // Canonical direction (IEEE div / sqrt chain, ~2k cycles of latency) and the mean halfspace, off the team's path.
  if (warp == kRgDirectorWarp) {
    int it = 0;
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++it) {
      const int par = it & 1, k3 = it % 3, use3 = it / 3;
      Ctl* ctl = ctl_base + k3;
      const RgRed* red = red_base + par;
      double e0 = 0.0, e1 = 0.0;
      if (a.h_in != nullptr) {
        e0 = a.h_in[2 * b];
        e1 = a.h_in[2 * b + 1];
      } else if (a.ego != nullptr) {
        e0 = a.ego[2 * b];
        e1 = a.ego[2 * b + 1];
      }
      bar_sync(kRgBarA + par, kRgTeam + 64);                             // red[par] is complete
      if (use3 > 0) mbar_wait_spin(&bars->fdone[k3], (use3 - 1) & 1);   // the finisher is done with ctl[k3] (halfspace it-3: long ago)
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) t[g] = red->T[g][j];
#pragma unroll
        for (int n = 8; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);  // adjacent-pair tree (canonical)
        w[j] = t[0];
      }
      // the lane sums were taken relative to the first sample
      const double m0 = __dadd_rn(static_cast<double>(red->first[0]), __ddiv_rn(w[0], static_cast<double>(N)));
      const double m1 = __dadd_rn(static_cast<double>(red->first[1]), __ddiv_rn(w[1], static_cast<double>(N)));
      int nonfinite = !(isfinite(m0) && isfinite(m1));
      int degenerate = 0;
      double h0, h1;
      if (a.h_in != nullptr) {
        h0 = e0;
        h1 = e1;
      } else {
        const double d0 = __dsub_rn(m0, e0), d1 = __dsub_rn(m1, e1);
        const double nrm = norm2_canon(d0, d1);
        if (nrm < 1e-10) {
          h0 = 1.0;
          h1 = 0.0;
          degenerate = 1;
        } else {
          h0 = __ddiv_rn(d0, nrm);
          h1 = __ddiv_rn(d1, nrm);
        }
      }
      nonfinite |= !(isfinite(h0) && isfinite(h1));
      if (lane == 0) {
        ctl->h0 = h0; ctl->h1 = h1; ctl->m0 = m0; ctl->m1 = m1;
        ctl->nonfinite = nonfinite;
        ctl->degenerate = degenerate;
      }
      __syncwarp();
      bar_arrive(kRgBarH + par, kRgTeam + 32);
      if (lane == 0) write_mean_outputs(a, b, m0, m1);
    }
    return;
  }

  // ============================================================================================ placer warp
  // Everything here only PLACES the window (speed, never the result) except the fp32 band, which carries rigorous error
  // bounds against the canonical direction the director computes meanwhile: the classification uses h_a = (h0f, h1f) with
  // |h_a - h| <= err_h per component.  Same derivation as halfspace_kernel.cuh.
  if (warp == kRgPlacerWarp) {
    int it = 0;
    const float inv_n_f = static_cast<float>(1.0 / static_cast<double>(N));
    const double inv_n = 1.0 / static_cast<double>(N);
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++it) {
      const int par = it & 1, k3 = it % 3, use3 = it / 3;
      Ctl* ctl = ctl_base + k3;
      const RgRed* red = red_base + par;
      double pre0 = 0.0, pre1 = 0.0;
      if (a.h_in != nullptr) {
        pre0 = a.h_in[2 * b];
        pre1 = a.h_in[2 * b + 1];
      } else if (a.ego != nullptr) {
        pre0 = a.ego[2 * b];
        pre1 = a.ego[2 * b + 1];
      }
      bar_sync(kRgBarA + par, kRgTeam + 64);
      // ctl[k3] carries the learned window of this chain (halfspaces it, it-3, it-6, ...): the finisher of it-3 wrote it
      if (use3 > 0) mbar_wait_spin(&bars->fdone[k3], (use3 - 1) & 1);
      double w0 = 0.0, w1 = 0.0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        w0 += red->T[g][0];
        w1 += red->T[g][1];
      }
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, b2 = 0.f;
#pragma unroll
      for (int g = 0; g < kRgWarps; ++g) {
        q0 += red->q[g][0];
        q1 += red->q[g][1];
        q2 += red->q[g][2];
        b2 = fmaxf(b2, red->q[g][3]);
      }
      const double f0 = static_cast<double>(red->first[0]), f1 = static_cast<double>(red->first[1]);
      const double mr0 = w0 * inv_n, mr1 = w1 * inv_n;   // mean relative to the first sample
      const double m0 = mr0 + f0, m1 = mr1 + f1;
      const float mr0f = static_cast<float>(mr0), mr1f = static_cast<float>(mr1);
      bool usable = isfinite(m0) && isfinite(m1);
      float h0f, h1f, err_h;
      if (a.h_in != nullptr) {
        h0f = static_cast<float>(pre0);
        h1f = static_cast<float>(pre1);
        err_h = fmaxf(fabsf(h0f), fabsf(h1f)) * 1.2e-7f + 1.5e-45f;
      } else {
        const double d0 = m0 - pre0, d1 = m1 - pre1;   // pre = ego
        const float d0f = static_cast<float>(d0), d1f = static_cast<float>(d1);
        const float n2 = fmaf(d0f, d0f, d1f * d1f);
        if (n2 > 0.99e-20f && n2 < 1.01e-20f) usable = false;   // too close to the degenerate-direction switch
        if (n2 < 1e-20f) {
          h0f = 1.0f;
          h1f = 0.0f;
          err_h = 0.f;
        } else {
          const float rn = rsqrtf(n2);
          h0f = d0f * rn;
          h1f = d1f * rn;
          // fp32 chain: 2 conversions, fma, rsqrt (2 ulp), multiply  ->  < 5e-7; plus the fp64 cancellation in m - ego
          const float mag = static_cast<float>(fabs(m0) + fabs(m1) + fabs(pre0) + fabs(pre1));
          err_h = 1e-6f + 4e-15f * mag * rn;
          usable = usable && isfinite(rn) && rn > 0.f && isfinite(mag);
        }
      }
      usable = usable && isfinite(h0f) && isfinite(h1f) && err_h < 1e-3f;
      const float cxx = q0 * inv_n_f - mr0f * mr0f, cyy = q1 * inv_n_f - mr1f * mr1f, cxy = q2 * inv_n_f - mr0f * mr1f;
      const float var_l = h0f * h0f * cxx + 2.0f * h0f * h1f * cxy + h1f * h1f * cyy;
      const float sigma = sqrt_approx(var_l);   // placement only
      int window_ok = a.use_window && usable && (var_l > 0.f) && isfinite(sigma);
      const float pm = fmaf(h1f, mr1f, h0f * mr0f);
      // window bounds in z units: the Gaussian plan of the host, or — after two consecutive misses in this chain, i.e.
      // samples that are evidently not Gaussian — the position learned from the chain's earlier halfspaces (finisher)
      float zlo = a.z_lo_f, zhi = a.z_hi_f;
      if (ctl->z_learned) {
        zlo = ctl->z_lo_use;
        zhi = ctl->z_hi_use;
      }
      const float a_lo = pm - zlo * sigma, a_hi = pm - zhi * sigma;
      const double c = static_cast<double>(h0f) * f0 + static_cast<double>(h1f) * f1;   // h_a . first
      const double t_lo = __dadd_rn(-static_cast<double>(a_lo) - c, 0.0);  // +0.0: never -0.0 (canonical losses are +0)
      const double t_hi = __dadd_rn(-static_cast<double>(a_hi) - c, 0.0);
      // fp32 classification of p32 = h_a . d, d = fl32(xi - first)  (p = h.xi = h.first + h.d = -L):
      //   |p32 - h_a.d_true| <= 5 * 2^-24 * (|h0f| + |h1f|) * max|d|; we allow 2^-19 (32x);
      //   |h.xi - (c + h_a.d_true)| <= err_h (|first| + max|d|) (1-norms); the fp64 roundings of c, t and L are ~1e-16 relative.
      //   p32 <  thr_above0  =>  L > t_hi  for sure;    p32 > thr_keep0  =>  L < t_lo  for sure.
      const float dmax = sqrt_approx(b2) * 1.0001f;
      const float af0 = fabsf(static_cast<float>(f0)) * 1.0001f, af1 = fabsf(static_cast<float>(f1)) * 1.0001f;
      const float habs = fabsf(h0f) + fabsf(h1f);
      const float eps = (habs * (af0 + af1 + dmax)) * 1e-15f + err_h * 1.5f * (af0 + af1 + 2.0f * dmax);
      const float bound = habs * dmax * 1.9073486e-06f + 1.1754944e-38f + eps * 1.0001f;
      const float thr_keep0 = a_lo + (bound + fabsf(a_lo) * 2.3841858e-07f);
      const float thr_above0 = a_hi - (bound + fabsf(a_hi) * 2.3841858e-07f);
      // sweep B evaluates q32 = fma(h1f, dy, fma(h0f, dx, -mid)): two more roundings than p32, each <= 2^-24 of a value
      // below |h_a|.max|d| + |mid|; the band [mid - half, mid + half] must contain [thr_above0, thr_keep0] widened by
      // that (we allow 2^-21 (|h_a| dmax + |mid|)) and by the roundings of mid and half themselves
      const float mid = 0.5f * (thr_above0 + thr_keep0);
      const float slack = (habs * dmax + fabsf(mid)) * 4.7683716e-07f + 1.1754944e-38f;
      const float half = (0.5f * (thr_keep0 - thr_above0) + slack) * 1.000001f + fabsf(mid) * 2.3841858e-07f;
      const unsigned long long klo = key_of(t_lo), khi = key_of(t_hi);
      const unsigned long long span = khi - klo;
      const int bits = span ? 64 - __clzll(static_cast<long long>(span)) : 0;
      window_ok = window_ok && isfinite(thr_keep0) && isfinite(thr_above0) && (khi >= klo) && (thr_above0 <= thr_keep0) &&
                  isfinite(t_lo) && isfinite(t_hi) && isfinite(mid) && isfinite(half);
      if (lane == 0) {
        ctl->f0 = f0; ctl->f1 = f1;
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->h0f = h0f; ctl->h1f = h1f; ctl->thr_keep = half; ctl->thr_above = mid;   // (band half-width, band centre)
        ctl->key_lo = klo;
        ctl->hist_shift = bits > 8 ? bits - 8 : 0;
        ctl->window_ok = window_ok;
        ctl->pl = Ctl::Place{pm, sigma, c};
      }
      __syncwarp();
      bar_arrive(kRgBarPlaced + par, kRgTeam + 32);
    }
    return;
  }

  // ============================================================================================ finisher warp
  if (warp >= kRgFinisherWarp) {
    const int par = warp - kRgFinisherWarp;   // this finisher takes the halfspaces of one parity
    int it = par;
    for (long long b = static_cast<long long>(blockIdx.x) + static_cast<long long>(par) * gridDim.x; b < a.B;
         b += 2ll * gridDim.x, it += 2) {
      const int k3 = it % 3;
      Ctl* ctl = ctl_base + k3;
      unsigned* hist = hist_base + par * kHistBuckets;
      const double* cand = cand_base + par * kRgWarps * kRgWarpCand;
      const RgFin* fin = fin_base + par;
      double* small = small_base + par * kResolveMax;
      bar_sync(kRgBarFull + par, kRgTeam + 32);   // the team hands halfspace b over
      int cnt_hi = 0, ncand = 0, ovf = 0;
#pragma unroll
      for (int w = 0; w < kRgWarps; ++w) {
        cnt_hi += fin->wc[w];
        ncand += fin->nc[w];
        ovf |= fin->ovf[w];
      }
      const bool clean = ctl->window_ok != 0 && ctl->nonfinite == 0 && ctl->degenerate == 0 && !ovf;
      const bool fast = clean && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
      if (fast) {
        const unsigned long long klo = ctl->key_lo;
        const int hshift = ctl->hist_shift;
        int bstar, r, cnt_in;
        scan_hist_warp(hist, a.kc - cnt_hi, lane, bstar, r, cnt_in);
        // pass over all candidates: above bucket b* -> counted / summed; bucket b* -> gathered for the exact ranking
        double s3 = 0.0;
        int c3 = 0, n_small = 0;
        for (int w = 0; w < kRgWarps; ++w) {
          const int nc = fin->nc[w];
          const double* wc = cand + w * kRgWarpCand;
          for (int j0 = 0; j0 < nc; j0 += 32) {
            const int j = j0 + lane;
            bool in_b = false;
            double L = 0.0;
            if (j < nc) {
              L = wc[j];
              const int bk = static_cast<int>((key_of(L) - klo) >> hshift);
              if (bk > bstar) {
                ++c3;
                s3 += L;
              }
              in_b = bk == bstar;
            }
            const unsigned bal = __ballot_sync(kFull, in_b);
            if (bal) {
              const int pos = n_small + __popc(bal & ((1u << lane) - 1u));
              if (in_b && pos < kResolveMax) small[pos] = L;
              n_small += __popc(bal);
            }
          }
        }
        __syncwarp();
        double T_thr, s4 = 0.0;
        int c4 = 0;
        if (cnt_in <= kResolveMax) {
          // all-pairs rank inside bucket b*; members above T are summed in rank order (deterministic)
          const double mineL = lane < cnt_in ? small[lane] : 0.0;
          const unsigned long long mine = lane < cnt_in ? key_of(mineL) : 0ull;
          int rank = 0;
          for (int j = 0; j < cnt_in; ++j) {
            const unsigned long long other = __shfl_sync(kFull, mine, j);
            rank += (other > mine) || (other == mine && j < lane);
          }
          const unsigned owner = __ballot_sync(kFull, lane < cnt_in && rank == r - 1);
          T_thr = __shfl_sync(kFull, mineL, __ffs(owner) - 1);
          __syncwarp();
          if (lane < cnt_in) small[rank] = mineL;
          __syncwarp();
          const bool mineAbove = lane < cnt_in && lane < r - 1 && key_of(small[lane]) > key_of(T_thr);
          s4 = warp_sum_any(mineAbove ? small[lane] : 0.0);
          c4 = __popc(__ballot_sync(kFull, mineAbove));
        } else {
          // dense / heavily tied bucket: narrow further inside the finisher warp
          const unsigned long long lo2 = klo + (static_cast<unsigned long long>(bstar) << hshift);
          unsigned long long hi2 = lo2 + ((1ull << hshift) - 1ull);
          const unsigned long long khi = key_of(ctl->t_hi);
          if (hi2 > khi || hi2 < lo2) hi2 = khi;
          auto each = [&](auto&& f) {
            for (int w = 0; w < kRgWarps; ++w) {
              const int nc = fin->nc[w];
              for (int j = lane; j < nc; j += 32) f(cand[w * kRgWarpCand + j]);
            }
          };
          T_thr = select_rank(each, [] { __syncwarp(); }, true, lane, 32, lo2, hi2, r, hist, small, ctl);
          each([&](double L) {
            const int bk = static_cast<int>((key_of(L) - klo) >> hshift);
            if (bk == bstar && L > T_thr) {
              ++c4;
              s4 += L;
            }
          });
          c4 = __reduce_add_sync(kFull, c4);
          s4 = warp_sum_any(s4);
        }
        const int c3t = __reduce_add_sync(kFull, c3);
        const double s3t = warp_sum_any(s3);
        double lx = 0.0, ly = 0.0;   // "surely above" coordinate sums: lane partials of the sweep warps, fixed order
#pragma unroll
        for (int w = 0; w < kRgWarps; ++w) {
          const float2 p = reinterpret_cast<const float2*>(cand + w * kRgWarpCand + kRgCandCap)[lane];
          lx += static_cast<double>(p.x);
          ly += static_cast<double>(p.y);
        }
        const double s_x = warp_sum_any(lx), s_y = warp_sum_any(ly);
        if (lane == 0) {
          // loss sum of the "surely above" set by linearity, xi_i = f + d_i:
          //   sum_i -(h.xi_i) = -(h0 (n f0 + sum dx) + h1 (n f1 + sum dy))
          const double n_lin = static_cast<double>(cnt_hi);
          const double s_lin = -(ctl->h0 * (n_lin * ctl->f0 + s_x) + ctl->h1 * (n_lin * ctl->f1 + s_y));
          const double s_tot = (s_lin + s3t) + s4;
          const int c_tot = cnt_hi + c3t + c4;
          write_risk_outputs(a, b, ctl, false, s_tot, c_tot, T_thr, 0);
          // learn where the threshold sits in z units: (T - mean loss) / sigma = (pm + T + c) / sigma
          ctl->z_missrun = 0;
          if (ctl->z_learned) {
            const float zT = (ctl->pl.pm + static_cast<float>(T_thr + ctl->pl.c_shift)) / ctl->pl.sigma;
            const float ze = 0.5f * (ctl->z_est + zT);
            if (isfinite(ze) && fabsf(ze) < 8.f) {
              ctl->z_est = ze;
              ctl->z_lo_use = ze - a.z_half_adapt_f;
              ctl->z_hi_use = ze + a.z_half_adapt_f;
            }
          }
        }
      } else if (lane == 0) {
        a.redo_list[b] = 1;   // the exact general path of the redo pass takes this halfspace
        // a placed window that missed: after two in a row in this chain (or already in learned mode) move a learned
        // centre past the window, towards the side the threshold is on
        if (clean) {
          const int run = ++ctl->z_missrun;
          if (run >= 2 || ctl->z_learned) {
            const float z_used = ctl->z_learned ? ctl->z_est : a.z_mid_f;
            const float half_used = ctl->z_learned ? a.z_half_adapt_f : a.z_half_f;
            const float z_new = z_used + (a.kc <= cnt_hi ? 2.0f : -2.0f) * half_used;
            if (isfinite(z_new) && fabsf(z_new) < 8.f) {
              ctl->z_learned = 1;
              ctl->z_est = z_new;
              ctl->z_lo_use = z_new - a.z_half_adapt_f;
              ctl->z_hi_use = z_new + a.z_half_adapt_f;
            }
          }
        }
      }
      // hand the buffers back: histogram zeroed, counters reset
      for (int i = lane; i < kHistBuckets; i += 32) hist[i] = 0;
      if (lane == 0) ctl->small_n = 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->fdone[k3]);
    }
    return;
  }

  // ============================================================================================ sweep team
  const uint32_t ring_s = smem_u32(smem_raw);
  const uint32_t toff = 16u * tid;
  float2* my_list = list_base + tid * kRgListStride;
  const uint32_t lbase = smem_u32(my_list);
  float2* wdense = dense_base + warp * kRgDenseCap;
  const int ngroups = (nch + 1) >> 1;             // groups of two chunks (four rows) per halfspace
  const int g_full = rows_full >> 2;              // groups whose four rows are complete

  // ------------------------------------------------------------------ sweep A of halfspace `n` (chunks from ring slot g0)
  // Row r = the 16-byte vector r*256 + tid: samples 2(r*256+tid)+{0,1} = fp32 lanes 2 slot + {0,1} of the canonical
  // 1024-lane sums, slot = (r & 1)*256 + tid; a chunk = rows (2c, 2c+1).
  auto sweep_a = [&](int n, int g0) {
    const int par = n & 1;
    RgRed* red = red_base + par;
    mbar_wait_spin(&bars->full[n & (kRgFullBars - 1)], (n / kRgFullBars) & 1);
    float fx, fy;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(fx), "=f"(fy) : "r"(ring_s + static_cast<uint32_t>(g0) * kRgChunkBytes) : "memory");
    const float2 nf = make_float2(-fx, -fy);
    float2 acc[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) acc[q][0] = acc[q][1] = make_float2(0.f, 0.f);
    float2 sq[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float sxy[2] = {0.f, 0.f};
    auto body = [&](const float4 x, int q) {
      const float2 d0 = __fadd2_rn(make_float2(x.x, x.y), nf), d1 = __fadd2_rn(make_float2(x.z, x.w), nf);
      acc[q][0] = __fadd2_rn(acc[q][0], d0);
      acc[q][1] = __fadd2_rn(acc[q][1], d1);
      sq[q] = __ffma2_rn(d0, d0, sq[q]);
      sq[q] = __ffma2_rn(d1, d1, sq[q]);
      sxy[q] = fmaf(d0.x, d0.y, sxy[q]);
      sxy[q] = fmaf(d1.x, d1.y, sxy[q]);
    };
    // groups of two chunks = four rows; the loads of group g+1 are in flight while group g is summed (two sweep warps per
    // scheduler cannot hide the shared-memory latency by themselves)
    int slot = g0;
    auto load = [&](float4 (&v)[4]) {
      const uint32_t s0 = ring_s + static_cast<uint32_t>(slot) * kRgChunkBytes + toff;
      slot = slot + 1 == kRgRing ? 0 : slot + 1;
      const uint32_t s1 = ring_s + static_cast<uint32_t>(slot) * kRgChunkBytes + toff;
      slot = slot + 1 == kRgRing ? 0 : slot + 1;
      v[0] = rg_lds128(s0);
      v[1] = rg_lds128(s0 + kRgRowBytes);
      v[2] = rg_lds128(s1);
      v[3] = rg_lds128(s1 + kRgRowBytes);
    };
    auto compute = [&](float4 (&v)[4], int g) {
      if (g >= g_full) {   // ragged end: vectors beyond the data count as the first sample, whose shifted value +0 adds nothing
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((4 * g + i) * kRgTeam + tid >= n_vec) v[i] = make_float4(fx, fy, fx, fy);
      }
      body(v[0], 0);
      body(v[1], 1);
      body(v[2], 0);
      body(v[3], 1);
    };
    {
      float4 va[4], vb[4];
      load(va);
      int g = 0;
#pragma unroll 1
      for (; g + 1 < ngroups; g += 2) {
        load(vb);
        compute(va, g);
        if (g + 2 < ngroups) load(va);
        compute(vb, g + 1);
      }
      if (g < ngroups) compute(va, g);
    }
    // adjacent fp32 lanes (2 slot, 2 slot + 1) widened and added in fp64, then slot tid + slot tid+256
    const double s0x = __dadd_rn(static_cast<double>(acc[0][0].x), static_cast<double>(acc[0][1].x));
    const double s0y = __dadd_rn(static_cast<double>(acc[0][0].y), static_cast<double>(acc[0][1].y));
    const double s1x = __dadd_rn(static_cast<double>(acc[1][0].x), static_cast<double>(acc[1][1].x));
    const double s1y = __dadd_rn(static_cast<double>(acc[1][0].y), static_cast<double>(acc[1][1].y));
    const double u_x = __dadd_rn(s0x, s1x), u_y = __dadd_rn(s0y, s1y);
    const double txy = warp_sum_canon_pair(u_x, u_y, lane);   // x total in even lanes, y total in odd lanes
    const float qxx = sq[0].x + sq[1].x, qyy = sq[0].y + sq[1].y, qxy = sxy[0] + sxy[1];
    const float mq = warp_sum_any4(qxx, qyy, qxy, 0.f, lane);   // lanes (lane & 3) = 0: qxx, 1: qxy, 2: qyy
    const unsigned bnd = __reduce_max_sync(kFull, __float_as_uint(qxx + qyy));
    if (lane < 2) red->T[warp][lane] = txy;
    if (lane < 3) red->q[warp][lane == 0 ? 0 : (lane == 1 ? 2 : 1)] = mq;
    if (lane == 3) red->q[warp][3] = __uint_as_float(bnd);
    if (tid == 4) {
      red->first[0] = fx;
      red->first[1] = fy;
    }
    __syncwarp();
    bar_arrive(kRgBarA + par, kRgTeam + 64);   // placer and director start on halfspace n
  };

  RG_PH_DECL
  int it = 0;
  int g0 = 0;   // ring slot of the first chunk of halfspace `it`
  if (static_cast<long long>(blockIdx.x) < a.B) sweep_a(0, 0);
  RG_PH_MARK(0)
  for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++it) {
    const int par = it & 1, k3 = it % 3;
    Ctl* ctl = ctl_base + k3;
    int g_next = g0 + nch;
    if (g_next >= kRgRing) g_next -= kRgRing;
    if (b + gridDim.x < a.B) sweep_a(it + 1, g_next);
    RG_PH_MARK(1)

    // ------------------------------------------------------------------ sweep B of halfspace `it`
    bar_sync(kRgBarPlaced + par, kRgTeam + 32);   // (the placer finished while sweep A of the next halfspace ran)
    RG_PH_MARK(2)
    float ax = 0.f, ay = 0.f, cf = 0.f;
    uint32_t lp = lbase;
    float fx, fy;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(fx), "=f"(fy) : "r"(ring_s + static_cast<uint32_t>(g0) * kRgChunkBytes) : "memory");
    const float2 nf = make_float2(-fx, -fy);
    bool swept = ctl->window_ok != 0;
#ifdef DRCVAR_PROFILE_PHASES
    if (a.debug & 1) swept = false;
#endif
    {
      const float h0f = ctl->h0f, h1f = ctl->h1f, half = ctl->thr_keep, neg_mid = -ctl->thr_above, neg_half = -half;
      const float inf_ = __int_as_float(0x7f800000);
      auto classify_row = [&](const float4 x) {
        const float2 d0 = __fadd2_rn(make_float2(x.x, x.y), nf), d1 = __fadd2_rn(make_float2(x.z, x.w), nf);
        const float q0 = fmaf(h1f, d0.y, fmaf(h0f, d0.x, neg_mid)), q1 = fmaf(h1f, d1.y, fmaf(h0f, d1.x, neg_mid));
        rg_classify(q0, neg_half, half, d0.x, d0.y, x.x, x.y, ax, ay, cf, lp);
        rg_classify(q1, neg_half, half, d1.x, d1.y, x.z, x.w, ax, ay, cf, lp);
      };
      auto classify_masked = [&](const float4 x, bool ok) {   // vectors beyond the data get q = +inf: neither above nor kept
        const float2 d0 = __fadd2_rn(make_float2(x.x, x.y), nf), d1 = __fadd2_rn(make_float2(x.z, x.w), nf);
        float q0 = fmaf(h1f, d0.y, fmaf(h0f, d0.x, neg_mid)), q1 = fmaf(h1f, d1.y, fmaf(h0f, d1.x, neg_mid));
        q0 = ok ? q0 : inf_;
        q1 = ok ? q1 : inf_;
        rg_classify(q0, neg_half, half, ok ? d0.x : 0.f, ok ? d0.y : 0.f, x.x, x.y, ax, ay, cf, lp);
        rg_classify(q1, neg_half, half, ok ? d1.x : 0.f, ok ? d1.y : 0.f, x.z, x.w, ax, ay, cf, lp);
      };
      int slot = g0;
      int sl[2][2];
      auto load = [&](float4 (&v)[4], int k) {
        sl[k][0] = slot;
        const uint32_t s0 = ring_s + static_cast<uint32_t>(slot) * kRgChunkBytes + toff;
        slot = slot + 1 == kRgRing ? 0 : slot + 1;
        sl[k][1] = slot;
        const uint32_t s1 = ring_s + static_cast<uint32_t>(slot) * kRgChunkBytes + toff;
        slot = slot + 1 == kRgRing ? 0 : slot + 1;
        if (swept) {
          v[0] = rg_lds128(s0);
          v[1] = rg_lds128(s0 + kRgRowBytes);
          v[2] = rg_lds128(s1);
          v[3] = rg_lds128(s1 + kRgRowBytes);
        }
      };
      auto compute = [&](float4 (&v)[4], int g, int k) {
        if (swept) {
          if (g < g_full) {
            classify_row(v[0]);
            classify_row(v[1]);
            classify_row(v[2]);
            classify_row(v[3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) classify_masked(v[i], (4 * g + i) * kRgTeam + tid < n_vec);
          }
        }
        // both chunks have been read by this warp (the classification consumed every loaded value): hand them back
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->empty[sl[k][0]]);
          if (2 * g + 1 < nch) mbar_arrive(&bars->empty[sl[k][1]]);
        }
      };
      {
        float4 va[4], vb[4];
        load(va, 0);
        int g = 0;
#pragma unroll 1
        for (; g + 1 < ngroups; g += 2) {
          load(vb, 1);
          compute(va, g, 0);
          if (g + 2 < ngroups) load(va, 0);
          compute(vb, g + 1, 1);
        }
        if (g < ngroups) compute(va, g, 0);
      }
    }
    RG_PH_MARK(3)

    // ------------------------------------------------------------------ exact phase of halfspace `it`
    // canonical fp64 loss of the kept samples; window candidates -> warp-private list + histogram on order-preserving keys
    {
      RgFin* fin = fin_base + par;
      unsigned* hist = hist_base + par * kHistBuckets;
      double* wcand = cand_base + par * kRgWarps * kRgWarpCand + warp * kRgWarpCand;
      // compact the per-thread lists of this warp (lane order, then list order: deterministic)
      const int mine_n = static_cast<int>((lp - lbase) >> 3);
      int incl = mine_n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
      }
      const int n_list = __shfl_sync(kFull, incl, 31);
      const int n_max = __reduce_max_sync(kFull, mine_n);
      bool overflow = n_max > kRgListCap || n_list > kRgDenseCap;
      if (!overflow) {
        float2* dst = wdense + (incl - mine_n);
        for (int k = 0; k < n_max; ++k)
          if (k < mine_n) dst[k] = my_list[k];
      }
      __syncwarp();
      bar_sync(kRgBarH + par, kRgTeam + 32);   // canonical h of this halfspace (the director finished it long ago)
      // cand / hist / fin of this parity were last used by halfspace it-2: its finisher is done (long ago)
      if (it >= 2) mbar_wait_spin(&bars->fdone[(it - 2) % 3], ((it - 2) / 3) & 1);
      const double h0 = ctl->h0, h1 = ctl->h1;
      const double t_lo = ctl->t_lo, t_hi = ctl->t_hi;
      const unsigned long long klo = ctl->key_lo;
      const int hshift = ctl->hist_shift;
      int nc = 0;
      if (!overflow) {
        for (int k0 = 0; k0 < n_list; k0 += 32) {
          const int k = k0 + lane;
          const bool active = k < n_list;
          float2 s = make_float2(fx, fy);
          double L = 0.0;
          if (active) {
            s = wdense[k];
            L = loss_of(h0, h1, static_cast<double>(s.x), static_cast<double>(s.y));
          }
          const bool up = active && (L > t_hi);
          const bool cd = active && !up && (L >= t_lo);
          if (up) {   // inside the fp32 uncertainty band but exactly above the window: joins the "above" set
            cf += 1.0f;
            ax += __fadd_rn(s.x, -fx);
            ay += __fadd_rn(s.y, -fy);
          }
          const unsigned bal = __ballot_sync(kFull, cd);
          if (bal) {
            const int pos = nc + __popc(bal & ((1u << lane) - 1u));
            if (cd && pos < kRgCandCap) {
              wcand[pos] = L;
              atomicAdd(&hist[static_cast<unsigned>((key_of(L) - klo) >> hshift)], 1u);
            }
            nc += __popc(bal);
          }
        }
        overflow = nc > kRgCandCap;
      }
      const int wc = __reduce_add_sync(kFull, static_cast<int>(cf));
      reinterpret_cast<float2*>(wcand + kRgCandCap)[lane] = make_float2(ax, ay);
      if (lane == 0) {
        fin->wc[warp] = wc;
        fin->nc[warp] = nc < kRgCandCap ? nc : kRgCandCap;
        fin->ovf[warp] = (overflow || !swept) ? 1 : 0;
      }
      __syncwarp();
      bar_arrive(kRgBarFull + par, kRgTeam + 32);
    }
    RG_PH_MARK(4)
    g0 = g_next;
  }
#ifdef DRCVAR_PROFILE_PHASES
  if (tid == 32 && a.phase_cycles)
    for (int k = 0; k < 8; ++k) a.phase_cycles[(blockIdx.x * 2 + 0) * 12 + k] = rg_t[k];
#endif
}

}  // namespace drcvar
