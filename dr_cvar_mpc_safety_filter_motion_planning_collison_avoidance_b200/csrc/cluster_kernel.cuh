// cluster_kernel.cuh — single-read kernel for sample counts beyond one CTA's shared memory (N > kOctantMinN; BASELINE
// config 5, N = 100 000): ONE THREAD-BLOCK CLUSTER PER HALFSPACE, the samples spread over the distributed shared memory
// of its 2 / 4 / 8 CTAs (one CTA per SM, each holding 8 / C octants of the canonical mean contract).
//
// Per CTA: 16 sweep warps (the team) + producer, director and finisher warps, persistent over the batch (cluster-strided).
//   producer  1-D TMA bulk copies (cp.async.bulk, 32 KB chunks, one transaction mbarrier per chunk) of this CTA's part of
//             halfspace b+1, each chunk issued the moment all 16 sweep warps have released it in sweep B of halfspace b
//   sweep A   trails the chunks as they land: canonical lane sums per octant (shifted by the first sample, packed fp32
//             partials, fp64 tree) + second moments
//   exchange1 every CTA sends its octant totals + moments to every CTA (st.async through DSMEM, completion counted on the
//             destination's mbarrier); all CTAs then compute the SAME window / fp32 thresholds (warp 0) and the SAME canonical
//             direction (director warp: IEEE div / sqrt chain, off the team's path)              core/geometry.py:35-53
//   sweep B   fp32 classification with the rigorous bound of the resident kernel: surely above the window (count +
//             shifted coordinate sums; the loss sum follows from linearity), surely below (ignored), or "needs the exact
//             fp64 loss" -> raw copy into a per-warp list; every chunk is handed back (and refilled by TMA) as soon as the last warp
//             is past it
//   phase 2b  canonical fp64 loss of the listed samples; window candidates are compacted per warp
//   exchange2 partial counts / sums and the candidates go to the halfspace's leader CTA (rotating) through DSMEM
//   finish    leader: exact rank among the candidates (radix narrowing), CVaR, offsets          core/risk_metrics.py:84-338
// A window miss (3e-5 of Gaussian halfspaces), an overflow or non-finite data put the halfspace on a redo list that the
// streaming kernel processes right after (exact general select, same arithmetic contract).  Samples that are not
// Gaussian: the finishers publish where the threshold really sits ({tag, learned, z} in one word, shared through
// exchange 1, newest wins in every CTA alike) and the cluster switches to a learned, wider window.  Generate mode (samples ==
// nullptr): the team draws this CTA's part into the slot (sample_gen.cuh, once per halfspace — the streaming kernel re-draws
// it in each of its passes) instead of waiting for TMA.  Tail indices stay on the streaming kernel.  fp32 inputs only.
#pragma once

#include "halfspace_kernel.cuh"
#undef DRCVAR_FILE_ID
#define DRCVAR_FILE_ID 4

#ifdef DRCVAR_PROFILE_PHASES
#define CL_DBG(bit) ((a.debug & (bit)) != 0)   // ablation switches of the profiling build (wrong results, timing only)
#else
#define CL_DBG(bit) false
#endif

namespace drcvar {

constexpr int kClTeamWarps = 16;
constexpr int kClTeam = kClTeamWarps * 32;        // 512: thread t owns slot t of the canonical tree
constexpr int kClThreads = kClTeam + 96;          // + producer, director and finisher warps
constexpr int kClProducerWarp = kClTeamWarps;
constexpr int kClDirectorWarp = kClTeamWarps + 1;
constexpr int kClFinisherWarp = kClTeamWarps + 2;
constexpr int kClBarDirector = 2;                 // named barriers 2, 3 (one per parity): warp 0 of the team -> director
constexpr int kClMaxCtas = 8;
constexpr int kClMaxChunks = 8;                   // 32 KB chunks of one CTA's part (<= 227 KB)
constexpr int kClWarpList = 40;                   // masked samples (raw copies) per sweep warp
constexpr int kClPool = 1408;                     // candidate losses gathered at the leader (all CTAs together)
constexpr int kClX1 = 14;                         // doubles per source CTA in exchange 1
constexpr uint32_t kClLaneRow = 8192;             // one 16-byte load per team thread

struct ClShared {
  double pool[kClPool];                           // leader: candidates [src][cap], read by the finisher warp
  double xch[512];                                // sweep A: slots 256..511 on their way to threads 0..255
  double x1[2][kClMaxCtas][kClX1];                // exchange 1 [parity][src]: octant totals (x, y)..., qxx, qyy, qxy, bound2
  double x2[kClMaxCtas][4];                       // exchange 2 at the leader [src]: n_above, sum dx, sum dy, ncand (-1: overflow)
  float2 list[kClTeamWarps][kClWarpList];
  double red[kClTeamWarps * 2];
  double octtot[kOctants * 2];
  alignas(16) double wsum[kClTeamWarps * 4];
  alignas(16) float mom[kClTeamWarps * 4];
  double dmom[kClTeamWarps * 6];                  // fp64 kernel: per-warp qxx, qyy, qxy, qdx, qdy, n_sub
  int wcnt[kClTeamWarps * 2];
  unsigned hist[kHistBuckets];
  double small[kResolveMax];
  Ctl ctl[2];
  float2 gfirst;                                  // generate mode: sample 0 of the current halfspace
  unsigned long long zpub;                        // learned window state of this CTA's finisher: {(tag << 1 | learned), z_est bits}
  int fin_missrun;                                // finisher: consecutive window misses it has seen
  Ctl fin_ctl;                                    // leader: copy of ctl[par] for the finisher warp (it may lag behind)
  unsigned long long fdone;                       // finisher -> team: pool / x2 / fin_ctl may be refilled
  unsigned long long full[kClMaxChunks];
  unsigned long long free_[kClMaxChunks];
  unsigned long long xbar1[2];
  unsigned long long xbar2;
  unsigned long long hdone[2];
  // parity (tail-index) mode: the leader's finisher tells every CTA the threshold, where its indices start and how many
  // threshold-valued samples it may still emit
  alignas(8) double tmsg[4];                      // T, first output position of this CTA, its tie quota, 1 = emit / 0 = redo
  unsigned long long tbar;
  int tcnt[kClTeamWarps * 2];                     // per sweep warp: samples above T, samples equal to T
};

__host__ __device__ inline size_t cluster_smem_bytes(long long n, int ctas, size_t elem_bytes = 4) {
  const size_t part = static_cast<size_t>(kOctants / ctas) * static_cast<size_t>(octant_bytes(n, elem_bytes));
  return ((part + 127) & ~static_cast<size_t>(127)) + sizeof(ClShared) + 128;
}

// ---------------------------------------------------------------------------------------------- cluster PTX helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Remote store whose completion is counted in bytes on the DESTINATION CTA's mbarrier (like a TMA copy): the data is
// visible to whoever observes the phase completion, and the sender needs no fence.  (A release.cluster arrive compiles
// to MEMBAR.ALL.GPU, microseconds under full HBM load: measured 12 k cycles per halfspace in the first version.)
__device__ __forceinline__ void st_async_f64(uint32_t addr, double v, uint32_t cluster_bar_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr),
               "l"(__double_as_longlong(v)), "r"(cluster_bar_addr)
               : "memory");
}
// One arrival on a (remote) mbarrier that also announces the bytes this sender's st.async stores will complete.
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t cluster_bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar_addr), "r"(bytes)
               : "memory");
}
// Wait for an exchange: the payload was written into THIS CTA's shared memory by st.async, whose bytes complete on the
// barrier exactly like a TMA copy, so the plain (acquire.cta) wait every TMA consumer uses is enough.  An
// acquire.cluster wait adds CCTL.IVALL (L1 invalidate) per warp: 19 % of all stall samples in the first profile.
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, uint32_t parity) { mbar_wait_idle(bar, parity); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {   // explicit shared-space load (no generic -> shared conversion in the loops)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cl_team_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kClTeam) : "memory"); }

__device__ __forceinline__ double pair_tree8(const double* t_in, int stride) {
  double t[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) t[g] = t_in[g * stride];
#pragma unroll
  for (int n = 8; n > 1; n >>= 1)
#pragma unroll
    for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);
  return t[0];
}

// ---------------------------------------------------------------------------------------------- the kernel
// kRawB: sweep B classifies and sums the RAW coordinates instead of the coordinates relative to the first sample (one FADD2
// less per sample; same bound and same hand-back rule as pipelined_kernel<float, 8, true>, see pipelined_kernel.cuh).
template <bool kGen, bool kTail = false, bool kRawB = false>
__global__ void __launch_bounds__(kClThreads, 1) cluster_kernel_f32(const KernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const int C = a.cl_ctas, O = kOctants / C;
  const int lgO = 31 - __clz(O);   // O = 1, 2 or 4
  const uint32_t rank = cluster_ctarank();
  const long long q = blockIdx.x / C, n_clusters = gridDim.x / C;
  const uint32_t oct_b = static_cast<uint32_t>(octant_bytes(N, 4));
  const uint32_t row_b = static_cast<uint32_t>(N) * 8u;
  const uint32_t part_cap = static_cast<uint32_t>(O) * oct_b;
  const uint32_t part_lo = rank * part_cap;
  const uint32_t part_b = row_b > part_lo ? (row_b - part_lo < part_cap ? row_b - part_lo : part_cap) : 0u;
  const int n_chunks = static_cast<int>((part_b + kBulkChunk - 1) / kBulkChunk);
  const size_t slot_bytes = (static_cast<size_t>(part_cap) + 127) & ~static_cast<size_t>(127);
  ClShared* sh = reinterpret_cast<ClShared*>(smem_raw + slot_bytes);
  const int cap = kClPool / C;   // candidates per source CTA in the leader's pool
  constexpr bool gen = kGen;   // generate mode is a separate instantiation: the stored-sample kernel's code is untouched

  if (tid == 0) {
    for (int j = 0; j < kClMaxChunks; ++j) {
      mbar_init(&sh->full[j], 1);
      mbar_init(&sh->free_[j], kClTeamWarps);
    }
    mbar_init(&sh->xbar1[0], C);
    mbar_init(&sh->xbar1[1], C);
    mbar_init(&sh->xbar2, kClTeamWarps * C);
    mbar_init(&sh->hdone[0], 1);
    mbar_init(&sh->hdone[1], 1);
    mbar_init(&sh->fdone, 1);
    mbar_init(&sh->tbar, 1);
    mbar_fence_init();
    sh->zpub = 0ull;
    sh->fin_missrun = 0;
  }
  if (tid < 2 * kClTeamWarps) sh->wcnt[tid] = 0;
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers exist before anybody arrives on them remotely

  // this CTA's part of halfspace b, chunk j: global -> shared, completion counted on full[j]
  auto issue_chunk = [&](long long b, int j) {
    const unsigned char* src =
        reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b) * a.stride_b * sizeof(float) + part_lo;
    const uint32_t off = static_cast<uint32_t>(j) * kBulkChunk;
    const uint32_t n = part_b - off < kBulkChunk ? part_b - off : kBulkChunk;
    DRCVAR_ASSERT(b >= 0 && b < a.B && j >= 0 && j < kClMaxChunks && off + n <= slot_bytes && n > 0u && (n & 15u) == 0u);
    mbar_expect_tx(&sh->full[j], n);
    bulk_g2s(smem_raw + off, src + off, n, &sh->full[j]);
  };
  // ============================================================================================ producer warp
  if (warp == kClProducerWarp) {
    if (lane == 0 && !gen) {
      int it = 0;
      for (long long b = q; b < a.B; b += n_clusters, ++it)
        for (int j = 0; j < n_chunks; ++j) {
          if (it > 0) mbar_wait(&sh->free_[j], (it - 1) & 1);
          if (!(CL_DBG(2) && it > 0)) issue_chunk(b, j);
        }
    }
    return;
  }

  // ============================================================================================ director warp
  if (warp == kClDirectorWarp) {
    int it = 0;
    for (long long b = q; b < a.B; b += n_clusters, ++it) {
      const int par = it & 1;
      Ctl* ctl = &sh->ctl[par];
      bar_sync(kClBarDirector + par, 64);   // warp 0 of the team has seen exchange 1 complete: x1[par] is in
      const double f0 = ctl->f0, f1 = ctl->f1;   // first sample (shift origin), left by warp 0 before the hand-off
      // octant o = src * O + k lives at x1[par][src][2k + j]
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[kOctants];
#pragma unroll
        for (int o = 0; o < kOctants; ++o) t[o] = sh->x1[par][o >> lgO][2 * (o & (O - 1)) + j];
#pragma unroll
        for (int n = kOctants; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);
        w[j] = t[0];
      }
      const double m0 = __dadd_rn(f0, __ddiv_rn(w[0], static_cast<double>(N)));
      const double m1 = __dadd_rn(f1, __ddiv_rn(w[1], static_cast<double>(N)));
      int nonfinite = !(isfinite(m0) && isfinite(m1));
      int degenerate = 0;
      double h0, h1;
      if (a.h_in != nullptr) {
        h0 = a.h_in[2 * b];
        h1 = a.h_in[2 * b + 1];
      } else {
        const double e0 = a.ego ? a.ego[2 * b] : 0.0, e1 = a.ego ? a.ego[2 * b + 1] : 0.0;
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
      if (lane == 0) {
        mbar_arrive(&sh->hdone[par]);
        if (static_cast<int>(rank) == it % C) write_mean_outputs(a, b, m0, m1);
      }
    }
    return;
  }

  // ============================================================================================ finisher warp
  // Leader CTA of halfspace it (it % C == rank): exact rank among the gathered window candidates, CVaR, offsets — off the
  // sweep team's path, which goes straight on to the next halfspace.                         core/risk_metrics.py:84-338
  if (warp == kClFinisherWarp) {
    int it = 0;
    uint32_t n_lead = 0;
    for (long long b = q; b < a.B; b += n_clusters, ++it) {
      if (it % C != static_cast<int>(rank)) continue;
      while (!mbar_try_wait(&sh->xbar2, n_lead & 1u)) __nanosleep(400);   // off everybody's path: poll slowly
      ++n_lead;
      Ctl* fc = &sh->fin_ctl;
      bool fast = fc->window_ok != 0 && fc->nonfinite == 0;
      double n_above = 0.0, sdx = 0.0, sdy = 0.0;
      int ncand = 0;
      if (fast) {
        for (int s = 0; s < C; ++s) {
          n_above += sh->x2[s][0];
          sdx += sh->x2[s][1];
          sdy += sh->x2[s][2];
          const double c = sh->x2[s][3];
          if (c < 0.0) fast = false; else ncand += static_cast<int>(c);
        }
      }
      const int cnt_hi = static_cast<int>(n_above);
      fast = fast && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
      if (fast) {
        auto each = [&](auto&& f) {
          for (int s = 0; s < C; ++s) {
            const int n_s = static_cast<int>(sh->x2[s][3]);
            for (int j = lane; j < n_s; j += 32) f(sh->pool[s * cap + j]);
          }
        };
        const double T_thr = select_rank(each, [] { __syncwarp(); }, true, lane, 32, key_of(fc->t_lo), key_of(fc->t_hi),
                                         a.kc - cnt_hi, sh->hist, sh->small, fc);
        int c4 = 0;
        double s4 = 0.0;
        each([&](double L) {
          if (L > T_thr) {
            ++c4;
            s4 += L;
          }
        });
        c4 = __reduce_add_sync(kFull, c4);
        s4 = warp_sum_any(s4);
        if (lane == 0) {
          // loss sum of the "surely above" set by linearity, xi_i = f + d_i:  sum_i -(h.xi_i) = -(h0 (n f0 + sum dx) + h1 (n f1 + sum dy))
          // (kRawB: the sums are of the raw coordinates)
          const double s_lin = kRawB ? -(fc->h0 * sdx + fc->h1 * sdy)
                                     : -(fc->h0 * (n_above * fc->f0 + sdx) + fc->h1 * (n_above * fc->f1 + sdy));
          write_risk_outputs(a, b, fc, false, s_lin + s4, cnt_hi + c4, T_thr, fc->degenerate ? kStatusDegenerate : 0);
        }
        if constexpr (kTail) {
          // parity mode: every CTA emits the tail indices of ITS samples (they are still in its shared memory).  The leader
          // knows, per source CTA s, how many of its samples lie above T — the "surely above" count it sent plus its window
          // candidates above T — and how many equal T (ties are window candidates too): that fixes where the indices of
          // CTA s start and how many threshold-valued samples it may emit (ties go to the LOWER index: earlier CTAs first).
          const int need_eq = a.kc - (cnt_hi + c4);
          int base = 0, eq_seen = 0, my_base = 0, my_quota = 0;
          for (int s = 0; s < C; ++s) {
            const int n_s = static_cast<int>(sh->x2[s][3]);
            int gt_s = 0, eq_s = 0;
            for (int j = lane; j < n_s; j += 32) {
              const double L = sh->pool[s * cap + j];
              gt_s += L > T_thr;
              eq_s += L == T_thr;
            }
            gt_s = __reduce_add_sync(kFull, gt_s) + static_cast<int>(sh->x2[s][0]);
            eq_s = __reduce_add_sync(kFull, eq_s);
            int quota = need_eq - eq_seen;
            quota = quota < 0 ? 0 : (quota > eq_s ? eq_s : quota);
            if (lane == s) {
              my_base = base;
              my_quota = quota;
            }
            base += gt_s + quota;
            eq_seen += eq_s;
          }
          if (lane < C) {   // lane d serves destination CTA d
            const uint32_t dst = mapa_u32(smem_u32(&sh->tmsg[0]), static_cast<uint32_t>(lane));
            const uint32_t bar = mapa_u32(smem_u32(&sh->tbar), static_cast<uint32_t>(lane));
            mbar_arrive_expect_tx_remote(bar, 32u);
            st_async_f64(dst, T_thr, bar);
            st_async_f64(dst + 8, static_cast<double>(my_base), bar);
            st_async_f64(dst + 16, static_cast<double>(my_quota), bar);
            st_async_f64(dst + 24, 1.0, bar);
          }
        }
        if (lane == 0) {   // where the threshold sits, in sigma units around the loss mean (as in the resident kernel)
          const float zT = (fc->pl.pm + static_cast<float>(T_thr + fc->pl.c_shift)) / fc->pl.sigma;
          sh->fin_missrun = 0;
          const float z_new = fc->z_learned ? 0.5f * (fc->z_est + zT) : zT;
          if (isfinite(z_new) && fabsf(z_new) < 8.f)
            sh->zpub = (static_cast<unsigned long long>((static_cast<unsigned>(it + 1) << 1) | (fc->z_learned ? 1u : 0u)) << 32) |
                       __float_as_uint(z_new);
        }
      } else {
       if constexpr (kTail) {   // the teams wait for a message whenever a window was placed: tell them there is nothing to emit
        if (fc->window_ok != 0 && lane < C) {
          const uint32_t dst = mapa_u32(smem_u32(&sh->tmsg[0]), static_cast<uint32_t>(lane));
          const uint32_t bar = mapa_u32(smem_u32(&sh->tbar), static_cast<uint32_t>(lane));
          mbar_arrive_expect_tx_remote(bar, 32u);
          for (int i = 0; i < 4; ++i) st_async_f64(dst + 8u * i, 0.0, bar);
        }
       }
       if (lane == 0) {
        a.redo_list[b] = 1;   // redo flag of halfspace b
        // a placed window that missed: after two in a row (or already in learned mode) move a learned centre past the
        // window, towards the side the threshold is on
        bool clean = fc->window_ok != 0 && fc->nonfinite == 0;
        for (int s = 0; s < C && clean; ++s) clean = sh->x2[s][3] >= 0.0;
        if (clean) {
          const int run = ++sh->fin_missrun;
          if (run >= 2 || fc->z_learned) {
            const float z_used = fc->z_learned ? fc->z_est : a.z_mid_f;
            const float half_used = fc->z_learned ? a.z_half_adapt_f : a.z_half_f;
            const float z_new = z_used + (a.kc <= cnt_hi ? 2.0f : -2.0f) * half_used;
            if (isfinite(z_new) && fabsf(z_new) < 8.f)
              sh->zpub = (static_cast<unsigned long long>((static_cast<unsigned>(it + 1) << 1) | 1u) << 32) | __float_as_uint(z_new);
          }
        }
       }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->fdone);
    }
    return;
  }

  // ============================================================================================ sweep team
  const uint32_t slot_s = smem_u32(smem_raw);
  const uint32_t toff = 16u * tid;
  const uint32_t woff = 16u * (tid & ~31);   // first byte of this warp inside a lane-row
  float2* wlist = sh->list[warp];
  int it = 0;
  PH_DECL
  uint32_t n_lead = 0;    // halfspaces this CTA has led so far (phase of fdone)
  uint32_t tphase = 0;    // parity mode: phase of tbar (one message per halfspace whose window was placed)
  for (long long b = q; b < a.B; b += n_clusters, ++it) {
    const int par = it & 1;
    const int leader = it % C;
    Ctl* ctl = &sh->ctl[par];
    float2 first;
    if constexpr (gen) {
      // generate mode: draw this CTA's part of halfspace b into the slot — nominal position + L z, z from Philox4x32-10 +
      // Box-Muller, one Philox call per pair of samples (sample_gen.cuh; same stream as the other kernels)
      const float gmx = static_cast<float>(a.gen_mean[2 * b]), gmy = static_cast<float>(a.gen_mean[2 * b + 1]);
      const float gl00 = static_cast<float>(a.gen_chol[3 * b]), gl10 = static_cast<float>(a.gen_chol[3 * b + 1]),
                  gl11 = static_cast<float>(a.gen_chol[3 * b + 2]);
      const unsigned long long gb = static_cast<unsigned long long>(b + a.gen_index_offset);
      const uint32_t k0 = static_cast<uint32_t>(a.gen_seed), k1 = static_cast<uint32_t>(a.gen_seed >> 32);
      auto gen_pair = [&](uint32_t j) {   // samples 2j, 2j+1 of the halfspace
        const Philox4 r = philox4x32_10(j, static_cast<uint32_t>(gb), static_cast<uint32_t>(gb >> 32), kGenStreamTag, k0, k1);
        const float2 s0 = gen_sample(r.x, r.y, gmx, gmy, gl00, gl10, gl11);
        const float2 s1 = gen_sample(r.z, r.w, gmx, gmy, gl00, gl10, gl11);
        return make_float4(s0.x, s0.y, s1.x, s1.y);
      };
      const uint32_t pair_lo = part_lo >> 4, n_pairs = (part_b + 15u) >> 4;
      float* dump = a.gen_samples_out ? a.gen_samples_out + static_cast<size_t>(b) * N * 2 : nullptr;
      for (uint32_t j = tid; j < n_pairs; j += kClTeam) {
        const float4 p = gen_pair(pair_lo + j);
        *reinterpret_cast<float4*>(smem_raw + 16u * j) = p;
        if (dump) {
          const size_t i0 = 2 * static_cast<size_t>(pair_lo + j);
          dump[2 * i0] = p.x;
          dump[2 * i0 + 1] = p.y;
          if (i0 + 1 < static_cast<size_t>(N)) {
            dump[2 * i0 + 2] = p.z;
            dump[2 * i0 + 3] = p.w;
          }
        }
      }
      if (tid == kClTeam - 1) {   // sample 0 for everybody (every CTA needs the shift origin)
        const float4 p = gen_pair(0);
        sh->gfirst = make_float2(p.x, p.y);
      }
      cl_team_sync();
      first = sh->gfirst;
    } else {
      const float* fp = reinterpret_cast<const float*>(a.samples) + b * a.stride_b;
      first = make_float2(__ldg(fp), __ldg(fp + 1));
    }
    double pre0 = 0.0, pre1 = 0.0;   // warp 0: ego (or the explicit normal), fetched early
    if (warp == 0) {
      if (a.h_in != nullptr) {
        pre0 = a.h_in[2 * b];
        pre1 = a.h_in[2 * b + 1];
      } else if (a.ego != nullptr) {
        pre0 = a.ego[2 * b];
        pre1 = a.ego[2 * b + 1];
      }
    }
    const float2 nf = make_float2(-first.x, -first.y);
    const uint32_t fpar = it & 1;
    int have = 0;   // chunks of halfspace b known to have landed
    auto wait_upto = [&](uint32_t byte_off) {
      if (gen || (CL_DBG(2) && it > 0)) return;
      const int c = static_cast<int>(byte_off >> 15);
      while (have <= c) {
        mbar_wait_spin(&sh->full[have], fpar);
        ++have;
      }
    };

    PH_MARK(0)
    // ------------------------------------------------------------------ sweep A: canonical lane sums per octant + moments
    float2 sq = make_float2(0.f, 0.f);
    float sxy = 0.f;
    for (int k = 0; k < O; ++k) {
      const uint32_t ob = static_cast<uint32_t>(k) * oct_b;
      const uint32_t oe = part_b < ob + oct_b ? part_b : ob + oct_b;
      const uint32_t len = oe > ob ? oe - ob : 0u;
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
      auto body = [&](const float4 v) {
        const float2 d0 = __fadd2_rn(make_float2(v.x, v.y), nf), d1 = __fadd2_rn(make_float2(v.z, v.w), nf);
        acc0 = __fadd2_rn(acc0, d0);
        acc1 = __fadd2_rn(acc1, d1);
        sq = __ffma2_rn(d0, d0, sq);
        sq = __ffma2_rn(d1, d1, sq);
        sxy = fmaf(d0.x, d0.y, sxy);
        sxy = fmaf(d1.x, d1.y, sxy);
      };
      const uint32_t n_lr = (len + kClLaneRow - 1) / kClLaneRow;
      uint32_t lr = 0;
      if (4 * kClLaneRow <= len) wait_upto(ob + toff + 3 * kClLaneRow);
      for (; (lr + 4) * kClLaneRow <= len; lr += 4) {
        const uint32_t base = ob + lr * kClLaneRow + toff;
        const float4 v0 = lds128(slot_s + base), v1 = lds128(slot_s + base + kClLaneRow),
                     v2 = lds128(slot_s + base + 2 * kClLaneRow), v3 = lds128(slot_s + base + 3 * kClLaneRow);
        // the (≈90-cycle) barrier test for the NEXT group overlaps these loads and the arithmetic below
        if ((lr + 8) * kClLaneRow <= len) wait_upto(base + 7 * kClLaneRow);
        body(v0);
        body(v1);
        body(v2);
        body(v3);
      }
      for (; lr < n_lr; ++lr) {   // last rows of the octant: possibly ragged; samples beyond the end count as `first` (d = +0)
        const uint32_t off = ob + lr * kClLaneRow + toff;
        float4 v = make_float4(first.x, first.y, first.x, first.y);
        if (off + 8 <= oe) {
          wait_upto(off);
          v = lds128(slot_s + off);
          if (off + 16 > oe) {
            v.z = first.x;
            v.w = first.y;
          }
        }
        body(v);
      }
      // slot tid = adjacent fp32 lanes widened and added; u[j] = s[j] + s[j + 256]; butterfly per group of 32; pair tree of 8
      const double s_x = __dadd_rn(static_cast<double>(acc0.x), static_cast<double>(acc1.x));
      const double s_y = __dadd_rn(static_cast<double>(acc0.y), static_cast<double>(acc1.y));
      PH_MARK(1)
      double* xch = sh->xch;
      if (tid >= 256) {
        xch[2 * (tid - 256)] = s_x;
        xch[2 * (tid - 256) + 1] = s_y;
      }
      cl_team_sync();
      if (tid < 256) {
        const double txy = warp_sum_canon_pair(__dadd_rn(s_x, xch[2 * tid]), __dadd_rn(s_y, xch[2 * tid + 1]), lane);
        if (lane < 2) sh->red[warp * 2 + lane] = txy;
      }
      cl_team_sync();
      if (tid < 2) sh->octtot[2 * k + tid] = pair_tree8(sh->red + tid, 2);
    }
    {
      const float mq = warp_sum_any4(sq.x, sq.y, sxy, 0.f, lane);   // lanes & 3: 0 qxx, 1 qxy, 2 qyy
      const unsigned bnd = __reduce_max_sync(kFull, __float_as_uint(sq.x + sq.y));
      if (lane < 3) sh->mom[warp * 4 + lane] = mq;
      if (lane == 3) sh->mom[warp * 4 + 3] = __uint_as_float(bnd);
    }
    cl_team_sync();
    PH_MARK(2)
    // ------------------------------------------------------------------ exchange 1: totals + moments to every CTA
    if (warp == 0) {
      // this CTA's moments: lane w < 16 holds sweep warp w's partials, xor-butterfly over the 16 lanes
      float4 mm = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < kClTeamWarps) mm = *reinterpret_cast<const float4*>(&sh->mom[lane * 4]);
      double qxx = static_cast<double>(mm.x), qxy = static_cast<double>(mm.y), qyy = static_cast<double>(mm.z);
      float b2 = mm.w;
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) {
        qxx += shfl_xor_d(qxx, m);
        qxy += shfl_xor_d(qxy, m);
        qyy += shfl_xor_d(qyy, m);
        b2 = fmaxf(b2, __shfl_xor_sync(kFull, b2, m));
      }
      // leader of this halfspace: our finisher must be done with the pool / x2 / fin_ctl of the halfspace we led last,
      // before any CTA can get past exchange 1 and send the next candidates
      if (static_cast<int>(rank) == leader && n_lead > 0) mbar_wait(&sh->fdone, (n_lead - 1) & 1u);
      // learned-window state of our finisher (ONE read, the same bits to every destination).  Only the LEADER's word is
      // used, and the leader has just waited for its finisher: the state is a deterministic function of the halfspaces
      // this CTA led before (it - C, it - 2C, ...), so results do not depend on how far the finishers lag.
      unsigned long long zp = lane == 0 ? *reinterpret_cast<volatile unsigned long long*>(&sh->zpub) : 0ull;
      zp = __shfl_sync(kFull, zp, 0);
      if (lane < C) {   // lane d serves destination CTA d
        const uint32_t dst = mapa_u32(smem_u32(&sh->x1[par][rank][0]), static_cast<uint32_t>(lane));
        const uint32_t bar = mapa_u32(smem_u32(&sh->xbar1[par]), static_cast<uint32_t>(lane));
        mbar_arrive_expect_tx_remote(bar, 8u * (2u * O + 5u));
        for (int i = 0; i < 2 * O; ++i) st_async_f64(dst + 8u * i, sh->octtot[i], bar);
        st_async_f64(dst + 64, qxx, bar);
        st_async_f64(dst + 72, qyy, bar);
        st_async_f64(dst + 80, qxy, bar);
        st_async_f64(dst + 88, static_cast<double>(b2), bar);
        st_async_f64(dst + 96, __longlong_as_double(static_cast<long long>(zp)), bar);
      }
    }
    mbar_wait_cluster(&sh->xbar1[par], (it >> 1) & 1);
    if (warp == 0) {
      if (lane == 0) {
        ctl->f0 = static_cast<double>(first.x);
        ctl->f1 = static_cast<double>(first.y);
      }
      __syncwarp();
      bar_arrive(kClBarDirector + par, 64);   // the director starts the canonical chain
    }
    PH_MARK(3)

    // ------------------------------------------------------------------ window placement (warp 0; identical in every CTA)
    if (warp == 0) {
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[kOctants];
#pragma unroll
        for (int o = 0; o < kOctants; ++o) t[o] = sh->x1[par][o >> lgO][2 * (o & (O - 1)) + j];
#pragma unroll
        for (int n = kOctants; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = t[2 * g] + t[2 * g + 1];
        w[j] = t[0];
      }
      double qd[3] = {0.0, 0.0, 0.0};
      float b2 = 0.f;
      for (int s = 0; s < C; ++s) {
        qd[0] += sh->x1[par][s][8];
        qd[1] += sh->x1[par][s][9];
        qd[2] += sh->x1[par][s][10];
        b2 = fmaxf(b2, static_cast<float>(sh->x1[par][s][11]));
      }
      const double inv_n = 1.0 / static_cast<double>(N);
      const float inv_nf = static_cast<float>(inv_n);
      const double f0 = static_cast<double>(first.x), f1 = static_cast<double>(first.y);
      const double mr0 = w[0] * inv_n, mr1 = w[1] * inv_n;   // mean relative to the first sample
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
        if (n2 < 1.01e-20f) usable = false;   // degenerate direction (or too close to the switch): redo list
        const float rn = rsqrtf(n2);
        h0f = d0f * rn;
        h1f = d1f * rn;
        // fp32 chain: 2 conversions, fma, rsqrt (2 ulp), multiply  ->  < 5e-7; plus the fp64 cancellation in m - ego
        const float mag = static_cast<float>(fabs(m0) + fabs(m1) + fabs(pre0) + fabs(pre1));
        err_h = 1e-6f + 4e-15f * mag * rn;
        usable = usable && isfinite(rn) && rn > 0.f && isfinite(mag);
      }
      usable = usable && isfinite(h0f) && isfinite(h1f) && err_h < 1e-3f;
      const float cxx = static_cast<float>(qd[0]) * inv_nf - mr0f * mr0f, cyy = static_cast<float>(qd[1]) * inv_nf - mr1f * mr1f,
                  cxy = static_cast<float>(qd[2]) * inv_nf - mr0f * mr1f;
      const float var_l = h0f * h0f * cxx + 2.0f * h0f * h1f * cxy + h1f * h1f * cyy;
      const float sigma = sqrt_approx(var_l);   // placement only
      int window_ok = a.use_window && usable && (var_l > 0.f) && isfinite(sigma);
      // thresholds in shifted coordinates, p = h_a.(xi - first):  a_lo <-> t_lo,  a_hi <-> t_hi  (a_hi <= a_lo)
      const float pm = fmaf(h1f, mr1f, h0f * mr0f);
      // Window bounds in z units: the Gaussian plan of the host, or — once the finisher of this halfspace's LEADER has seen
      // two window misses in a row on the halfspaces it led (samples that are evidently not Gaussian) — its LEARNED centre
      // with a wider window.  Every CTA received the leader's word in exchange 1, so all CTAs place the same window, and
      // the word is a deterministic function of earlier halfspaces (one chain per leader, like the resident kernel's
      // parity chains).  Speed only: a wrong window is detected and the halfspace redone.
      const unsigned long long zs = static_cast<unsigned long long>(__double_as_longlong(sh->x1[par][leader][12]));
      const float z_state = __uint_as_float(static_cast<unsigned>(zs));
      const int z_learned = static_cast<int>((zs >> 32) & 1ull);
      float zlo = a.z_lo_f, zhi = a.z_hi_f;
      if (z_learned) {
        zlo = z_state - a.z_half_adapt_f;
        zhi = z_state + a.z_half_adapt_f;
      }
      const float a_lo = pm - zlo * sigma, a_hi = pm - zhi * sigma;
      const double c = static_cast<double>(h0f) * f0 + static_cast<double>(h1f) * f1;   // h_a . first
      const double t_lo = __dadd_rn(-static_cast<double>(a_lo) - c, 0.0);  // +0.0: never -0.0 (canonical losses are +0)
      const double t_hi = __dadd_rn(-static_cast<double>(a_hi) - c, 0.0);
      // rigorous fp32 classification bound: see halfspace_kernel.cuh (window placement); dmax from the per-thread sums
      const float dmax = sqrt_approx(b2) * 1.0001f;
      const float af0 = fabsf(static_cast<float>(f0)) * 1.0001f, af1 = fabsf(static_cast<float>(f1)) * 1.0001f;
      const float habs = fabsf(h0f) + fabsf(h1f);
      const float eps = (habs * (af0 + af1 + dmax)) * 1e-15f + err_h * 1.5f * (af0 + af1 + 2.0f * dmax);
      float thr_keep, thr_above;
      bool raw_ok = true;
      if constexpr (kRawB) {   // thresholds in the space of the raw projection; m = kc / (C x 512) adds per thread (pipelined_kernel.cuh)
        const float vmax = af0 + af1 + dmax;
        const float m_adds = fmaxf(1.0f, static_cast<float>(a.kc) / static_cast<float>(C * kClTeam));
        raw_ok = vmax * m_adds <= 256.0f;
        const float bound = habs * vmax * 2.3841858e-07f + 1.1754944e-38f + eps * 1.0001f;
        const float r_lo = static_cast<float>(static_cast<double>(a_lo) + c), r_hi = static_cast<float>(static_cast<double>(a_hi) + c);
        thr_keep = r_lo + (bound + fabsf(r_lo) * 4.7683716e-07f);
        thr_above = r_hi - (bound + fabsf(r_hi) * 4.7683716e-07f);
      } else {
        const float bound = habs * dmax * 1.9073486e-06f + 1.1754944e-38f + eps * 1.0001f;
        thr_keep = a_lo + (bound + fabsf(a_lo) * 2.3841858e-07f);
        thr_above = a_hi - (bound + fabsf(a_hi) * 2.3841858e-07f);
      }
      const unsigned long long klo = key_of(t_lo), khi = key_of(t_hi);
      window_ok = window_ok && isfinite(thr_keep) && isfinite(thr_above) && (khi >= klo) && (thr_above <= thr_keep) &&
                  isfinite(t_lo) && isfinite(t_hi) && raw_ok;
      if (lane == 0) {
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->h0f = h0f; ctl->h1f = h1f; ctl->thr_keep = thr_keep; ctl->thr_above = thr_above;
        ctl->key_lo = klo;
        ctl->window_ok = window_ok;
        ctl->z_learned = z_learned;
        ctl->z_est = z_state;
        ctl->pl = Ctl::Place{pm, sigma, c};
      }
    }
    cl_team_sync();  // S2
    PH_MARK(4)
    const bool window = ctl->window_ok != 0;

    int rel = 0;   // chunks this warp has handed back (warp-uniform)
    auto release_now = [&](int c) {   // chunks [rel, c) are no longer read by this warp
      if (c > rel) {
        __syncwarp();
        if (lane == 0)
          for (int j = rel; j < c; ++j) mbar_arrive(&sh->free_[j]);
        rel = c;
      }
    };
    auto release_upto = [&](int c) {   // parity mode keeps the whole part resident until the tail indices are written
      if constexpr (!kTail) release_now(c);
    };
    if (!window) {
      // no usable window (degenerate / non-finite / tiny tails): the streaming kernel redoes this halfspace
      release_now(n_chunks);
      mbar_wait(&sh->hdone[par], (it >> 1) & 1);   // the director is done with x1[par] / ctl[par] before we run ahead
      if (lane == 0) {   // the leader's finisher puts the halfspace on the redo list (fin_ctl.window_ok == 0)
        if (tid == 0 && static_cast<int>(rank) == leader) {
          sh->fin_ctl = *ctl;
          mbar_arrive(&sh->xbar2);   // local arrive: release.cta orders the copy
        } else {
          mbar_arrive_expect_tx_remote(mapa_u32(smem_u32(&sh->xbar2), static_cast<uint32_t>(leader)), 0u);
        }
      }
      if (static_cast<int>(rank) == leader) ++n_lead;
      continue;
    }

    // ------------------------------------------------------------------ sweep B: fp32 classification, raw copies of the rest
    float ax = 0.f, ay = 0.f, cf = 0.f;   // "surely above": shifted coordinate sums and count
    int n_list = 0;                       // warp-uniform
    {
      const float h0f = ctl->h0f, h1f = ctl->h1f, thr_keep = ctl->thr_keep, thr_above = ctl->thr_above;
      const unsigned lt_mask = (1u << lane) - 1u;
      auto extract = [&](unsigned mk, uint32_t base, const float4& v0, const float4& v1, const float4& v2, const float4& v3) {
        // mask bit 2u + e <-> sample e of row u (this thread's 16 bytes at base + u * 8 KB).
        const unsigned bal = __ballot_sync(kFull, mk != 0u);
        if (__all_sync(kFull, (mk & (mk - 1u)) == 0u)) {   // usual case: no lane holds more than one masked sample
          if (mk) {
            const unsigned e = static_cast<unsigned>(__ffs(static_cast<int>(mk))) - 1u;
            const int pos = n_list + __popc(bal & lt_mask);
            DRCVAR_ASSERT(pos >= 0 && base + (e >> 1) * kClLaneRow + (e & 1u) * 8u + 8u <= slot_bytes);
            if (pos < kClWarpList)
              wlist[pos] = lds64(slot_s + base + (e >> 1) * kClLaneRow + (e & 1u) * 8u);
          }
          n_list += __popc(bal);
          return;
        }
        // general case: exclusive prefix of the per-lane counts (0..8) by ballot planes
        const int mine = __popc(mk);
        const unsigned p0 = __ballot_sync(kFull, mine & 1), p1 = __ballot_sync(kFull, mine & 2),
                       p2 = __ballot_sync(kFull, mine & 4), p3 = __ballot_sync(kFull, mine & 8);
        int pos = n_list + __popc(p0 & lt_mask) + 2 * __popc(p1 & lt_mask) + 4 * __popc(p2 & lt_mask) + 8 * __popc(p3 & lt_mask);
        n_list += __popc(p0) + 2 * __popc(p1) + 4 * __popc(p2) + 8 * __popc(p3);
        if (mk) {
          const float4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (mk & (1u << (2 * u))) {
              if (pos < kClWarpList) wlist[pos] = make_float2(vv[u].x, vv[u].y);
              ++pos;
            }
            if (mk & (2u << (2 * u))) {
              if (pos < kClWarpList) wlist[pos] = make_float2(vv[u].z, vv[u].w);
              ++pos;
            }
          }
        }
      };
      auto classify4 = [&](const float4& v, unsigned& mk, unsigned bit, bool ok0, bool ok1) {
        const float2 d0 = kRawB ? make_float2(v.x, v.y) : __fadd2_rn(make_float2(v.x, v.y), nf);
        const float2 d1 = kRawB ? make_float2(v.z, v.w) : __fadd2_rn(make_float2(v.z, v.w), nf);
        float p0 = fmaf(h1f, d0.y, h0f * d0.x), p1 = fmaf(h1f, d1.y, h0f * d1.x);
        if (!ok0) p0 = __int_as_float(0x7f800000);
        if (!ok1) p1 = __int_as_float(0x7f800000);
        classify_f32(p0, thr_above, thr_keep, d0.x, d0.y, ax, ay, cf, mk, bit);
        classify_f32(p1, thr_above, thr_keep, d1.x, d1.y, ax, ay, cf, mk, bit + bit);
      };
      for (int k = 0; k < (CL_DBG(4) ? 0 : O); ++k) {
        const uint32_t ob = static_cast<uint32_t>(k) * oct_b;
        const uint32_t oe = part_b < ob + oct_b ? part_b : ob + oct_b;
        const uint32_t len = oe > ob ? oe - ob : 0u;
        const uint32_t n_lr = (len + kClLaneRow - 1) / kClLaneRow;
        uint32_t lr = 0;
        for (; (lr + 4) * kClLaneRow <= len; lr += 4) {
          const uint32_t base = ob + lr * kClLaneRow + toff;
          release_upto(static_cast<int>((ob + lr * kClLaneRow + woff) >> 15));
          const float4 v0 = lds128(slot_s + base), v1 = lds128(slot_s + base + kClLaneRow),
                       v2 = lds128(slot_s + base + 2 * kClLaneRow), v3 = lds128(slot_s + base + 3 * kClLaneRow);
          unsigned mk = 0u;
          classify4(v0, mk, 1u, true, true);
          classify4(v1, mk, 4u, true, true);
          classify4(v2, mk, 16u, true, true);
          classify4(v3, mk, 64u, true, true);
          if (!CL_DBG(1) && __any_sync(kFull, mk != 0u)) extract(mk, base, v0, v1, v2, v3);
        }
        if (lr < n_lr) {   // last (up to 3 whole + 1 ragged) rows of the octant, one at a time
          release_upto(static_cast<int>((ob + lr * kClLaneRow + woff) >> 15));
          const uint32_t base = ob + lr * kClLaneRow + toff;
          const float4 f4 = make_float4(first.x, first.y, first.x, first.y);
          float4 vv[4] = {f4, f4, f4, f4};
          unsigned mk = 0u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (lr + u < n_lr) {   // warp-uniform
              const uint32_t off = base + u * kClLaneRow;
              if (off + 16 <= oe) {           // whole 16 bytes inside the octant (every row but the ragged last one)
                vv[u] = lds128(slot_s + off);
                classify4(vv[u], mk, 1u << (2 * u), true, true);
              } else if (off + 8 <= oe) {     // N odd: only the first sample of the pair exists
                vv[u] = lds128(slot_s + off);
                classify4(vv[u], mk, 1u << (2 * u), true, false);
              }
            }
          }
          if (!CL_DBG(1) && __any_sync(kFull, mk != 0u)) extract(mk, base, vv[0], vv[1], vv[2], vv[3]);
        }
      }
    }
    PH_MARK(5)
    release_upto(n_chunks);   // the whole part is back with the producer: halfspace b+1 streams in behind us

    // ------------------------------------------------------------------ phase 2b: exact loss of the listed samples
    mbar_wait(&sh->hdone[par], (it >> 1) & 1);
    PH_MARK(6)
    const double h0 = ctl->h0, h1 = ctl->h1;
    const bool nonfinite = ctl->nonfinite != 0;
    const double t_lo = ctl->t_lo, t_hi = ctl->t_hi;
    bool overflow = n_list > kClWarpList || nonfinite;
    int nc = 0;   // window candidates of this warp, compacted in place at the head of its list (as doubles)
    if (!overflow) {
      double* wcand = reinterpret_cast<double*>(wlist);
      for (int k0 = 0; k0 < n_list; k0 += 32) {
        const int kk = k0 + lane;
        const bool active = kk < n_list;
        double L = 0.0;
        float2 v = first;
        if (active) {
          v = wlist[kk];
          L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
        }
        const bool up = active && (L > t_hi);
        const bool cd = active && !up && (L >= t_lo);
        if (up) {  // inside the fp32 uncertainty band but exactly above the window: joins the "above" set
          cf += 1.0f;
          ax += kRawB ? v.x : v.x - first.x;
          ay += kRawB ? v.y : v.y - first.y;
        }
        const unsigned bal = __ballot_sync(kFull, cd);   // (all lanes have read their entry: the ballot orders the overwrite)
        if (cd) wcand[nc + __popc(bal & ((1u << lane) - 1u))] = L;
        nc += __popc(bal);
      }
    }
    {
      const int wc = __reduce_add_sync(kFull, static_cast<int>(cf));
      const double sx = warp_sum_any(static_cast<double>(ax)), sy = warp_sum_any(static_cast<double>(ay));
      if (lane == 0) {
        sh->wsum[warp * 4] = static_cast<double>(wc);
        sh->wsum[warp * 4 + 1] = sx;
        sh->wsum[warp * 4 + 2] = sy;
        sh->wcnt[warp * 2] = nc;
        sh->wcnt[warp * 2 + 1] = overflow ? 1 : 0;
      }
    }
    PH_MARK(7)
    cl_team_sync();  // S3
    PH_MARK(8)
    // ------------------------------------------------------------------ exchange 2: everything goes to the leader CTA
    {
      int before = 0, total = 0, ovf = 0;
#pragma unroll
      for (int w = 0; w < kClTeamWarps; ++w) {
        const int c = sh->wcnt[w * 2];
        before += w < warp ? c : 0;
        total += c;
        ovf |= sh->wcnt[w * 2 + 1];
      }
      if (total > cap) ovf = 1;
      const uint32_t bar = mapa_u32(smem_u32(&sh->xbar2), static_cast<uint32_t>(leader));
      if (lane == 0) {
        const uint32_t tx = (ovf ? 0u : 8u * nc) + (tid == 0 ? 32u : 0u);
        if (tid == 0 && static_cast<int>(rank) == leader) {
          // what our finisher needs of ctl[par], which the team reuses; the LOCAL arrive releases the copy (release.cta)
          sh->fin_ctl = *ctl;
          mbar_expect_tx(&sh->xbar2, tx);
        } else {
          mbar_arrive_expect_tx_remote(bar, tx);
        }
      }
      if (!ovf) {
        const double* wcand = reinterpret_cast<const double*>(wlist);
        DRCVAR_ASSERT(before >= 0 && static_cast<int>(rank) * cap + before <= kClPool);
        const uint32_t dst = mapa_u32(smem_u32(&sh->pool[rank * cap + before]), static_cast<uint32_t>(leader));
        for (int j = lane; j < nc; j += 32) st_async_f64(dst + 8u * j, wcand[j], bar);
      }
      if (warp == 0) {   // CTA totals of the "surely above" set: lane w < 16 holds warp w's partials, fixed butterfly
        double n_above = 0.0, sdx = 0.0, sdy = 0.0;
        if (lane < kClTeamWarps) {
          n_above = sh->wsum[lane * 4];
          sdx = sh->wsum[lane * 4 + 1];
          sdy = sh->wsum[lane * 4 + 2];
        }
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1) {
          n_above += shfl_xor_d(n_above, m);
          sdx += shfl_xor_d(sdx, m);
          sdy += shfl_xor_d(sdy, m);
        }
        if (lane == 0) {
          const uint32_t dst = mapa_u32(smem_u32(&sh->x2[rank][0]), static_cast<uint32_t>(leader));
          st_async_f64(dst, n_above, bar);
          st_async_f64(dst + 8, sdx, bar);
          st_async_f64(dst + 16, sdy, bar);
          st_async_f64(dst + 24, ovf ? -1.0 : static_cast<double>(total), bar);
        }
      }
      if (static_cast<int>(rank) == leader) ++n_lead;
    }
    if constexpr (kTail) {
      // ---------------------------------------------------------------- parity mode: tail indices from the resident samples
      // Sweep warp w owns a contiguous byte range of this CTA's part (so the indices it emits are ascending and the warps'
      // ranges follow each other): pass 1 counts its samples above / equal to T, one team barrier, pass 2 writes the
      // indices at (CTA start from the leader) + (what the warps before it emit).  Ties go to the lower index.
      mbar_wait(&sh->tbar, tphase);
      tphase ^= 1u;
      const double T_thr = sh->tmsg[0];
      const bool emit = sh->tmsg[3] != 0.0 && a.tail_idx_out != nullptr;
      if (emit) {
        const long long cta_base = static_cast<long long>(sh->tmsg[1]);
        const int cta_quota = static_cast<int>(sh->tmsg[2]);
        const uint32_t seg = ((part_b + kClTeamWarps - 1) / kClTeamWarps + 511u) & ~511u;   // whole warp-rows of 32 x 16 bytes
        const uint32_t lo = warp * seg < part_b ? warp * seg : part_b;
        const uint32_t hi = lo + seg < part_b ? lo + seg : part_b;
        const unsigned lt_mask = (1u << lane) - 1u;
        auto losses = [&](uint32_t off, bool& ok0, bool& ok1, double& L0, double& L1) {
          ok0 = off + 8u <= hi;
          ok1 = off + 16u <= hi;
          L0 = L1 = 0.0;
          if (ok0) {
            const float4 v = ok1 ? lds128(slot_s + off) : make_float4(lds64(slot_s + off).x, lds64(slot_s + off).y, 0.f, 0.f);
            L0 = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
            if (ok1) L1 = loss_of(h0, h1, static_cast<double>(v.z), static_cast<double>(v.w));
          }
        };
        int n_gt = 0, n_eq = 0;
        for (uint32_t off = lo + 16u * lane; off < hi; off += 512u) {
          bool ok0, ok1;
          double L0, L1;
          losses(off, ok0, ok1, L0, L1);
          n_gt += (ok0 && L0 > T_thr) + (ok1 && L1 > T_thr);
          n_eq += (ok0 && L0 == T_thr) + (ok1 && L1 == T_thr);
        }
        n_gt = __reduce_add_sync(kFull, n_gt);
        n_eq = __reduce_add_sync(kFull, n_eq);
        if (lane == 0) {
          sh->tcnt[warp * 2] = n_gt;
          sh->tcnt[warp * 2 + 1] = n_eq;
        }
        cl_team_sync();
        int gt_before = 0, eq_before = 0;
#pragma unroll
        for (int w = 0; w < kClTeamWarps; ++w) {
          gt_before += w < warp ? sh->tcnt[w * 2] : 0;
          eq_before += w < warp ? sh->tcnt[w * 2 + 1] : 0;
        }
        long long run = cta_base + gt_before + (eq_before < cta_quota ? eq_before : cta_quota);
        int eq_run = eq_before;   // threshold-valued samples of this CTA before the current position
        int* out = a.tail_idx_out + b * static_cast<long long>(a.kc);
        for (uint32_t off0 = lo; off0 < hi; off0 += 512u) {   // warp-uniform trip count
          const uint32_t off = off0 + 16u * lane;
          bool ok0, ok1;
          double L0, L1;
          losses(off, ok0, ok1, L0, L1);
          const bool eq0 = ok0 && L0 == T_thr, eq1 = ok1 && L1 == T_thr;
          const unsigned be0 = __ballot_sync(kFull, eq0), be1 = __ballot_sync(kFull, eq1);
          const int rank0 = eq_run + __popc(be0 & lt_mask) + __popc(be1 & lt_mask), rank1 = rank0 + (eq0 ? 1 : 0);
          const bool sel0 = ok0 && (L0 > T_thr || (eq0 && rank0 < cta_quota));
          const bool sel1 = ok1 && (L1 > T_thr || (eq1 && rank1 < cta_quota));
          const unsigned bs0 = __ballot_sync(kFull, sel0), bs1 = __ballot_sync(kFull, sel1);
          const long long pos0 = run + __popc(bs0 & lt_mask) + __popc(bs1 & lt_mask);
          const int i0 = static_cast<int>((part_lo + off) >> 3);
          DRCVAR_ASSERT(!(sel0 || sel1) || (pos0 >= 0 && pos0 + (sel0 && sel1 ? 1 : 0) < a.kc));
          if (sel0) out[pos0] = i0;
          if (sel1) out[pos0 + (sel0 ? 1 : 0)] = i0 + 1;
          run += __popc(bs0) + __popc(bs1);
          eq_run += __popc(be0) + __popc(be1);
        }
      }
      release_now(n_chunks);   // only now may the next halfspace stream in
    }
    PH_MARK(9)
    // (list / wsum / wcnt are next written behind the team barriers of the next halfspace's octant trees)
    PH_MARK(10)
  }
#ifdef DRCVAR_PROFILE_PHASES
  if (tid == 64 && a.phase_cycles)
    for (int k = 0; k < 12; ++k) a.phase_cycles[(blockIdx.x * 2 + 0) * 12 + k] = ph_t[k];
#endif
  cluster_sync_all();   // nobody leaves while a peer may still write into its shared memory
}

}  // namespace drcvar
