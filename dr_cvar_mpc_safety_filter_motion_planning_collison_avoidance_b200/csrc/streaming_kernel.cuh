// streaming_kernel.cuh — kernel for sample counts that do not fit one CTA's shared memory
// (N > drcvar_max_samples(): e.g. BASELINE config 5, N = 100 000).
//
// Same arithmetic contract and outputs as halfspace_kernel, but the samples stay in global memory.  One 512-thread
// CTA per halfspace, grid-stride over the batch:
//   pass 1  canonical lane sums + second moments            (read 1: HBM)
//   warp 0  canonical mean / direction, mean halfspace, statistical window [t_lo, t_hi] (placement only)
//   pass 2  exact canonical fp64 loss of every sample: above the window -> count + sum, inside -> per-warp
//           candidate lists in shared memory                (read 2: HBM or L2)
//   select  exact rank among the candidates (range-narrowing radix select in shared memory), CVaR, offsets
// If the window misses (or GENERAL_ONLY): the exact multi-pass radix select over all samples (2-4 more reads).
// fp32 samples with 32768 < N <~ 205000 normally run the single-read cluster / DSMEM kernel (cluster_kernel.cuh); this kernel
// serves everything else — fp64, tail indices, strided or unaligned views, 24000 < N <= 32768 — and the halfspaces the
// cluster kernel hands back (flags `redo_list[b]`).  After two consecutive window misses a CTA learns the window centre.
#pragma once

#include "halfspace_kernel.cuh"
#undef DRCVAR_FILE_ID
#define DRCVAR_FILE_ID 3

namespace drcvar {

constexpr int kStreamThreads = 512;
constexpr int kStreamWarps = kStreamThreads / 32;
constexpr int kStreamCand = 256;    // window candidates per warp (doubles)
constexpr int kStreamRedoCap = 512;   // flagged halfspaces a CTA lists up front in a redo pass
constexpr int kStreamUnroll = 8;    // 16-byte loads in flight per thread in pass 2

template <typename T, bool kTail, bool kGen = false>
__global__ void __launch_bounds__(kStreamThreads, 1) streaming_kernel(const KernelArgs a) {
  using V2 = typename Vec2<T>::type;
  constexpr bool kF32 = sizeof(T) == 4;
  constexpr int kPerLoad = kF32 ? 2 : 1;                  // samples per 16-byte load
  constexpr int kRowSamples = kSweepThreads * kPerLoad;   // canonical row: 256 loads (the resident kernel's lane map)
  constexpr int kW = kStreamWarps;
  __shared__ unsigned hist[kHistBuckets];
  __shared__ double small[kResolveMax];
  __shared__ double red[kW * 8];
  __shared__ float mom[kW * 8];
  __shared__ double xch[kSweepThreads * 2];
  __shared__ double cand[kW * kStreamCand];
  __shared__ int iscr[4 * kW];
  __shared__ double oct_tot[2 * kOctants];
  __shared__ Ctl ctl_s;
  Ctl* ctl = &ctl_s;
  // Window centre learned from this CTA's earlier halfspaces (as in the resident kernel): after two CONSECUTIVE window
  // misses — samples that are evidently not Gaussian — the exact threshold of the general select, in sigma units around
  // the loss mean, becomes the centre of a 2.5 x wider window and is tracked from then on.  Isolated misses (3e-5 of
  // Gaussian halfspaces) change nothing.  Speed only: T, the tail set and the status of the others never depend on it.
  if (threadIdx.x == 0) {
    ctl_s.z_learned = 0;
    ctl_s.z_missrun = 0;
    ctl_s.z_est = 0.f;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = tid >> 8, lt = tid & (kSweepThreads - 1);   // half 0: even canonical rows, half 1: odd rows
  const int N = a.N;
  auto sync = [] { __syncthreads(); };

  // work: the whole batch, or — redo pass of the cluster / register-file kernels — the halfspaces whose redo flag is set.
  // The assignment of halfspaces to CTAs stays the static one (b = blockIdx.x + i gridDim.x, increasing i); a redo pass
  // first gathers the flagged i of this CTA with all threads (a dependent flag load per halfspace would cost more than
  // the few flagged halfspaces themselves at B = 655 360) and sorts the short list.
  __shared__ int redo_i[kStreamRedoCap];
  __shared__ int redo_n;
  const long long n_mine = blockIdx.x < a.B ? (a.B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  long long n_work = n_mine;
  bool listed = false;
  if (a.redo_list != nullptr) {
    if (threadIdx.x == 0) redo_n = 0;
    __syncthreads();
    for (long long i = threadIdx.x; i < n_mine; i += kStreamThreads)
      if (a.redo_list[blockIdx.x + i * gridDim.x] != 0) {
        const int pos = atomicAdd(&redo_n, 1);
        if (pos < kStreamRedoCap) redo_i[pos] = static_cast<int>(i);
      }
    __syncthreads();
    if (redo_n <= kStreamRedoCap) {   // (more than that: nearly everything is flagged, the plain scan below costs nothing extra)
      listed = true;
      n_work = redo_n;
      if (threadIdx.x == 0)   // insertion sort: the list is short, its order must not depend on the atomics
        for (int i = 1; i < redo_n; ++i) {
          const int key = redo_i[i];
          int j = i - 1;
          for (; j >= 0 && redo_i[j] > key; --j) redo_i[j + 1] = redo_i[j];
          redo_i[j + 1] = key;
        }
      __syncthreads();
    }
  }
  for (long long wi = 0; wi < n_work; ++wi) {
    const long long b = blockIdx.x + (listed ? static_cast<long long>(redo_i[wi]) : wi) * gridDim.x;
    if (!listed && a.redo_list != nullptr && a.redo_list[b] == 0) continue;
    const T* base = reinterpret_cast<const T*>(a.samples) + b * a.stride_b;
    const bool vec = kGen || a.bulk != 0;   // contiguous (x, y) pairs, 16-byte aligned rows: vector loads
    // generate mode (fp32): the samples are drawn on the fly in EVERY pass (sample_gen.cuh; one Philox call per pair)
    float gmx = 0.f, gmy = 0.f, gl00 = 0.f, gl10 = 0.f, gl11 = 0.f;
    unsigned long long gb = 0;
    if constexpr (kGen) {
      gmx = static_cast<float>(a.gen_mean[2 * b]);
      gmy = static_cast<float>(a.gen_mean[2 * b + 1]);
      gl00 = static_cast<float>(a.gen_chol[3 * b]);
      gl10 = static_cast<float>(a.gen_chol[3 * b + 1]);
      gl11 = static_cast<float>(a.gen_chol[3 * b + 2]);
      gb = static_cast<unsigned long long>(b + a.gen_index_offset);
    }
    auto gen_pair = [&](int j) {   // samples 2j, 2j+1
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(j), static_cast<uint32_t>(gb), static_cast<uint32_t>(gb >> 32),
                                      kGenStreamTag, static_cast<uint32_t>(a.gen_seed),
                                      static_cast<uint32_t>(a.gen_seed >> 32));
      const float2 s0 = gen_sample(r.x, r.y, gmx, gmy, gl00, gl10, gl11);
      const float2 s1 = gen_sample(r.z, r.w, gmx, gmy, gl00, gl10, gl11);
      return make_float4(s0.x, s0.y, s1.x, s1.y);
    };
    auto load4 = [&](int j) {      // 16-byte item j (vec path only)
      if constexpr (kGen) {
        return gen_pair(j);
      } else {
        return reinterpret_cast<const float4*>(base)[j];
      }
    };
    auto load = [&](int i) {
      if constexpr (kGen) {
        const float4 p = gen_pair(i >> 1);
        V2 v;
        v.x = (i & 1) ? p.z : p.x;
        v.y = (i & 1) ? p.w : p.y;
        return v;
      } else {
        if (vec) return reinterpret_cast<const V2*>(base)[i];
        const T* p = base + static_cast<long long>(i) * a.stride_n;
        V2 v;
        v.x = p[0];
        v.y = p[a.stride_c];
        return v;
      }
    };
    if constexpr (kGen) {
      if (a.gen_samples_out) {   // parity tests: dump the generated samples (extra pass)
        float* dump = a.gen_samples_out + static_cast<size_t>(b) * N * 2;
        for (int j = tid; j < (N + 1) / 2; j += kStreamThreads) {
          const float4 p = gen_pair(j);
          dump[4 * j] = p.x;
          dump[4 * j + 1] = p.y;
          if (2 * j + 1 < N) {
            dump[4 * j + 2] = p.z;
            dump[4 * j + 3] = p.w;
          }
        }
      }
    }
    // ------------------------------------------------------------------ pass 1: canonical lane sums + moments
    // Thread (half, lt) owns slot half*256 + lt: rows r = half, half+2, ... in increasing order (the canonical order).
    // N > kOctantMinN: the contract cuts the samples into 8 octants of whole 4 KB rows, each with its own slot sums and
    // tree (oracle/closed_form.py) — the split the cluster kernel distributes over its CTAs.
    const V2 first = load(0);
    const int n_oct = N > kOctantMinN ? kOctants : 1;
    const int oct_len = n_oct == 1 ? N : octant_samples(N, sizeof(T));
    float q_xx = 0.f, q_yy = 0.f, q_xy = 0.f, q_dx = 0.f, q_dy = 0.f;
    int n_sub_i = 0;   // fp64 inputs: samples behind the second moments (rows 0, 4, 8, ... of every octant)
    double sx[kOctants], sy[kOctants];   // this thread's slot sums, per octant
#pragma unroll
    for (int oc = 0; oc < kOctants; ++oc) sx[oc] = sy[oc] = 0.0;
    if (n_oct == 1) {
      const int full_rows = N / kRowSamples;
      const int rows_all = (N + kRowSamples - 1) / kRowSamples;
      if constexpr (kF32) {
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f), sq = make_float2(0.f, 0.f);
        float sxy = 0.f;
        const float2 nf = make_float2(-first.x, -first.y);
        auto body = [&](float2 v, float2& acc) {
          const float2 d = __fadd2_rn(v, nf);
          acc = __fadd2_rn(acc, d);
          sq = __ffma2_rn(d, d, sq);
          sxy = fmaf(d.x, d.y, sxy);
        };
        int r = half;
        if (vec) {
#pragma unroll 8
          for (; r < full_rows; r += 2) {
            const float4 v = load4(r * kSweepThreads + lt);
            body(make_float2(v.x, v.y), acc0);
            body(make_float2(v.z, v.w), acc1);
          }
        } else {
          for (; r < full_rows; r += 2) {
            const int i0 = r * kRowSamples + 2 * lt;
            body(load(i0), acc0);
            body(load(i0 + 1), acc1);
          }
        }
        if (r == full_rows && full_rows < rows_all) {   // ragged last row (only the half with its parity)
          const int i0 = r * kRowSamples + 2 * lt;
          if (i0 < N) body(load(i0), acc0);
          if (i0 + 1 < N) body(load(i0 + 1), acc1);
        }
        sx[0] = __dadd_rn(static_cast<double>(acc0.x), static_cast<double>(acc1.x));   // lanes 2 slot, 2 slot + 1
        sy[0] = __dadd_rn(static_cast<double>(acc0.y), static_cast<double>(acc1.y));
        q_xx = sq.x;
        q_yy = sq.y;
        q_xy = sxy;
      } else {
        double s_x = 0.0, s_y = 0.0;
        double dxx = 0.0, dyy = 0.0, dxy = 0.0, ddx = 0.0, ddy = 0.0;
#pragma unroll 8
        for (int r = half; r < rows_all; r += 2) {
          const int i = r * kRowSamples + lt;
          if (i < N) {
            const V2 v = load(i);
            s_x = __dadd_rn(s_x, v.x);
            s_y = __dadd_rn(s_y, v.y);
            if ((r & 3) == 0) {   // second moments on every 4th row (as in the resident kernel: n_sigma on the host)
              const double dx = v.x - first.x, dy = v.y - first.y;
              ddx += dx;
              ddy += dy;
              dxx = fma(dx, dx, dxx);
              dyy = fma(dy, dy, dyy);
              dxy = fma(dx, dy, dxy);
            }
          }
        }
        sx[0] = s_x;
        sy[0] = s_y;
        q_xx = static_cast<float>(dxx);
        q_yy = static_cast<float>(dyy);
        q_xy = static_cast<float>(dxy);
        q_dx = static_cast<float>(ddx);
        q_dy = static_cast<float>(ddy);
        const int r4 = (rows_all + 3) / 4;
        const int last = (r4 - 1) * 4 * kRowSamples;
        n_sub_i = (r4 - 1) * kSweepThreads + (N - last < kSweepThreads ? N - last : kSweepThreads);
      }
    } else {
      // the 8 octants side by side: row r of every octant in one iteration (8 independent 16-byte loads in flight per
      // thread), each octant with its own accumulators; within an octant the rows are still added in increasing order
      const int rows_oct = oct_len / kRowSamples;   // octants are whole 4 KB rows
      if constexpr (kF32) {
        float2 a0[kOctants], a1[kOctants];
#pragma unroll
        for (int oc = 0; oc < kOctants; ++oc) a0[oc] = a1[oc] = make_float2(0.f, 0.f);
        float2 sq = make_float2(0.f, 0.f);
        float sxy = 0.f;
        const float2 nf = make_float2(-first.x, -first.y);
        auto body = [&](float2 v, float2& acc) {
          const float2 d = __fadd2_rn(v, nf);
          acc = __fadd2_rn(acc, d);
          sq = __ffma2_rn(d, d, sq);
          sxy = fmaf(d.x, d.y, sxy);
        };
        for (int r = half; r < rows_oct; r += 2) {
          float4 v[kOctants];
#pragma unroll
          for (int oc = 0; oc < kOctants; ++oc) {
            const int i0 = oc * oct_len + r * kRowSamples + 2 * lt;
            v[oc] = make_float4(first.x, first.y, first.x, first.y);   // samples beyond N count as `first` (shifted value +0)
            if (i0 + 1 < N) {
              if (vec) {
                v[oc] = load4(i0 >> 1);
              } else {
                const V2 p0 = load(i0), p1 = load(i0 + 1);
                v[oc] = make_float4(p0.x, p0.y, p1.x, p1.y);
              }
            } else if (i0 < N) {
              const V2 p0 = load(i0);
              v[oc].x = p0.x;
              v[oc].y = p0.y;
            }
          }
#pragma unroll
          for (int oc = 0; oc < kOctants; ++oc) {
            body(make_float2(v[oc].x, v[oc].y), a0[oc]);
            body(make_float2(v[oc].z, v[oc].w), a1[oc]);
          }
        }
#pragma unroll
        for (int oc = 0; oc < kOctants; ++oc) {
          sx[oc] = __dadd_rn(static_cast<double>(a0[oc].x), static_cast<double>(a1[oc].x));
          sy[oc] = __dadd_rn(static_cast<double>(a0[oc].y), static_cast<double>(a1[oc].y));
        }
        q_xx = sq.x;
        q_yy = sq.y;
        q_xy = sxy;
      } else {
        double dxx = 0.0, dyy = 0.0, dxy = 0.0, ddx = 0.0, ddy = 0.0;
        for (int r = half; r < rows_oct; r += 2) {
          V2 v[kOctants];
          bool ok[kOctants];
#pragma unroll
          for (int oc = 0; oc < kOctants; ++oc) {
            const int i = oc * oct_len + r * kRowSamples + lt;
            ok[oc] = i < N;
            v[oc] = first;
            if (ok[oc]) v[oc] = load(i);
          }
#pragma unroll
          for (int oc = 0; oc < kOctants; ++oc) {
            if (ok[oc]) {
              sx[oc] = __dadd_rn(sx[oc], v[oc].x);
              sy[oc] = __dadd_rn(sy[oc], v[oc].y);
              if ((r & 3) == 0) {   // second moments on every 4th row of every octant
                const double dx = v[oc].x - first.x, dy = v[oc].y - first.y;
                ddx += dx;
                ddy += dy;
                dxx = fma(dx, dx, dxx);
                dyy = fma(dy, dy, dyy);
                dxy = fma(dx, dy, dxy);
              }
            }
          }
        }
        q_xx = static_cast<float>(dxx);
        q_yy = static_cast<float>(dyy);
        q_xy = static_cast<float>(dxy);
        q_dx = static_cast<float>(ddx);
        q_dy = static_cast<float>(ddy);
        for (int oc = 0; oc < kOctants; ++oc) {
          const int n_k = N - oc * oct_len < oct_len ? (N - oc * oct_len > 0 ? N - oc * oct_len : 0) : oct_len;
          if (n_k > 0) {
            const int rows_k = (n_k + kRowSamples - 1) / kRowSamples;
            const int r4 = (rows_k + 3) / 4;
            const int last = (r4 - 1) * 4 * kRowSamples;
            n_sub_i += (r4 - 1) * kSweepThreads + (n_k - last < kSweepThreads ? n_k - last : kSweepThreads);
          }
        }
      }
    }
    // per octant: u[j] = s[j] + s[j + 256], canonical butterfly inside each group of 32, adjacent-pair tree over the 8 groups
#pragma unroll
    for (int oc = 0; oc < kOctants; ++oc) {
      if (oc < n_oct) {
        if (half == 1) {
          xch[2 * lt] = sx[oc];
          xch[2 * lt + 1] = sy[oc];
        }
        __syncthreads();
        if (half == 0) {
          const double tx = warp_sum_canon(__dadd_rn(sx[oc], xch[2 * lt]));
          const double ty = warp_sum_canon(__dadd_rn(sy[oc], xch[2 * lt + 1]));
          if (lane == 0) {
            red[warp * 8] = tx;
            red[warp * 8 + 1] = ty;
          }
        }
        __syncthreads();
        if (tid < 2) {
          double t[kSweepWarps];
#pragma unroll
          for (int g = 0; g < kSweepWarps; ++g) t[g] = red[g * 8 + tid];
#pragma unroll
          for (int n = kSweepWarps; n > 1; n >>= 1)
#pragma unroll
            for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);
          oct_tot[2 * oc + tid] = t[0];
        }
      }
    }
    {
      const float mxx = warp_sum_any(q_xx), myy = warp_sum_any(q_yy), mxy = warp_sum_any(q_xy);
      const float mdx = warp_sum_any(q_dx), mdy = warp_sum_any(q_dy);
      if (lane == 0) {
        float* w = mom + warp * 8;
        w[0] = mxx; w[1] = myy; w[2] = mxy; w[3] = mdx; w[4] = mdy;
      }
    }
    __syncthreads();
    // ------------------------------------------------------------------ warp 0: canonical mean / direction, window
    if (warp == 0) {
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (n_oct == 1) {
          w[j] = oct_tot[j];
        } else {   // adjacent-pair tree over the 8 octant totals
          double t[kOctants];
#pragma unroll
          for (int g = 0; g < kOctants; ++g) t[g] = oct_tot[2 * g + j];
#pragma unroll
          for (int n = kOctants; n > 1; n >>= 1)
#pragma unroll
            for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);
          w[j] = t[0];
        }
      }
      const double f0 = static_cast<double>(first.x), f1 = static_cast<double>(first.y);
      double m0 = __ddiv_rn(w[0], static_cast<double>(N)), m1 = __ddiv_rn(w[1], static_cast<double>(N));
      if constexpr (kF32) {
        m0 = __dadd_rn(f0, m0);
        m1 = __dadd_rn(f1, m1);
      }
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
      // window placement (heuristic: affects speed only): loss mean and sigma from the sample moments
      double q[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int g = 0; g < kW; ++g)
#pragma unroll
        for (int j = 0; j < 5; ++j) q[j] += static_cast<double>(mom[g * 8 + j]);
      double n_sub, ex, ey;
      if (kF32) {
        n_sub = static_cast<double>(N);
        ex = m0 - f0;
        ey = m1 - f1;
      } else {
        n_sub = static_cast<double>(n_sub_i > 0 ? n_sub_i : 1);
        ex = q[3] / n_sub;
        ey = q[4] / n_sub;
      }
      const double cxx = q[0] / n_sub - ex * ex, cyy = q[1] / n_sub - ey * ey, cxy = q[2] / n_sub - ex * ey;
      const double var_l = h0 * h0 * cxx + 2.0 * h0 * h1 * cxy + h1 * h1 * cyy;
      const double sigma = sqrt(var_l);
      const double mu_l = -(h0 * m0 + h1 * m1);
      double z_lo = a.z_lo, z_hi = a.z_hi;
      if (ctl->z_learned) {   // written two barriers ago by thread 0: CTA-uniform
        z_lo = static_cast<double>(ctl->z_est) - static_cast<double>(a.z_half_adapt_f);
        z_hi = static_cast<double>(ctl->z_est) + static_cast<double>(a.z_half_adapt_f);
      }
      const double t_lo = __dadd_rn(mu_l + z_lo * sigma, 0.0);   // +0.0: never -0.0 (canonical losses are +0)
      const double t_hi = __dadd_rn(mu_l + z_hi * sigma, 0.0);
      const int window_ok = a.use_window && !nonfinite && n_sub >= 256.0 && var_l > 0.0 && isfinite(t_lo) &&
                            isfinite(t_hi) && t_lo <= t_hi;
      if (lane == 0) {
        ctl->h0 = h0; ctl->h1 = h1; ctl->m0 = m0; ctl->m1 = m1;
        ctl->nonfinite = nonfinite;
        ctl->degenerate = degenerate;
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->window_ok = window_ok;
        ctl->pl = Ctl::Place{static_cast<float>(mu_l), static_cast<float>(sigma), 0.0};
        write_mean_outputs(a, b, m0, m1);
      }
    }
    __syncthreads();
    const double h0 = ctl->h0, h1 = ctl->h1;
    const bool nonfinite = ctl->nonfinite != 0;
    int status = (nonfinite ? kStatusNonfinite : 0) | (ctl->degenerate ? kStatusDegenerate : 0);
    auto loss_at = [&](int i) {
      const V2 v = load(i);
      return loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
    };

    double T_thr = 0.0;
    int c_tot = 0;
    double s_tot = 0.0;
    bool fast = false;
    if (ctl->window_ok && !nonfinite) {
      // ---------------------------------------------------------------- pass 2: exact classification against the window
      const double t_lo = ctl->t_lo, t_hi = ctl->t_hi;
      double* wcand = cand + warp * kStreamCand;
      int c_gt = 0, nc = 0;
      double s_gt = 0.0;
      auto classify = [&](bool valid, double L) {
        const bool up = valid && (L > t_hi);
        const bool cd = valid && !up && (L >= t_lo);
        if (up) {
          ++c_gt;
          s_gt += L;
        }
        const unsigned bal = __ballot_sync(kFull, cd);
        if (bal) {
          const int pos = nc + __popc(bal & ((1u << lane) - 1u));
          DRCVAR_ASSERT(pos >= 0);
          if (cd && pos < kStreamCand) wcand[pos] = L;
          nc += __popc(bal);
        }
      };
      const int n_items = kF32 ? (N + 1) / 2 : N;   // 16-byte items: sample pairs (fp32) / samples (fp64)
      for (int it0 = 0; it0 < n_items; it0 += kStreamThreads * kStreamUnroll) {
        V2 v[kStreamUnroll][kPerLoad];
        bool ok[kStreamUnroll][kPerLoad];
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u) {
          const int it = it0 + u * kStreamThreads + tid;
#pragma unroll
          for (int e = 0; e < kPerLoad; ++e) {
            ok[u][e] = false;
            v[u][e] = first;
          }
          if (it < n_items) {
            if constexpr (kF32) {
              const int i0 = 2 * it;
              if (vec && i0 + 1 < N) {
                const float4 p = kGen ? gen_pair(it) : reinterpret_cast<const float4*>(base)[it];
                v[u][0] = make_float2(p.x, p.y);
                v[u][1] = make_float2(p.z, p.w);
                ok[u][0] = ok[u][1] = true;
              } else {
                v[u][0] = load(i0);
                ok[u][0] = true;
                if (i0 + 1 < N) {
                  v[u][1] = load(i0 + 1);
                  ok[u][1] = true;
                }
              }
            } else {
              v[u][0] = load(it);
              ok[u][0] = true;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kStreamUnroll; ++u)
#pragma unroll
          for (int e = 0; e < kPerLoad; ++e)
            classify(ok[u][e], loss_of(h0, h1, static_cast<double>(v[u][e].x), static_cast<double>(v[u][e].y)));
      }
      const int wc = __reduce_add_sync(kFull, c_gt);
      const double ws = warp_sum_any(s_gt);
      if (lane == 0) {
        iscr[warp] = wc;
        iscr[kW + warp] = nc;
        red[warp] = ws;
      }
      __syncthreads();
      int cnt_hi = 0, ncand = 0, ovf = 0;
      double s_hi = 0.0;
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        cnt_hi += iscr[w];
        ncand += iscr[kW + w];
        ovf |= iscr[kW + w] > kStreamCand;
        s_hi += red[w];
      }
      fast = !ovf && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
      if (fast) {
        const int my_nc = iscr[kW + warp];
        auto each = [&](auto&& f) {
          for (int j = lane; j < my_nc; j += 32) f(wcand[j]);
        };
        T_thr = select_rank(each, sync, warp == 0, tid, kStreamThreads, key_of(t_lo), key_of(t_hi), a.kc - cnt_hi, hist,
                            small, ctl);
        int c4 = 0;
        double s4 = 0.0;
        each([&](double L) {
          if (L > T_thr) {
            ++c4;
            s4 += L;
          }
        });
        const int wc4 = __reduce_add_sync(kFull, c4);
        const double ws4 = warp_sum_any(s4);
        __syncthreads();   // iscr / red were read above by everybody
        if (lane == 0) {
          iscr[2 * kW + warp] = wc4;
          red[kW + warp] = ws4;
        }
        __syncthreads();
        c_tot = cnt_hi;
        double s_c = 0.0;
#pragma unroll
        for (int w = 0; w < kW; ++w) {
          c_tot += iscr[2 * kW + w];
          s_c += red[kW + w];
        }
        s_tot = s_hi + s_c;
      }
      __syncthreads();
    }

    if (!fast) {
      // ------------------------------------------------------------------ general path: exact multi-pass radix select
      status |= kStatusGeneral;
      int c_gt = 0;
      double s_gt = 0.0;
      if (!nonfinite) {
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int i = tid; i < N; i += kStreamThreads) {
          const unsigned long long k = key_of(loss_at(i));
          kmin = k < kmin ? k : kmin;
          kmax = k > kmax ? k : kmax;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
          const unsigned long long o1 = __shfl_xor_sync(kFull, kmin, m), o2 = __shfl_xor_sync(kFull, kmax, m);
          kmin = o1 < kmin ? o1 : kmin;
          kmax = o2 > kmax ? o2 : kmax;
        }
        unsigned long long* kred = reinterpret_cast<unsigned long long*>(xch);
        __syncthreads();
        if (lane == 0) {
          kred[warp * 2] = kmin;
          kred[warp * 2 + 1] = kmax;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kW; ++w) {
          kmin = kred[w * 2] < kmin ? kred[w * 2] : kmin;
          kmax = kred[w * 2 + 1] > kmax ? kred[w * 2 + 1] : kmax;
        }
        T_thr = select_rank(
            [&](auto&& f) {
              for (int i = tid; i < N; i += kStreamThreads) f(loss_at(i));
            },
            sync, warp == 0, tid, kStreamThreads, kmin, kmax, a.kc, hist, small, ctl);
        for (int i = tid; i < N; i += kStreamThreads) {
          const double L = loss_at(i);
          if (L > T_thr) {
            ++c_gt;
            s_gt += L;
          }
        }
      }
      const int wc = __reduce_add_sync(kFull, c_gt);
      const double ws = warp_sum_any(s_gt);
      __syncthreads();
      if (lane == 0) {
        iscr[warp] = wc;
        red[warp] = ws;
      }
      __syncthreads();
      c_tot = 0;
      s_tot = 0.0;
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        c_tot += iscr[w];
        s_tot += red[w];
      }
    }
    if (tid == 0) {
      write_risk_outputs(a, b, ctl, nonfinite, s_tot, c_tot, T_thr, status);
      if (ctl->window_ok && !nonfinite) {   // where the threshold really sits, in sigma units around the loss mean
        const float zT = (static_cast<float>(T_thr) - ctl->pl.pm) / ctl->pl.sigma;
        if (fast) {
          ctl->z_missrun = 0;
          if (ctl->z_learned && isfinite(zT)) ctl->z_est = 0.5f * (ctl->z_est + zT);
        } else if (isfinite(zT)) {
          ctl->z_est = zT;
          if (++ctl->z_missrun >= 2) ctl->z_learned = 1;
        }
      }
    }

    if (kTail && a.tail_idx_out != nullptr) {
      int* out = a.tail_idx_out + b * static_cast<long long>(a.kc);
      if (nonfinite) {
        for (int i = tid; i < a.kc; i += kStreamThreads) out[i] = -1;
      } else {
        const int need = a.kc - c_tot;
        int run_eq = 0, run_out = 0;
        int* weq = iscr + kW;
        int* wsel = iscr + 2 * kW;
        __syncthreads();
        for (int bb = 0; bb < N; bb += kStreamThreads) {
          const int i = bb + tid;
          const bool valid = i < N;
          const double L = valid ? loss_at(i) : 0.0;
          const bool gt = valid && (L > T_thr), eq = valid && (L == T_thr);
          const unsigned meq = __ballot_sync(kFull, eq);
          if (lane == 0) weq[warp] = __popc(meq);
          __syncthreads();
          int eq_before = run_eq, tile_eq = 0;
#pragma unroll
          for (int w = 0; w < kW; ++w) {
            if (w < warp) eq_before += weq[w];
            tile_eq += weq[w];
          }
          const int eq_rank = eq_before + __popc(meq & ((1u << lane) - 1u));
          const bool sel = gt || (eq && eq_rank < need);
          const unsigned msel = __ballot_sync(kFull, sel);
          if (lane == 0) wsel[warp] = __popc(msel);
          __syncthreads();
          int out_before = run_out, tile_sel = 0;
#pragma unroll
          for (int w = 0; w < kW; ++w) {
            if (w < warp) out_before += wsel[w];
            tile_sel += wsel[w];
          }
          if (sel) out[out_before + __popc(msel & ((1u << lane) - 1u))] = i;
          run_eq += tile_eq;
          run_out += tile_sel;
          __syncthreads();
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace drcvar
