// streaming_kernel.cuh — fallback kernel for sample counts that do not fit one CTA's shared memory
// (N > drcvar_max_samples(): e.g. BASELINE config 5, N = 100 000).
//
// Same arithmetic contract and outputs as halfspace_kernel, but the samples stay in global memory and are streamed
// several times (canonical sums, key range, 2-4 histogram passes of the exact radix select, final sum): correct for any
// N, ~6 reads of the samples instead of one.  One 256-thread CTA per halfspace, grid-stride over the batch.
// The single-read cluster / DSMEM kernel for large N is the planned replacement.
#pragma once

#include "halfspace_kernel.cuh"

namespace drcvar {

constexpr int kStreamThreads = 256;

template <typename T, bool kTail>
__global__ void __launch_bounds__(kStreamThreads) streaming_kernel(const KernelArgs a) {
  using V2 = typename Vec2<T>::type;
  constexpr bool kF32 = sizeof(T) == 4;
  constexpr int kPerLoad = kF32 ? 2 : 1;
  constexpr int kRowSamples = kStreamThreads * kPerLoad;
  constexpr int kW = kStreamThreads / 32;
  __shared__ unsigned hist[kHistBuckets];
  __shared__ double small[kResolveMax];
  __shared__ double red[kW * 8];
  __shared__ int iscr[4 * kW];
  __shared__ Ctl ctl_s;
  Ctl* ctl = &ctl_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const int rows_all = (N + kRowSamples - 1) / kRowSamples;
  auto sync = [] { __syncthreads(); };

  for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
    const T* base = reinterpret_cast<const T*>(a.samples) + b * a.stride_b;
    const bool vec = a.bulk != 0;   // contiguous (x, y) pairs aligned to sizeof(V2): one vector load per sample
    auto load = [&](int i) {
      if (vec) return reinterpret_cast<const V2*>(base)[i];
      const T* p = base + static_cast<long long>(i) * a.stride_n;
      V2 v;
      v.x = p[0];
      v.y = p[a.stride_c];
      return v;
    };
    // ------------------------------------------------------------------ canonical lane sums (same lane map as the resident kernel)
    const V2 first = load(0);
    double u_x, u_y;
    if constexpr (kF32) {
      float2 acc[2][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) acc[q][0] = acc[q][1] = make_float2(0.f, 0.f);
      const float2 nf = make_float2(-first.x, -first.y);
      for (int r = 0; r < rows_all; ++r) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = r * kRowSamples + 2 * tid + e;
          if (i < N) {
            const float2 d = __fadd2_rn(load(i), nf);
            if ((r & 1) == 0) acc[0][e] = __fadd2_rn(acc[0][e], d); else acc[1][e] = __fadd2_rn(acc[1][e], d);
          }
        }
      }
      const double s0x = __dadd_rn(static_cast<double>(acc[0][0].x), static_cast<double>(acc[0][1].x));
      const double s0y = __dadd_rn(static_cast<double>(acc[0][0].y), static_cast<double>(acc[0][1].y));
      const double s1x = __dadd_rn(static_cast<double>(acc[1][0].x), static_cast<double>(acc[1][1].x));
      const double s1y = __dadd_rn(static_cast<double>(acc[1][0].y), static_cast<double>(acc[1][1].y));
      u_x = __dadd_rn(s0x, s1x);
      u_y = __dadd_rn(s0y, s1y);
    } else {
      double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
      for (int r = 0; r < rows_all; ++r) {
        const int i = r * kRowSamples + tid;
        if (i < N) {
          const V2 v = load(i);
          if ((r & 1) == 0) {
            s00 = __dadd_rn(s00, v.x);
            s01 = __dadd_rn(s01, v.y);
          } else {
            s10 = __dadd_rn(s10, v.x);
            s11 = __dadd_rn(s11, v.y);
          }
        }
      }
      u_x = __dadd_rn(s00, s10);
      u_y = __dadd_rn(s01, s11);
    }
    {
      const double tx = warp_sum_canon(u_x), ty = warp_sum_canon(u_y);
      if (lane == 0) {
        red[warp * 8] = tx;
        red[warp * 8 + 1] = ty;
      }
    }
    __syncthreads();
    if (warp == 0) {
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[kW];
#pragma unroll
        for (int g = 0; g < kW; ++g) t[g] = red[g * 8 + j];
#pragma unroll
        for (int n = kW; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);
        w[j] = t[0];
      }
      double m0 = __ddiv_rn(w[0], static_cast<double>(N)), m1 = __ddiv_rn(w[1], static_cast<double>(N));
      if constexpr (kF32) {
        m0 = __dadd_rn(static_cast<double>(first.x), m0);
        m1 = __dadd_rn(static_cast<double>(first.y), m1);
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
      if (lane == 0) {
        ctl->h0 = h0; ctl->h1 = h1; ctl->m0 = m0; ctl->m1 = m1;
        ctl->nonfinite = nonfinite;
        ctl->degenerate = degenerate;
        write_mean_outputs(a, b, m0, m1);
      }
    }
    __syncthreads();
    const double h0 = ctl->h0, h1 = ctl->h1;
    const bool nonfinite = ctl->nonfinite != 0;
    const int status = (nonfinite ? kStatusNonfinite : 0) | (ctl->degenerate ? kStatusDegenerate : 0) | kStatusGeneral;
    auto loss_at = [&](int i) {
      const V2 v = load(i);
      return loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
    };

    double T_thr = 0.0, s_gt = 0.0;
    int c_gt = 0;
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
      unsigned long long* kred = reinterpret_cast<unsigned long long*>(red);
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
    int c_tot = 0;
    double s_tot = 0.0;
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      c_tot += iscr[w];
      s_tot += red[w];
    }
    if (tid == 0) write_risk_outputs(a, b, ctl, nonfinite, s_tot, c_tot, T_thr, status);

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
