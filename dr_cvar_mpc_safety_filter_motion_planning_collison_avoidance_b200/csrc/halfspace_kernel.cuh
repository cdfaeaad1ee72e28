// halfspace_kernel.cuh — sm_100a device code of the risk-bounded safe-halfspace path.
//
// One CTA (512 threads) per (scenario, obstacle, step) halfspace, persistent over the batch.  Per halfspace:
//   stage   N samples -> shared memory with cp.async.bulk (TMA bulk copy, mbarrier completion) or a strided loader
//   sweep A canonical 512-lane fp64 sums of x,y (+ heuristic second moments, max |coordinate|)  -> mean m
//   h       = unit(m - ego)                                                            core/geometry.py:35-53
//   sweep B fp32 inputs: a rigorous fp32 screen keeps only samples that can reach the candidate window, then
//           the survivors get the canonical fp64 loss L_i = -(h.xi_i) (no FMA); losses above the window are
//           counted/summed, losses inside it go to warp-private candidate lists
//   select  exact kc-th largest loss T by adaptive range-narrowing radix select on order-preserving u64 keys
//   finish  CVaR = (sum_{L>T} L + (k_f - #{L>T}) T) / k_f  and the three offsets        core/risk_metrics.py:84-338
// The window and the screen only decide HOW FAST the exact answer is found; a miss is detected and the general
// multi-sweep select runs instead.  The arithmetic contract (what is bit-identical to oracle/closed_form.py) is
// in DESIGN.md.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace drcvar {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kLanes = 512;          // canonical summation lanes (one per thread)
constexpr int kWarpCand = 128;       // candidate slots per warp (window path)
constexpr int kHistBuckets = 256;
constexpr int kResolveMax = 32;      // a bucket this small is resolved by one warp
constexpr int kMaxPerThread = 64;    // screening mask is 64 bits: N <= 64 * 512
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kBulkChunk = 32768;

constexpr int kStatusNonfinite = 1;
constexpr int kStatusGeneral = 2;
constexpr int kStatusDegenerate = 4;

struct KernelArgs {
  const void* samples;
  long long B;
  int N;
  long long stride_b, stride_n, stride_c;  // elements
  const double* ego;
  const double* h_in;
  double delta, eoa, R, k_f;
  int kc;
  int use_window;
  double z_lo, z_hi;
  int bulk;
  double* h_out;
  double* h_mean_out;
  double* g_out;
  double* cvar_out;
  double* var_out;
  double* gstar_out;
  int* status_out;
  int* tail_idx_out;
};

struct Ctl {
  unsigned long long mbar;
  double T;
  double h0, h1, t_lo, t_hi, m0, m1;
  float h0f, h1f, screen_thr, pad0;
  int bstar, rprime, cnt_in, small_n;
  int window_ok, nonfinite, degenerate, pad1;
};

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// shared-memory footprint of one CTA (host and device must agree)
__host__ __device__ inline size_t slot_bytes_for(long long n, size_t elem_bytes) {
  return (static_cast<size_t>(n) * 2 * elem_bytes + 127) & ~static_cast<size_t>(127);
}
constexpr int kRedDoubles = kWarps * 8;  // per buffer
__host__ __device__ inline size_t fixed_smem_bytes() {
  return sizeof(double) * kWarpCand * kWarps        // cand
         + sizeof(unsigned) * kHistBuckets          // hist
         + sizeof(double) * 2 * kRedDoubles         // red (double-buffered)
         + sizeof(double) * kResolveMax             // small
         + sizeof(int) * 4 * kWarps                 // ired
         + sizeof(Ctl);
}

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA bulk copy global -> shared::cta, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------- small helpers
__device__ __forceinline__ unsigned long long key_of(double v) {
  unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double(static_cast<long long>(u));
}
// canonical loss  L = 0 - (rn(h0*x) + rn(h1*y))   (never fused)
__device__ __forceinline__ double loss_of(double h0, double h1, double x, double y) {
  return __dsub_rn(0.0, __dadd_rn(__dmul_rn(h0, x), __dmul_rn(h1, y)));
}
__device__ __forceinline__ double norm2_canon(double a, double b) {
  return __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double warp_sum_any(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += shfl_xor_d(v, m);
  return v;
}
__device__ __forceinline__ float warp_sum_any(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(kFull, v, m);
  return v;
}
__device__ __forceinline__ double warp_sum_canon(double v) {  // xor 1,2,4,8,16 — part of the arithmetic contract
#pragma unroll
  for (int m = 1; m <= 16; m <<= 1) v = __dadd_rn(v, shfl_xor_d(v, m));
  return v;
}

// Exact r-th largest (1-based) among the enumerated losses whose keys lie in [lo, hi].
// for_each(f) must call f(L) for every candidate owned by the calling thread; all threads must call this.
// `hist` must be zero on entry (it is left dirty).
template <class ForEach>
__device__ double select_rank(ForEach&& for_each, unsigned long long lo, unsigned long long hi, int r,
                              unsigned* hist, double* small, Ctl* ctl, bool hist_clean) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    const unsigned long long span = hi - lo;
    if (span == 0) return value_of(lo);
    const int bits = 64 - __clzll(static_cast<long long>(span));
    const int shift = bits > 8 ? bits - 8 : 0;  // (span >> shift) < 256
    if (!hist_clean) {
      if (tid < kHistBuckets) hist[tid] = 0;
      if (tid == 0) ctl->small_n = 0;
      __syncthreads();
    }
    hist_clean = false;
    for_each([&](double L) {
      const unsigned long long k = key_of(L);
      if (k >= lo && k <= hi) atomicAdd(&hist[static_cast<unsigned>((k - lo) >> shift)], 1u);
    });
    __syncthreads();
    if (warp == 0) {
      // rows of 32 buckets, row 0 = top; lane l of row i is bucket 255 - (32 i + l)
      int run = 0, row = -1, r_row = 0;
#pragma unroll
      for (int i = 0; i < kHistBuckets / 32; ++i) {
        const int tot = __reduce_add_sync(kFull, static_cast<int>(hist[kHistBuckets - 1 - (32 * i + lane)]));
        if (row < 0 && run + tot >= r) {
          row = i;
          r_row = r - run;
        }
        run += tot;
      }
      if (row < 0) { row = kHistBuckets / 32 - 1; r_row = 1; }  // unreachable when r <= #candidates in range
      const int bucket = kHistBuckets - 1 - (32 * row + lane);
      const int c = static_cast<int>(hist[bucket]);
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
      }
      const int excl = incl - c;
      if (excl < r_row && r_row <= incl) {
        ctl->bstar = bucket;
        ctl->rprime = r_row - excl;
        ctl->cnt_in = c;
      }
    }
    __syncthreads();
    const int bstar = ctl->bstar;
    r = ctl->rprime;
    const int cnt_in = ctl->cnt_in;
    const unsigned long long nlo = lo + (static_cast<unsigned long long>(bstar) << shift);
    unsigned long long nhi = nlo + ((1ull << shift) - 1ull);
    if (nhi > hi || nhi < nlo) nhi = hi;
    lo = nlo;
    hi = nhi;
    if (cnt_in <= kResolveMax) {
      for_each([&](double L) {
        const unsigned long long k = key_of(L);
        if (k >= lo && k <= hi) {
          const int pos = atomicAdd(&ctl->small_n, 1);
          if (pos < kResolveMax) small[pos] = L;
        }
      });
      __syncthreads();
      if (warp == 0) {
        const unsigned long long mine = lane < cnt_in ? key_of(small[lane]) : 0ull;
        int rank = 0;
        for (int j = 0; j < cnt_in; ++j) {
          const unsigned long long other = __shfl_sync(kFull, mine, j);
          rank += (other > mine) || (other == mine && j < lane);
        }
        if (lane < cnt_in && rank == r - 1) ctl->T = value_of(mine);
      }
      __syncthreads();
      return ctl->T;
    }
  }
}

// ---------------------------------------------------------------------------------------------- the kernel
template <typename T, bool kTail>
__global__ void __launch_bounds__(kThreads, 2) halfspace_kernel(const KernelArgs a) {
  using V2 = typename Vec2<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const size_t slot_bytes = slot_bytes_for(N, sizeof(T));
  V2* sm = reinterpret_cast<V2*>(smem_raw);
  double* cand = reinterpret_cast<double*>(smem_raw + slot_bytes);
  unsigned* hist = reinterpret_cast<unsigned*>(cand + kWarpCand * kWarps);
  double* red_base = reinterpret_cast<double*>(hist + kHistBuckets);
  double* small = red_base + 2 * kRedDoubles;
  int* ired = reinterpret_cast<int*>(small + kResolveMax);
  Ctl* ctl = reinterpret_cast<Ctl*>(ired + 4 * kWarps);
  double* wcand = cand + warp * kWarpCand;

  const uint32_t copy_bytes = static_cast<uint32_t>(static_cast<size_t>(N) * sizeof(V2));
  auto issue_bulk = [&](long long b) {
    const unsigned char* src =
        reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b) * a.stride_b * sizeof(T);
    mbar_expect_tx(&ctl->mbar, copy_bytes);
    for (uint32_t off = 0; off < copy_bytes; off += kBulkChunk) {
      const uint32_t n = copy_bytes - off < kBulkChunk ? copy_bytes - off : kBulkChunk;
      bulk_g2s(smem_raw + off, src + off, n, &ctl->mbar);
    }
  };

  if (a.bulk) {
    if (tid == 0) {
      mbar_init(&ctl->mbar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && static_cast<long long>(blockIdx.x) < a.B) issue_bulk(blockIdx.x);
  }
  uint32_t phase = 0;
  int iter = 0;

  for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
    double* red = red_base + (iter & 1) * kRedDoubles;
    bool next_issued = false;
    const long long b_next = b + gridDim.x;

    // ------------------------------------------------------------------ stage
    if (a.bulk) {
      mbar_wait(&ctl->mbar, phase);
      phase ^= 1u;
    } else {
      const T* base = reinterpret_cast<const T*>(a.samples) + b * a.stride_b;
      for (int i = tid; i < N; i += kThreads) {
        const T* p = base + static_cast<long long>(i) * a.stride_n;
        V2 v;
        v.x = p[0];
        v.y = p[a.stride_c];
        sm[i] = v;
      }
      __syncthreads();
    }

    // ------------------------------------------------------------------ sweep A: canonical lane sums (+ heuristics)
    double sx = 0.0, sy = 0.0;
    T qdx = 0, qdy = 0, qxx = 0, qxy = 0, qyy = 0;  // second moments on every 4th tile, shifted by the first sample
    T amax = 0;                                     // max |coordinate| (bounds the fp32 screening error)
    const V2 first = sm[0];
    {
      auto moments = [&](const V2 v) {
        const T dx = v.x - first.x, dy = v.y - first.y;
        qdx += dx;
        qdy += dy;
        qxx = fma(dx, dx, qxx);
        qxy = fma(dx, dy, qxy);
        qyy = fma(dy, dy, qyy);
      };
      auto accum = [&](const V2 v) {
        sx = __dadd_rn(sx, static_cast<double>(v.x));
        sy = __dadd_rn(sy, static_cast<double>(v.y));
        amax = fmax(amax, fmax(fabs(v.x), fabs(v.y)));
      };
      int i = tid;
      for (; i + 3 * kThreads < N; i += 4 * kThreads) {
        const V2 v0 = sm[i], v1 = sm[i + kThreads], v2 = sm[i + 2 * kThreads], v3 = sm[i + 3 * kThreads];
        accum(v0);
        accum(v1);
        accum(v2);
        accum(v3);
        moments(v0);
      }
      bool lead = true;
      for (; i < N; i += kThreads) {
        const V2 v = sm[i];
        accum(v);
        if (lead) moments(v);
        lead = false;
      }
    }
    {
      // canonical: xor-butterfly inside each warp (= group of 32 lanes); the 16 warp totals are tree-added below
      const double tx = warp_sum_canon(sx);
      const double ty = warp_sum_canon(sy);
      const float mdx = warp_sum_any(static_cast<float>(qdx)), mdy = warp_sum_any(static_cast<float>(qdy));
      const float mxx = warp_sum_any(static_cast<float>(qxx)), mxy = warp_sum_any(static_cast<float>(qxy));
      const float myy = warp_sum_any(static_cast<float>(qyy));
      float mx = static_cast<float>(amax);
      if (static_cast<T>(mx) < amax) mx = __uint_as_float(__float_as_uint(mx) + 1);  // round up (T = double)
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, m));
      if (lane == 0) {
        double* w = red + warp * 8;
        w[0] = tx; w[1] = ty; w[2] = mdx; w[3] = mdy; w[4] = mxx; w[5] = mxy; w[6] = myy; w[7] = mx;
      }
      if (warp == 1 && lane < kHistBuckets / 32) {
        // nothing: hist is cleared below by warps that do not compute the direction
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ direction + window (warp 0); others clear hist
    if (warp == 0) {
      double w[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        double t[kWarps];
#pragma unroll
        for (int q = 0; q < kWarps; ++q) t[q] = red[q * 8 + j];
#pragma unroll
        for (int n = kWarps; n > 1; n >>= 1)
#pragma unroll
          for (int q = 0; q < n / 2; ++q) t[q] = __dadd_rn(t[2 * q], t[2 * q + 1]);  // adjacent-pair tree
        w[j] = t[0];
      }
      float mx = 0.f;
#pragma unroll
      for (int q = 0; q < kWarps; ++q) mx = fmaxf(mx, static_cast<float>(red[q * 8 + 7]));
      const double m0 = __ddiv_rn(w[0], static_cast<double>(N));
      const double m1 = __ddiv_rn(w[1], static_cast<double>(N));
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
      // heuristic window around the expected kc-th largest loss (affects speed only, never the result)
      const int groups = N / (4 * kThreads), rem = N - groups * 4 * kThreads;
      const double n_sub = static_cast<double>(groups * kThreads + (rem < kThreads ? rem : kThreads));
      const double ex = w[2] / n_sub, ey = w[3] / n_sub;
      const double cxx = w[4] / n_sub - ex * ex, cxy = w[5] / n_sub - ex * ey, cyy = w[6] / n_sub - ey * ey;
      const double var_l = h0 * h0 * cxx + 2.0 * h0 * h1 * cxy + h1 * h1 * cyy;
      const double mu_l = -(h0 * m0 + h1 * m1);
      const double sigma = static_cast<double>(sqrtf(static_cast<float>(var_l)));
      const int window_ok = a.use_window && (var_l > 0.0) && isfinite(sigma) && !nonfinite && (N <= kMaxPerThread * kThreads);
      const double t_lo = __dadd_rn(mu_l + a.z_lo * sigma, 0.0);  // +0.0: never -0.0 (canonical losses are +0)
      const double t_hi = __dadd_rn(mu_l + a.z_hi * sigma, 0.0);
      // fp32 screen: keep sample iff p32 <= thr, where p32 = fma(h1f, y, h0f*x) approximates p = h.xi = -L.
      // |p32 - p| <= 4 * 2^-24 * (|h0| + |h1|) * max|coord|; we allow 2^-19 (32x) plus the rounding of -t_lo.
      const float h0f = static_cast<float>(h0), h1f = static_cast<float>(h1);
      const float bound = (fabsf(h0f) + fabsf(h1f)) * mx * 1.9073486e-06f + fabsf(static_cast<float>(t_lo)) * 2.3841858e-07f;
      const float thr = static_cast<float>(-t_lo) + bound + 1.1754944e-38f;
      if (lane == 0) {
        ctl->h0 = h0; ctl->h1 = h1; ctl->m0 = m0; ctl->m1 = m1;
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->h0f = h0f; ctl->h1f = h1f; ctl->screen_thr = thr;
        ctl->window_ok = window_ok && isfinite(thr);
        ctl->nonfinite = nonfinite;
        ctl->degenerate = degenerate;
        ctl->small_n = 0;
      }
    } else if (warp >= kWarps - kHistBuckets / 32) {
      hist[tid - (kThreads - kHistBuckets)] = 0;  // last 8 warps clear the 256-bucket histogram
    }
    __syncthreads();
    const double h0 = ctl->h0, h1 = ctl->h1;
    const bool nonfinite = ctl->nonfinite != 0;
    const bool window = ctl->window_ok != 0;
    const double t_lo = ctl->t_lo, t_hi = ctl->t_hi;
    int status = (nonfinite ? kStatusNonfinite : 0) | (ctl->degenerate ? kStatusDegenerate : 0);

    double T_thr = 0.0;
    int c_gt = 0;
    double s_gt = 0.0;
    bool fast = false;

    if (!nonfinite) {
      // ---------------------------------------------------------------- sweep B (window path)
      if (window) {
        int nc = 0;  // candidates of this warp (warp-uniform)
        auto classify = [&](bool active, double L) {
          const bool hi = active && (L > t_hi);
          const bool cd = active && !hi && (L >= t_lo);
          if (hi) {
            ++c_gt;
            s_gt += L;
          }
          const unsigned bal = __ballot_sync(kFull, cd);
          if (bal) {
            const int pos = nc + __popc(bal & ((1u << lane) - 1u));
            if (cd && pos < kWarpCand) wcand[pos] = L;
            nc += __popc(bal);
          }
        };
        if constexpr (sizeof(T) == 4) {
          // phase 1: fp32 screen -> per-thread survivor mask (bit j <-> sample tid + 512 j)
          const float h0f = ctl->h0f, h1f = ctl->h1f, thr = ctl->screen_thr;
          unsigned mlo = 0, mhi = 0;
          {
            int i = tid;
            unsigned bit = 1u;
            int j = 0;
            for (; i + 3 * kThreads < N && j < 32; i += 4 * kThreads, j += 4, bit <<= 4) {
              const V2 v0 = sm[i], v1 = sm[i + kThreads], v2 = sm[i + 2 * kThreads], v3 = sm[i + 3 * kThreads];
              if (fmaf(h1f, v0.y, h0f * v0.x) <= thr) mlo |= bit;
              if (fmaf(h1f, v1.y, h0f * v1.x) <= thr) mlo |= bit << 1;
              if (fmaf(h1f, v2.y, h0f * v2.x) <= thr) mlo |= bit << 2;
              if (fmaf(h1f, v3.y, h0f * v3.x) <= thr) mlo |= bit << 3;
            }
            for (; i < N && j < 32; i += kThreads, ++j, bit <<= 1)
              if (fmaf(h1f, sm[i].y, h0f * sm[i].x) <= thr) mlo |= bit;
            bit = 1u;
            for (; i + 3 * kThreads < N; i += 4 * kThreads, bit <<= 4) {
              const V2 v0 = sm[i], v1 = sm[i + kThreads], v2 = sm[i + 2 * kThreads], v3 = sm[i + 3 * kThreads];
              if (fmaf(h1f, v0.y, h0f * v0.x) <= thr) mhi |= bit;
              if (fmaf(h1f, v1.y, h0f * v1.x) <= thr) mhi |= bit << 1;
              if (fmaf(h1f, v2.y, h0f * v2.x) <= thr) mhi |= bit << 2;
              if (fmaf(h1f, v3.y, h0f * v3.x) <= thr) mhi |= bit << 3;
            }
            for (; i < N; i += kThreads, bit <<= 1)
              if (fmaf(h1f, sm[i].y, h0f * sm[i].x) <= thr) mhi |= bit;
          }
          // phase 2: canonical fp64 loss of the survivors, warp-lockstep over each lane's k-th survivor
          for (int word = 0; word < 2; ++word) {
            unsigned m = word ? mhi : mlo;
            while (__any_sync(kFull, m != 0)) {
              const bool active = m != 0;
              double L = 0.0;
              if (active) {
                const int j = __ffs(m) - 1 + 32 * word;
                m &= m - 1;
                const V2 v = sm[tid + kThreads * j];
                L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
              }
              classify(active, L);
            }
          }
        } else {
          // fp64 inputs: no conversions to save, every sample takes the canonical path
          for (int base = 0; base < N; base += kThreads) {
            const int i = base + tid;
            const bool active = i < N;
            double L = 0.0;
            if (active) {
              const V2 v = sm[i];
              L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
            }
            classify(active, L);
          }
        }
        const int wc = __reduce_add_sync(kFull, c_gt);
        if (lane == 0) {
          ired[warp * 2] = wc;
          ired[warp * 2 + 1] = nc < kWarpCand ? nc : kWarpCand;
        }
        const int ovf = __syncthreads_or(nc > kWarpCand);
        int cnt_hi = 0, ncand = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          cnt_hi += ired[w * 2];
          ncand += ired[w * 2 + 1];
        }
        fast = !ovf && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
        if (fast) {
          // the sample slot is dead from here on: prefetch the next halfspace under the select phase
          if (!kTail && a.bulk && tid == 0 && b_next < a.B) issue_bulk(b_next);
          next_issued = !kTail && a.bulk;
          T_thr = select_rank(
              [&](auto&& f) {
                for (int j = lane; j < nc; j += 32) f(wcand[j]);
              },
              key_of(t_lo), key_of(t_hi), a.kc - cnt_hi, hist, small, ctl, true);
          for (int j = lane; j < nc; j += 32) {
            const double L = wcand[j];
            if (L > T_thr) {
              ++c_gt;
              s_gt += L;
            }
          }
        }
      }
      if (!fast) {
        // -------------------------------------------------------------- general path: sweeps over all samples
        status |= kStatusGeneral;
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int i = tid; i < N; i += kThreads) {
          const V2 v = sm[i];
          const unsigned long long k = key_of(loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y)));
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
        __syncthreads();  // red[] of this iteration was consumed by warp 0 above
        if (lane == 0) {
          kred[warp * 2] = kmin;
          kred[warp * 2 + 1] = kmax;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          kmin = kred[w * 2] < kmin ? kred[w * 2] : kmin;
          kmax = kred[w * 2 + 1] > kmax ? kred[w * 2 + 1] : kmax;
        }
        T_thr = select_rank(
            [&](auto&& f) {
              for (int i = tid; i < N; i += kThreads) {
                const V2 v = sm[i];
                f(loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y)));
              }
            },
            kmin, kmax, a.kc, hist, small, ctl, false);
        c_gt = 0;
        s_gt = 0.0;
        for (int i = tid; i < N; i += kThreads) {
          const V2 v = sm[i];
          const double L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
          if (L > T_thr) {
            ++c_gt;
            s_gt += L;
          }
        }
      }
    }

    // ------------------------------------------------------------------ block totals of (c_gt, s_gt)
    {
      const int wc = __reduce_add_sync(kFull, c_gt);
      const double ws = warp_sum_any(s_gt);
      __syncthreads();  // protects ired / red reuse
      if (lane == 0) {
        ired[2 * kWarps + warp] = wc;
        red[warp] = ws;
      }
      __syncthreads();
    }
    int c_tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) c_tot += ired[2 * kWarps + w];

    // ------------------------------------------------------------------ tail indices (parity mode only)
    if (kTail && a.tail_idx_out != nullptr) {
      int* out = a.tail_idx_out + b * static_cast<long long>(a.kc);
      if (nonfinite) {
        for (int i = tid; i < a.kc; i += kThreads) out[i] = -1;
      } else {
        const int need = a.kc - c_tot;
        int run_eq = 0, run_out = 0;
        int* weq = ired;           // [kWarps]
        int* wsel = ired + kWarps; // [kWarps]
        for (int bb = 0; bb < N; bb += kThreads) {
          const int i = bb + tid;
          const bool valid = i < N;
          double L = 0.0;
          if (valid) {
            const V2 v = sm[i];
            L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
          }
          const bool gt = valid && (L > T_thr), eq = valid && (L == T_thr);
          const unsigned meq = __ballot_sync(kFull, eq);
          if (lane == 0) weq[warp] = __popc(meq);
          __syncthreads();
          int eq_before = run_eq, tile_eq = 0;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) {
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
          for (int w = 0; w < kWarps; ++w) {
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

    // ------------------------------------------------------------------ epilogue (one thread)
    if (tid == 0) {
      const double m0 = ctl->m0, m1 = ctl->m1;
      double s_tot = red[0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) s_tot += red[w];
      const double hn = norm2_canon(h0, h1);
      const double r = __dmul_rn(a.R, hn);
      double cvar, g_cvar, g_star, g_dr, var_t;
      if (nonfinite) {
        cvar = __longlong_as_double(0x7ff8000000000000ll);
        var_t = cvar;
        g_cvar = 100.0;
        g_star = 100.0;
        g_dr = __dsub_rn(100.0, r);
      } else {
        const double S = __dadd_rn(s_tot, __dmul_rn(__dsub_rn(a.k_f, static_cast<double>(c_tot)), T_thr));
        cvar = __ddiv_rn(S, a.k_f);
        var_t = T_thr;
        const double cr = __dadd_rn(cvar, r);
        g_cvar = __dsub_rn(cr, a.delta);
        g_star = __dsub_rn(__dadd_rn(cr, a.eoa), a.delta);
        g_dr = __dsub_rn(g_star, r);
      }
      // mean halfspace: direction from the ORIGIN (core/halfspaces.py:88)
      double hm0, hm1;
      const double mn = norm2_canon(m0, m1);
      if (mn < 1e-10) {
        hm0 = 1.0;
        hm1 = 0.0;
      } else {
        hm0 = __ddiv_rn(m0, mn);
        hm1 = __ddiv_rn(m1, mn);
      }
      const double hmn = norm2_canon(hm0, hm1);
      const double g_mean =
          -__dsub_rn(__dadd_rn(__dmul_rn(hm0, m0), __dmul_rn(hm1, m1)), __dmul_rn(a.R, hmn));
      a.h_out[2 * b] = h0;
      a.h_out[2 * b + 1] = h1;
      if (a.h_mean_out) {
        a.h_mean_out[2 * b] = hm0;
        a.h_mean_out[2 * b + 1] = hm1;
      }
      a.g_out[3 * b] = g_mean;
      a.g_out[3 * b + 1] = g_cvar;
      a.g_out[3 * b + 2] = g_dr;
      if (a.cvar_out) a.cvar_out[b] = cvar;
      if (a.var_out) a.var_out[b] = var_t;
      if (a.gstar_out) a.gstar_out[b] = g_star;
      if (a.status_out) a.status_out[b] = status;
    }

    // ------------------------------------------------------------------ release the slot / prefetch
    if (!next_issued) {
      __syncthreads();
      if (a.bulk && tid == 0 && b_next < a.B) issue_bulk(b_next);
    }
  }
}

}  // namespace drcvar
