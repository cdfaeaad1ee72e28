// halfspace_kernel.cuh — sm_100a device code of the risk-bounded safe-halfspace path.
//
// One CTA (512 threads) per (scenario, obstacle, step) halfspace, persistent over the batch.  Per halfspace:
//   stage   N samples -> shared memory with cp.async.bulk (TMA bulk copy, mbarrier completion) or a strided loader
//   sweep A canonical lane sums of x,y (fp32 inputs: packed fp32 lane partials, fp64 cross-lane tree;
//           fp64 inputs: fp64 throughout) + heuristic second moments + max |coordinate|          -> mean m
//   h       = unit(m - ego)                                                            core/geometry.py:35-53
//   sweep B classify every sample against a statistical window [t_lo, t_hi] around the expected kc-th largest
//           loss.  fp32 inputs: a rigorous fp32 bound decides "surely above" (count + raw coordinate sums; the
//           loss sum follows from linearity), "surely below" (ignored) or "needs the exact fp64 loss" (a bit in a
//           per-thread mask).  fp64 inputs: exact canonical loss for every sample.
//   phase 2 masked samples get the canonical fp64 loss L_i = -(h.xi_i) (no FMA); window losses go to warp-private
//           candidate lists and a 256-bucket histogram over the window
//   select  exact kc-th largest loss T: histogram scan -> bucket -> rank resolve by one warp (general fallback:
//           adaptive range-narrowing radix select on order-preserving u64 keys over all samples)
//   finish  CVaR = (sum_{L>T} L + (k_f - #{L>T}) T) / k_f  and the three offsets        core/risk_metrics.py:84-338
// The window and the fp32 bound only decide HOW FAST the exact threshold is found; a miss is detected and the
// general multi-sweep select runs instead.  The arithmetic contract is in DESIGN.md / oracle/closed_form.py.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace drcvar {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kLanes = 512;          // canonical cross-lane tree width (fp32 inputs: 1024 fp32 lanes, paired)
constexpr int kWarpCand = 96;        // candidate losses per warp (window path)
constexpr int kWarpList = 160;       // masked sample indices per warp (window path)
constexpr int kHistBuckets = 256;
constexpr int kResolveMax = 32;      // a bucket this small is resolved by one warp
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
  unsigned long long key_lo;       // key(t_lo): histogram origin
  double T;
  double h0, h1, t_lo, t_hi, m0, m1;
  float h0f, h1f, thr_above, thr_keep;
  int hist_shift;
  int bstar, rprime, cnt_in, small_n;
  int window_ok, nonfinite, degenerate, c_tot, pad;
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
  return sizeof(double) * kWarpCand * kWarps          // cand
         + sizeof(unsigned short) * kWarpList * kWarps  // list
         + sizeof(unsigned) * kHistBuckets            // hist
         + sizeof(double) * 2 * kRedDoubles           // red (double-buffered by iteration parity)
         + sizeof(double) * 2 * kRedDoubles           // fin (double-buffered): per-warp finals
         + sizeof(double) * 2 * kResolveMax           // small (double-buffered)
         + sizeof(int) * 4 * kWarps                   // ired
         + 2 * sizeof(Ctl);                           // ctl (double-buffered)
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
__device__ __forceinline__ float absmax3(float m, float a, float b) { return fmaxf(m, fmaxf(fabsf(a), fabsf(b))); }

// Exact r-th largest (1-based) among the enumerated losses whose keys lie in [lo, hi] (general machinery).
// for_each(f) must call f(L) for every candidate owned by the calling thread; all threads must call this.
template <class ForEach>
__device__ double select_rank(ForEach&& for_each, unsigned long long lo, unsigned long long hi, int r,
                              unsigned* hist, double* small, Ctl* ctl) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    const unsigned long long span = hi - lo;
    if (span == 0) return value_of(lo);
    const int bits = 64 - __clzll(static_cast<long long>(span));
    const int shift = bits > 8 ? bits - 8 : 0;  // (span >> shift) < 256
    if (tid < kHistBuckets) hist[tid] = 0;
    if (tid == 0) ctl->small_n = 0;
    __syncthreads();
    for_each([&](double L) {
      const unsigned long long k = key_of(L);
      if (k >= lo && k <= hi) atomicAdd(&hist[static_cast<unsigned>((k - lo) >> shift)], 1u);
    });
    __syncthreads();
    if (warp == 0) {
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
  constexpr bool kF32 = sizeof(T) == 4;
  // samples one thread touches per row: fp32 -> one float4 = samples (2t, 2t+1) of a 1024-sample tile;
  // fp64 -> one double2 = sample t of a 512-sample tile
  constexpr int kPerRow = kF32 ? 2 : 1;
  constexpr int kTile = kThreads * kPerRow;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const size_t slot_bytes = slot_bytes_for(N, sizeof(T));
  V2* sm = reinterpret_cast<V2*>(smem_raw);
  double* cand = reinterpret_cast<double*>(smem_raw + slot_bytes);
  unsigned short* list = reinterpret_cast<unsigned short*>(cand + kWarpCand * kWarps);
  unsigned* hist = reinterpret_cast<unsigned*>(list + kWarpList * kWarps);
  double* red_base = reinterpret_cast<double*>(hist + kHistBuckets);
  double* fin_base = red_base + 2 * kRedDoubles;
  double* small_base = fin_base + 2 * kRedDoubles;
  int* ired = reinterpret_cast<int*>(small_base + 2 * kResolveMax);
  Ctl* ctl_base = reinterpret_cast<Ctl*>(ired + 4 * kWarps);
  unsigned long long* mbar = &ctl_base[0].mbar;
  double* wcand = cand + warp * kWarpCand;
  unsigned short* wlist = list + warp * kWarpList;

  const uint32_t copy_bytes = static_cast<uint32_t>(static_cast<size_t>(N) * sizeof(V2));
  auto issue_bulk = [&](long long b) {
    const unsigned char* src =
        reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b) * a.stride_b * sizeof(T);
    mbar_expect_tx(mbar, copy_bytes);
    for (uint32_t off = 0; off < copy_bytes; off += kBulkChunk) {
      const uint32_t n = copy_bytes - off < kBulkChunk ? copy_bytes - off : kBulkChunk;
      bulk_g2s(smem_raw + off, src + off, n, mbar);
    }
  };

  if (a.bulk) {
    if (tid == 0) {
      mbar_init(mbar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && static_cast<long long>(blockIdx.x) < a.B) issue_bulk(blockIdx.x);
  }
  uint32_t phase = 0;
  int iter = 0;

  for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
    const int par = iter & 1;
    double* red = red_base + par * kRedDoubles;
    double* fin = fin_base + par * kRedDoubles;
    double* small = small_base + par * kResolveMax;
    Ctl* ctl = ctl_base + par;
    bool next_issued = false;
    const long long b_next = b + gridDim.x;

    // ------------------------------------------------------------------ stage
    if (a.bulk) {
      mbar_wait(mbar, phase);
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
    const int rows = (N + kTile - 1) / kTile;
    const int full_rows = N / kTile;
    double sx, sy;                                  // this thread's contribution to the 512-wide fp64 tree
    T qdx = 0, qdy = 0, qxx = 0, qxy = 0, qyy = 0;  // second moments (warps 0-3 only), shifted by the first sample
    float amax = 0.f;                               // max |coordinate| (bounds the fp32 classification error)
    const V2 first = sm[0];
    if constexpr (kF32) {
      const float4* sm4 = reinterpret_cast<const float4*>(smem_raw);
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);  // lanes 2t and 2t+1, fp32 partial sums
      auto accum = [&](const float4 v) {
        a0 = __fadd2_rn(a0, make_float2(v.x, v.y));
        a1 = __fadd2_rn(a1, make_float2(v.z, v.w));
        amax = absmax3(absmax3(amax, v.x, v.y), v.z, v.w);
      };
      auto moments = [&](const float4 v) {
        const float dx0 = v.x - first.x, dy0 = v.y - first.y, dx1 = v.z - first.x, dy1 = v.w - first.y;
        qdx += dx0 + dx1;
        qdy += dy0 + dy1;
        qxx = fmaf(dx0, dx0, fmaf(dx1, dx1, qxx));
        qxy = fmaf(dx0, dy0, fmaf(dx1, dy1, qxy));
        qyy = fmaf(dy0, dy0, fmaf(dy1, dy1, qyy));
      };
      if (warp < 4) {
#pragma unroll 2
        for (int m = 0; m < full_rows; ++m) {
          const float4 v = sm4[m * kThreads + tid];
          accum(v);
          moments(v);
        }
      } else {
#pragma unroll 4
        for (int m = 0; m < full_rows; ++m) accum(sm4[m * kThreads + tid]);
      }
      if (full_rows < rows) {  // ragged last tile: element-wise
        const int i0 = full_rows * kTile + 2 * tid;
        if (i0 < N) {
          const float2 v = sm[i0];
          a0 = __fadd2_rn(a0, v);
          amax = absmax3(amax, v.x, v.y);
        }
        if (i0 + 1 < N) {
          const float2 v = sm[i0 + 1];
          a1 = __fadd2_rn(a1, v);
          amax = absmax3(amax, v.x, v.y);
        }
      }
      sx = __dadd_rn(static_cast<double>(a0.x), static_cast<double>(a1.x));  // adjacent lanes, in fp64
      sy = __dadd_rn(static_cast<double>(a0.y), static_cast<double>(a1.y));
    } else {
      sx = 0.0;
      sy = 0.0;
      auto moments = [&](const V2 v) {
        const T dx = v.x - first.x, dy = v.y - first.y;
        qdx += dx;
        qdy += dy;
        qxx = fma(dx, dx, qxx);
        qxy = fma(dx, dy, qxy);
        qyy = fma(dy, dy, qyy);
      };
      if (warp < 4) {
#pragma unroll 2
        for (int i = tid; i < N; i += kThreads) {
          const V2 v = sm[i];
          sx = __dadd_rn(sx, v.x);
          sy = __dadd_rn(sy, v.y);
          moments(v);
        }
      } else {
#pragma unroll 4
        for (int i = tid; i < N; i += kThreads) {
          const V2 v = sm[i];
          sx = __dadd_rn(sx, v.x);
          sy = __dadd_rn(sy, v.y);
        }
      }
    }
    {
      // canonical: xor-butterfly inside each warp (= group of 32 values); the 16 warp totals are tree-added below
      const double tx = warp_sum_canon(sx);
      const double ty = warp_sum_canon(sy);
      const unsigned mxb = __reduce_max_sync(kFull, __float_as_uint(amax));
      double* w = red + warp * 8;
      if (warp < 4) {
        const float mdx = warp_sum_any(static_cast<float>(qdx)), mdy = warp_sum_any(static_cast<float>(qdy));
        const float mxx = warp_sum_any(static_cast<float>(qxx)), mxy = warp_sum_any(static_cast<float>(qxy));
        const float myy = warp_sum_any(static_cast<float>(qyy));
        if (lane == 0) {
          w[2] = mdx; w[3] = mdy; w[4] = mxx; w[5] = mxy; w[6] = myy;
        }
      }
      if (lane == 0) {
        w[0] = tx; w[1] = ty; w[7] = static_cast<double>(__uint_as_float(mxb));
      }
    }
    __syncthreads();  // S1

    // ------------------------------------------------------------------ direction + window (warp 0); warps 8-15 clear hist
    if (warp == 0) {
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[kWarps];
#pragma unroll
        for (int q = 0; q < kWarps; ++q) t[q] = red[q * 8 + j];
#pragma unroll
        for (int n = kWarps; n > 1; n >>= 1)
#pragma unroll
          for (int q = 0; q < n / 2; ++q) t[q] = __dadd_rn(t[2 * q], t[2 * q + 1]);  // adjacent-pair tree
        w[j] = t[0];
      }
      double q[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) q[j] = (red[0 * 8 + 2 + j] + red[1 * 8 + 2 + j]) + (red[2 * 8 + 2 + j] + red[3 * 8 + 2 + j]);
      float mx = 0.f;
#pragma unroll
      for (int qq = 0; qq < kWarps; ++qq) mx = fmaxf(mx, static_cast<float>(red[qq * 8 + 7]));
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
      // samples that entered the moments: warps 0-3 (threads 0..127), full rows only for fp32 inputs
      const int n_sub_i = kF32 ? full_rows * 128 * 2
                               : 128 * (N / kThreads) + ((N % kThreads) < 128 ? (N % kThreads) : 128);
      const double n_sub = static_cast<double>(n_sub_i > 0 ? n_sub_i : 1);
      const double ex = q[0] / n_sub, ey = q[1] / n_sub;
      const double cxx = q[2] / n_sub - ex * ex, cxy = q[3] / n_sub - ex * ey, cyy = q[4] / n_sub - ey * ey;
      const double var_l = h0 * h0 * cxx + 2.0 * h0 * h1 * cxy + h1 * h1 * cyy;
      const double mu_l = -(h0 * m0 + h1 * m1);
      const double sigma = static_cast<double>(sqrtf(static_cast<float>(var_l)));
      int window_ok = a.use_window && (n_sub_i >= 256) && (var_l > 0.0) && isfinite(sigma) && !nonfinite &&
                      (rows * kPerRow <= 64);
      const double t_lo = __dadd_rn(mu_l + a.z_lo * sigma, 0.0);  // +0.0: never -0.0 (canonical losses are +0)
      const double t_hi = __dadd_rn(mu_l + a.z_hi * sigma, 0.0);
      // fp32 classification of p32 = fma(h1f, y, h0f*x)  (p = h.xi = -L):
      //   |p32 - p| <= 4 * 2^-24 * (|h0| + |h1|) * max|coord|; we allow 2^-19 (32x) plus the rounding of the thresholds.
      //   p32 <  thr_above  =>  L > t_hi  for sure;    p32 > thr_keep  =>  L < t_lo  for sure.
      const float h0f = static_cast<float>(h0), h1f = static_cast<float>(h1);
      const float bound = (fabsf(h0f) + fabsf(h1f)) * mx * 1.9073486e-06f + 1.1754944e-38f;
      const float thr_keep = static_cast<float>(-t_lo) + (bound + fabsf(static_cast<float>(t_lo)) * 2.3841858e-07f);
      const float thr_above = static_cast<float>(-t_hi) - (bound + fabsf(static_cast<float>(t_hi)) * 2.3841858e-07f);
      const unsigned long long klo = key_of(t_lo), khi = key_of(t_hi);
      const unsigned long long span = khi - klo;
      const int bits = span ? 64 - __clzll(static_cast<long long>(span)) : 0;
      window_ok = window_ok && isfinite(thr_keep) && isfinite(thr_above) && (khi >= klo);
      if (lane == 0) {
        ctl->h0 = h0; ctl->h1 = h1; ctl->m0 = m0; ctl->m1 = m1;
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->h0f = h0f; ctl->h1f = h1f; ctl->thr_keep = thr_keep; ctl->thr_above = thr_above;
        ctl->key_lo = klo;
        ctl->hist_shift = bits > 8 ? bits - 8 : 0;
        ctl->window_ok = window_ok;
        ctl->nonfinite = nonfinite;
        ctl->degenerate = degenerate;
        ctl->small_n = 0;
      }
    } else if (warp >= kWarps - kHistBuckets / 32) {
      hist[tid - (kThreads - kHistBuckets)] = 0;  // last 8 warps clear the 256-bucket histogram
    }
    __syncthreads();  // S2
    const double h0 = ctl->h0, h1 = ctl->h1;
    const bool nonfinite = ctl->nonfinite != 0;
    const bool window = ctl->window_ok != 0;
    const double t_lo = ctl->t_lo, t_hi = ctl->t_hi;
    int status = (nonfinite ? kStatusNonfinite : 0) | (ctl->degenerate ? kStatusDegenerate : 0);

    double T_thr = 0.0;
    int c_gt = 0;        // exact-classified losses above the threshold (this thread)
    double s_gt = 0.0;   // their sum
    bool fast = false;
    bool finished_by_warp0 = false;

    if (!nonfinite) {
      if (window) {
        // -------------------------------------------------------------- sweep B: classify, build the exact-needed mask
        unsigned mlo = 0, mhi = 0;
        int c32 = 0;                                 // "surely above" by the fp32 bound
        float2 ab = make_float2(0.f, 0.f);           // their raw coordinate sums (loss sum follows by linearity)
        if constexpr (kF32) {
          const float4* sm4 = reinterpret_cast<const float4*>(smem_raw);
          const float h0f = ctl->h0f, h1f = ctl->h1f, thr_keep = ctl->thr_keep, thr_above = ctl->thr_above;
          auto classify2 = [&](const float4 v, unsigned& mask, const unsigned bit) {
            const float p0 = fmaf(h1f, v.y, h0f * v.x), p1 = fmaf(h1f, v.w, h0f * v.z);
            const bool up0 = p0 < thr_above, up1 = p1 < thr_above;
            if (up0) { ab = __fadd2_rn(ab, make_float2(v.x, v.y)); ++c32; }
            if (up1) { ab = __fadd2_rn(ab, make_float2(v.z, v.w)); ++c32; }
            if (!up0 && p0 <= thr_keep) mask |= bit;
            if (!up1 && p1 <= thr_keep) mask |= bit << 1;
          };
          int m = 0;
          const int lo_rows = full_rows < 16 ? full_rows : 16;
          unsigned bit = 1u;
#pragma unroll 4
          for (; m < lo_rows; ++m, bit <<= 2) classify2(sm4[m * kThreads + tid], mlo, bit);
          bit = 1u;
#pragma unroll 4
          for (; m < full_rows; ++m, bit <<= 2) classify2(sm4[m * kThreads + tid], mhi, bit);
          if (full_rows < rows) {  // ragged last tile
            const int i0 = full_rows * kTile + 2 * tid;
            const unsigned bitr = 1u << (2 * (full_rows & 15));
            unsigned& mask = full_rows < 16 ? mlo : mhi;
            if (i0 < N) {
              const float2 v = sm[i0];
              const float p = fmaf(h1f, v.y, h0f * v.x);
              if (p < thr_above) { ab = __fadd2_rn(ab, v); ++c32; }
              else if (p <= thr_keep) mask |= bitr;
            }
            if (i0 + 1 < N) {
              const float2 v = sm[i0 + 1];
              const float p = fmaf(h1f, v.y, h0f * v.x);
              if (p < thr_above) { ab = __fadd2_rn(ab, v); ++c32; }
              else if (p <= thr_keep) mask |= bitr << 1;
            }
          }
        } else {
          unsigned bit = 1u;
          int m = 0;
          for (int i = tid; i < N; i += kThreads, ++m, bit = (bit << 1) | (bit >> 31)) {
            const V2 v = sm[i];
            const double L = loss_of(h0, h1, v.x, v.y);
            if (L > t_hi) {
              ++c_gt;
              s_gt += L;
            } else if (L >= t_lo) {
              if (m < 32) mlo |= bit; else mhi |= bit;
            }
          }
        }

        // -------------------------------------------------------------- phase 2: compact the masked samples per warp
        const int mine_n = __popc(mlo) + __popc(mhi);
        int incl = mine_n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += t;
        }
        const int n_list = __shfl_sync(kFull, incl, 31);
        bool overflow = n_list > kWarpList;
        if (!overflow) {
          int pos = incl - mine_n;
          unsigned mm = mlo;
          while (mm) {
            const int bitpos = __ffs(mm) - 1;
            mm &= mm - 1;
            const int i = kF32 ? ((bitpos >> 1) * kTile + 2 * tid + (bitpos & 1)) : (bitpos * kThreads + tid);
            wlist[pos++] = static_cast<unsigned short>(i);
          }
          mm = mhi;
          while (mm) {
            const int bitpos = __ffs(mm) - 1;
            mm &= mm - 1;
            const int i = kF32 ? (((bitpos >> 1) + 16) * kTile + 2 * tid + (bitpos & 1)) : ((bitpos + 32) * kThreads + tid);
            wlist[pos++] = static_cast<unsigned short>(i);
          }
        }
        __syncwarp();
        // exact canonical loss of the listed samples, dense over the warp; window losses -> candidates + histogram
        int nc = 0;  // candidates of this warp (warp-uniform)
        const unsigned long long klo = ctl->key_lo;
        const int hshift = ctl->hist_shift;
        if (!overflow) {
          for (int k0 = 0; k0 < n_list; k0 += 32) {
            const int k = k0 + lane;
            const bool active = k < n_list;
            double L = 0.0;
            if (active) {
              const V2 v = sm[wlist[k]];
              L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
            }
            const bool up = active && (L > t_hi);
            const bool cd = active && !up && (L >= t_lo);
            if (kF32 && up) {  // (fp64 inputs were classified exactly in sweep B; the list holds candidates only)
              ++c_gt;
              s_gt += L;
            }
            const unsigned bal = __ballot_sync(kFull, cd);
            if (bal) {
              const int pos = nc + __popc(bal & ((1u << lane) - 1u));
              if (cd && pos < kWarpCand) {
                wcand[pos] = L;
                atomicAdd(&hist[static_cast<unsigned>((key_of(L) - klo) >> hshift)], 1u);
              }
              nc += __popc(bal);
            }
          }
          overflow = nc > kWarpCand;
        }
        // per-warp partials -> fin[]: {exact sum, raw x sum, raw y sum}, ired[]: {exact count + fp32 count, candidates}
        {
          const int wc = __reduce_add_sync(kFull, c_gt + c32);
          const double ws = warp_sum_any(s_gt);
          const double wx = warp_sum_any(static_cast<double>(ab.x));
          const double wy = warp_sum_any(static_cast<double>(ab.y));
          if (lane == 0) {
            ired[warp * 2] = wc;
            ired[warp * 2 + 1] = nc;
            fin[warp * 8 + 0] = ws;
            fin[warp * 8 + 1] = wx;
            fin[warp * 8 + 2] = wy;
          }
        }
        const int ovf = __syncthreads_or(overflow);  // S3
        int cnt_hi = 0, ncand = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          cnt_hi += ired[w * 2];
          ncand += ired[w * 2 + 1];
        }
        fast = !ovf && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
        if (fast) {
          // the sample slot is dead from here on (unless tail indices are wanted): prefetch the next halfspace
          if (!kTail && a.bulk && tid == 0 && b_next < a.B) issue_bulk(b_next);
          next_issued = !kTail && a.bulk;
          // warp 0: scan the histogram from the top for the bucket holding rank r
          if (warp == 0) {
            const int r = a.kc - cnt_hi;
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
            if (row < 0) { row = kHistBuckets / 32 - 1; r_row = 1; }
            const int bucket = kHistBuckets - 1 - (32 * row + lane);
            const int c = static_cast<int>(hist[bucket]);
            int inc2 = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int t = __shfl_up_sync(kFull, inc2, d);
              if (lane >= d) inc2 += t;
            }
            const int excl = inc2 - c;
            if (excl < r_row && r_row <= inc2) {
              ctl->bstar = bucket;
              ctl->rprime = r_row - excl;
              ctl->cnt_in = c;
            }
          }
          __syncthreads();  // S4
          const int bstar = ctl->bstar, cnt_in = ctl->cnt_in;
          if (cnt_in <= kResolveMax) {
            // candidates above bucket b* are above T; bucket b* goes to one warp for exact ranking
            double s3 = 0.0;
            int c3 = 0;
            for (int j = lane; j < nc; j += 32) {
              const double L = wcand[j];
              const int bk = static_cast<int>((key_of(L) - klo) >> hshift);
              if (bk > bstar) {
                ++c3;
                s3 += L;
              } else if (bk == bstar) {
                const int pos = atomicAdd(&ctl->small_n, 1);
                if (pos < kResolveMax) small[pos] = L;
              }
            }
            const int wc3 = __reduce_add_sync(kFull, c3);
            const double ws3 = warp_sum_any(s3);
            if (lane == 0) {
              ired[2 * kWarps + warp] = wc3;
              fin[warp * 8 + 3] = ws3;
            }
            __syncthreads();  // S5
            finished_by_warp0 = true;
            if (warp == 0) {
              // resolve T inside bucket b*: all-pairs rank; then sum the bucket members above T in rank order
              const int r = ctl->rprime;
              const double mineL = lane < cnt_in ? small[lane] : 0.0;
              const unsigned long long mine = lane < cnt_in ? key_of(mineL) : 0ull;
              int rank = 0;
              for (int j = 0; j < cnt_in; ++j) {
                const unsigned long long other = __shfl_sync(kFull, mine, j);
                rank += (other > mine) || (other == mine && j < lane);
              }
              const unsigned owner = __ballot_sync(kFull, lane < cnt_in && rank == r - 1);
              const double Tval = __shfl_sync(kFull, mineL, __ffs(owner) - 1);
              // deterministic: place by rank, then butterfly
              __syncwarp();
              if (lane < cnt_in) small[rank] = mineL;
              __syncwarp();
              const double byrank = (lane < cnt_in && lane < r - 1) ? small[lane] : 0.0;  // ranks 0..r-2 are > or == T
              const bool strictly = (lane < cnt_in && lane < r - 1) && (key_of(byrank) > key_of(Tval));
              const double s4 = warp_sum_any(strictly ? byrank : 0.0);
              const int c4 = __popc(__ballot_sync(kFull, strictly));
              T_thr = Tval;
              // totals
              double s_exact = 0.0, s_x = 0.0, s_y = 0.0, s_c = 0.0;
              int c_all = cnt_hi + c4;
#pragma unroll
              for (int w = 0; w < kWarps; ++w) {
                s_exact += fin[w * 8 + 0];
                s_x += fin[w * 8 + 1];
                s_y += fin[w * 8 + 2];
                s_c += fin[w * 8 + 3];
                c_all += ired[2 * kWarps + w];
              }
              // loss sum of the "surely above" set by linearity: sum_i -(h.xi_i) = -(h0 sum x + h1 sum y)
              const double s_lin = -(h0 * s_x + h1 * s_y);
              s_gt = ((s_exact + s_lin) + s_c) + s4;
              c_gt = c_all;
              if (kTail && lane == 0) {
                ctl->T = T_thr;
                ctl->c_tot = c_all;
              }
            }
          } else {
            // a dense / heavily tied bucket: finish with the general narrowing loop on the candidates
            const unsigned long long lo2 = klo + (static_cast<unsigned long long>(bstar) << hshift);
            unsigned long long hi2 = lo2 + ((1ull << hshift) - 1ull);
            const unsigned long long khi = key_of(t_hi);
            if (hi2 > khi || hi2 < lo2) hi2 = khi;
            // everything above bucket b* is above T
            int above_b = 0;
            for (int j = lane; j < nc; j += 32) above_b += static_cast<int>((key_of(wcand[j]) - klo) >> hshift) > bstar;
            (void)above_b;
            T_thr = select_rank(
                [&](auto&& f) {
                  for (int j = lane; j < nc; j += 32) f(wcand[j]);
                },
                lo2, hi2, ctl->rprime, hist, small, ctl);
            for (int j = lane; j < nc; j += 32) {
              const double L = wcand[j];
              if (L > T_thr) {
                ++c_gt;
                s_gt += L;
              }
            }
            // fold the fp32 "surely above" set in (thread 0 of each warp carries the warp's linear part)
            c_gt += c32;
            const double wx = warp_sum_any(static_cast<double>(ab.x));
            const double wy = warp_sum_any(static_cast<double>(ab.y));
            if (lane == 0) s_gt += -(h0 * wx + h1 * wy);
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
            kmin, kmax, a.kc, hist, small, ctl);
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

    // ------------------------------------------------------------------ block totals of (c_gt, s_gt) unless warp 0 has them
    int c_tot = c_gt;
    double s_tot = s_gt;
    if (!finished_by_warp0) {
      const int wc = __reduce_add_sync(kFull, c_gt);
      const double ws = warp_sum_any(s_gt);
      __syncthreads();  // protects ired / fin reuse
      if (lane == 0) {
        ired[3 * kWarps + warp] = wc;
        fin[warp * 8 + 4] = ws;
      }
      __syncthreads();
      c_tot = 0;
      s_tot = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        c_tot += ired[3 * kWarps + w];
        s_tot += fin[w * 8 + 4];
      }
    } else if (kTail) {
      __syncthreads();  // T and the total count were produced by warp 0
      T_thr = ctl->T;
      c_tot = ctl->c_tot;
    }

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
        __syncthreads();
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
