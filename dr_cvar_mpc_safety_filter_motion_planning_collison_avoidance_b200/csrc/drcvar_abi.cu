// drcvar_abi.cu — host side of libdrcvar.so: argument checks, launch planning, host<->device staging.
// The C ABI is declared in include/drcvar.h (each entry cites the reference interface it replaces).
#include "../../include/drcvar.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "halfspace_kernel.cuh"
#include "streaming_kernel.cuh"
#include "cluster_kernel.cuh"
#include "cluster_kernel_f64.cuh"
#include "pipelined_kernel.cuh"
#ifndef DRCVAR_PIPELINE_F64_DEFAULT
#define DRCVAR_PIPELINE_F64_DEFAULT 0   // fp64 samples at resident sizes: 0 = halfspace_kernel<double>, 16 = pipelined_kernel<double, 16>
#endif

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(DRCVAR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                        \
  } while (0)

// ---- tail size (mirrors oracle/closed_form.py:tail_count) -------------------------------------------------
bool tail_count(double alpha, long long n, double* k_f_out, long long* kc_out) {
  if (!(alpha > 0.0) || !(alpha <= 1.0) || n < 1) return false;
  double k_f = alpha * static_cast<double>(n);
  const double kr = std::nearbyint(k_f);
  if (std::fabs(k_f - kr) <= 1e-9 * std::max(1.0, kr)) k_f = kr;
  if (k_f > static_cast<double>(n)) k_f = static_cast<double>(n);
  long long kc = static_cast<long long>(std::ceil(k_f));
  kc = std::min(std::max(kc, 1LL), n);
  *k_f_out = k_f;
  *kc_out = kc;
  return true;
}

// ---- inverse normal CDF (Acklam's rational approximation, |rel err| < 1.2e-9): speed heuristic only ---------
double inv_norm_cdf(double p) {
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01,  -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  const double plow = 0.02425, phigh = 1 - plow;
  if (p < plow) {
    const double q = std::sqrt(-2 * std::log(p));
    return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  if (p > phigh) {
    const double q = std::sqrt(-2 * std::log(1 - p));
    return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
           ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  const double q = p - 0.5, r = q * q;
  return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
         (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
}

// Candidate window [z_lo, z_hi] (in loss-sigma units around the loss mean) expected to bracket the kc-th
// largest loss.  Purely a speed heuristic: a miss is detected on the device and the general select runs.
bool plan_window(long long n, long long kc, long long n_sigma, double cand_capacity, double* z_lo, double* z_hi) {
  const double p = static_cast<double>(kc) / static_cast<double>(n);  // upper-tail fraction
  if (n < 1024 || kc < 16 || n - kc < 16 || p < 1e-3 || p > 0.999) return false;
  const double z = inv_norm_cdf(1.0 - p);
  const double phi = std::exp(-0.5 * z * z) / std::sqrt(2.0 * M_PI);
  const double sq = std::sqrt(p * (1.0 - p) / static_cast<double>(n)) / phi;  // std of the sample quantile
  // (actual - predicted) quantile in z units, predicted = mean^ + z sigma^ from the sample moments.  For Gaussian losses
  // the sample mean and sigma are independent of the studentised quantile (Basu), so their variances SUBTRACT from the
  // quantile noise: var = p(1-p)/(n phi^2) - 1/n - z^2/(2n).  5 sigma + slack for non-normality; floor at half the
  // quantile noise.  (A miss only costs a re-fetch + the general select, never the result.)
  const double nn = static_cast<double>(n);
  // sigma^ taken from a subsample of n_sigma <= n samples (fp64 inputs: every 4th row) adds z^2 (1/(2 n_sigma) - 1/(2n))
  const double extra = 0.5 * z * z * (1.0 / static_cast<double>(n_sigma) - 1.0 / nn);
  const double var = std::max(sq * sq - 1.0 / nn - 0.5 * z * z / nn, 0.25 * sq * sq) + std::max(extra, 0.0);
  const double w = 4.0 * std::sqrt(var) + 0.002;
  const double expect = static_cast<double>(n) * 2.0 * w * phi;
  if (expect > cand_capacity) return false;  // per-warp candidate lists: leave headroom for the spread between warps
  *z_lo = z - w;
  *z_hi = z + w;
  return true;
}

// CTAs per halfspace of the cluster kernels: the smallest cluster whose per-CTA part + scratch fits one SM (0: none does)
int cluster_ctas_for(long long n, size_t elem_bytes, size_t smem_optin) {
  for (int cc = 2; cc <= drcvar::kClMaxCtas; cc *= 2)
    if (drcvar::cluster_smem_bytes(n, cc, elem_bytes) <= smem_optin) return cc;
  return 0;
}

#ifdef DRCVAR_PROFILE_PHASES
long long* g_phase_cycles = nullptr;  // device buffer [4096][2][12], set by drcvar_debug_phase_buffer()
#endif

struct DeviceInfo {
  bool ready = false;
  int sms = 0;
  int max_smem_optin = 0;
};
DeviceInfo g_dev[64];
std::mutex g_dev_mu;

int device_info(int device, DeviceInfo** out) {
  if (device < 0 || device >= 64) return fail(DRCVAR_ERR_INVALID, "device index %d out of range", device);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DeviceInfo& d = g_dev[device];
  if (!d.ready) {
    CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // the cluster path takes its redo list from the stream-ordered allocator: keep freed blocks in the pool instead of
    // handing them back to the driver at every synchronisation (a cudaMallocAsync per call would cost milliseconds)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    d.ready = true;
  }
  *out = &d;
  return DRCVAR_OK;
}

struct Call {
  const void* samples;
  long long B, N, stride_b, stride_n, stride_c;
  const double* ego;
  const double* h_in;
  double alpha, delta, epsilon, r_robot, r_obs;
  uint32_t flags;
  double *h_out, *h_mean_out, *g_out, *cvar_out, *var_out, *gstar_out;
  int32_t* status_out;
  int32_t* tail_idx_out;
  // generate mode (samples == nullptr): see drcvar_halfspaces_generated_f32
  const double* gen_mean = nullptr;
  const double* gen_chol = nullptr;
  uint64_t gen_seed = 0;
  long long gen_index_offset = 0;
  float* gen_samples_out = nullptr;
};

template <typename T>
int launch_on_device(const Call& c, int device, cudaStream_t stream) {
  using namespace drcvar;
  DeviceInfo* di = nullptr;
  int rc = device_info(device, &di);
  if (rc) return rc;
  double k_f;
  long long kc;
  if (!tail_count(c.alpha, c.N, &k_f, &kc)) return fail(DRCVAR_ERR_INVALID, "alpha must be in (0,1] and N >= 1");
  const size_t smem = slot_bytes_for(c.N, sizeof(T)) + fixed_smem_bytes(sizeof(T));
  // N beyond one CTA's shared memory (or on request): samples stay in global memory and are streamed several times
  const bool streaming = (c.flags & DRCVAR_FLAG_FORCE_STREAMING) || smem > static_cast<size_t>(di->max_smem_optin);
  if (c.B == 0) return DRCVAR_OK;

  KernelArgs a{};
  a.samples = c.samples;
  a.B = c.B;
  a.N = static_cast<int>(c.N);
  a.stride_b = c.stride_b;
  a.stride_n = c.stride_n;
  a.stride_c = c.stride_c;
  a.ego = c.ego;
  a.h_in = c.h_in;
  a.delta = c.delta;
  a.eoa = c.epsilon / c.alpha;
  a.R = c.r_robot + c.r_obs;
  a.k_f = k_f;
  a.kc = static_cast<int>(kc);
  a.use_window = 0;
  long long n_sigma = c.N;   // samples behind the kernel's second moments: all (fp32) / rows 0, 4, 8, ... of 256 (fp64)
  if (sizeof(T) == 8) {
    const long long rows = (c.N + 255) / 256, r4 = (rows + 3) / 4, last = (r4 - 1) * 4 * 256;
    n_sigma = (r4 - 1) * 256 + std::min<long long>(256, c.N - last);
  }
  const double cand_capacity = streaming ? 0.5 * kStreamCand * kStreamWarps : 0.6 * kCandCap * kSweepWarps;
  if (!(c.flags & DRCVAR_FLAG_GENERAL_ONLY))
    a.use_window = plan_window(c.N, kc, n_sigma, cand_capacity, &a.z_lo, &a.z_hi) ? 1 : 0;
  a.z_mid_f = static_cast<float>(0.5 * (a.z_lo + a.z_hi));
  a.z_half_f = static_cast<float>(0.5 * (a.z_hi - a.z_lo));
  a.z_lo_f = a.z_mid_f - a.z_half_f;
  a.z_hi_f = a.z_mid_f + a.z_half_f;
  // learned-centre mode (non-Gaussian samples): the Basu cancellation does not hold and the density at the quantile is
  // unknown -> 2.5 x the Gaussian half-width, capped by what the per-warp lists hold
  a.z_half_adapt_f = 2.5f * a.z_half_f;
  const size_t row_bytes = static_cast<size_t>(c.N) * 2 * sizeof(T);
  const bool contiguous = (c.stride_c == 1 && c.stride_n == 2);
  a.bulk = contiguous && !(c.flags & DRCVAR_FLAG_NO_BULK) && (reinterpret_cast<uintptr_t>(c.samples) % 16 == 0) &&
           ((static_cast<size_t>(c.stride_b) * sizeof(T)) % 16 == 0 || c.B == 1) && (row_bytes % 16 == 0) &&
           row_bytes < (1u << 20);
  a.h_out = c.h_out;
  a.h_mean_out = c.h_mean_out;
  a.g_out = c.g_out;
  a.cvar_out = c.cvar_out;
  a.var_out = c.var_out;
  a.gstar_out = c.gstar_out;
  a.status_out = c.status_out;
  a.tail_idx_out = c.tail_idx_out;
  a.gen_mean = c.gen_mean;
  a.gen_chol = c.gen_chol;
  a.gen_seed = c.gen_seed;
  a.gen_index_offset = c.gen_index_offset;
  a.gen_samples_out = c.gen_samples_out;
  if (c.gen_mean != nullptr) {
    if (sizeof(T) != 4) return fail(DRCVAR_ERR_UNSUPPORTED, "generate mode is fp32 only");
    a.bulk = 0;   // the sweep team fills the slot itself
  }

#ifdef DRCVAR_PROFILE_PHASES
  a.phase_cycles = g_phase_cycles;
  if (const char* e = getenv("DRCVAR_DEBUG_FLAGS")) a.debug = atoi(e);
#else
  a.phase_cycles = nullptr;
#endif
  const bool tail = c.tail_idx_out != nullptr;
  a.cl_ctas = 0;
  a.redo_list = nullptr;
  // (tail indices — parity mode — on the cluster path: fp32 stored samples only; every CTA emits the indices of its resident part)
  const bool cluster_tail_ok = !tail || (sizeof(T) == 4 && c.gen_mean == nullptr && c.N < 0x7fffffffLL / 2);
  if (streaming && !(c.flags & (DRCVAR_FLAG_FORCE_STREAMING | DRCVAR_FLAG_NO_CLUSTER | DRCVAR_FLAG_GENERAL_ONLY)) &&
      cluster_tail_ok && c.N > kOctantMinN && c.B < 0x7fffffffLL &&
      (sizeof(T) == 4 || (c.gen_mean == nullptr && (c.flags & DRCVAR_FLAG_FORCE_CLUSTER)))) {
    // (fp64 samples: the two-pass streaming kernel is faster — 1.96 vs 1.73 M halfspaces/s at N = 100 000: 1.6 MB per
    //  halfspace leaves no room for a second halfspace in flight per cluster — so the fp64 cluster kernel is opt-in)
    // ---- cluster / DSMEM kernel: one cluster of 2 / 4 / 8 CTAs per halfspace, every sample read from HBM once
    const bool gen_mode = c.gen_mean != nullptr;   // samples drawn by the kernel: no TMA, no alignment to ask for
    const bool bulk_ok = gen_mode ? (c.N % 2 == 0)
                                  : (contiguous && (reinterpret_cast<uintptr_t>(c.samples) % 16 == 0) &&
                                     ((static_cast<size_t>(c.stride_b) * sizeof(T)) % 16 == 0 || c.B == 1) &&
                                     (row_bytes % 16 == 0));
    const int ctas = cluster_ctas_for(c.N, sizeof(T), static_cast<size_t>(di->max_smem_optin));
    double zl = 0, zh = 0;
    // stored fp32 samples, no tail indices: raw-coordinate sweep B unless the caller's frame is far from the origin
    static const bool env_large_cl = getenv("DRCVAR_LARGE_COORDS") != nullptr;
    const bool raw_b = !(c.flags & DRCVAR_FLAG_LARGE_COORDS) && !env_large_cl;
    void (*ck)(const KernelArgs) = sizeof(T) == 4 ? (gen_mode ? cluster_kernel_f32<true>
                                                              : (tail ? cluster_kernel_f32<false, true>
                                                                      : (raw_b ? cluster_kernel_f32<false, false, true> : cluster_kernel_f32<false>)))
                                                  : cluster_kernel_f64;
    const long long cl_n_sigma = sizeof(T) == 4 ? c.N : std::max<long long>(1, c.N / 4);   // fp64: moments on every 4th row
    if (bulk_ok && ctas && plan_window(c.N, kc, cl_n_sigma, 0.7 * kClPool, &zl, &zh)) {
      KernelArgs ca = a;
      ca.use_window = 1;
      ca.z_lo = zl;
      ca.z_hi = zh;
      ca.z_mid_f = static_cast<float>(0.5 * (zl + zh));
      ca.z_half_f = static_cast<float>(0.5 * (zh - zl));
      ca.z_lo_f = ca.z_mid_f - ca.z_half_f;
      ca.z_hi_f = ca.z_mid_f + ca.z_half_f;
      ca.z_half_adapt_f = 1.6f * ca.z_half_f;   // learned-centre mode: what the leader's candidate pool still holds
      ca.bulk = gen_mode ? 0 : 1;
      ca.cl_ctas = ctas;
      const size_t csmem = cluster_smem_bytes(c.N, ctas, sizeof(T));
      CUDA_TRY(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(csmem)));
      int* redo = nullptr;   // redo flag of halfspace b at redo[b]; stream-ordered allocation, freed on every path below
      CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&redo), sizeof(int) * static_cast<size_t>(c.B), stream));
      auto cluster_pass = [&]() -> cudaError_t {   // returns cudaErrorNotReady when no cluster fits (fall through to streaming)
        cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(int) * static_cast<size_t>(c.B), stream);
        if (e != cudaSuccess) return e;
        ca.redo_list = redo;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(ctas);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(static_cast<unsigned>(ctas), 1, 1);
        cfg.blockDim = dim3(kClThreads, 1, 1);
        cfg.dynamicSmemBytes = csmem;
        cfg.stream = stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int max_clusters = 0;
        e = cudaOccupancyMaxActiveClusters(&max_clusters, ck, &cfg);
        if (e != cudaSuccess) return e;
#ifdef DRCVAR_PROFILE_PHASES
        fprintf(stderr, "[drcvar] cluster kernel: N=%lld ctas=%d smem=%zu max_active_clusters=%d\n", c.N, ctas, csmem, max_clusters);
#endif
        if (max_clusters < 1) return cudaErrorNotReady;
        const long long n_cl = std::min<long long>(c.B, max_clusters);
        cfg.gridDim = dim3(static_cast<unsigned>(n_cl * ctas), 1, 1);
        e = cudaLaunchKernelEx(&cfg, ck, ca);
        if (e != cudaSuccess) return e;
        g_launches.fetch_add(1);
        // the halfspaces it handed back (window miss, overflow, non-finite data): exact general select, streaming kernel
        KernelArgs ra = a;   // the streaming kernel's own window plan; a CTA that misses twice in a row learns the centre
        ra.redo_list = redo;
        ra.bulk = 1;
        auto rk = tail ? streaming_kernel<T, true> : streaming_kernel<T, false>;   // (parity mode: the redo pass writes the indices too)
        if constexpr (sizeof(T) == 4) {
          if (gen_mode) rk = streaming_kernel<float, false, true>;
        }
        const long long rgrid = std::min<long long>(c.B, di->sms);
        rk<<<static_cast<unsigned>(rgrid), kStreamThreads, 0, stream>>>(ra);
        e = cudaGetLastError();
        if (e == cudaSuccess) g_launches.fetch_add(1);
        return e;
      };
      const cudaError_t ce = cluster_pass();
      const cudaError_t fe = cudaFreeAsync(redo, stream);
      if (ce == cudaSuccess && fe == cudaSuccess) return DRCVAR_OK;
      if (ce != cudaErrorNotReady)
        return fail(DRCVAR_ERR_CUDA, "cluster kernel path failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : fe));
      cudaGetLastError();
    }
  }
  if (streaming) {
    auto sk = tail ? streaming_kernel<T, true> : streaming_kernel<T, false>;
    if constexpr (sizeof(T) == 4) {
      if (c.gen_mean != nullptr) sk = tail ? streaming_kernel<float, true, true> : streaming_kernel<float, false, true>;
    }
    a.bulk = contiguous && (reinterpret_cast<uintptr_t>(c.samples) % 16 == 0) &&
             ((static_cast<size_t>(c.stride_b) * sizeof(T)) % 16 == 0 || c.B == 1);   // 16-byte vector loads
    int s_per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_per_sm, sk, kStreamThreads, 0));
    if (s_per_sm < 1) return fail(DRCVAR_ERR_UNSUPPORTED, "streaming kernel does not fit on an SM");
    const long long sgrid = std::min<long long>(c.B, static_cast<long long>(di->sms) * s_per_sm);
    sk<<<static_cast<unsigned>(sgrid), kStreamThreads, 0, stream>>>(a);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return DRCVAR_OK;
  }
  // ---- pipelined resident kernel (pipelined_kernel.cuh): contiguous samples, window planned, no tail indices.  Halfspaces
  // it hands back (window miss / not placeable / overflow) are computed by the streaming kernel's redo pass on the same stream.
  {
    static const bool env_off = getenv("DRCVAR_NO_PIPELINE") != nullptr;
    static const int env_f64 = getenv("DRCVAR_PIPELINE_F64") ? atoi(getenv("DRCVAR_PIPELINE_F64")) : DRCVAR_PIPELINE_F64_DEFAULT;
    const long long per_load = 16 / (2 * sizeof(T));
    // returns DRCVAR_OK after launching, a negative error, or 1 when the path does not apply (fall through)
    auto run_pipelined = [&](auto pk, int threads, size_t psmem, long long row_samples, const KernelArgs& base, int words = kMaskWords) -> int {
      const long long rows_all = (c.N + row_samples - 1) / row_samples;
      if (rows_all * per_load > 32 * words || psmem > static_cast<size_t>(di->max_smem_optin)) return 1;
      CUDA_TRY(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(psmem)));
      int p_per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p_per_sm, pk, threads, psmem));
      if (p_per_sm < 1) return 1;
      int* redo = nullptr;   // redo flag of halfspace b at redo[b]; stream-ordered allocation, freed on every path below
      CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&redo), sizeof(int) * static_cast<size_t>(c.B), stream));
      auto pass = [&]() -> cudaError_t {
        cudaError_t e = cudaMemsetAsync(redo, 0, sizeof(int) * static_cast<size_t>(c.B), stream);
        if (e != cudaSuccess) return e;
        KernelArgs pa = base;
        pa.redo_list = redo;
        const long long pgrid = std::min<long long>(c.B, static_cast<long long>(p_per_sm) * di->sms);
        pk<<<static_cast<unsigned>(pgrid), threads, psmem, stream>>>(pa);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        g_launches.fetch_add(1);
        KernelArgs ra = a;   // the streaming kernel places its own windows (and learns them after two misses in a row)
        ra.redo_list = redo;
        ra.bulk = 1;
        {
          // most hand-backs are statistical misses of the planned window: the second attempt takes twice the width, so that it
          // finds the threshold among its candidates instead of falling through to the multi-pass general select
          // (config 4: 55 -> 31 us per redo pass; speed only, the threshold is exact either way)
          const double mid = 0.5 * (ra.z_lo + ra.z_hi), half = ra.z_hi - ra.z_lo;
          ra.z_lo = mid - half;
          ra.z_hi = mid + half;
          ra.z_mid_f = static_cast<float>(mid);
          ra.z_half_f = static_cast<float>(half);
          ra.z_lo_f = ra.z_mid_f - ra.z_half_f;
          ra.z_hi_f = ra.z_mid_f + ra.z_half_f;
        }
        const long long rgrid = std::min<long long>(c.B, di->sms);
        streaming_kernel<T, false><<<static_cast<unsigned>(rgrid), kStreamThreads, 0, stream>>>(ra);
        e = cudaGetLastError();
        if (e == cudaSuccess) g_launches.fetch_add(1);
        return e;
      };
      const cudaError_t pe = pass();
      const cudaError_t fe = cudaFreeAsync(redo, stream);
      if (pe != cudaSuccess || fe != cudaSuccess)
        return fail(DRCVAR_ERR_CUDA, "pipelined kernel path failed: %s", cudaGetErrorString(pe != cudaSuccess ? pe : fe));
      return DRCVAR_OK;
    };
    const bool applies = !tail && c.gen_mean == nullptr && a.bulk && a.use_window && !(c.flags & DRCVAR_FLAG_NO_PIPELINE) && !env_off &&
                         c.B < 0x7fffffffLL;
    if constexpr (sizeof(T) == 4) {
      if (applies) {
        // raw-coordinate sweep B unless the caller says the coordinates are far from the origin (or DRCVAR_LARGE_COORDS=1):
        // halfspaces too large for raw fp32 sums are handed to the redo pass, so the flag only ever changes the speed
        static const bool env_large = getenv("DRCVAR_LARGE_COORDS") != nullptr;
        const bool large = (c.flags & DRCVAR_FLAG_LARGE_COORDS) || env_large;
        // N <= 16 384 (64 samples per thread): the raw instantiation with two mask words per thread (less code between the hot loops)
        const bool two_words = !large && ((c.N + 511) / 512) * 2 <= 64;
        const int r = large ? run_pipelined(pipelined_kernel<float, 8, false>, pipelined_threads<8>(),
                                            slot_bytes_for(c.N, 4) + pipelined_fixed_smem_bytes<8>(4), 512, a)
                      : two_words ? run_pipelined(pipelined_kernel<float, 8, true, 2>, pipelined_threads<8>(),
                                                  slot_bytes_for(c.N, 4) + pipelined_fixed_smem_bytes<8>(4), 512, a, 2)
                                  : run_pipelined(pipelined_kernel<float, 8, true>, pipelined_threads<8>(),
                                                  slot_bytes_for(c.N, 4) + pipelined_fixed_smem_bytes<8>(4), 512, a);
        if (r <= 0) return r;
      }
    } else {
      // fp64 samples: 16 sweep warps + placer warp in the one CTA per SM the 160 KB slot allows (DRCVAR_PIPELINE_F64=16);
      // =8 runs the two-CTA layout of the fp32 path (measured slower than halfspace_kernel<double>), 0 = halfspace_kernel<double>
      if (applies && env_f64 == 16) {
        KernelArgs pa = a;   // second moments on rows 0, 4, 8, ... of 512 samples: re-plan the window for that subsample
        const long long rows = (c.N + 511) / 512, r4 = (rows + 3) / 4, last = (r4 - 1) * 4 * 512;
        const long long ns16 = (r4 - 1) * 512 + std::min<long long>(512, c.N - last);
        if (plan_window(c.N, kc, ns16, 0.6 * PCaps<16>::kCandCap * 16, &pa.z_lo, &pa.z_hi)) {
          pa.z_mid_f = static_cast<float>(0.5 * (pa.z_lo + pa.z_hi));
          pa.z_half_f = static_cast<float>(0.5 * (pa.z_hi - pa.z_lo));
          pa.z_lo_f = pa.z_mid_f - pa.z_half_f;
          pa.z_hi_f = pa.z_mid_f + pa.z_half_f;
          pa.z_half_adapt_f = 2.5f * pa.z_half_f;
          const int r = run_pipelined(pipelined_kernel<double, 16>, pipelined_threads<16>(),
                                      slot_bytes_for(c.N, 8) + pipelined_fixed_smem_bytes<16>(8), 512, pa);
          if (r <= 0) return r;
        }
      } else if (applies && env_f64 == 8) {
        const int r = run_pipelined(pipelined_kernel<double, 8>, pipelined_threads<8>(),
                                    slot_bytes_for(c.N, 8) + pipelined_fixed_smem_bytes<8>(8), 256, a);
        if (r <= 0) return r;
      }
    }
  }
  auto kern = tail ? halfspace_kernel<T, true> : halfspace_kernel<T, false>;
  if constexpr (sizeof(T) == 4) {
    if (c.gen_mean != nullptr) kern = tail ? halfspace_kernel<float, true, true> : halfspace_kernel<float, false, true>;
  }
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) return fail(DRCVAR_ERR_UNSUPPORTED, "kernel does not fit on an SM for N=%lld", c.N);
  long long grid = std::min<long long>(c.B, static_cast<long long>(per_sm) * di->sms);
#ifdef DRCVAR_PROFILE_PHASES
  if (const char* e = getenv("DRCVAR_DEBUG_GRID")) grid = std::min<long long>(grid, atoll(e));
#endif
  kern<<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return DRCVAR_OK;
}

int check_common(const Call& c) {
  if (c.B < 0) return fail(DRCVAR_ERR_INVALID, "B must be >= 0");
  if (c.N < 1) return fail(DRCVAR_ERR_INVALID, "N must be >= 1");
  if (c.N > 0x7fffff00LL) return fail(DRCVAR_ERR_UNSUPPORTED, "N too large");
  if (!(c.alpha > 0.0) || !(c.alpha <= 1.0)) return fail(DRCVAR_ERR_INVALID, "alpha must be in (0, 1]");
  if (c.B > 0 && ((!c.samples && !c.gen_mean) || !c.h_out || !c.g_out))
    return fail(DRCVAR_ERR_INVALID, "samples, h_out and g_out must be non-null");
  if (c.gen_mean && !c.gen_chol) return fail(DRCVAR_ERR_INVALID, "generate mode needs mean and chol");
  if (c.stride_n < 0 || c.stride_c < 0 || c.stride_b < 0) return fail(DRCVAR_ERR_INVALID, "negative strides are not supported");
  return DRCVAR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// DRCVAR_HOST path: chunked, double-buffered staging through pinned memory, kernels overlapped with copies.
struct HostCtx {
  std::mutex mu;
  int device = -1;
  cudaStream_t streams[2] = {nullptr, nullptr};
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  void* d_samples[2] = {nullptr, nullptr};
  size_t d_samples_cap[2] = {0, 0};
  void* h_pack[2] = {nullptr, nullptr};  // pinned, for strided host inputs
  size_t h_pack_cap[2] = {0, 0};
  unsigned char* d_io[2] = {nullptr, nullptr};  // ego/h_in + outputs
  size_t d_io_cap[2] = {0, 0};
  void* h_out_stage[2] = {nullptr, nullptr};    // pinned: the whole output block of a chunk comes back in ONE copy
  size_t h_out_cap[2] = {0, 0};
  unsigned char* h_small = nullptr;   // pinned: inputs | outputs of a small call, one copy each way
  unsigned char* d_small = nullptr;
  size_t small_cap = 0;
  double last_stage_ms = 0, last_kernel_ms = 0;
  long long last_h2d = 0, last_d2h = 0;
};
constexpr size_t kSmallCallBytes = static_cast<size_t>(1) << 20;   // host calls up to this many input bytes take run_host_small
// One context per device: a process that drives several GPUs through DRCVAR_HOST (cudaSetDevice between calls, or one
// thread per GPU) keeps every device's streams, events and staging buffers; nothing is torn down on a device switch and
// calls on different devices do not serialise on each other.
HostCtx g_host[64];

int host_ctx(HostCtx** out) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(DRCVAR_ERR_INVALID, "device index %d out of range", dev);
  *out = &g_host[dev];
  return DRCVAR_OK;
}

// (called with hc.mu held, on the device the context belongs to)
int ensure_host_ctx(HostCtx& hc) {
  if (hc.device < 0) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&hc.streams[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&hc.ev_start));
    CUDA_TRY(cudaEventCreate(&hc.ev_stop));
    hc.device = dev;
  }
  return DRCVAR_OK;
}

int grow_dev(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return DRCVAR_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = std::max(need, static_cast<size_t>(1) << 16);
  CUDA_TRY(cudaMalloc(p, want));
  *cap = want;
  return DRCVAR_OK;
}
int grow_pinned(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return DRCVAR_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = std::max(need, static_cast<size_t>(1) << 16);
  CUDA_TRY(cudaMallocHost(p, want));
  *cap = want;
  return DRCVAR_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Small host calls (the reference's own use: one halfspace per call, evaluation/timing_analysis.py:74,101, or one
// trajectory, simulation/environment.py:60-106): everything the kernel reads is packed into ONE pinned block and goes over
// in one copy, everything it writes comes back in one copy, one stream synchronisation, no events — the call is pure
// latency, and every CUDA API call in it costs more than the arithmetic.
template <typename T>
int run_host_small(const Call& c, HostCtx& hc, long long kc) {
  const size_t dbl = sizeof(double);
  const size_t nb = static_cast<size_t>(c.B);
  const size_t row_pitch = align_up(static_cast<size_t>(c.N) * 2 * sizeof(T), 16);
  const size_t off_s = 0, off_ego = off_s + nb * row_pitch, off_hin = off_ego + nb * 2 * dbl, in_bytes = off_hin + nb * 2 * dbl;
  const size_t off_h = align_up(in_bytes, 16), off_hm = off_h + nb * 2 * dbl, off_g = off_hm + nb * 2 * dbl,
               off_cvar = off_g + nb * 3 * dbl, off_var = off_cvar + nb * dbl, off_gs = off_var + nb * dbl,
               off_st = off_gs + nb * dbl, off_tail = align_up(off_st + nb * sizeof(int32_t), 16),
               total = off_tail + (c.tail_idx_out ? nb * static_cast<size_t>(kc) * sizeof(int32_t) : 0);
  if (hc.small_cap < total) {
    if (hc.h_small) cudaFreeHost(hc.h_small);
    if (hc.d_small) cudaFree(hc.d_small);
    hc.h_small = hc.d_small = nullptr;
    hc.small_cap = 0;
    const size_t want = std::max(total, static_cast<size_t>(1) << 18);
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&hc.h_small), want));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&hc.d_small), want));
    hc.small_cap = want;
  }
  unsigned char* hb = hc.h_small;
  unsigned char* db = hc.d_small;
  const bool rows_contig = (c.stride_c == 1 && c.stride_n == 2);
  for (size_t b = 0; b < nb; ++b) {   // (strided views, e.g. traj[:, t, :] of simulation/environment.py:88, are packed here)
    T* row = reinterpret_cast<T*>(hb + off_s + b * row_pitch);
    const T* sb = reinterpret_cast<const T*>(c.samples) + static_cast<long long>(b) * c.stride_b;
    if (rows_contig) {
      std::memcpy(row, sb, static_cast<size_t>(c.N) * 2 * sizeof(T));
    } else {
      for (long long i = 0; i < c.N; ++i) {
        row[2 * i] = sb[i * c.stride_n];
        row[2 * i + 1] = sb[i * c.stride_n + c.stride_c];
      }
    }
  }
  if (c.ego) std::memcpy(hb + off_ego, c.ego, nb * 2 * dbl);
  if (c.h_in) std::memcpy(hb + off_hin, c.h_in, nb * 2 * dbl);
  cudaStream_t st = hc.streams[0];
  CUDA_TRY(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, st));
  Call d = c;
  d.samples = db + off_s;
  d.stride_b = static_cast<long long>(row_pitch / sizeof(T));
  d.stride_n = 2;
  d.stride_c = 1;
  d.ego = c.ego ? reinterpret_cast<const double*>(db + off_ego) : nullptr;
  d.h_in = c.h_in ? reinterpret_cast<const double*>(db + off_hin) : nullptr;
  d.h_out = reinterpret_cast<double*>(db + off_h);
  d.h_mean_out = reinterpret_cast<double*>(db + off_hm);
  d.g_out = reinterpret_cast<double*>(db + off_g);
  d.cvar_out = reinterpret_cast<double*>(db + off_cvar);
  d.var_out = reinterpret_cast<double*>(db + off_var);
  d.gstar_out = reinterpret_cast<double*>(db + off_gs);
  d.status_out = reinterpret_cast<int32_t*>(db + off_st);
  d.tail_idx_out = c.tail_idx_out ? reinterpret_cast<int32_t*>(db + off_tail) : nullptr;
  const int rc = launch_on_device<T>(d, hc.device, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(hb + off_h, db + off_h, total - off_h, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  hc.last_h2d = static_cast<long long>(in_bytes);
  hc.last_d2h = static_cast<long long>(total - off_h);
  auto put = [&](void* host, size_t off, size_t bytes) {
    if (host) std::memcpy(host, hb + off, bytes);
  };
  put(c.h_out, off_h, nb * 2 * dbl);
  put(c.h_mean_out, off_hm, nb * 2 * dbl);
  put(c.g_out, off_g, nb * 3 * dbl);
  put(c.cvar_out, off_cvar, nb * dbl);
  put(c.var_out, off_var, nb * dbl);
  put(c.gstar_out, off_gs, nb * dbl);
  put(c.status_out, off_st, nb * sizeof(int32_t));
  put(c.tail_idx_out, off_tail, nb * static_cast<size_t>(kc) * sizeof(int32_t));
  return DRCVAR_OK;
}

template <typename T>
int run_host(const Call& c) {
  HostCtx* hcp = nullptr;
  int rc = host_ctx(&hcp);
  if (rc) return rc;
  HostCtx& hc = *hcp;
  std::lock_guard<std::mutex> lk(hc.mu);
  rc = ensure_host_ctx(hc);
  if (rc) return rc;
  double k_f;
  long long kc;
  if (!tail_count(c.alpha, c.N, &k_f, &kc)) return fail(DRCVAR_ERR_INVALID, "alpha must be in (0,1] and N >= 1");
  hc.last_stage_ms = hc.last_kernel_ms = 0;
  hc.last_h2d = hc.last_d2h = 0;
  if (c.B == 0) return DRCVAR_OK;
  if (static_cast<size_t>(c.B) * static_cast<size_t>(c.N) * 2 * sizeof(T) <= kSmallCallBytes &&
      (!c.tail_idx_out || static_cast<size_t>(c.B) * static_cast<size_t>(kc) * 4 <= kSmallCallBytes))
    return run_host_small<T>(c, hc, kc);

  const size_t row_bytes = static_cast<size_t>(c.N) * 2 * sizeof(T);
  const size_t row_pitch = align_up(row_bytes, 16);  // device rows are 16-B aligned so the bulk loader applies
  const bool contiguous = (c.stride_c == 1 && c.stride_n == 2 && c.stride_b * sizeof(T) == row_bytes && row_pitch == row_bytes);
  const size_t target_chunk = static_cast<size_t>(64) << 20;
  long long chunk_b = static_cast<long long>(std::max<size_t>(1, target_chunk / row_pitch));
  chunk_b = std::min(chunk_b, c.B);
  const long long n_chunks = (c.B + chunk_b - 1) / chunk_b;

  // is the caller's sample array page-locked?  (numpy arrays are not; drcvar_host_alloc / cudaHostRegister memory is)
  bool pageable = true;
  {
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, c.samples) == cudaSuccess) pageable = pa.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
  }
  static const int stage_threads = [] {
    const char* e = getenv("DRCVAR_STAGE_THREADS");
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (e) return std::max(1, atoi(e));
    const char* lw = getenv("LOCAL_WORLD_SIZE");   // torchrun: one process per GPU shares the host's cores
    const int ranks = lw ? std::max(1, atoi(lw)) : 1;
    return std::max(1, std::min(16, hw / ranks));   // 16 threads: 35 GB/s of staging on a 16-core box (profiles/r2_pageable_e2e.txt)
  }();
  struct Pending {
    bool live;
    long long b0, nb;
    size_t off_h, off_hm, off_g, off_cvar, off_var, off_gs, off_st, off_tail;
  } pending[2] = {};
  auto drain = [&](int s) {   // scatter the staged output block of slot s into the caller's arrays
    Pending& q = pending[s];
    if (!q.live) return;
    const unsigned char* base = static_cast<const unsigned char*>(hc.h_out_stage[s]) - q.off_h;
    const size_t dbl = sizeof(double);
    auto put = [&](void* host, size_t off, size_t bytes) {
      if (host) std::memcpy(host, base + off, bytes);
    };
    put(c.h_out + 2 * q.b0, q.off_h, q.nb * 2 * dbl);
    put(c.h_mean_out ? c.h_mean_out + 2 * q.b0 : nullptr, q.off_hm, q.nb * 2 * dbl);
    put(c.g_out + 3 * q.b0, q.off_g, q.nb * 3 * dbl);
    put(c.cvar_out ? c.cvar_out + q.b0 : nullptr, q.off_cvar, q.nb * dbl);
    put(c.var_out ? c.var_out + q.b0 : nullptr, q.off_var, q.nb * dbl);
    put(c.gstar_out ? c.gstar_out + q.b0 : nullptr, q.off_gs, q.nb * dbl);
    put(c.status_out ? c.status_out + q.b0 : nullptr, q.off_st, q.nb * sizeof(int32_t));
    put(c.tail_idx_out ? c.tail_idx_out + q.b0 * kc : nullptr, q.off_tail, static_cast<size_t>(q.nb) * kc * sizeof(int32_t));
    q.live = false;
  };

  CUDA_TRY(cudaEventRecord(hc.ev_start, hc.streams[0]));
  for (long long ci = 0; ci < n_chunks; ++ci) {
    const int s = static_cast<int>(ci & 1);
    cudaStream_t st = hc.streams[s];
    const long long b0 = ci * chunk_b, nb = std::min(chunk_b, c.B - b0);
    // buffers of this slot are reused: wait for the chunk that used them two iterations ago and hand its results over
    if (ci >= 2) {
      CUDA_TRY(cudaStreamSynchronize(st));
      drain(s);
    }
    rc = grow_dev(&hc.d_samples[s], &hc.d_samples_cap[s], static_cast<size_t>(nb) * row_pitch);
    if (rc) return rc;
    // io block layout: ego[nb,2] h_in[nb,2] | h[nb,2] hm[nb,2] g[nb,3] cvar[nb] var[nb] gstar[nb] status[nb] tail[nb,kc]
    const size_t dbl = sizeof(double);
    const size_t off_ego = 0, off_hin = off_ego + nb * 2 * dbl, off_h = off_hin + nb * 2 * dbl,
                 off_hm = off_h + nb * 2 * dbl, off_g = off_hm + nb * 2 * dbl, off_cvar = off_g + nb * 3 * dbl,
                 off_var = off_cvar + nb * dbl, off_gs = off_var + nb * dbl, off_st = off_gs + nb * dbl,
                 off_tail = align_up(off_st + nb * sizeof(int32_t), 16),
                 io_bytes = off_tail + (c.tail_idx_out ? static_cast<size_t>(nb) * kc * sizeof(int32_t) : 0);
    rc = grow_dev(reinterpret_cast<void**>(&hc.d_io[s]), &hc.d_io_cap[s], io_bytes);
    if (rc) return rc;
    unsigned char* io = hc.d_io[s];

    // ---- samples H2D
    const T* src = reinterpret_cast<const T*>(c.samples) + b0 * c.stride_b;
    if (contiguous && !pageable) {
      CUDA_TRY(cudaMemcpyAsync(hc.d_samples[s], src, static_cast<size_t>(nb) * row_bytes, cudaMemcpyHostToDevice, st));
    } else {
      // pageable numpy arrays (what the reference's callers hand over) and strided host views (traj[:, t, :],
      // simulation/environment.py:88) are staged into this slot's pinned buffer by a few host threads; the copy of chunk
      // ci then runs by DMA while the threads stage chunk ci + 1.  (A pageable cudaMemcpyAsync is staged by the driver on
      // ONE thread at ~11 GB/s and blocks the caller meanwhile.)
      rc = grow_pinned(&hc.h_pack[s], &hc.h_pack_cap[s], static_cast<size_t>(nb) * row_pitch);
      if (rc) return rc;
      unsigned char* dst = static_cast<unsigned char*>(hc.h_pack[s]);
      const Call* cp = &c;
      auto stage_rows = [=](long long lo, long long hi) {
        if (contiguous) {
          std::memcpy(dst + static_cast<size_t>(lo) * row_pitch, reinterpret_cast<const unsigned char*>(src) + static_cast<size_t>(lo) * row_bytes,
                      static_cast<size_t>(hi - lo) * row_bytes);
          return;
        }
        for (long long b = lo; b < hi; ++b) {
          T* row = reinterpret_cast<T*>(dst + static_cast<size_t>(b) * row_pitch);
          const T* sb = src + b * cp->stride_b;
          for (long long i = 0; i < cp->N; ++i) {
            row[2 * i] = sb[i * cp->stride_n];
            row[2 * i + 1] = sb[i * cp->stride_n + cp->stride_c];
          }
        }
      };
      const size_t chunk_bytes = static_cast<size_t>(nb) * row_bytes;
      const long long n_thr = std::max<long long>(1, std::min<long long>({static_cast<long long>(stage_threads), nb,
                                                                         static_cast<long long>(chunk_bytes >> 21)}));
      if (n_thr == 1) {
        stage_rows(0, nb);
      } else {
        std::vector<std::thread> pool;
        for (long long k = 1; k < n_thr; ++k) pool.emplace_back(stage_rows, nb * k / n_thr, nb * (k + 1) / n_thr);
        stage_rows(0, nb / n_thr);
        for (auto& th : pool) th.join();
      }
      CUDA_TRY(cudaMemcpyAsync(hc.d_samples[s], dst, static_cast<size_t>(nb) * row_pitch, cudaMemcpyHostToDevice, st));
    }
    hc.last_h2d += static_cast<long long>(nb) * static_cast<long long>(row_bytes);
    if (c.ego) {
      CUDA_TRY(cudaMemcpyAsync(io + off_ego, c.ego + 2 * b0, nb * 2 * dbl, cudaMemcpyHostToDevice, st));
      hc.last_h2d += nb * 2 * dbl;
    }
    if (c.h_in) {
      CUDA_TRY(cudaMemcpyAsync(io + off_hin, c.h_in + 2 * b0, nb * 2 * dbl, cudaMemcpyHostToDevice, st));
      hc.last_h2d += nb * 2 * dbl;
    }

    Call d = c;
    d.samples = hc.d_samples[s];
    d.B = nb;
    d.stride_b = static_cast<long long>(row_pitch / sizeof(T));
    d.stride_n = 2;
    d.stride_c = 1;
    d.ego = c.ego ? reinterpret_cast<const double*>(io + off_ego) : nullptr;
    d.h_in = c.h_in ? reinterpret_cast<const double*>(io + off_hin) : nullptr;
    d.h_out = reinterpret_cast<double*>(io + off_h);
    d.h_mean_out = reinterpret_cast<double*>(io + off_hm);
    d.g_out = reinterpret_cast<double*>(io + off_g);
    d.cvar_out = reinterpret_cast<double*>(io + off_cvar);
    d.var_out = reinterpret_cast<double*>(io + off_var);
    d.gstar_out = reinterpret_cast<double*>(io + off_gs);
    d.status_out = reinterpret_cast<int32_t*>(io + off_st);
    d.tail_idx_out = c.tail_idx_out ? reinterpret_cast<int32_t*>(io + off_tail) : nullptr;
    rc = launch_on_device<T>(d, hc.device, st);
    if (rc) return rc;

    // ---- results D2H: one copy of the contiguous output block into pinned memory, scattered to the caller's arrays
    //      once the chunk's stream has been synchronised (drain)
    const size_t out_bytes = io_bytes - off_h;
    rc = grow_pinned(&hc.h_out_stage[s], &hc.h_out_cap[s], out_bytes);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(hc.h_out_stage[s], io + off_h, out_bytes, cudaMemcpyDeviceToHost, st));
    hc.last_d2h += static_cast<long long>(out_bytes);
    pending[s] = {true, b0, nb, off_h, off_hm, off_g, off_cvar, off_var, off_gs, off_st, off_tail};
  }
  CUDA_TRY(cudaStreamSynchronize(hc.streams[0]));
  CUDA_TRY(cudaStreamSynchronize(hc.streams[1]));
  drain(0);
  drain(1);
  CUDA_TRY(cudaEventRecord(hc.ev_stop, hc.streams[0]));
  CUDA_TRY(cudaEventSynchronize(hc.ev_stop));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, hc.ev_start, hc.ev_stop));
  hc.last_kernel_ms = ms;  // device-side span of the whole pipelined call (copies + kernels)
  return DRCVAR_OK;
}

// DRCVAR_HOST path of the generate mode: the inputs are tiny (mean, chol, ego per halfspace), one stream, no chunking.
int run_host_generated(const Call& c) {
  HostCtx* hcp = nullptr;
  int rc = host_ctx(&hcp);
  if (rc) return rc;
  HostCtx& hc = *hcp;
  std::lock_guard<std::mutex> lk(hc.mu);
  rc = ensure_host_ctx(hc);
  if (rc) return rc;
  double k_f;
  long long kc;
  if (!tail_count(c.alpha, c.N, &k_f, &kc)) return fail(DRCVAR_ERR_INVALID, "alpha must be in (0,1] and N >= 1");
  hc.last_stage_ms = hc.last_kernel_ms = 0;
  hc.last_h2d = hc.last_d2h = 0;
  if (c.B == 0) return DRCVAR_OK;
  cudaStream_t st = hc.streams[0];
  const size_t nb = static_cast<size_t>(c.B), dbl = sizeof(double);
  const size_t off_mean = 0, off_chol = off_mean + nb * 2 * dbl, off_ego = off_chol + nb * 3 * dbl,
               off_hin = off_ego + nb * 2 * dbl, off_h = off_hin + nb * 2 * dbl, off_hm = off_h + nb * 2 * dbl,
               off_g = off_hm + nb * 2 * dbl, off_cvar = off_g + nb * 3 * dbl, off_var = off_cvar + nb * dbl,
               off_gs = off_var + nb * dbl, off_st = off_gs + nb * dbl,
               off_tail = align_up(off_st + nb * sizeof(int32_t), 16),
               io_bytes = off_tail + (c.tail_idx_out ? nb * kc * sizeof(int32_t) : 0);
  rc = grow_dev(reinterpret_cast<void**>(&hc.d_io[0]), &hc.d_io_cap[0], io_bytes);
  if (rc) return rc;
  unsigned char* io = hc.d_io[0];
  const size_t dump_bytes = c.gen_samples_out ? nb * static_cast<size_t>(c.N) * 2 * sizeof(float) : 0;
  if (dump_bytes) {
    rc = grow_dev(&hc.d_samples[0], &hc.d_samples_cap[0], dump_bytes);
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(hc.ev_start, st));
  auto up = [&](size_t off, const double* host, size_t bytes) -> cudaError_t {
    if (!host) return cudaSuccess;
    hc.last_h2d += static_cast<long long>(bytes);
    return cudaMemcpyAsync(io + off, host, bytes, cudaMemcpyHostToDevice, st);
  };
  CUDA_TRY(up(off_mean, c.gen_mean, nb * 2 * dbl));
  CUDA_TRY(up(off_chol, c.gen_chol, nb * 3 * dbl));
  CUDA_TRY(up(off_ego, c.ego, nb * 2 * dbl));
  CUDA_TRY(up(off_hin, c.h_in, nb * 2 * dbl));
  Call d = c;
  d.gen_mean = reinterpret_cast<const double*>(io + off_mean);
  d.gen_chol = reinterpret_cast<const double*>(io + off_chol);
  d.ego = c.ego ? reinterpret_cast<const double*>(io + off_ego) : nullptr;
  d.h_in = c.h_in ? reinterpret_cast<const double*>(io + off_hin) : nullptr;
  d.h_out = reinterpret_cast<double*>(io + off_h);
  d.h_mean_out = reinterpret_cast<double*>(io + off_hm);
  d.g_out = reinterpret_cast<double*>(io + off_g);
  d.cvar_out = reinterpret_cast<double*>(io + off_cvar);
  d.var_out = reinterpret_cast<double*>(io + off_var);
  d.gstar_out = reinterpret_cast<double*>(io + off_gs);
  d.status_out = reinterpret_cast<int32_t*>(io + off_st);
  d.tail_idx_out = c.tail_idx_out ? reinterpret_cast<int32_t*>(io + off_tail) : nullptr;
  d.gen_samples_out = dump_bytes ? static_cast<float*>(hc.d_samples[0]) : nullptr;
  rc = launch_on_device<float>(d, hc.device, st);
  if (rc) return rc;
  const size_t out_bytes = io_bytes - off_h;
  rc = grow_pinned(&hc.h_out_stage[0], &hc.h_out_cap[0], out_bytes);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(hc.h_out_stage[0], io + off_h, out_bytes, cudaMemcpyDeviceToHost, st));
  hc.last_d2h += static_cast<long long>(out_bytes);
  if (dump_bytes) {
    CUDA_TRY(cudaMemcpyAsync(c.gen_samples_out, hc.d_samples[0], dump_bytes, cudaMemcpyDeviceToHost, st));
    hc.last_d2h += static_cast<long long>(dump_bytes);
  }
  CUDA_TRY(cudaEventRecord(hc.ev_stop, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const unsigned char* base = static_cast<const unsigned char*>(hc.h_out_stage[0]) - off_h;
  auto put = [&](void* host, size_t off, size_t bytes) {
    if (host) std::memcpy(host, base + off, bytes);
  };
  put(c.h_out, off_h, nb * 2 * dbl);
  put(c.h_mean_out, off_hm, nb * 2 * dbl);
  put(c.g_out, off_g, nb * 3 * dbl);
  put(c.cvar_out, off_cvar, nb * dbl);
  put(c.var_out, off_var, nb * dbl);
  put(c.gstar_out, off_gs, nb * dbl);
  put(c.status_out, off_st, nb * sizeof(int32_t));
  put(c.tail_idx_out, off_tail, nb * kc * sizeof(int32_t));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, hc.ev_start, hc.ev_stop));
  hc.last_kernel_ms = ms;
  return DRCVAR_OK;
}

template <typename T>
int entry(const T* samples, int64_t B, int64_t N, int64_t stride_b, int64_t stride_n, int64_t stride_c,
          const double* ego, const double* h_in, double alpha, double delta, double epsilon, double r_robot,
          double r_obs, uint32_t flags, double* h_out, double* h_mean_out, double* g_out, double* cvar_out,
          double* var_out, double* gstar_out, int32_t* status_out, int32_t* tail_idx_out, int device, void* stream) {
  Call c{samples, B, N, stride_b, stride_n, stride_c, ego, h_in, alpha, delta, epsilon, r_robot, r_obs, flags,
         h_out, h_mean_out, g_out, cvar_out, var_out, gstar_out, status_out, tail_idx_out};
  int rc = check_common(c);
  if (rc) return rc;
  if (device == DRCVAR_HOST) return run_host<T>(c);
  int prev = 0;
  CUDA_TRY(cudaGetDevice(&prev));
  if (prev != device) CUDA_TRY(cudaSetDevice(device));
  rc = launch_on_device<T>(c, device, static_cast<cudaStream_t>(stream));
  if (rc == DRCVAR_OK && (flags & DRCVAR_FLAG_SYNC)) {
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) rc = fail(DRCVAR_ERR_CUDA, "stream synchronize failed: %s", cudaGetErrorString(e));
  }
  if (prev != device) cudaSetDevice(prev);
  return rc;
}

}  // namespace

extern "C" {

int drcvar_version(void) { return DRCVAR_ABI_VERSION; }
const char* drcvar_last_error(void) { return g_err; }
int drcvar_reduction_lanes(void) { return drcvar::kSlots; }

int drcvar_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t drcvar_tail_count(double alpha, int64_t n_samples, double* k_f_out) {
  double k_f;
  long long kc;
  if (!tail_count(alpha, n_samples, &k_f, &kc)) return fail(DRCVAR_ERR_INVALID, "alpha must be in (0,1] and N >= 1");
  if (k_f_out) *k_f_out = k_f;
  return kc;
}

int64_t drcvar_max_samples(int elem_bytes, int device) {
  if (elem_bytes != 4 && elem_bytes != 8) return fail(DRCVAR_ERR_INVALID, "elem_bytes must be 4 or 8");
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return fail(DRCVAR_ERR_CUDA, "no CUDA device");
  DeviceInfo* di = nullptr;
  int rc = device_info(dev, &di);
  if (rc) return rc;
  const size_t avail = static_cast<size_t>(di->max_smem_optin) - drcvar::fixed_smem_bytes(static_cast<size_t>(elem_bytes));
  return static_cast<int64_t>((avail & ~static_cast<size_t>(127)) / (2 * static_cast<size_t>(elem_bytes)));
}

int drcvar_halfspaces_f32(const float* samples, int64_t B, int64_t N, int64_t stride_b, int64_t stride_n,
                          int64_t stride_c, const double* ego, const double* h_in, double alpha, double delta,
                          double epsilon, double r_robot, double r_obs, uint32_t flags, double* h_out,
                          double* h_mean_out, double* g_out, double* cvar_out, double* var_out, double* gstar_out,
                          int32_t* status_out, int32_t* tail_idx_out, int device, void* stream) {
  return entry<float>(samples, B, N, stride_b, stride_n, stride_c, ego, h_in, alpha, delta, epsilon, r_robot, r_obs,
                      flags, h_out, h_mean_out, g_out, cvar_out, var_out, gstar_out, status_out, tail_idx_out, device,
                      stream);
}

int drcvar_halfspaces_f64(const double* samples, int64_t B, int64_t N, int64_t stride_b, int64_t stride_n,
                          int64_t stride_c, const double* ego, const double* h_in, double alpha, double delta,
                          double epsilon, double r_robot, double r_obs, uint32_t flags, double* h_out,
                          double* h_mean_out, double* g_out, double* cvar_out, double* var_out, double* gstar_out,
                          int32_t* status_out, int32_t* tail_idx_out, int device, void* stream) {
  return entry<double>(samples, B, N, stride_b, stride_n, stride_c, ego, h_in, alpha, delta, epsilon, r_robot, r_obs,
                       flags, h_out, h_mean_out, g_out, cvar_out, var_out, gstar_out, status_out, tail_idx_out, device,
                       stream);
}

int drcvar_halfspaces_generated_f32(const double* mean, const double* chol, uint64_t seed, int64_t index_offset, int64_t B,
                                    int64_t N, const double* ego, const double* h_in, double alpha, double delta,
                                    double epsilon, double r_robot, double r_obs, uint32_t flags, double* h_out,
                                    double* h_mean_out, double* g_out, double* cvar_out, double* var_out,
                                    double* gstar_out, int32_t* status_out, int32_t* tail_idx_out, float* samples_out,
                                    int device, void* stream) {
  Call c{nullptr, B, N, 2 * N, 2, 1, ego, h_in, alpha, delta, epsilon, r_robot, r_obs, flags,
         h_out, h_mean_out, g_out, cvar_out, var_out, gstar_out, status_out, tail_idx_out};
  c.gen_mean = mean;
  c.gen_chol = chol;
  c.gen_seed = seed;
  c.gen_index_offset = index_offset;
  c.gen_samples_out = samples_out;
  if (B > 0 && (!mean || !chol)) return fail(DRCVAR_ERR_INVALID, "mean and chol must be non-null");
  int rc = check_common(c);
  if (rc) return rc;
  if (device == DRCVAR_HOST) return run_host_generated(c);
  int prev = 0;
  CUDA_TRY(cudaGetDevice(&prev));
  if (prev != device) CUDA_TRY(cudaSetDevice(device));
  rc = launch_on_device<float>(c, device, static_cast<cudaStream_t>(stream));
  if (rc == DRCVAR_OK && (flags & DRCVAR_FLAG_SYNC)) {
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) rc = fail(DRCVAR_ERR_CUDA, "stream synchronize failed: %s", cudaGetErrorString(e));
  }
  if (prev != device) cudaSetDevice(prev);
  return rc;
}

int drcvar_trajectory_f64(const double* const* traj, int64_t n_obs, int64_t N, int64_t T1, int64_t n_steps,
                          const double* ego_steps, double alpha, double delta, double epsilon, double r_robot,
                          double r_obs, uint32_t flags, double* h_out, double* h_mean_out, double* g_out,
                          int32_t* status_out) {
  if (!traj || n_obs < 1 || N < 1 || T1 < 1 || n_steps < 0 || n_steps > T1 || !ego_steps || !h_out || !g_out)
    return fail(DRCVAR_ERR_INVALID, "bad trajectory arguments");
  // pack halfspace (t, i) = traj[i][:, t, :] into one [n_steps*n_obs, N, 2] batch, ego repeated per obstacle.
  // The batch lives in a pinned buffer that is kept between calls (no page faults, full-speed H2D from the host path) and
  // is filled obstacle-major by a few threads: every source array is read once, front to back (with many obstacles — all
  // the runs of a Monte-Carlo batch — a step-major gather would stream the whole input through the cache once per step).
  const int64_t B = n_steps * n_obs;
  if (B == 0) return DRCVAR_OK;
  for (int64_t i = 0; i < n_obs; ++i)
    if (!traj[i]) return fail(DRCVAR_ERR_INVALID, "null trajectory pointer");
  static std::mutex pack_mu;
  static void* pack_buf = nullptr;
  static size_t pack_cap = 0;
  std::lock_guard<std::mutex> lk(pack_mu);
  const size_t pack_bytes = static_cast<size_t>(B) * N * 2 * sizeof(double);
  int rc = grow_pinned(&pack_buf, &pack_cap, pack_bytes);
  if (rc) return rc;
  double* pack = static_cast<double*>(pack_buf);
  std::vector<double> ego(static_cast<size_t>(B) * 2);
  auto pack_range = [&](int64_t i_lo, int64_t i_hi) {
    for (int64_t i = i_lo; i < i_hi; ++i)
      for (int64_t s = 0; s < N; ++s) {
        const double* src = traj[i] + s * T1 * 2;
        for (int64_t t = 0; t < n_steps; ++t) {
          double* dst = pack + (static_cast<size_t>(t * n_obs + i) * N + s) * 2;
          dst[0] = src[2 * t];
          dst[1] = src[2 * t + 1];
        }
      }
  };
  const int64_t n_thr = std::max<int64_t>(1, std::min<int64_t>({8, static_cast<int64_t>(std::thread::hardware_concurrency()),
                                                                n_obs, static_cast<int64_t>(pack_bytes >> 22)}));
  if (n_thr <= 1) {
    pack_range(0, n_obs);
  } else {
    std::vector<std::thread> pool;
    for (int64_t k = 0; k < n_thr; ++k) pool.emplace_back(pack_range, k * n_obs / n_thr, (k + 1) * n_obs / n_thr);
    for (auto& th : pool) th.join();
  }
  for (int64_t t = 0; t < n_steps; ++t)
    for (int64_t i = 0; i < n_obs; ++i) {
      ego[static_cast<size_t>(t * n_obs + i) * 2] = ego_steps[2 * t];
      ego[static_cast<size_t>(t * n_obs + i) * 2 + 1] = ego_steps[2 * t + 1];
    }
  return drcvar_halfspaces_f64(pack, B, N, N * 2, 2, 1, ego.data(), nullptr, alpha, delta, epsilon, r_robot,
                               r_obs, flags, h_out, h_mean_out, g_out, nullptr, nullptr, nullptr, status_out, nullptr,
                               DRCVAR_HOST, nullptr);
}

void* drcvar_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    fail(DRCVAR_ERR_NOMEM, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void drcvar_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int drcvar_cluster_ctas(int64_t n_samples, int elem_bytes, int64_t smem_optin_bytes) {
  if (n_samples <= drcvar::kOctantMinN || (elem_bytes != 4 && elem_bytes != 8) || smem_optin_bytes <= 0) return 0;
  return cluster_ctas_for(n_samples, static_cast<size_t>(elem_bytes), static_cast<size_t>(smem_optin_bytes));
}

int64_t drcvar_launch_count(void) { return g_launches.load(); }

int64_t drcvar_debug_check_failures(int32_t* first_site) {
  if (first_site) *first_site = 0;
#ifdef DRCVAR_CHECKED
  unsigned long long n = 0;
  int site = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(&n, drcvar_check_fail_count, sizeof(n)) != cudaSuccess) return -2;
  if (cudaMemcpyFromSymbol(&site, drcvar_check_first_site, sizeof(site)) != cudaSuccess) return -2;
  if (first_site) *first_site = site;
  return static_cast<int64_t>(n);
#else
  return -1;   // not a checked build (make -C csrc checked)
#endif
}

#ifdef DRCVAR_PROFILE_PHASES
// profiling builds only (not part of include/drcvar.h): device buffer receiving per-CTA phase cycle counts
void drcvar_debug_phase_buffer(long long* dev_buf) { g_phase_cycles = dev_buf; }
#endif

int drcvar_last_host_call_stats(double* stage_ms, double* kernel_ms, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  HostCtx* hc = nullptr;   // the context of the CURRENT device (the one the last DRCVAR_HOST call of this thread ran on)
  const int rc = host_ctx(&hc);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(hc->mu);
  if (stage_ms) *stage_ms = hc->last_stage_ms;
  if (kernel_ms) *kernel_ms = hc->last_kernel_ms;
  if (h2d_bytes) *h2d_bytes = hc->last_h2d;
  if (d2h_bytes) *d2h_bytes = hc->last_d2h;
  return DRCVAR_OK;
}

}  // extern "C"
