// halfspace_kernel.cuh — sm_100a device code of the risk-bounded safe-halfspace path.
//
// One CTA per (scenario, obstacle, step) halfspace, persistent over the batch; two CTAs per SM.
// A CTA is a team of 8 "sweep" warps plus a "finisher" warp and a "director" warp, pipelined over consecutive
// halfspaces through two parity buffers, so the select + epilogue of halfspace j, the slow IEEE div/sqrt chain of the
// canonical direction and the bulk copy of halfspace j+1 overlap the sweeps around them.  Synchronisation: transaction
// mbarriers for the TMA copies, mbarriers empty[]/hdone[] towards the team, named barrier 1 inside the team, named
// barriers 2-7 (bar.arrive / bar.sync) towards the two helper warps (they block in hardware, no polling).
//
// sweep team, per halfspace:
//   stage   N samples -> shared memory: cp.async.bulk (TMA, issued by the director as soon as the slot is free), a strided
//           loader, or — generate mode — drawn in place (sample_gen.cuh)
//   sweep A starts on the first 32 KB chunk; canonical lane sums of the coordinates (fp32 inputs: shifted by the first
//           sample, packed fp32 lane partials, fp64 cross-lane tree; fp64 inputs: fp64 throughout) + second moments
//   window  warp 0: fp32 direction h_a with a rigorous bound |h_a - h| <= err_h, statistical window [t_lo, t_hi] around
//           the predicted kc-th largest loss, fp32 classification thresholds
//           director warp (concurrently): the canonical h = unit(m - ego) with IEEE div/sqrt  core/geometry.py:35-53
//           and the mean halfspace                                                            core/halfspaces.py:70-106
//   sweep B classify every sample.  fp32 inputs: the rigorous fp32 bound decides "surely above" (count + shifted
//           coordinate sums; the loss sum follows from linearity), "surely below" (ignored) or "needs the exact fp64
//           loss" (a bit in a per-thread mask).  fp64 inputs: exact canonical loss for every sample.
//   phase 2 masked samples are copied to per-warp lists (2a), THE SLOT IS RELEASED, then they get the canonical fp64
//           loss L_i = -(h.xi_i) (no FMA) (2b); window losses go to warp-private candidate lists and a 256-bucket
//           histogram
// finisher warp, per halfspace:
//   select  exact kc-th largest loss T: histogram scan -> bucket -> all-pairs rank inside the bucket
//   finish  CVaR = (sum_{L>T} L + (k_f - #{L>T}) T) / k_f  and the CVaR / DR-CVaR offsets  core/risk_metrics.py:84-338
// The window and the fp32 bound only decide HOW FAST the exact threshold is found; a miss is detected and the
// general multi-sweep radix select (sweep team, all samples) runs instead.  Arithmetic contract: DESIGN.md.
//
// Pipe budget notes (measured on B200, profiles/ubench): FFMA/FADD/FMUL/IADD issue at 4 warp-inst/clk/SM; FSETP, SEL,
// LOP3, SHF, FMNMX, FADD2, IMAD, DADD/DMUL at 2; SHFL ~1; POPC/FLO/REDUX/F2F ~0.5; LDS.128 moves 128 B/clk/SM.  The hot
// loops are written to keep the half-rate "ALU" ops to the two FSETPs per sample.
#pragma once

#include <cuda_runtime.h>

#include "sample_gen.cuh"
#include <stdint.h>

// Optional phase timing (profiling builds only: -DDRCVAR_PROFILE_PHASES; see profiles/phase_cycles.py)
#ifdef DRCVAR_PROFILE_PHASES
#define PH_DECL long long ph_t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long ph_last = clock64();
#define PH_MARK(k) { const long long ph_now = clock64(); ph_t[k] += ph_now - ph_last; ph_last = ph_now; }
#else
#define PH_DECL
#define PH_MARK(k)
#endif

#ifndef DRCVAR_PF_WHERE
#define DRCVAR_PF_WHERE 0
#endif
#ifndef DRCVAR_PF_DIST
#define DRCVAR_PF_DIST 1
#endif

// Checked builds (make -C csrc checked -> libdrcvar_checked.so, never loaded by the product path): every index into a
// shared-memory list / histogram / candidate pool and every staged byte range is asserted in range; failures are counted
// in device globals read back by drcvar_debug_check_failures().  compute-sanitizer is not available on the GPU pool this
// was developed on; tests/test_gpu_checked_build.py runs every kernel path through this build instead.
#ifdef DRCVAR_CHECKED
__device__ unsigned long long drcvar_check_fail_count = 0ull;
__device__ int drcvar_check_first_site = 0;   // 100000 * file id + line of the first failed assertion
#define DRCVAR_ASSERT_AT(fid, cond)                                                                  \
  do {                                                                                               \
    if (!(cond) && atomicAdd(&drcvar_check_fail_count, 1ull) == 0ull) drcvar_check_first_site = 100000 * (fid) + __LINE__; \
  } while (0)
#else
#define DRCVAR_ASSERT_AT(fid, cond) do { } while (0)
#endif
#define DRCVAR_ASSERT(cond) DRCVAR_ASSERT_AT(DRCVAR_FILE_ID, cond)
#define DRCVAR_FILE_ID 1   // halfspace_kernel.cuh (redefined at the top of the other kernel files)

namespace drcvar {

constexpr int kSweepWarps = 8;
constexpr int kSweepThreads = kSweepWarps * 32;      // 256
constexpr int kThreads = kSweepThreads + 64;         // + finisher warp + director warp
constexpr int kFinisherWarp = kSweepWarps;
constexpr int kDirectorWarp = kSweepWarps + 1;
// (a second finisher warp — one per parity buffer — was measured: the single finisher is busy 99 % of the time since its loops
//  are rolled, but 11 warps leave 80 registers per thread and the step got 2.5 % slower: profiles/r2_ab_two_finishers.txt)
constexpr int kSlots = 512;                          // canonical cross-lane tree width (2 slots per sweep thread)
constexpr int kWarpCand = 128;                       // doubles per sweep warp in the candidate buffer ...
constexpr int kCandCap = 96;                         // ... of which candidate losses; the last 32 hold per-lane (sum dx, sum dy)
constexpr int kWarpList = 160;                       // masked samples (raw copies) per sweep warp
constexpr int kHistBuckets = 256;
constexpr int kResolveMax = 32;                      // a bucket this small is ranked by one warp
constexpr int kMaskWords = 4;                        // per-thread "needs exact loss" mask: 128 bits
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kBulkChunk = 32768;
constexpr int kRowsPerChunk = kBulkChunk / (16 * 256);   // a row = one 16-byte load per sweep thread = 4 KB

// Canonical mean for N > kOctantMinN: 8 octants of whole 4 KB rows, each with its own slot sums + tree; the octant totals
// are combined by an adjacent-pair tree (oracle/closed_form.py).  This is the split the cluster kernel
// distributes over its CTAs; the resident kernel never sees such N.
constexpr int kOctantMinN = 32768;
constexpr int kOctants = 8;
__host__ __device__ inline long long octant_bytes(long long n, size_t elem_bytes) {
  const long long row = n * 2 * static_cast<long long>(elem_bytes);
  return (row + kOctants * 4096 - 1) / (kOctants * 4096) * 4096;
}
__host__ __device__ inline int octant_samples(long long n, size_t elem_bytes) {
  return static_cast<int>(octant_bytes(n, elem_bytes) / (2 * static_cast<long long>(elem_bytes)));
}

constexpr int kStatusNonfinite = 1;
constexpr int kStatusGeneral = 2;
constexpr int kStatusDegenerate = 4;

constexpr int kModeFinish = 0;   // finisher computes the result from the candidate lists
constexpr int kModeSkip = 1;     // the sweep team already wrote the result (general path / non-finite input)

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
  float z_mid_f, z_half_f;   // window centre / half-width in z units (Gaussian plan, host), fp32
  float z_lo_f, z_hi_f;      // = z_mid_f -/+ z_half_f
  float z_half_adapt_f;      // half-width used once the window centre is LEARNED from earlier halfspaces (non-Gaussian data)
  int bulk;
  double* h_out;
  double* h_mean_out;
  double* g_out;
  double* cvar_out;
  double* var_out;
  double* gstar_out;
  int* status_out;
  int* tail_idx_out;
  long long* phase_cycles;   // profiling builds: [grid][2][12] accumulated cycles (sweep warp 2, finisher)
  // generate mode (fp32 only, samples == nullptr): samples are drawn on the device, see sample_gen.cuh
  const double* gen_mean;    // [B,2] nominal obstacle position of each halfspace
  const double* gen_chol;    // [B,3] lower Cholesky factor (l00, l10, l11) of the noise covariance
  unsigned long long gen_seed;
  long long gen_index_offset;   // global index of halfspace 0 of this launch (shards reproduce the unsharded stream)
  float* gen_samples_out;    // optional [B,N,2] dump of the generated samples (parity tests)
  // cluster kernel (cluster_kernel.cuh): CTAs per halfspace; halfspaces it could not finish (window miss, overflow,
  // non-finite data) get redo_list[b] = 1 (one flag per halfspace) and are processed by the streaming kernel, which then
  // skips every halfspace whose flag is 0 — a static assignment of halfspaces to CTAs, so its learned windows (and the
  // last bits of the sums) do not depend on the order in which the cluster kernel found the misses
  int cl_ctas;
  int* redo_list;
  int debug;   // profiling builds only: ablation switches
};

struct Ctl {                        // one per parity buffer
  unsigned long long key_lo;       // key(t_lo): histogram origin
  double T;
  double h0, h1, t_lo, t_hi, m0, m1;
  double hn;                       // ||h|| (canonical), written by the director with h
  double f0, f1;                   // first sample (fp32 path: shift origin)
  float h0f, h1f, thr_above, thr_keep;
  int hist_shift;
  int bstar, rprime, cnt_in, small_n;
  int window_ok, nonfinite, degenerate, c_tot;
  int mode, cnt_hi, status;
  int acc_hi, acc_nc;              // team totals (shared atomics): sure-above count, window candidates
  int z_learned;                   // the window centre is taken from z_est (set after two consecutive misses; persists)
  int z_missrun;                   // consecutive window misses of this parity chain
  float z_est;                     // learned (T - mean loss) / sigma of the CTA's earlier halfspaces
  float z_lo_use, z_hi_use;        // window bounds in z units for the NEXT halfspace of this parity (Gaussian plan or learned)
  float pad_f;
  struct __align__(16) Place {     // this halfspace (one 16-byte store): mean and sigma of p = h_a.(xi - first), h_a . first
    float pm, sigma;
    double c_shift;
  } pl;
};

struct Bars {
  unsigned long long data;         // bulk copy landed (all chunks after the first)
  unsigned long long data0;        // first 32 KB chunk landed: sweep A starts on it while the rest is in flight
  unsigned long long empty[2];     // finisher -> sweep team
  unsigned long long hdone[2];     // director -> sweep team: canonical h / mean / flags are in ctl[par]
};

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// shared-memory footprint of one CTA (host and device must agree)
__host__ __device__ inline size_t slot_bytes_for(long long n, size_t elem_bytes) {
  return (static_cast<size_t>(n) * 2 * elem_bytes + 127) & ~static_cast<size_t>(127);
}
constexpr int kRedDoubles = kSweepWarps * 8;                     // per warp: {tx, ty, qxx, qyy, qxy, bound, mdx, mdy}
constexpr int kFinDoubles = kSweepWarps * 4;                     // per parity, per warp: {sum dx, sum dy, exact sum, n}
__host__ __device__ inline size_t fixed_smem_bytes(size_t elem_bytes) {
  return sizeof(double) * 2 * kWarpCand * kSweepWarps   // cand   [2][warps][kWarpCand]
         + sizeof(unsigned) * 2 * kHistBuckets          // hist   [2][256]
         + sizeof(double) * 2 * kRedDoubles             // red    [2]
         + sizeof(double) * 2 * kFinDoubles             // fin    [2]
         + sizeof(double) * 2 * kResolveMax             // small  [2]
         + sizeof(int) * 2 * 2 * kSweepWarps            // ired   [2][warps][2]
         + sizeof(int) * 4 * kSweepWarps                // iscr   (team scratch)
         + 2 * elem_bytes * kWarpList * kSweepWarps     // raw copies of the masked samples [warps][kWarpList]
         + 2 * sizeof(Ctl) + sizeof(Bars);
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void mbar_wait_spin(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// helper warps (finisher, director): potentially-blocking wait with a suspend-time hint so that the idle warp does not
// burn issue slots of the sweep team while it spins
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity);
__device__ __forceinline__ void mbar_wait_idle(unsigned long long* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) { mbar_wait_idle(bar, parity); }
// TMA bulk copy global -> shared::cta, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of a halfspace that is still one slot release away (no shared-memory destination needed): the later bulk
// copy then reads L2 instead of waiting for HBM
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// Named barriers 2..7 (one per parity): handshakes towards the two helper warps.  The producer side arrives without
// waiting (bar.arrive), the helper blocks in hardware (bar.sync) instead of polling an mbarrier.  A barrier id is
// re-armed two halfspaces later, after the waits on empty[] / hdone[] have proved that the helper consumed it.
constexpr int kBarFull = 2;       // warp 0 of the team (32) + finisher (32)
constexpr int kBarADone = 4;      // warp 0 of the team (32) + director (32)
constexpr int kBarSlotFree = 6;   // the 8 sweep warps (256) + director (32)
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// barrier 1: the sweep team only (the finisher warp never joins it)
__device__ __forceinline__ void team_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSweepThreads) : "memory"); }
__device__ __forceinline__ int team_sync_or(int pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.s32 p, %1, 0;\n\t"
      "bar.red.or.pred q, 1, %2, p;\n\t"
      "selp.s32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"(pred), "n"(kSweepThreads)
      : "memory");
  return r;
}
// fp32 classification of one sample (p = h.(xi - first) in fp32):  up = p < thr_above;  keep = !up && p <= thr_keep.
//   up:   accumulate the shifted coordinates and a float count (full-rate FADDs, fma pipe)
//   keep: add `bit` into the mask (bits are distinct, so add == or; full-rate IADD)
__device__ __forceinline__ void classify_f32(float p, float thr_above, float thr_keep, float dx, float dy, float& ax,
                                             float& ay, float& cnt, unsigned& mask, unsigned bit) {
  asm("{\n\t.reg .pred u, k;\n\t"
      "setp.lt.f32 u, %4, %5;\n\t"
      "setp.le.and.f32 k, %4, %6, !u;\n\t"
      "@u add.f32 %0, %0, %7;\n\t"
      "@u add.f32 %1, %1, %8;\n\t"
      "@u add.f32 %2, %2, 0f3F800000;\n\t"
      "@k add.u32 %3, %3, %9;\n\t}"
      : "+f"(ax), "+f"(ay), "+f"(cnt), "+r"(mask)
      : "f"(p), "f"(thr_above), "f"(thr_keep), "f"(dx), "f"(dy), "r"(bit));
}

// exact classification of one fp64 loss, branch-free (the if / else-if form spent 14 % of the fp64 kernel's warp samples
// resolving divergent branches):  up = L > t_hi: sum and count;  keep = !up && L >= t_lo: `bit` into the mask
// The same on the canonical PROJECTION p (L = 0 - p, a negation: exact): up = p < -t_hi, keep = !up && p <= -t_lo, and the sum
// of the projections of the "up" set, whose negative is the sum of their losses bit for bit (round-to-nearest is symmetric).
// One DADD less per sample than forming L first.
__device__ __forceinline__ void classify_f64_proj(double p, double nt_hi, double nt_lo, double& sp_gt, int& c_gt, unsigned& mask,
                                                  unsigned bit) {
  asm("{\n\t.reg .pred u, k;\n\t"
      "setp.lt.f64 u, %3, %4;\n\t"
      "setp.le.and.f64 k, %3, %5, !u;\n\t"
      "@u add.rn.f64 %0, %0, %3;\n\t"
      "@u add.s32 %1, %1, 1;\n\t"
      "@k add.u32 %2, %2, %6;\n\t}"
      : "+d"(sp_gt), "+r"(c_gt), "+r"(mask)
      : "d"(p), "d"(nt_hi), "d"(nt_lo), "r"(bit));
}
__device__ __forceinline__ void classify_f64(double L, double t_hi, double t_lo, double& s_gt, int& c_gt, unsigned& mask,
                                             unsigned bit) {
  asm("{\n\t.reg .pred u, k;\n\t"
      "setp.gt.f64 u, %3, %4;\n\t"
      "setp.ge.and.f64 k, %3, %5, !u;\n\t"
      "@u add.rn.f64 %0, %0, %3;\n\t"
      "@u add.s32 %1, %1, 1;\n\t"
      "@k add.u32 %2, %2, %6;\n\t}"
      : "+d"(s_gt), "+r"(c_gt), "+r"(mask)
      : "d"(L), "d"(t_hi), "d"(t_lo), "r"(bit));
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
// explicit shared-space copies for the compaction loop of phase 2a (no generic -> shared conversion, no pointer pairs)
__device__ __forceinline__ void smem_copy8(uint32_t dst, uint32_t src) {
  asm volatile("{\n\t.reg .b64 t;\n\tld.shared.b64 t, [%1];\n\tst.shared.b64 [%0], t;\n\t}" ::"r"(dst), "r"(src) : "memory");
}
__device__ __forceinline__ void smem_copy16(uint32_t dst, uint32_t src) {
  asm volatile("{\n\t.reg .b64 t, u;\n\tld.shared.v2.b64 {t, u}, [%1];\n\tst.shared.v2.b64 [%0], {t, u};\n\t}" ::"r"(dst), "r"(src) : "memory");
}
// canonical loss  L = 0 - (rn(h0*x) + rn(h1*y))   (never fused)
__device__ __forceinline__ double loss_of(double h0, double h1, double x, double y) {
  return __dsub_rn(0.0, __dadd_rn(__dmul_rn(h0, x), __dmul_rn(h1, y)));
}
// IEEE fp64 division / square root are ~25-instruction software sequences.  The per-halfspace scalar code uses eleven of
// them (mean, direction, mean halfspace, offsets); inlined they were ~270 instructions of a per-halfspace code footprint
// that has to stay inside the 32 KB instruction cache next to the sweeps (profiles/README.md, round 2), so they are
// shared, never inlined.  Same instructions, same bits.
__device__ __noinline__ double ddiv_canon(double a, double b) { return __ddiv_rn(a, b); }
__device__ __noinline__ double dsqrt_canon(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double norm2_canon(double a, double b) {
  return dsqrt_canon(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
}
__device__ __forceinline__ float sqrt_approx(float x) {   // one MUFU.SQRT (<= 2 ulp), no fix-up / slow path
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
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

// Two canonical warp sums for the price of one: the first butterfly stage swaps x against y between lane pairs, so that
// even lanes go on with the x tree and odd lanes with the y tree.  Every addition has the same operands as in
// warp_sum_canon (fp addition commutes), hence the same bits.  Result: sum of x in even lanes, sum of y in odd lanes.
__device__ __forceinline__ double warp_sum_canon_pair(double x, double y, int lane) {
  const bool odd = lane & 1;
  double v = __dadd_rn(odd ? y : x, shfl_xor_d(odd ? x : y, 1));
#pragma unroll
  for (int m = 2; m <= 16; m <<= 1) v = __dadd_rn(v, shfl_xor_d(v, m));
  return v;
}
// Four (non-canonical) float warp sums in 6 shuffles: totals of a, c, b, d end up in lanes with (lane & 3) = 0, 1, 2, 3.
__device__ __forceinline__ float warp_sum_any4(float a, float b, float c, float d, int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  const float k0 = (b0 ? c : a) + __shfl_xor_sync(kFull, b0 ? a : c, 1);
  const float k1 = (b0 ? d : b) + __shfl_xor_sync(kFull, b0 ? b : d, 1);
  float v = (b1 ? k1 : k0) + __shfl_xor_sync(kFull, b1 ? k0 : k1, 2);
#pragma unroll
  for (int m = 4; m <= 16; m <<= 1) v += __shfl_xor_sync(kFull, v, m);
  return v;
}

// Histogram scan by ONE warp: finds the bucket (from the top) that holds rank r (1-based).  Lane l takes the eight buckets
// 255-8l .. 248-8l (two 16-byte loads), ONE inclusive scan over the 32 lane totals locates the lane, the lane walks its
// eight counts.  (The earlier form — eight dependent rounds of load + warp reduction — was ~900 cycles of the finisher's
// serial chain.)
__device__ __forceinline__ void scan_hist_warp(const unsigned* hist, int r, int lane, int& bstar, int& rprime, int& cnt_in) {
  static_assert(kHistBuckets == 256, "lane l owns buckets 255-8l .. 248-8l");
  const uint4* h4 = reinterpret_cast<const uint4*>(hist);
  const uint4 lo = h4[62 - 2 * lane], hi = h4[63 - 2 * lane];   // hist[248-8l .. 251-8l], hist[252-8l .. 255-8l]
  const int c[8] = {static_cast<int>(hi.w), static_cast<int>(hi.z), static_cast<int>(hi.y), static_cast<int>(hi.x),
                    static_cast<int>(lo.w), static_cast<int>(lo.z), static_cast<int>(lo.y), static_cast<int>(lo.x)};   // from the top
  const int tot = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
  int incl = tot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += t;
  }
  const int excl = incl - tot;
  const unsigned hit = __ballot_sync(kFull, excl < r && r <= incl);
  int bk = 0, rp = 1, ci = c[7];   // (no lane holds the rank: unreachable when r <= #candidates in range)
  int run = excl;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (run < r && r <= run + c[i]) {
      bk = kHistBuckets - 1 - (8 * lane + i);
      rp = r - run;
      ci = c[i];
    }
    run += c[i];
  }
  const int src = hit ? __ffs(hit) - 1 : 31;
  bstar = __shfl_sync(kFull, bk, src);
  rprime = __shfl_sync(kFull, rp, src);
  cnt_in = __shfl_sync(kFull, ci, src);
}

// Four (non-canonical) fp64 warp sums at once (transposed butterfly, 9 shuffles instead of 20): returns (a, b, c, d) totals
// in every lane.
__device__ __forceinline__ void warp_sum_any4d(double& a, double& b, double& c, double& d, int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  const double k0 = (b0 ? c : a) + shfl_xor_d(b0 ? a : c, 1);   // even lanes: a, odd lanes: c   (pairs)
  const double k1 = (b0 ? d : b) + shfl_xor_d(b0 ? b : d, 1);   // even lanes: b, odd lanes: d
  double v = (b1 ? k1 : k0) + shfl_xor_d(b1 ? k0 : k1, 2);      // lane & 3 = 0: a, 1: c, 2: b, 3: d   (quads)
#pragma unroll
  for (int m = 4; m <= 16; m <<= 1) v += shfl_xor_d(v, m);
  a = __shfl_sync(kFull, v, 0);
  c = __shfl_sync(kFull, v, 1);
  b = __shfl_sync(kFull, v, 2);
  d = __shfl_sync(kFull, v, 3);
}

// Exact r-th largest (1-based) among the enumerated losses whose keys lie in [lo, hi] — generic narrowing loop.
// SYNC() synchronises the participating threads (team barrier, or __syncwarp for one warp); `leader` is true for
// the first warp of the group, `gtid`/`gsize` index the group.  for_each(f) enumerates this thread's candidates.
template <class ForEach, class Sync>
__device__ double select_rank(ForEach&& for_each, Sync&& SYNC, bool leader, int gtid, int gsize, unsigned long long lo,
                              unsigned long long hi, int r, unsigned* hist, double* small, Ctl* ctl) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    const unsigned long long span = hi - lo;
    if (span == 0) return value_of(lo);
    const int bits = 64 - __clzll(static_cast<long long>(span));
    const int shift = bits > 8 ? bits - 8 : 0;  // (span >> shift) < 256
    for (int i = gtid; i < kHistBuckets; i += gsize) hist[i] = 0;
    if (gtid == 0) ctl->small_n = 0;
    SYNC();
    for_each([&](double L) {
      const unsigned long long k = key_of(L);
      if (k >= lo && k <= hi) {
        DRCVAR_ASSERT(((k - lo) >> shift) < static_cast<unsigned long long>(kHistBuckets));
        atomicAdd(&hist[static_cast<unsigned>((k - lo) >> shift)], 1u);
      }
    });
    SYNC();
    if (leader) {
      int bstar, rprime, cnt_in;
      scan_hist_warp(hist, r, lane, bstar, rprime, cnt_in);
      if (lane == 0) {
        ctl->bstar = bstar;
        ctl->rprime = rprime;
        ctl->cnt_in = cnt_in;
      }
    }
    SYNC();
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
      SYNC();
      if (leader) {
        const unsigned long long mine = lane < cnt_in ? key_of(small[lane]) : 0ull;
        int rank = 0;
        for (int j = 0; j < cnt_in; ++j) {
          const unsigned long long other = __shfl_sync(kFull, mine, j);
          rank += (other > mine) || (other == mine && j < lane);
        }
        if (lane < cnt_in && rank == r - 1) ctl->T = value_of(mine);
      }
      SYNC();
      return ctl->T;
    }
  }
}

// Mean halfspace of one (scenario, obstacle, step): direction from the ORIGIN, core/halfspaces.py:70-106.  One thread.
__device__ __forceinline__ void write_mean_outputs(const KernelArgs& a, long long b, double m0, double m1) {
  double hm0, hm1;
  const double mn = norm2_canon(m0, m1);
  if (mn < 1e-10) {
    hm0 = 1.0;
    hm1 = 0.0;
  } else {
    hm0 = ddiv_canon(m0, mn);
    hm1 = ddiv_canon(m1, mn);
  }
  const double hmn = norm2_canon(hm0, hm1);
  const double g_mean = -__dsub_rn(__dadd_rn(__dmul_rn(hm0, m0), __dmul_rn(hm1, m1)), __dmul_rn(a.R, hmn));
  if (a.h_mean_out) {
    a.h_mean_out[2 * b] = hm0;
    a.h_mean_out[2 * b + 1] = hm1;
  }
  a.g_out[3 * b] = g_mean;
}

// CVaR / DR-CVaR offsets of one halfspace from (h, CVaR ingredients); one thread.      core/risk_metrics.py:84-338
// hn = ||h|| by the canonical chain (norm2_canon(h0, h1)): the resident kernel's director warp has it ready in ctl->hn.
template <class CtlT>
__device__ __forceinline__ void write_risk_outputs(const KernelArgs& a, long long b, const CtlT* ctl, bool nonfinite,
                                                   double s_tot, int c_tot, double T_thr, int status, double hn) {
  const double h0 = ctl->h0, h1 = ctl->h1;
  const double r = __dmul_rn(a.R, hn);
  double cvar, g_cvar, g_star, g_dr, var_t;
  if (nonfinite) {
    cvar = __longlong_as_double(0x7ff8000000000000ll);
    var_t = cvar;
    g_cvar = 100.0;  // solver-failure sentinel, core/risk_metrics.py:177,265,303,338
    g_star = 100.0;
    g_dr = __dsub_rn(100.0, r);
  } else {
    const double S = __dadd_rn(s_tot, __dmul_rn(__dsub_rn(a.k_f, static_cast<double>(c_tot)), T_thr));
    cvar = ddiv_canon(S, a.k_f);
    var_t = T_thr;
    const double cr = __dadd_rn(cvar, r);
    g_cvar = __dsub_rn(cr, a.delta);
    g_star = __dsub_rn(__dadd_rn(cr, a.eoa), a.delta);
    g_dr = __dsub_rn(g_star, r);
  }
  DRCVAR_ASSERT(b >= 0 && b < a.B);
  a.h_out[2 * b] = h0;
  a.h_out[2 * b + 1] = h1;
  a.g_out[3 * b + 1] = g_cvar;
  a.g_out[3 * b + 2] = g_dr;
  if (a.cvar_out) a.cvar_out[b] = cvar;
  if (a.var_out) a.var_out[b] = var_t;
  if (a.gstar_out) a.gstar_out[b] = g_star;
  if (a.status_out) a.status_out[b] = status;
}
__device__ __forceinline__ void write_risk_outputs(const KernelArgs& a, long long b, const Ctl* ctl, bool nonfinite,
                                                   double s_tot, int c_tot, double T_thr, int status) {
  write_risk_outputs(a, b, ctl, nonfinite, s_tot, c_tot, T_thr, status, norm2_canon(ctl->h0, ctl->h1));
}

// ---------------------------------------------------------------------------------------------- the kernel
template <typename T, bool kTail, bool kGen = false>
__global__ void __launch_bounds__(kThreads, 2) halfspace_kernel(const KernelArgs a) {
  using V2 = typename Vec2<T>::type;
  constexpr bool kF32 = sizeof(T) == 4;
  constexpr int kPerLoad = kF32 ? 2 : 1;          // samples per 16-byte shared load
  constexpr int kRowSamples = kSweepThreads * kPerLoad;   // samples per row of 16-byte loads (512 / 256)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const size_t slot_bytes = slot_bytes_for(N, sizeof(T));
  V2* sm = reinterpret_cast<V2*>(smem_raw);
  double* cand_base = reinterpret_cast<double*>(smem_raw + slot_bytes);
  unsigned* hist_base = reinterpret_cast<unsigned*>(cand_base + 2 * kWarpCand * kSweepWarps);
  double* red_base = reinterpret_cast<double*>(hist_base + 2 * kHistBuckets);
  double* fin_base = red_base + 2 * kRedDoubles;
  double* small_base = fin_base + 2 * kFinDoubles;
  int* ired_base = reinterpret_cast<int*>(small_base + 2 * kResolveMax);
  int* iscr = ired_base + 2 * 2 * kSweepWarps;
  V2* list_base = reinterpret_cast<V2*>(iscr + 4 * kSweepWarps);
  Ctl* ctl_base = reinterpret_cast<Ctl*>(list_base + kWarpList * kSweepWarps);
  Bars* bars = reinterpret_cast<Bars*>(ctl_base + 2);

  if (tid == 0) {
    mbar_init(&bars->data, 1);
    mbar_init(&bars->data0, 1);
    mbar_init(&bars->empty[0], 1);
    mbar_init(&bars->empty[1], 1);
    mbar_init(&bars->hdone[0], 1);
    mbar_init(&bars->hdone[1], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 2 * kHistBuckets; i += kThreads) hist_base[i] = 0;
  if (tid < 2) {
    ctl_base[tid].small_n = 0;
    ctl_base[tid].z_learned = 0;
    ctl_base[tid].z_missrun = 0;
    ctl_base[tid].z_lo_use = a.z_mid_f - a.z_half_f;
    ctl_base[tid].z_hi_use = a.z_mid_f + a.z_half_f;
    ctl_base[tid].z_est = 0.f;
  }
  __syncthreads();

  const uint32_t copy_bytes = static_cast<uint32_t>(static_cast<size_t>(N) * sizeof(V2));
  auto issue_bulk = [&](long long b) {
    const unsigned char* src =
        reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b) * a.stride_b * sizeof(T);
    const uint32_t n0 = copy_bytes < kBulkChunk ? copy_bytes : kBulkChunk;
    DRCVAR_ASSERT(b >= 0 && b < a.B && copy_bytes <= slot_bytes && (copy_bytes & 15u) == 0u);
    mbar_expect_tx(&bars->data0, n0);
    bulk_g2s(smem_raw, src, n0, &bars->data0);
    if (copy_bytes > n0) {
      mbar_expect_tx(&bars->data, copy_bytes - n0);
#pragma unroll 1
      for (uint32_t off = n0; off < copy_bytes; off += kBulkChunk) {
        const uint32_t n = copy_bytes - off < kBulkChunk ? copy_bytes - off : kBulkChunk;
        DRCVAR_ASSERT(off + n <= slot_bytes);
        bulk_g2s(smem_raw + off, src + off, n, &bars->data);
      }
    }
  };

  // ============================================================================================ finisher warp
  if (warp == kFinisherWarp) {
    int iter = 0;
    PH_DECL
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
      const int par = iter & 1;
      Ctl* ctl = ctl_base + par;
      unsigned* hist = hist_base + par * kHistBuckets;
      const double* cand = cand_base + par * kWarpCand * kSweepWarps;
      const double* fin = fin_base + par * kFinDoubles;
      const int* ired = ired_base + par * 2 * kSweepWarps;
      double* small = small_base + par * kResolveMax;
      bar_sync(kBarFull + par, 64);   // warp 0 of the team hands halfspace b over
      PH_MARK(0)
      if (ctl->mode == kModeFinish) {
        const int cnt_hi = ctl->cnt_hi;
        const unsigned long long klo = ctl->key_lo;
        const int hshift = ctl->hist_shift;
        int bstar, r, cnt_in;
        scan_hist_warp(hist, a.kc - cnt_hi, lane, bstar, r, cnt_in);
        // pass over all candidates: above bucket b* -> counted/summed; bucket b* -> gathered for exact ranking
        double s3 = 0.0;
        int c3 = 0, n_small = 0;
        // (rolled on purpose, here and below: the finisher is off the team's path, but its code shares the 32 KB
        //  instruction cache with the sweeps — the per-halfspace footprint of the kernel was 35 KB)
#pragma unroll 1
        for (int w = 0; w < kSweepWarps; ++w) {
          const int nc = ired[w * 2 + 1];
          const double* wc = cand + w * kWarpCand;
#pragma unroll 1
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
          s4 = mineAbove ? small[lane] : 0.0;   // (lane partial: reduced with the other sums below)
          c4 = __popc(__ballot_sync(kFull, mineAbove));
        } else {
          // dense / heavily tied bucket: narrow further inside the finisher warp
          const unsigned long long lo2 = klo + (static_cast<unsigned long long>(bstar) << hshift);
          unsigned long long hi2 = lo2 + ((1ull << hshift) - 1ull);
          const unsigned long long khi = key_of(ctl->t_hi);
          if (hi2 > khi || hi2 < lo2) hi2 = khi;
          auto each = [&](auto&& f) {
            for (int w = 0; w < kSweepWarps; ++w) {
              const int nc = ired[w * 2 + 1];
              for (int j = lane; j < nc; j += 32) f(cand[w * kWarpCand + j]);
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
        }
        const int c3t = __reduce_add_sync(kFull, c3);
        double lx = 0.0, ly = 0.0;   // "surely above" coordinate sums: lane partials of the 8 sweep warps, fixed order
        if constexpr (kF32) {
          double lx2 = 0.0, ly2 = 0.0;   // (two chains: half the latency of the rolled loop)
#pragma unroll 1
          for (int w = 0; w < kSweepWarps; w += 2) {
            const float2 p = reinterpret_cast<const float2*>(cand + w * kWarpCand + kCandCap)[lane];
            const float2 q = reinterpret_cast<const float2*>(cand + (w + 1) * kWarpCand + kCandCap)[lane];
            lx += static_cast<double>(p.x);
            ly += static_cast<double>(p.y);
            lx2 += static_cast<double>(q.x);
            ly2 += static_cast<double>(q.y);
          }
          lx += lx2;
          ly += ly2;
        }
        double s3t = s3, s_x = lx, s_y = ly;
        warp_sum_any4d(s3t, s4, s_x, s_y, lane);   // the four sums of the result in one transposed butterfly
        if (lane == 0) {
          double s_e = 0.0, n_lin = static_cast<double>(cnt_hi);   // fp32 inputs: every "above" sample is in the linear sums
          if constexpr (!kF32) {
            n_lin = 0.0;
#pragma unroll 1
            for (int w = 0; w < kSweepWarps; ++w) s_e += fin[w * 4 + 2];
          }
          // loss sum of the "surely above" set by linearity (fp32 inputs), xi_i = f + d_i:
          //   sum_i -(h.xi_i) = -(h0 (n f0 + sum dx) + h1 (n f1 + sum dy))
          const double s_lin = -(ctl->h0 * (n_lin * ctl->f0 + s_x) + ctl->h1 * (n_lin * ctl->f1 + s_y));
          const double s_tot = ((s_e + s_lin) + s3t) + s4;
          const int c_tot = cnt_hi + c3t + c4;
          write_risk_outputs(a, b, ctl, false, s_tot, c_tot, T_thr, ctl->status, ctl->hn);
          {   // learn where the threshold sits in z units: (T - mean loss) / sigma = (pm + T + c) / sigma
            ctl->z_missrun = 0;
            if (ctl->z_learned) {
              const float zT = (ctl->pl.pm + static_cast<float>(T_thr + ctl->pl.c_shift)) / ctl->pl.sigma;
              const float ze = 0.5f * (ctl->z_est + zT);
              ctl->z_est = ze;
              if (isfinite(ze)) {
                ctl->z_lo_use = ze - a.z_half_adapt_f;
                ctl->z_hi_use = ze + a.z_half_adapt_f;
              }
            }
          }
          if (kTail) {
            ctl->T = T_thr;
            ctl->c_tot = c_tot;
          }
        }
      }
      // hand the parity buffers back: histogram zeroed, counters reset
      for (int i = lane; i < kHistBuckets; i += 32) hist[i] = 0;
      if (lane == 0) ctl->small_n = 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->empty[par]);
      PH_MARK(1)
    }
#ifdef DRCVAR_PROFILE_PHASES
    if (lane == 0 && a.phase_cycles)
      for (int k = 0; k < 12; ++k) a.phase_cycles[(blockIdx.x * 2 + 1) * 12 + k] = ph_t[k];
#endif
    return;
  }

  // ============================================================================================ director warp
  // Canonical direction (IEEE div / sqrt chain, ~2k cycles of latency) and the mean halfspace, off the team's path.
  if (warp == kDirectorWarp) {
    int iter = 0;
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
      const int par = iter & 1;
      Ctl* ctl = ctl_base + par;
      const double* red = red_base + par * kRedDoubles;
#if DRCVAR_PF_WHERE == 0
      if (a.bulk && lane == 0) {
        const long long b_pf = b + static_cast<long long>(DRCVAR_PF_DIST) * gridDim.x;
        if (b_pf < a.B)
          bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b_pf) * a.stride_b * sizeof(T),
                           copy_bytes & ~15u);
      }
#endif
      bar_sync(kBarADone + par, 64);   // red[par] is complete
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[kSweepWarps];
#pragma unroll
        for (int g = 0; g < kSweepWarps; ++g) t[g] = red[g * 8 + j];
#pragma unroll
        for (int n = kSweepWarps; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);  // adjacent-pair tree
        w[j] = t[0];
      }
      double m0 = ddiv_canon(w[0], static_cast<double>(N));
      double m1 = ddiv_canon(w[1], static_cast<double>(N));
      if constexpr (kF32) {  // fp32 inputs: the lane sums were taken relative to the first sample
        m0 = __dadd_rn(red[6], m0);
        m1 = __dadd_rn(red[7], m1);
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
          h0 = ddiv_canon(d0, nrm);
          h1 = ddiv_canon(d1, nrm);
        }
      }
      nonfinite |= !(isfinite(h0) && isfinite(h1));
      const double hn = norm2_canon(h0, h1);   // ||h|| for the offsets (the team needs h only after sweep B: time to spare)
      if (lane == 0) {
        ctl->h0 = h0; ctl->h1 = h1; ctl->m0 = m0; ctl->m1 = m1;
        ctl->hn = hn;
        ctl->nonfinite = nonfinite;
        ctl->degenerate = degenerate;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->hdone[par]);
#if DRCVAR_PF_WHERE == 1
      if (a.bulk && lane == 0) {
        const long long b_pf = b + static_cast<long long>(DRCVAR_PF_DIST) * gridDim.x;
        if (b_pf < a.B)
          bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b_pf) * a.stride_b * sizeof(T),
                           copy_bytes & ~15u);
      }
#endif
      // TMA producer: as soon as all 8 sweep warps are done with the sample slot, fetch the next halfspace
      bar_sync(kBarSlotFree + par, kSweepThreads + 32);
      if (a.bulk) {
        const long long b_next = b + gridDim.x;
        if (lane == 0 && b_next < a.B) issue_bulk(b_next);
      }
      if (lane == 0) write_mean_outputs(a, b, m0, m1);
    }
    return;
  }

  // ============================================================================================ sweep team
  if (a.bulk && tid == 0 && static_cast<long long>(blockIdx.x) < a.B) issue_bulk(blockIdx.x);
  uint32_t phase = 0;
  int iter = 0;
  const int full_rows = N / kRowSamples;                 // rows of 16-byte loads fully inside the data
  const int rows_all = (N + kRowSamples - 1) / kRowSamples;
  V2* wlist = list_base + warp * kWarpList;
  const uint32_t wlist_s = smem_u32(wlist), tslot_s = smem_u32(smem_raw) + 16u * static_cast<uint32_t>(tid);   // shared-space addresses (phase 2a)
  const double inv_n = 1.0 / static_cast<double>(N);
  double inv_sub = inv_n;   // 1 / (#samples in the second moments): all samples (fp32) / every 4th row (fp64)
  if (!kF32) {
    const int r4 = (rows_all + 3) / 4;
    const int last = (r4 - 1) * 4 * kRowSamples;
    const int n_sub0 = (r4 - 1) * kSweepThreads + (N - last < kSweepThreads ? N - last : kSweepThreads);
    inv_sub = 1.0 / static_cast<double>(n_sub0 > 0 ? n_sub0 : 1);
  }
  const float inv_sub_f = static_cast<float>(inv_sub);
  // Non-bulk staging of halfspace b into the slot by the team: strided loader for arbitrary element strides, or — in
  // generate mode (fp32) — the samples themselves: nominal position + L z, z from Philox4x32-10 + Box-Muller
  // (sample_gen.cuh; one Philox call per pair of samples), optionally dumped for the parity tests.
  auto stage_generic = [&](long long b) {
    if constexpr (kF32 && kGen) {   // separate instantiation: the resident kernel's code is untouched
      {
        const float mx = static_cast<float>(a.gen_mean[2 * b]), my = static_cast<float>(a.gen_mean[2 * b + 1]);
        const float l00 = static_cast<float>(a.gen_chol[3 * b]), l10 = static_cast<float>(a.gen_chol[3 * b + 1]),
                    l11 = static_cast<float>(a.gen_chol[3 * b + 2]);
        const unsigned long long gb = static_cast<unsigned long long>(b + a.gen_index_offset);
        const uint32_t k0 = static_cast<uint32_t>(a.gen_seed), k1 = static_cast<uint32_t>(a.gen_seed >> 32);
        const int n_pairs = (N + 1) >> 1;
        float4* sm4w = reinterpret_cast<float4*>(smem_raw);
        float* dump = a.gen_samples_out ? a.gen_samples_out + static_cast<size_t>(b) * N * 2 : nullptr;
        for (int j = tid; j < n_pairs; j += kSweepThreads) {
          const Philox4 r = philox4x32_10(static_cast<uint32_t>(j), static_cast<uint32_t>(gb),
                                          static_cast<uint32_t>(gb >> 32), kGenStreamTag, k0, k1);
          const float2 s0 = gen_sample(r.x, r.y, mx, my, l00, l10, l11);
          const float2 s1 = gen_sample(r.z, r.w, mx, my, l00, l10, l11);
          sm4w[j] = make_float4(s0.x, s0.y, s1.x, s1.y);   // an odd N leaves >= 8 spare bytes in the 128-byte rounded slot
          if (dump) {
            dump[4 * j] = s0.x;
            dump[4 * j + 1] = s0.y;
            if (2 * j + 1 < N) {
              dump[4 * j + 2] = s1.x;
              dump[4 * j + 3] = s1.y;
            }
          }
        }
        return;
      }
    }
    const T* base = reinterpret_cast<const T*>(a.samples) + b * a.stride_b;
    for (int i = tid; i < N; i += kSweepThreads) {
      const T* p = base + static_cast<long long>(i) * a.stride_n;
      V2 v;
      v.x = p[0];
      v.y = p[a.stride_c];
      sm[i] = v;
    }
  };
  PH_DECL

  for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
    const int par = iter & 1, use = iter >> 1;
    Ctl* ctl = ctl_base + par;
    unsigned* hist = hist_base + par * kHistBuckets;
    double* wcand = cand_base + par * kWarpCand * kSweepWarps + warp * kWarpCand;
    double* fin = fin_base + par * kFinDoubles;
    int* ired = ired_base + par * 2 * kSweepWarps;
    double* small = small_base + par * kResolveMax;
    double* red = red_base + par * kRedDoubles;
    double pre0 = 0.0, pre1 = 0.0;   // warp 0: ego (or the explicit normal) of halfspace b, fetched early
    if (warp == 0) {
      if (a.h_in != nullptr) {
        pre0 = a.h_in[2 * b];
        pre1 = a.h_in[2 * b + 1];
      } else if (a.ego != nullptr) {
        pre0 = a.ego[2 * b];
        pre1 = a.ego[2 * b + 1];
      }
    }
    bool released = false;   // this warp has told the director that it no longer reads the sample slot
    bool redo_bulk = false;
    const long long b_next = b + gridDim.x;

    // ------------------------------------------------------------------ stage
    if (a.bulk) {
      mbar_wait(&bars->data0, phase);   // the first chunk; the rest is awaited inside sweep A
    } else {
      team_sync();   // every warp is done with the previous halfspace's samples
      stage_generic(b);
      team_sync();
    }
    PH_MARK(9)   // (profiling builds: wait for the first chunk)
    // parity buffers must have been handed back by the finisher (two iterations ago)
    if (use > 0) mbar_wait(&bars->empty[par], (use - 1) & 1);
    PH_MARK(0)   // (wait for the finisher)

    // ------------------------------------------------------------------ sweep A: canonical lane sums + second moments
    // Row r = the 16-byte vector r*256 + tid.  fp32: samples 2(r*256+tid)+{0,1} = lanes 2 slot + {0,1} of tile r/2,
    // slot = (r&1)*256 + tid.  fp64: sample r*256 + tid = slot (r&1)*256 + tid of tile r/2.
    bool rest_pending = a.bulk != 0;
    auto wait_rest = [&]() {   // chunks 1.. of halfspace b (a no-op when the whole copy fits the first chunk)
      if (rest_pending) {
        PH_MARK(1)
        if (copy_bytes > kBulkChunk) mbar_wait(&bars->data, phase);
        PH_MARK(10)
        phase ^= 1u;
        rest_pending = false;
      }
    };
    const V2 first = sm[0];
    double u_x, u_y;                 // this thread's value in the 256-wide tree: slot tid + slot tid+256
    double q_xx, q_yy, q_xy;         // second moments of (xi - first): all samples (fp32) / every 4th row (fp64)
    double q_dx = 0.0, q_dy = 0.0;   // fp64 inputs only: first moments of the same subset
    float bound2 = 0.f;              // fp32 inputs: per-thread sum of |xi - first|^2 >= max |xi - first|^2
    if constexpr (kF32) {
      const float4* sm4 = reinterpret_cast<const float4*>(smem_raw);
      const float2 nf = make_float2(-first.x, -first.y);
      float2 acc[2][2];
#pragma unroll
      for (int q = 0; q < 2; ++q) acc[q][0] = acc[q][1] = make_float2(0.f, 0.f);
      float2 sq = make_float2(0.f, 0.f);
      float sxy = 0.f;
      auto body = [&](const float4 v, int q) {
        const float2 d0 = __fadd2_rn(make_float2(v.x, v.y), nf), d1 = __fadd2_rn(make_float2(v.z, v.w), nf);
        acc[q][0] = __fadd2_rn(acc[q][0], d0);
        acc[q][1] = __fadd2_rn(acc[q][1], d1);
        sq = __ffma2_rn(d0, d0, sq);
        sq = __ffma2_rn(d1, d1, sq);
        sxy = fmaf(d0.x, d0.y, sxy);
        sxy = fmaf(d1.x, d1.y, sxy);
      };
      int r = 0;
      if (full_rows >= kRowsPerChunk) {   // the first chunk while the others land
#pragma unroll
        for (; r < kRowsPerChunk; r += 2) {
          const float4 va = sm4[r * kSweepThreads + tid], vb = sm4[(r + 1) * kSweepThreads + tid];
          body(va, 0);
          body(vb, 1);
        }
      }
      wait_rest();
#pragma unroll 1
      for (; r + 1 < full_rows; r += 2) {
        const float4 va = sm4[r * kSweepThreads + tid], vb = sm4[(r + 1) * kSweepThreads + tid];
        body(va, 0);
        body(vb, 1);
      }
      // remaining rows (a last pair whose second row is ragged, and/or a single last row): same vector path; samples
      // beyond N are replaced by the first sample, whose shifted value is +0 and adds nothing (the slot is followed by
      // other shared arrays, so the 16-byte load itself stays inside the CTA's allocation).  r is even here, so the
      // tile parity of every row is static.
      auto masked = [&](int row) {
        float4 v = sm4[row * kSweepThreads + tid];
        const int i0 = row * kRowSamples + 2 * tid;
        if (i0 >= N) { v.x = first.x; v.y = first.y; }
        if (i0 + 1 >= N) { v.z = first.x; v.w = first.y; }
        return v;
      };
      if (r + 1 < rows_all) {
        const float4 va = sm4[r * kSweepThreads + tid], vb = masked(r + 1);
        body(va, 0);
        body(vb, 1);
        r += 2;
      }
      if (r < rows_all) body(masked(r), 0);
      // adjacent fp32 lanes (2 slot, 2 slot + 1) widened and added in fp64, then slot tid + slot tid+256
      const double s0x = __dadd_rn(static_cast<double>(acc[0][0].x), static_cast<double>(acc[0][1].x));
      const double s0y = __dadd_rn(static_cast<double>(acc[0][0].y), static_cast<double>(acc[0][1].y));
      const double s1x = __dadd_rn(static_cast<double>(acc[1][0].x), static_cast<double>(acc[1][1].x));
      const double s1y = __dadd_rn(static_cast<double>(acc[1][0].y), static_cast<double>(acc[1][1].y));
      u_x = __dadd_rn(s0x, s1x);
      u_y = __dadd_rn(s0y, s1y);
      q_xx = sq.x;
      q_yy = sq.y;
      q_xy = sxy;
      bound2 = sq.x + sq.y;
    } else {
      double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
      q_xx = q_yy = q_xy = 0.0;
      // one row = one sample per thread; even rows feed slot tid, odd rows slot tid + 256, each in increasing row order
      // (canonical); second moments on every 4th row.  Groups of four rows with the loads up front: with one CTA
      // per SM (160 KB slot) there are two sweep warps per scheduler and nothing else hides the shared-memory latency.
      auto acc_row = [&](const V2 v, bool odd, bool mom) {
        if (!odd) {
          s00 = __dadd_rn(s00, v.x);
          s01 = __dadd_rn(s01, v.y);
        } else {
          s10 = __dadd_rn(s10, v.x);
          s11 = __dadd_rn(s11, v.y);
        }
        if (mom) {
          const double dx = v.x - first.x, dy = v.y - first.y;
          q_dx += dx;
          q_dy += dy;
          q_xx = fma(dx, dx, q_xx);
          q_yy = fma(dy, dy, q_yy);
          q_xy = fma(dx, dy, q_xy);
        }
      };
      auto rows = [&](int r_lo, int r_hi) {   // r_lo is a multiple of 4
        int r = r_lo;
        const int g_hi = r_lo + (((r_hi < full_rows ? r_hi : full_rows) - r_lo) & ~3);
#pragma unroll 1
        for (; r < g_hi; r += 4) {
          V2 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = sm[(r + k) * kRowSamples + tid];
#pragma unroll
          for (int k = 0; k < 4; ++k) acc_row(v[k], (k & 1) != 0, k == 0);
        }
        for (; r < r_hi; ++r) {   // (a masked group of four was measured for these last rows: slower)
          const int i = r * kRowSamples + tid;
          if (i < N) acc_row(sm[i], (r & 1) != 0, (r & 3) == 0);
        }
      };
      const int r_first = rows_all < kRowsPerChunk ? rows_all : kRowsPerChunk;
      rows(0, r_first);        // the first chunk while the others land
      wait_rest();
      rows(r_first, rows_all);
      u_x = __dadd_rn(s00, s10);
      u_y = __dadd_rn(s01, s11);
    }
    wait_rest();
    {
      // canonical: xor-butterfly inside each group of 32 (x in even lanes, y in odd lanes); the 8 group totals are
      // tree-added by the director (canonical) and by warp 0 (window placement)
      const double txy = warp_sum_canon_pair(u_x, u_y, lane);
      const float mq = warp_sum_any4(static_cast<float>(q_xx), static_cast<float>(q_yy), static_cast<float>(q_xy), 0.f, lane);
      const unsigned bnd = __reduce_max_sync(kFull, __float_as_uint(bound2));
      float mdx = 0.f, mdy = 0.f;
      if constexpr (!kF32) {
        mdx = warp_sum_any(static_cast<float>(q_dx));
        mdy = warp_sum_any(static_cast<float>(q_dy));
      }
      double* w = red + warp * 8;
      float* wf = reinterpret_cast<float*>(w);
      if (lane < 2) w[lane] = txy;                                   // lane 0: x total, lane 1: y total
      if (lane < 3) wf[4 + ((lane & 1) << 1) + (lane >> 1)] = mq;    // lane 0: qxx -> [4], lane 1: qxy -> [6], lane 2: qyy -> [5]
      if (lane == 0) {
        wf[7] = __uint_as_float(bnd);
        wf[8] = mdx; wf[9] = mdy;
        if (kF32) {   // shift origin for the director
          w[6] = static_cast<double>(first.x);
          w[7] = static_cast<double>(first.y);
        }
      }
    }
    PH_MARK(1)
    team_sync();  // S1
    PH_MARK(2)

    // ------------------------------------------------------------------ window placement (warp 0); canonical h: director warp
    if (warp == 0) bar_arrive(kBarADone + par, 64);   // red[par] is complete (S1): the director starts the canonical chain
    if (warp == 0) {
      // Everything here only PLACES the window (speed, never the result) except the fp32 thresholds, which carry
      // rigorous error bounds against the canonical direction the director is computing meanwhile.  The direction used
      // for the classification is h_a = (h0f, h1f), an fp32 approximation: |h_a - h| <= err_h per component.
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[kSweepWarps];
#pragma unroll
        for (int g = 0; g < kSweepWarps; ++g) t[g] = red[g * 8 + j];
#pragma unroll
        for (int n = kSweepWarps; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = t[2 * g] + t[2 * g + 1];
        w[j] = t[0];
      }
      const float* redf = reinterpret_cast<const float*>(red);   // warp g: floats 4..9 of its 16 = qxx,qyy,qxy,bound,mdx,mdy
      float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      float b2 = 0.f;
#pragma unroll   // (warp 0's window placement is the one serial section on the team's path: rolling this loop costs 3 %)
      for (int g = 0; g < kSweepWarps; ++g) {
        q[0] += redf[g * 16 + 4];
        q[1] += redf[g * 16 + 5];
        q[2] += redf[g * 16 + 6];
        b2 = fmaxf(b2, redf[g * 16 + 7]);
        if (!kF32) {
          q[3] += redf[g * 16 + 8];
          q[4] += redf[g * 16 + 9];
        }
      }
      const double f0 = static_cast<double>(first.x), f1 = static_cast<double>(first.y);
      double m0 = w[0] * inv_n, m1 = w[1] * inv_n;   // fp32 inputs: mean relative to the first sample
      double mr0 = m0, mr1 = m1;
      if (kF32) {
        m0 += f0;
        m1 += f1;
      } else {
        mr0 = m0 - f0;
        mr1 = m1 - f1;
      }
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
      // heuristic window around the expected kc-th largest loss
      int n_sub_i;
      float ex, ey;
      if (kF32) {
        n_sub_i = N;
        ex = mr0f;
        ey = mr1f;
      } else {
        const int r4 = (rows_all + 3) / 4;                       // rows 0, 4, 8, ...
        const int last = (r4 - 1) * 4 * kRowSamples;             // first sample of the last such row
        n_sub_i = (r4 - 1) * kSweepThreads + (N - last < kSweepThreads ? N - last : kSweepThreads);
        ex = q[3] * inv_sub_f;
        ey = q[4] * inv_sub_f;
      }
      const float cxx = q[0] * inv_sub_f - ex * ex, cyy = q[1] * inv_sub_f - ey * ey, cxy = q[2] * inv_sub_f - ex * ey;
      const float var_l = h0f * h0f * cxx + 2.0f * h0f * h1f * cxy + h1f * h1f * cyy;
      const float sigma = sqrt_approx(var_l);   // placement only
      int window_ok = a.use_window && usable && (n_sub_i >= 256) && (var_l > 0.f) && isfinite(sigma) &&
                      (rows_all * kPerLoad <= 32 * kMaskWords);
      // thresholds in shifted coordinates, p = h_a.(xi - first):  a_lo <-> t_lo,  a_hi <-> t_hi  (a_hi <= a_lo)
      const float pm = fmaf(h1f, mr1f, h0f * mr0f);
      // window bounds in z units: the Gaussian plan of the host, or — after two CONSECUTIVE misses in this CTA's parity
      // chain, i.e. samples that are evidently not Gaussian — the position measured on the earlier halfspaces with a wider
      // window (z_lo_use / z_hi_use, maintained by the finisher and the fallback path).  Isolated misses, 3e-5 of
      // Gaussian halfspaces, change nothing: the window of a halfspace then depends on its own samples only and results
      // are independent of how the batch is composed.  The state was written two halfspaces ago in this parity buffer,
      // so the choice is deterministic.  Speed only, never T or the tail set.
      float zlo = a.z_lo_f, zhi = a.z_hi_f;
      if (ctl->z_learned) {   // rare, warp-uniform
        zlo = ctl->z_lo_use;
        zhi = ctl->z_hi_use;
      }
      const float a_lo = pm - zlo * sigma, a_hi = pm - zhi * sigma;
      const double c = static_cast<double>(h0f) * f0 + static_cast<double>(h1f) * f1;   // h_a . first
      const double t_lo = __dadd_rn(-static_cast<double>(a_lo) - c, 0.0);  // +0.0: never -0.0 (canonical losses are +0)
      const double t_hi = __dadd_rn(-static_cast<double>(a_hi) - c, 0.0);
      // fp32 classification of p32 = fma(h1f, dy, h0f*dx), d = fl32(xi - first)  (p = h.xi = h.first + h.d = -L):
      //   |p32 - h_a.d_true| <= 5 * 2^-24 * (|h0f| + |h1f|) * max|d|; we allow 2^-19 (32x);
      //   |h.xi - (c + h_a.d_true)| <= err_h (|first| + max|d|) (1-norms); the fp64 roundings of c, t and L are ~1e-16 relative.
      //   p32 <  thr_above  =>  L > t_hi  for sure;    p32 > thr_keep  =>  L < t_lo  for sure.
      const float dmax = sqrt_approx(b2) * 1.0001f;   // 2 ulp, inside the 1e-4 margin
      const float af0 = fabsf(static_cast<float>(f0)) * 1.0001f, af1 = fabsf(static_cast<float>(f1)) * 1.0001f;
      const float habs = fabsf(h0f) + fabsf(h1f);
      const float eps = (habs * (af0 + af1 + dmax)) * 1e-15f + err_h * 1.5f * (af0 + af1 + 2.0f * dmax);
      const float bound = habs * dmax * 1.9073486e-06f + 1.1754944e-38f + eps * 1.0001f;
      const float thr_keep = a_lo + (bound + fabsf(a_lo) * 2.3841858e-07f);
      const float thr_above = a_hi - (bound + fabsf(a_hi) * 2.3841858e-07f);
      const unsigned long long klo = key_of(t_lo), khi = key_of(t_hi);
      const unsigned long long span = khi - klo;
      const int bits = span ? 64 - __clzll(static_cast<long long>(span)) : 0;
      window_ok = window_ok && isfinite(thr_keep) && isfinite(thr_above) && (khi >= klo) && (thr_above <= thr_keep) &&
                  isfinite(t_lo) && isfinite(t_hi);
      if (lane == 0) {
        ctl->f0 = f0; ctl->f1 = f1;
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->h0f = h0f; ctl->h1f = h1f; ctl->thr_keep = thr_keep; ctl->thr_above = thr_above;
        ctl->key_lo = klo;
        ctl->hist_shift = bits > 8 ? bits - 8 : 0;
        ctl->window_ok = window_ok;
        ctl->pl = Ctl::Place{pm, sigma, c};
        ctl->acc_hi = 0;
        ctl->acc_nc = 0;
      }
    }
    team_sync();  // S2
    PH_MARK(3)
    const bool window = ctl->window_ok != 0;
    const double t_lo = ctl->t_lo, t_hi = ctl->t_hi;
    // canonical direction and flags arrive from the director warp; fp32 inputs only need them for phase 2b
    double h0 = 0.0, h1 = 0.0;
    bool nonfinite = false;
    int status = 0;
    bool have_h = false;
    auto need_h = [&]() {
      if (!have_h) {
        mbar_wait(&bars->hdone[par], use & 1);
        h0 = ctl->h0;
        h1 = ctl->h1;
        nonfinite = ctl->nonfinite != 0;
        status = (nonfinite ? kStatusNonfinite : 0) | (ctl->degenerate ? kStatusDegenerate : 0);
        have_h = true;
      }
    };
    if (!kF32 || !window) need_h();

    double T_thr = 0.0;
    int c_gt = 0;        // exact-classified losses above the threshold (this thread)
    double s_gt = 0.0;   // their sum
    bool fast = false;

    if (window && !nonfinite) {
      // ---------------------------------------------------------------- sweep B: classify, build the exact-needed mask
      // Mask bit P = kPerLoad r + e  <->  sample r*kRowSamples + kPerLoad tid + e.
      unsigned mask[kMaskWords] = {0u, 0u, 0u, 0u};
      float ax = 0.f, ay = 0.f, cf = 0.f;            // "surely above": shifted coordinate sums and count
      constexpr int kRowsPerWord = 32 / kPerLoad;     // 16 (fp32) / 32 (fp64)
      if constexpr (kF32) {
        const float4* sm4 = reinterpret_cast<const float4*>(smem_raw);
        const float h0f = ctl->h0f, h1f = ctl->h1f, thr_keep = ctl->thr_keep, thr_above = ctl->thr_above;
        const float2 nf = make_float2(-first.x, -first.y);
#pragma unroll
        for (int wd = 0; wd < kMaskWords; ++wd) {   // all complete rows (groups of four, then the one to three left over)
          const int r_lo = wd * kRowsPerWord;
          const int r_hi = full_rows < r_lo + kRowsPerWord ? full_rows : r_lo + kRowsPerWord;
          unsigned bit = 1u;
#pragma unroll 4   // (8 rows per iteration issue no better and cost 1.3 KB of instruction-cache footprint)
          for (int r = r_lo; r < r_hi; ++r) {
            const float4 v = sm4[r * kSweepThreads + tid];
            const float2 d0 = __fadd2_rn(make_float2(v.x, v.y), nf), d1 = __fadd2_rn(make_float2(v.z, v.w), nf);
            const float p0 = fmaf(h1f, d0.y, h0f * d0.x), p1 = fmaf(h1f, d1.y, h0f * d1.x);
            classify_f32(p0, thr_above, thr_keep, d0.x, d0.y, ax, ay, cf, mask[wd], bit);
            classify_f32(p1, thr_above, thr_keep, d1.x, d1.y, ax, ay, cf, mask[wd], bit + bit);
            bit <<= 2;
          }
        }
        if (full_rows < rows_all) {
          // the ragged last row: samples beyond N get p = +inf (neither "above" nor kept).  (Pushing every row after the last
          // complete group of four through this masked code cost the pipelined kernel 2.7 %.)
          const float4 v = sm4[full_rows * kSweepThreads + tid];
          const int i0 = full_rows * kRowSamples + 2 * tid;
          const float2 d0 = __fadd2_rn(make_float2(v.x, v.y), nf), d1 = __fadd2_rn(make_float2(v.z, v.w), nf);
          float p0 = fmaf(h1f, d0.y, h0f * d0.x), p1 = fmaf(h1f, d1.y, h0f * d1.x);
          if (i0 >= N) p0 = __int_as_float(0x7f800000);
          if (i0 + 1 >= N) p1 = __int_as_float(0x7f800000);
          unsigned mk = 0;
          const unsigned bit = 1u << ((2 * full_rows) & 31);
          classify_f32(p0, thr_above, thr_keep, d0.x, d0.y, ax, ay, cf, mk, bit);
          classify_f32(p1, thr_above, thr_keep, d1.x, d1.y, ax, ay, cf, mk, bit + bit);
          const int wg = (2 * full_rows) >> 5;
#pragma unroll
          for (int w2 = 0; w2 < kMaskWords; ++w2) mask[w2] |= (w2 == wg) ? mk : 0u;   // selects keep mask[] in registers
        }
      } else {
        const double nt_hi = -t_hi, nt_lo = -t_lo;   // window edges for the projection p = -L
        double sp_gt = 0.0;                          // sum of the projections of the losses above the window
#pragma unroll
        for (int wd = 0; wd < kMaskWords; ++wd) {
          const int r_lo = wd * kRowsPerWord;
          const int r_hi = rows_all < r_lo + kRowsPerWord ? rows_all : r_lo + kRowsPerWord;
          unsigned bit = 1u;
          auto one = [&](const V2 v, unsigned bt) {   // exact canonical projection of one sample (rows in increasing order: the sum is deterministic)
            const double p = __dadd_rn(__dmul_rn(h0, v.x), __dmul_rn(h1, v.y));
            classify_f64_proj(p, nt_hi, nt_lo, sp_gt, c_gt, mask[wd], bt);
          };
          int r = r_lo;
          const int g_hi = r_lo + (((r_hi < full_rows ? r_hi : full_rows) - r_lo) & ~3);
#pragma unroll 1
          for (; r < g_hi; r += 4, bit <<= 4) {   // four complete rows, loads up front
            V2 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = sm[(r + k) * kRowSamples + tid];
#pragma unroll
            for (int k = 0; k < 4; ++k) one(v[k], bit << k);
          }
          for (; r < r_hi; ++r, bit <<= 1) {
            const int i = r * kRowSamples + tid;
            if (i < N) one(sm[i], bit);
          }
        }
        s_gt = __dsub_rn(0.0, sp_gt);   // sum of the losses, same bits as adding L = 0 - p in the same order (and never -0)
      }
      PH_MARK(4)

      // -------------------------------------------------------------- phase 2a: compact the masked samples per warp
      int mine_n = 0;
#pragma unroll
      for (int wd = 0; wd < kMaskWords; ++wd)
        if (wd * 32 < rows_all * kPerLoad) mine_n += __popc(mask[wd]);
      // exclusive prefix over the lanes.  Counts are tiny (~0.6 per lane): three independent ballots (one per bit plane)
      // instead of a dependent 5-stage shuffle scan; the shuffle scan remains for counts >= 8.
      int excl, n_list;
      const unsigned lt_mask = (1u << lane) - 1u;
      if (__ballot_sync(kFull, mine_n >= 8) == 0u) {
        const unsigned b0 = __ballot_sync(kFull, mine_n & 1), b1 = __ballot_sync(kFull, mine_n & 2),
                       b2 = __ballot_sync(kFull, mine_n & 4);
        excl = __popc(b0 & lt_mask) + 2 * __popc(b1 & lt_mask) + 4 * __popc(b2 & lt_mask);
        n_list = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
      } else {
        int incl = mine_n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += t;
        }
        n_list = __shfl_sync(kFull, incl, 31);
        excl = incl - mine_n;
      }
      bool overflow = n_list > kWarpList;
      if (!overflow) {
        // 32-bit shared-space addresses; byte address of mask bit P = 32 wd + bp:  fp32: (P>>1) rows of 4 KB + 16 tid + 8 (P&1);
        // fp64: P rows + 16 tid
        uint32_t dst_s = wlist_s + static_cast<uint32_t>(sizeof(V2)) * static_cast<uint32_t>(excl);
#pragma unroll
        for (int wd = 0; wd < kMaskWords; ++wd) {
          if (wd * 32 < rows_all * kPerLoad) {
            unsigned mm = mask[wd];
            const uint32_t wb = tslot_s + (kF32 ? 16u : 32u) * 4096u * wd;
            while (mm) {
              const unsigned bp = 31u - static_cast<unsigned>(__clz(static_cast<int>(mm)));   // highest set bit (one FLO)
              mm ^= 1u << bp;
              const uint32_t off = kF32 ? (((bp << 11) & 0xF000u) | ((bp << 3) & 8u)) : (bp << 12);
              DRCVAR_ASSERT(dst_s + sizeof(V2) <= wlist_s + sizeof(V2) * kWarpList &&
                            (wb - tslot_s + 16u * tid + off) + sizeof(V2) <= slot_bytes);
              if constexpr (kF32) smem_copy8(dst_s, wb + off);
              else smem_copy16(dst_s, wb + off);
              dst_s += static_cast<uint32_t>(sizeof(V2));
            }
          }
        }
      }
      __syncwarp();
      if (!kTail) {
        // the masked samples were copied out: the slot can be refilled while phase 2b and the select run
        bar_arrive(kBarSlotFree + par, kSweepThreads + 32);
        released = true;
      }
      need_h();
      if (nonfinite) overflow = true;   // (cannot happen when the window was placed; keeps the flow uniform)
      // -------------------------------------------------------------- phase 2b: exact loss of the listed samples, dense
      int nc = 0;  // candidates of this warp (warp-uniform)
      const unsigned long long klo = ctl->key_lo;
      const int hshift = ctl->hist_shift;
      if (!overflow) {
        for (int k0 = 0; k0 < n_list; k0 += 32) {
          const int k = k0 + lane;
          const bool active = k < n_list;
          double L = 0.0;
          V2 v = first;
          if (active) {
            v = wlist[k];
            L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
          }
          const bool up = active && (L > t_hi);
          const bool cd = active && !up && (L >= t_lo);
          if constexpr (kF32) {
            if (up) {  // inside the fp32 uncertainty band but exactly above the window: joins the "above" set
              cf += 1.0f;
              ax += static_cast<float>(v.x) - static_cast<float>(first.x);
              ay += static_cast<float>(v.y) - static_cast<float>(first.y);
            }
          }
          const unsigned bal = __ballot_sync(kFull, cd);
          if (bal) {
            const int pos = nc + __popc(bal & ((1u << lane) - 1u));
            if (cd && pos < kCandCap) {
              DRCVAR_ASSERT(pos >= 0 && ((key_of(L) - klo) >> hshift) < static_cast<unsigned long long>(kHistBuckets));
              wcand[pos] = L;
              atomicAdd(&hist[static_cast<unsigned>((key_of(L) - klo) >> hshift)], 1u);
            }
            nc += __popc(bal);
          }
        }
        overflow = nc > kCandCap;
      }
      PH_MARK(5)
      // per-warp partials (fp32 inputs: the per-lane coordinate sums go to the finisher as they are)
      {
        const int wc = __reduce_add_sync(kFull, c_gt + static_cast<int>(cf));
        double pe = 0.0;
        if constexpr (kF32) {
          reinterpret_cast<float2*>(wcand + kCandCap)[lane] = make_float2(ax, ay);
        } else {
          pe = warp_sum_any(s_gt);
        }
        if (lane == 0) {
          const int ncc = nc < kCandCap ? nc : kCandCap;
          ired[warp * 2 + 1] = ncc;
          atomicAdd(&ctl->acc_hi, wc);
          atomicAdd(&ctl->acc_nc, ncc);
          fin[warp * 4 + 2] = pe;
          fin[warp * 4 + 3] = kF32 ? static_cast<double>(wc) : 0.0;
        }
      }
      PH_MARK(6)
      const int ovf = team_sync_or(overflow);  // S3: the sample slot is no longer read on the fast path
      PH_MARK(7)
      const int cnt_hi = ctl->acc_hi, ncand = ctl->acc_nc;
      fast = !ovf && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
      if (fast) {
        if (warp == 0) {
          if (lane == 0) {
            ctl->mode = kModeFinish;
            ctl->cnt_hi = cnt_hi;
            ctl->status = status;
          }
          __syncwarp();
          bar_arrive(kBarFull + par, 64);   // hand halfspace b to the finisher warp
        }
        if (kTail) {
          // parity mode: the sweep team needs T and the total count to emit the tail indices
          mbar_wait(&bars->empty[par], use & 1);
          T_thr = ctl->T;
          c_gt = 0;
        }
      }
    }

    int c_tot = 0;
    if (!fast) {
      need_h();
      if (released && !nonfinite) {
        // rare: the window was placed but missed, and the slot has already been handed back -> fetch halfspace b again
        if (a.bulk) {
          if (b_next < a.B) {   // the director's prefetch of b_next is landing in the slot: let it finish, then overwrite
            mbar_wait(&bars->data0, phase);
            if (copy_bytes > kBulkChunk) mbar_wait(&bars->data, phase);
            phase ^= 1u;
          }
          team_sync();
          if (tid == 0) issue_bulk(b);
          mbar_wait(&bars->data0, phase);
          if (copy_bytes > kBulkChunk) mbar_wait(&bars->data, phase);
          phase ^= 1u;
          redo_bulk = true;
        } else {
          team_sync();
          stage_generic(b);
          team_sync();
        }
      }
      if (!nonfinite) {
        // -------------------------------------------------------------- general path: sweeps over all samples
        status |= kStatusGeneral;
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (int i = tid; i < N; i += kSweepThreads) {
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
        team_sync();  // red[] was consumed by warps 0-1 above; hist/small of this parity are ours until the handoff
        if (lane == 0) {
          kred[warp * 2] = kmin;
          kred[warp * 2 + 1] = kmax;
        }
        team_sync();
#pragma unroll
        for (int w = 0; w < kSweepWarps; ++w) {
          kmin = kred[w * 2] < kmin ? kred[w * 2] : kmin;
          kmax = kred[w * 2 + 1] > kmax ? kred[w * 2 + 1] : kmax;
        }
        T_thr = select_rank(
            [&](auto&& f) {
              for (int i = tid; i < N; i += kSweepThreads) {
                const V2 v = sm[i];
                f(loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y)));
              }
            },
            [] { team_sync(); }, warp == 0, tid, kSweepThreads, kmin, kmax, a.kc, hist, small, ctl);
        c_gt = 0;
        s_gt = 0.0;
        for (int i = tid; i < N; i += kSweepThreads) {
          const V2 v = sm[i];
          const double L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
          if (L > T_thr) {
            ++c_gt;
            s_gt += L;
          }
        }
      } else {
        c_gt = 0;
        s_gt = 0.0;
      }
      // team totals of (c_gt, s_gt)
      const int wc = __reduce_add_sync(kFull, c_gt);
      const double ws = warp_sum_any(s_gt);
      team_sync();
      if (lane == 0) {
        iscr[warp] = wc;
        red[warp] = ws;
      }
      team_sync();
      double s_tot = 0.0;
#pragma unroll
      for (int w = 0; w < kSweepWarps; ++w) {
        c_tot += iscr[w];
        s_tot += red[w];
      }
      if (tid == 0) {
        write_risk_outputs(a, b, ctl, nonfinite, s_tot, c_tot, T_thr, status);
        if (ctl->window_ok && !nonfinite) {   // the window was placed and missed: measure where the threshold really is
          const float ze = (ctl->pl.pm + static_cast<float>(T_thr + ctl->pl.c_shift)) / ctl->pl.sigma;
          ctl->z_est = ze;
          if (++ctl->z_missrun >= 2) ctl->z_learned = 1;
          if (ctl->z_learned && isfinite(ze)) {
            ctl->z_lo_use = ze - a.z_half_adapt_f;
            ctl->z_hi_use = ze + a.z_half_adapt_f;
          }
        }
      }
    } else if (kTail) {
      c_tot = ctl->c_tot;
    }

    // ------------------------------------------------------------------ tail indices (parity mode only)
    if (kTail && a.tail_idx_out != nullptr) {
      int* out = a.tail_idx_out + b * static_cast<long long>(a.kc);
      if (nonfinite) {
        for (int i = tid; i < a.kc; i += kSweepThreads) out[i] = -1;
      } else {
        const int need = a.kc - c_tot;
        int run_eq = 0, run_out = 0;
        int* weq = iscr + kSweepWarps;       // [kSweepWarps]
        int* wsel = iscr + 2 * kSweepWarps;  // [kSweepWarps]
        team_sync();
        for (int bb = 0; bb < N; bb += kSweepThreads) {
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
          team_sync();
          int eq_before = run_eq, tile_eq = 0;
#pragma unroll
          for (int w = 0; w < kSweepWarps; ++w) {
            if (w < warp) eq_before += weq[w];
            tile_eq += weq[w];
          }
          const int eq_rank = eq_before + __popc(meq & ((1u << lane) - 1u));
          const bool sel = gt || (eq && eq_rank < need);
          const unsigned msel = __ballot_sync(kFull, sel);
          if (lane == 0) wsel[warp] = __popc(msel);
          team_sync();
          int out_before = run_out, tile_sel = 0;
#pragma unroll
          for (int w = 0; w < kSweepWarps; ++w) {
            if (w < warp) out_before += wsel[w];
            tile_sel += wsel[w];
          }
          DRCVAR_ASSERT(!sel || out_before + __popc(msel & ((1u << lane) - 1u)) < a.kc);
          if (sel) out[out_before + __popc(msel & ((1u << lane) - 1u))] = i;
          run_eq += tile_eq;
          run_out += tile_sel;
          team_sync();
        }
      }
    }

    PH_MARK(8)
    // ------------------------------------------------------------------ release the slot / keep the pipeline in step
    if (!fast) {
      team_sync();
      if (warp == 0) {
        if (lane == 0) {
          ctl->mode = kModeSkip;   // result already written by the team: the finisher only recycles the buffers
          if (redo_bulk && b_next < a.B) issue_bulk(b_next);   // the re-fetch of b displaced the director's prefetch
        }
        __syncwarp();
        bar_arrive(kBarFull + par, 64);
      }
    }
    if (!released) {
      __syncwarp();
      bar_arrive(kBarSlotFree + par, kSweepThreads + 32);
    }
  }
#ifdef DRCVAR_PROFILE_PHASES
  if (tid == 64 && a.phase_cycles)
    for (int k = 0; k < 12; ++k) a.phase_cycles[(blockIdx.x * 2 + 0) * 12 + k] = ph_t[k];
#endif
}

}  // namespace drcvar
