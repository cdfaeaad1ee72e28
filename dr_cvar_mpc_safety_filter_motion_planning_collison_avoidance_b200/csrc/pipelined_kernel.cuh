// pipelined_kernel.cuh — the resident kernel of halfspace_kernel.cuh with the per-halfspace serial chain taken apart.
//
// Same arithmetic contract, same shared-memory slot, same three roles (8 sweep warps, finisher warp, director warp), two
// CTAs per SM.  What changed is WHO waits for WHOM:
//   * the window placement (≈ 1 500 cycles of one warp) is done by the DIRECTOR warp, before its canonical div/sqrt chain;
//   * the exact phase (2b) of halfspace b-1 is deferred into that gap: a sweep warp goes
//         sweep A(b) -> arrive -> 2b(b-1) -> arrive -> [window(b) ready] -> sweep B(b) -> 2a(b) -> arrive (slot free)
//     so the window placement is hidden behind useful work instead of 7 warps waiting at a barrier;
//   * the sweep warps never synchronise with each other: every hand-over is an arrive on a named barrier that only the
//     helper warp waits on (director: lane sums complete / slot free; finisher: candidates complete).  The finisher adds up
//     the per-warp counts itself and decides whether the window held the threshold;
//   * there is no general path in here: a halfspace whose window could not be placed, missed or overflowed gets its redo
//     flag set (`redo_list[b] = 1`) and the follow-up launch of the streaming kernel computes it (the mechanism of the
//     cluster kernel).  Launches that disable the window, want tail indices, generate samples or use strided views stay
//     on halfspace_kernel.
// fp32 samples, default instantiation <float, 8, true>: sweep B projects and sums the RAW coordinates (thresholds moved into that
// space with a rigorous rounding bound; coordinates too large for raw fp32 sums -> redo pass); <float, 8, false> keeps the
// coordinates relative to the first sample (DRCVAR_FLAG_LARGE_COORDS).  Same h, T and tail set either way.
// Learned window (non-Gaussian samples): the finisher keeps ONE chain of states per CTA, state(i) after its i-th
// halfspace, in a ring of four; the director places window(i) with state(i-3) after waiting for finisher(i-3): a fixed
// lag, hence deterministic.  A miss moves the centre two half-widths towards the side the threshold is on (known from the
// counts), a hit in learned mode tracks (T - mean loss) / sigma.
#pragma once

#include "halfspace_kernel.cuh"
#undef DRCVAR_FILE_ID
#define DRCVAR_FILE_ID 2

namespace drcvar {

struct PHand {                     // what the finisher needs of a halfspace: copied out of PWin at the hand-over
  double h0, h1, hn, f0, f1, t_hi, c_shift;   // (f0, f1): origin of the fp32 "above" sums = first sample, or (0, 0) in raw mode
  unsigned long long key_lo;
  float pm, sigma;
  int hist_shift, degenerate;
  int window_ok, nonfinite;
  int z_learned_used;
  float z_used;
};
static_assert(sizeof(PHand) == 96, "PHand is copied as twelve 8-byte words");
struct PWin {                      // per parity: written by the director (window part, then canonical part)
  PHand hand;
  double t_lo;
  float h0f, h1f, thr_above, thr_keep;
};
struct PZState {                   // learned-window state after the i-th halfspace of this CTA (ring slot i & 3)
  float z_est;
  int learned, missrun, pad;
};
struct PBars {
  unsigned long long data, data0, empty[2], hdone[2], wdone[2];
};
constexpr int kPBarCount = kSweepThreads + 32;   // every named barrier: the 8 sweep warps arrive, one helper warp waits
constexpr int kPOverflowBit = 1 << 30;

// Sweep warps per CTA: 8 (two CTAs per SM: fp32 samples) or 16 (fp64 samples: the 160 KB slot allows ONE CTA per SM, so the
// CTA itself brings the warps that hide the sweeps' latency; per-warp list capacities halve with the per-warp share).
template <int W> struct PCaps {
  static constexpr int kWarpCand = W == 8 ? drcvar::kWarpCand : 80;   // doubles per sweep warp in the candidate buffer ...
  static constexpr int kCandCap = W == 8 ? drcvar::kCandCap : 48;     // ... of which candidate losses (the rest: per-lane sums)
  static constexpr int kWarpList = W == 8 ? drcvar::kWarpList : 80;   // masked samples (raw copies) per sweep warp
};
template <int W = 8>
__host__ __device__ inline size_t pipelined_fixed_smem_bytes(size_t elem_bytes) {
  return sizeof(double) * 2 * PCaps<W>::kWarpCand * W   // cand   [2][warps][kWarpCand]
         + sizeof(unsigned) * 2 * kHistBuckets          // hist   [2][256]
         + sizeof(double) * 2 * (W * 8)                 // red    [2]
         + sizeof(double) * 2 * (W * 4)                 // fin    [2]
         + sizeof(double) * 2 * kResolveMax             // small  [2]
         + sizeof(int) * 2 * 2 * W                      // ired   [2][warps][2]
         + 2 * elem_bytes * PCaps<W>::kWarpList * W     // raw copies of the masked samples [warps][kWarpList]
         + (W == 16 ? sizeof(double) * 2 * 256 : 0)     // xch    (W = 16: slots 256..511 on their way to threads 0..255)
         + 2 * sizeof(PWin) + 2 * sizeof(PHand) + 4 * sizeof(PZState) + sizeof(Ctl) + sizeof(PBars);
}
template <int W> __host__ __device__ constexpr int pipelined_threads() { return W * 32 + 64 + (W == 16 ? 32 : 0); }   // W = 16: + placer warp

// kRawB (fp32 samples): sweep B classifies and sums the RAW coordinates instead of the coordinates relative to the first sample
// (one FADD2 less per sample).  Raw fp32 partial sums are only as good as the coordinates are small, so a halfspace whose
// coordinates are too large for them (see `raw_ok` in the window placement) is handed to the redo pass; callers whose frames
// are far from the origin select the shifted instantiation (DRCVAR_FLAG_LARGE_COORDS), which has no such limit.
// kWords: mask words per thread the instantiation carries (drcvar::kMaskWords = 4 covers every size; 2 covers 64 samples per thread,
// i.e. N <= 16 384 fp32 samples, and leaves the unrolled sweep-B / phase-2a copies of words 2 and 3 out of the instruction stream).
template <typename T, int W = 8, bool kRawB = false, int kWords = drcvar::kMaskWords>
__global__ void __launch_bounds__(pipelined_threads<W>(), W == 8 ? 2 : 1) pipelined_kernel(const KernelArgs a) {
  using V2 = typename Vec2<T>::type;
  constexpr bool kF32 = sizeof(T) == 4;
  constexpr int kMaskWords = kWords;   // (shadows the namespace-level constant)
  static_assert(kWords >= 1 && kWords <= drcvar::kMaskWords, "mask words per thread");
  static_assert(!kRawB || kF32, "raw-coordinate sweep B: fp32 samples");
  static_assert(W == 8 || (W == 16 && !kF32), "16 sweep warps: fp64 samples (one thread per slot of the 512-wide canonical tree)");
  // the names of halfspace_kernel.cuh, for THIS instantiation's team size (they shadow the namespace-level constants)
  constexpr int kSweepWarps = W, kSweepThreads = W * 32, kThreads = pipelined_threads<W>();
  constexpr int kFinisherWarp = W, kDirectorWarp = W + 1, kPlacerWarp = W + 2;   // (placer: W = 16 only)
  constexpr int kPBarCount = kSweepThreads + 32;
  constexpr int kWarpCand = PCaps<W>::kWarpCand, kCandCap = PCaps<W>::kCandCap, kWarpList = PCaps<W>::kWarpList;
  constexpr int kRedDoubles = W * 8, kFinDoubles = W * 4;
  constexpr uint32_t kRowBytes = 16u * kSweepThreads;                 // one 16-byte load per sweep thread
  constexpr int kRowsPerChunk = kBulkChunk / kRowBytes;
  constexpr bool kPlacer = W == 16;                                   // window placement on its own warp, next to the director's chain
  constexpr int kPerLoad = kF32 ? 2 : 1;
  constexpr int kRowSamples = kSweepThreads * kPerLoad;
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
  V2* list_base = reinterpret_cast<V2*>(ired_base + 2 * 2 * kSweepWarps);
  double* xch = reinterpret_cast<double*>(list_base + kWarpList * kSweepWarps);
  PWin* win_base = reinterpret_cast<PWin*>(xch + (W == 16 ? 2 * 256 : 0));
  PHand* hand_base = reinterpret_cast<PHand*>(win_base + 2);
  PZState* zring = reinterpret_cast<PZState*>(hand_base + 2);
  Ctl* fscr = reinterpret_cast<Ctl*>(zring + 4);   // finisher-private scratch of select_rank (dense buckets only)
  PBars* bars = reinterpret_cast<PBars*>(fscr + 1);

  if (tid == 0) {
    mbar_init(&bars->data, 1);
    mbar_init(&bars->data0, 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(&bars->empty[k], 1);
      mbar_init(&bars->hdone[k], 1);
      mbar_init(&bars->wdone[k], 1);
    }
    mbar_fence_init();
    fscr->small_n = 0;
  }
  for (int i = tid; i < 2 * kHistBuckets; i += kThreads) hist_base[i] = 0;
  __syncthreads();

  const uint32_t copy_bytes = static_cast<uint32_t>(static_cast<size_t>(N) * sizeof(V2));
  auto src_of = [&](long long b) {
    return reinterpret_cast<const unsigned char*>(a.samples) + static_cast<size_t>(b) * a.stride_b * sizeof(T);
  };
  auto issue_bulk = [&](long long b) {
    const unsigned char* src = src_of(b);
    const uint32_t n0 = copy_bytes < kBulkChunk ? copy_bytes : kBulkChunk;
    DRCVAR_ASSERT(b >= 0 && b < a.B && copy_bytes <= slot_bytes && (copy_bytes & 15u) == 0u);
    mbar_expect_tx(&bars->data0, n0);
    bulk_g2s(smem_raw, src, n0, &bars->data0);
    if (copy_bytes > n0) {
      mbar_expect_tx(&bars->data, copy_bytes - n0);
#pragma unroll 1
      for (uint32_t off = n0; off < copy_bytes; off += kBulkChunk) {
        const uint32_t n = copy_bytes - off < kBulkChunk ? copy_bytes - off : kBulkChunk;
        bulk_g2s(smem_raw + off, src + off, n, &bars->data);
      }
    }
  };
  const int full_rows = N / kRowSamples;
  const int rows_all = (N + kRowSamples - 1) / kRowSamples;

  // ============================================================================================ finisher warp
  if (warp == kFinisherWarp) {
    int iter = 0;
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
      const int par = iter & 1;
      const PHand* hd = hand_base + par;
      unsigned* hist = hist_base + par * kHistBuckets;
      const double* cand = cand_base + par * kWarpCand * kSweepWarps;
      const double* fin = fin_base + par * kFinDoubles;
      const int* ired = ired_base + par * 2 * kSweepWarps;
      double* small = small_base + par * kResolveMax;
      bar_sync(kBarFull + par, kPBarCount);   // all 8 sweep warps have delivered halfspace b
      // team totals: sure-above count, window candidates, any overflow
      DRCVAR_ASSERT(lane >= kSweepWarps || (ired[lane * 2 + 1] >= 0 && ired[lane * 2 + 1] <= kCandCap));
      const int my_hi = lane < kSweepWarps ? ired[lane * 2] : 0;
      const int my_nc = lane < kSweepWarps ? ired[lane * 2 + 1] : 0;
      const int ovf = __reduce_or_sync(kFull, static_cast<unsigned>(my_hi & kPOverflowBit)) != 0u;
      const int cnt_hi = __reduce_add_sync(kFull, my_hi & (kPOverflowBit - 1));
      const int ncand = __reduce_add_sync(kFull, my_nc);
      const bool placed = hd->window_ok != 0 && hd->nonfinite == 0;
      const bool fast = placed && !ovf && cnt_hi < a.kc && a.kc <= cnt_hi + ncand;
      PZState zs = iter > 0 ? zring[(iter - 1) & 3] : PZState{0.f, 0, 0, 0};
      if (fast) {
        const unsigned long long klo = hd->key_lo;
        const int hshift = hd->hist_shift;
        int bstar, r, cnt_in;
        scan_hist_warp(hist, a.kc - cnt_hi, lane, bstar, r, cnt_in);
        double s3 = 0.0;
        int c3 = 0, n_small = 0;
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
          s4 = mineAbove ? small[lane] : 0.0;
          c4 = __popc(__ballot_sync(kFull, mineAbove));
        } else {
          // dense / heavily tied bucket: narrow further inside the finisher warp
          const unsigned long long lo2 = klo + (static_cast<unsigned long long>(bstar) << hshift);
          unsigned long long hi2 = lo2 + ((1ull << hshift) - 1ull);
          const unsigned long long khi = key_of(hd->t_hi);
          if (hi2 > khi || hi2 < lo2) hi2 = khi;
          auto each = [&](auto&& f) {
            for (int w = 0; w < kSweepWarps; ++w) {
              const int nc = ired[w * 2 + 1];
              for (int j = lane; j < nc; j += 32) f(cand[w * kWarpCand + j]);
            }
          };
          T_thr = select_rank(each, [] { __syncwarp(); }, true, lane, 32, lo2, hi2, r, hist, small, fscr);
          each([&](double L) {
            const int bk = static_cast<int>((key_of(L) - klo) >> hshift);
            if (bk == bstar && L > T_thr) {
              ++c4;
              s4 += L;
            }
          });
          c4 = __reduce_add_sync(kFull, c4);
          if (lane == 0) fscr->small_n = 0;
        }
        const int c3t = __reduce_add_sync(kFull, c3);
        double lx = 0.0, ly = 0.0;
        if constexpr (kF32) {
          double lx2 = 0.0, ly2 = 0.0;
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
        warp_sum_any4d(s3t, s4, s_x, s_y, lane);
        if (lane == 0) {
          double s_e = 0.0, n_lin = static_cast<double>(cnt_hi);
          if constexpr (!kF32) {
            n_lin = 0.0;
#pragma unroll 1
            for (int w = 0; w < kSweepWarps; ++w) s_e += fin[w * 4 + 2];
          }
          // fp32 samples: the "above" set through linearity; its coordinate sums are relative to (f0, f1) = the first sample,
          // or to the origin when sweep B ran on raw coordinates
          const double s_lin = -(hd->h0 * (n_lin * hd->f0 + s_x) + hd->h1 * (n_lin * hd->f1 + s_y));
          const double s_tot = ((s_e + s_lin) + s3t) + s4;
          const int c_tot = cnt_hi + c3t + c4;
          write_risk_outputs(a, b, hd, false, s_tot, c_tot, T_thr, hd->degenerate ? kStatusDegenerate : 0, hd->hn);
          zs.missrun = 0;
          if (zs.learned) {   // track where the threshold sits: (T - mean loss) / sigma = (pm + T + c) / sigma
            const float zT = (hd->pm + static_cast<float>(T_thr + hd->c_shift)) / hd->sigma;
            const float ze = 0.5f * (zs.z_est + zT);
            if (isfinite(ze) && fabsf(ze) < 8.f) zs.z_est = ze;
          }
        }
      } else if (lane == 0) {
        DRCVAR_ASSERT(b >= 0 && b < a.B);
        a.redo_list[b] = 1;   // the streaming kernel's redo pass computes this halfspace
        if (placed && !ovf) {
          // a placed window that missed: after two in a row (or already in learned mode) move the centre past the window,
          // towards the side the threshold is on
          const int run = ++zs.missrun;
          if (run >= 2 || zs.learned) {
            const float z_used = hd->z_learned_used ? hd->z_used : a.z_mid_f;
            const float half_used = hd->z_learned_used ? a.z_half_adapt_f : a.z_half_f;
            const float z_new = z_used + (a.kc <= cnt_hi ? 2.0f : -2.0f) * half_used;
            if (isfinite(z_new) && fabsf(z_new) < 8.f) {
              zs.z_est = z_new;
              zs.learned = 1;
            }
          }
        }
      }
      if (lane == 0) zring[iter & 3] = zs;
      for (int i = lane; i < kHistBuckets; i += 32) hist[i] = 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->empty[par]);
    }
    return;
  }

  // ============================================================================================ director warp
  // window placement first (the team is waiting for it), then the canonical direction and the mean halfspace
  if (warp == kDirectorWarp || (kPlacer && warp == kPlacerWarp)) {
    const bool do_window = !kPlacer || warp == kPlacerWarp;   // no placer warp: the director does both, window first
    const bool do_canon = warp == kDirectorWarp;
    constexpr int kADoneCount = kPBarCount + (kPlacer ? 32 : 0);
    const double inv_n = 1.0 / static_cast<double>(N);
    double inv_sub = inv_n;   // 1 / (#samples in the second moments): all samples (fp32) / every 4th row (fp64)
    if (!kF32) {
      const int r4 = (rows_all + 3) / 4;
      const int last = (r4 - 1) * 4 * kRowSamples;
      const int n_sub0 = (r4 - 1) * kSweepThreads + (N - last < kSweepThreads ? N - last : kSweepThreads);
      inv_sub = 1.0 / static_cast<double>(n_sub0 > 0 ? n_sub0 : 1);
    }
    const float inv_sub_f = static_cast<float>(inv_sub);
    int iter = 0;
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
      const int par = iter & 1;
      PWin* win = win_base + par;
      const double* red = red_base + par * kRedDoubles;
      if (lane == 0 && do_canon) {
        const long long b_pf = b + gridDim.x;
        if (b_pf < a.B) bulk_prefetch_l2(src_of(b_pf), copy_bytes & ~15u);
      }
      double pre0 = 0.0, pre1 = 0.0;   // ego (or the explicit normal) of halfspace b, fetched early
      if (a.h_in != nullptr) {
        pre0 = a.h_in[2 * b];
        pre1 = a.h_in[2 * b + 1];
      } else if (a.ego != nullptr) {
        pre0 = a.ego[2 * b];
        pre1 = a.ego[2 * b + 1];
      }
      // learned-window state with a fixed lag of three halfspaces (deterministic)
      PZState zs{0.f, 0, 0, 0};
      if (do_window && iter >= 3) {
        mbar_wait(&bars->empty[(iter - 3) & 1], ((iter - 3) >> 1) & 1);
        zs = zring[(iter - 3) & 3];
      }
      bar_sync(kBarADone + par, kADoneCount);   // red[par] is complete
      // ---------------------------------------------------------------- window placement (speed only, never the result)
      double w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double t[8];   // the 8 group totals of the 256-wide canonical tree (sweep warps 0..7 in both team sizes)
#pragma unroll
        for (int g = 0; g < 8; ++g) t[g] = red[g * 8 + j];
#pragma unroll
        for (int n = 8; n > 1; n >>= 1)
#pragma unroll
          for (int g = 0; g < n / 2; ++g) t[g] = __dadd_rn(t[2 * g], t[2 * g + 1]);  // adjacent-pair tree (canonical)
        w[j] = t[0];
      }
      const double f0 = red[6], f1 = red[7];   // first sample (warp 0's record)
      // ---------------------------------------------------------------- canonical direction (IEEE div / sqrt chain)
      double m0 = 0.0, m1 = 0.0;
      auto canonical_chain = [&]() {
        m0 = ddiv_canon(w[0], static_cast<double>(N));
        m1 = ddiv_canon(w[1], static_cast<double>(N));
        if constexpr (kF32) {  // fp32 inputs: the lane sums were taken relative to the first sample
          m0 = __dadd_rn(f0, m0);
          m1 = __dadd_rn(f1, m1);
        }
        int nonfinite = !(isfinite(m0) && isfinite(m1));
        int degenerate = 0;
        double h0, h1;
        if (a.h_in != nullptr) {
          h0 = pre0;
          h1 = pre1;
        } else {
          const double d0 = __dsub_rn(m0, pre0), d1 = __dsub_rn(m1, pre1);
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
        const double hn = norm2_canon(h0, h1);
        if (lane == 0) {
          win->hand.h0 = h0; win->hand.h1 = h1;
          win->hand.hn = hn;
          win->hand.nonfinite = nonfinite;
          win->hand.degenerate = degenerate;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->hdone[par]);
      };
      if (!kF32 && !kPlacer) canonical_chain();  // fp64 samples classify with the canonical direction: it goes first
      if (do_window) {
        const float* redf = reinterpret_cast<const float*>(red);   // warp g: floats 4..9 of its 16 = qxx,qyy,qxy,bound,mdx,mdy
        float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        float b2 = 0.f;
#pragma unroll
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
            const float mag = static_cast<float>(fabs(m0) + fabs(m1) + fabs(pre0) + fabs(pre1));
            err_h = 1e-6f + 4e-15f * mag * rn;
            usable = usable && isfinite(rn) && rn > 0.f && isfinite(mag);
          }
        }
        usable = usable && isfinite(h0f) && isfinite(h1f) && err_h < 1e-3f;
        int n_sub_i;
        float ex, ey;
        if (kF32) {
          n_sub_i = N;
          ex = mr0f;
          ey = mr1f;
        } else {
          const int r4 = (rows_all + 3) / 4;
          const int last = (r4 - 1) * 4 * kRowSamples;
          n_sub_i = (r4 - 1) * kSweepThreads + (N - last < kSweepThreads ? N - last : kSweepThreads);
          ex = q[3] * inv_sub_f;
          ey = q[4] * inv_sub_f;
        }
        const float cxx = q[0] * inv_sub_f - ex * ex, cyy = q[1] * inv_sub_f - ey * ey, cxy = q[2] * inv_sub_f - ex * ey;
        const float var_l = h0f * h0f * cxx + 2.0f * h0f * h1f * cxy + h1f * h1f * cyy;
        const float sigma = sqrt_approx(var_l);   // placement only
        int window_ok = a.use_window && usable && (n_sub_i >= 256) && (var_l > 0.f) && isfinite(sigma) &&
                        (rows_all * kPerLoad <= 32 * kMaskWords);
        const float pm = fmaf(h1f, mr1f, h0f * mr0f);
        float zlo = a.z_lo_f, zhi = a.z_hi_f;
        if (zs.learned) {
          zlo = zs.z_est - a.z_half_adapt_f;
          zhi = zs.z_est + a.z_half_adapt_f;
        }
        const float a_lo = pm - zlo * sigma, a_hi = pm - zhi * sigma;
        const double c = static_cast<double>(h0f) * f0 + static_cast<double>(h1f) * f1;   // h_a . first
        const double t_lo = __dadd_rn(-static_cast<double>(a_lo) - c, 0.0);  // +0.0: never -0.0 (canonical losses are +0)
        const double t_hi = __dadd_rn(-static_cast<double>(a_hi) - c, 0.0);
        // rigorous fp32 classification bounds: see halfspace_kernel.cuh (window placement of warp 0), same expressions
        const float dmax = sqrt_approx(b2) * 1.0001f;
        const float af0 = fabsf(static_cast<float>(f0)) * 1.0001f, af1 = fabsf(static_cast<float>(f1)) * 1.0001f;
        const float habs = fabsf(h0f) + fabsf(h1f);
        const float eps = (habs * (af0 + af1 + dmax)) * 1e-15f + err_h * 1.5f * (af0 + af1 + 2.0f * dmax);
        float thr_keep, thr_above;
        bool raw_ok = true;
        if constexpr (kRawB) {
          // Sweep B classifies p = fma(h1f, y, h0f * x) of the RAW coordinates and sums them as they are.
          //   |p - h_a.xi| <= 2^-24 (|h0f x| + |p|) <= 2^-23 habs vmax (1 + 2^-24), vmax = |first| + dmax; taken twice over.
          //   The window edges move into that space through c = h_a . first (double) and one rounding to fp32 (2^-24 |edge|;
          //   2^-21 |edge| is allowed for it and for the additions below).
          // The per-thread sums of the "above" set are chains of m = kc / 256 fp32 adds, each within 2^-24 of a partial sum
          // <= m vmax; the finisher adds the 256 of them in fp64: the CVaR moves by about 2^-24 (m vmax / 2) / sqrt(kc).  With
          // m vmax <= 256 that is 2.4e-7 at kc = 1 000 (1e-8 typical at |xi| < 8, config 4).  Larger coordinates: redo pass.
          const float vmax = af0 + af1 + dmax;
          const float m_adds = fmaxf(1.0f, static_cast<float>(a.kc) * (1.0f / kSweepThreads));
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
        const unsigned long long span = khi - klo;
        const int bits = span ? 64 - __clzll(static_cast<long long>(span)) : 0;
        window_ok = window_ok && isfinite(thr_keep) && isfinite(thr_above) && (khi >= klo) && (thr_above <= thr_keep) &&
                    isfinite(t_lo) && isfinite(t_hi) && raw_ok;
        if (lane == 0) {
          win->hand.f0 = kRawB ? 0.0 : f0; win->hand.f1 = kRawB ? 0.0 : f1;
          win->t_lo = t_lo;
          win->hand.t_hi = t_hi;
          win->h0f = h0f; win->h1f = h1f; win->thr_keep = thr_keep; win->thr_above = thr_above;
          win->hand.key_lo = klo;
          win->hand.hist_shift = bits > 8 ? bits - 8 : 0;
          win->hand.window_ok = window_ok;
          win->hand.pm = pm; win->hand.sigma = sigma; win->hand.c_shift = c;
          win->hand.z_learned_used = zs.learned;
          win->hand.z_used = zs.z_est;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->wdone[par]);
      }
      if (!do_canon) continue;   // (placer warp: done with this halfspace)
      if (kF32 || kPlacer) canonical_chain();   // fp32 samples: after the window (the team classifies with the fp32 direction)
      // TMA producer: as soon as all 8 sweep warps are done with the sample slot, fetch the next halfspace
      bar_sync(kBarSlotFree + par, kPBarCount);
      {
        const long long b_next = b + gridDim.x;
        if (lane == 0 && b_next < a.B) issue_bulk(b_next);
      }
      if (lane == 0) write_mean_outputs(a, b, m0, m1);
    }
    return;
  }

  // ============================================================================================ sweep team
  if (tid == 0 && static_cast<long long>(blockIdx.x) < a.B) issue_bulk(blockIdx.x);
  uint32_t phase = 0;
  const long long n_it = (a.B - static_cast<long long>(blockIdx.x) + gridDim.x - 1) / gridDim.x;
  V2* wlist = list_base + warp * kWarpList;
  const uint32_t wlist_s = smem_u32(wlist), tslot_s = smem_u32(smem_raw) + 16u * static_cast<uint32_t>(tid);   // shared-space addresses (phase 2a)
  // state of the previous halfspace carried into its deferred exact phase
  int n_list_prev = -1;                         // -1: nothing usable (window not placed / list overflow)
  float ax_prev = 0.f, ay_prev = 0.f, cf_prev = 0.f;
  int c_gt_prev = 0;                            // fp64 inputs: exact-classified losses above the window
  double s_gt_prev = 0.0;

  for (long long it = 0; it <= n_it; ++it) {
    const int par = static_cast<int>(it & 1);
    V2 first;           // first sample of halfspace `it` (shift origin), valid once its data has landed
    first.x = 0;
    first.y = 0;
    // ------------------------------------------------------------------ sweep A of halfspace it
    if (it < n_it) {
      double* red = red_base + par * kRedDoubles;
      mbar_wait(&bars->data0, phase);
      bool rest_pending = true;
      auto wait_rest = [&]() {
        if (rest_pending) {
          if (copy_bytes > kBulkChunk) mbar_wait(&bars->data, phase);
          phase ^= 1u;
          rest_pending = false;
        }
      };
      first = sm[0];
      double u_x, u_y;
      double q_xx, q_yy, q_xy;
      double q_dx = 0.0, q_dy = 0.0;
      float bound2 = 0.f;
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
      } else if constexpr (W == 16) {
        // 512 sweep threads: thread t IS slot t of the canonical 512-slot sum (samples i = t mod 512 in increasing i: one
        // chain per coordinate); slots 256..511 then travel through shared memory to threads 0..255 (u[j] = s[j] + s[j+256])
        double s0 = 0.0, s1 = 0.0;
        q_xx = q_yy = q_xy = 0.0;
        auto acc_row = [&](const V2 v, bool mom) {
          s0 = __dadd_rn(s0, v.x);
          s1 = __dadd_rn(s1, v.y);
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
            for (int k = 0; k < 4; ++k) acc_row(v[k], k == 0);
          }
          for (; r < r_hi; ++r) {
            const int i = r * kRowSamples + tid;
            if (i < N) acc_row(sm[i], (r & 3) == 0);
          }
        };
        const int r_first = rows_all < kRowsPerChunk ? rows_all : kRowsPerChunk;
        rows(0, r_first);
        wait_rest();
        rows(r_first, rows_all);
        if (tid >= 256) {
          xch[2 * (tid - 256)] = s0;
          xch[2 * (tid - 256) + 1] = s1;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSweepThreads) : "memory");   // the one barrier inside this team (xch is reused a halfspace later)
        u_x = tid < 256 ? __dadd_rn(s0, xch[2 * tid]) : 0.0;
        u_y = tid < 256 ? __dadd_rn(s1, xch[2 * tid + 1]) : 0.0;
      } else {
        double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
        q_xx = q_yy = q_xy = 0.0;
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
          for (; r < r_hi; ++r) {
            const int i = r * kRowSamples + tid;
            if (i < N) acc_row(sm[i], (r & 1) != 0, (r & 3) == 0);
          }
        };
        const int r_first = rows_all < kRowsPerChunk ? rows_all : kRowsPerChunk;
        rows(0, r_first);
        wait_rest();
        rows(r_first, rows_all);
        u_x = __dadd_rn(s00, s10);
        u_y = __dadd_rn(s01, s11);
      }
      wait_rest();
      {
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
          w[6] = static_cast<double>(first.x);   // (only warp 0's record is read)
          w[7] = static_cast<double>(first.y);
        }
      }
      __syncwarp();
      bar_arrive(kBarADone + par, kPBarCount + (kPlacer ? 32 : 0));   // director (and placer) take over from here
    }

    // ------------------------------------------------------------------ deferred exact phase (2b) of halfspace it-1
    if (it > 0) {
      const int q = par ^ 1;
      const int useq = static_cast<int>((it - 1) >> 1);
      const PWin* win = win_base + q;
      unsigned* hist = hist_base + q * kHistBuckets;
      double* wcand = cand_base + q * kWarpCand * kSweepWarps + warp * kWarpCand;
      double* fin = fin_base + q * kFinDoubles;
      int* ired = ired_base + q * 2 * kSweepWarps;
      if (useq > 0) mbar_wait(&bars->empty[q], (useq - 1) & 1);   // finisher is done with this parity's buffers
      mbar_wait(&bars->hdone[q], useq & 1);                        // canonical h (long done)
      if (warp == 0) {   // hand-over record for the finisher
        if (lane < 12)
          reinterpret_cast<unsigned long long*>(hand_base + q)[lane] = reinterpret_cast<const unsigned long long*>(&win->hand)[lane];
      }
      const double h0 = win->hand.h0, h1 = win->hand.h1;
      const double t_lo = win->t_lo, t_hi = win->hand.t_hi;
      const unsigned long long klo = win->hand.key_lo;
      const int hshift = win->hand.hist_shift;
      const float pfx = static_cast<float>(win->hand.f0), pfy = static_cast<float>(win->hand.f1);   // origin of it-1's "above" sums
      bool overflow = n_list_prev < 0;
      int nc = 0;
      float ax = ax_prev, ay = ay_prev, cf = cf_prev;
      if (!overflow) {
        for (int k0 = 0; k0 < n_list_prev; k0 += 32) {
          const int k = k0 + lane;
          const bool active = k < n_list_prev;
          double L = 0.0;
          V2 v;
          v.x = 0;
          v.y = 0;
          if (active) {
            DRCVAR_ASSERT(k < kWarpList);
            v = wlist[k];
            L = loss_of(h0, h1, static_cast<double>(v.x), static_cast<double>(v.y));
          }
          const bool up = active && (L > t_hi);
          const bool cd = active && !up && (L >= t_lo);
          if constexpr (kF32) {
            if (up) {  // inside the fp32 uncertainty band but exactly above the window: joins the "above" set
              cf += 1.0f;
              ax += static_cast<float>(v.x) - pfx;
              ay += static_cast<float>(v.y) - pfy;
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
      {
        const int wc = __reduce_add_sync(kFull, c_gt_prev + static_cast<int>(cf));
        double pe = 0.0;
        if constexpr (kF32) {
          reinterpret_cast<float2*>(wcand + kCandCap)[lane] = make_float2(ax, ay);
        } else {
          pe = warp_sum_any(s_gt_prev);
        }
        if (lane == 0) {
          ired[warp * 2] = wc | (overflow ? kPOverflowBit : 0);
          ired[warp * 2 + 1] = nc < kCandCap ? nc : kCandCap;
          fin[warp * 4 + 2] = pe;
        }
      }
      __syncwarp();
      bar_arrive(kBarFull + q, kPBarCount);   // the finisher takes halfspace it-1 from here
    }

    // ------------------------------------------------------------------ sweep B + phase 2a of halfspace it
    if (it < n_it) {
      const PWin* win = win_base + par;
      const int use = static_cast<int>(it >> 1);
      mbar_wait(&bars->wdone[par], use & 1);
      n_list_prev = -1;
      ax_prev = ay_prev = cf_prev = 0.f;
      c_gt_prev = 0;
      s_gt_prev = 0.0;
      if (win->hand.window_ok) {
        unsigned mask[kMaskWords];
#pragma unroll
        for (int w2 = 0; w2 < kMaskWords; ++w2) mask[w2] = 0u;
        float ax = 0.f, ay = 0.f, cf = 0.f;
        int c_gt = 0;
        double s_gt = 0.0;
        constexpr int kRowsPerWord = 32 / kPerLoad;
        if constexpr (kF32) {
          const float4* sm4 = reinterpret_cast<const float4*>(smem_raw);
          const float h0f = win->h0f, h1f = win->h1f, thr_keep = win->thr_keep, thr_above = win->thr_above;
          // kRawB: coordinates as they are (thresholds are in that space); otherwise relative to the first sample
          const float2 nf = make_float2(-first.x, -first.y);
          auto shifted = [&](float x, float y) { return kRawB ? make_float2(x, y) : __fadd2_rn(make_float2(x, y), nf); };
#pragma unroll
          for (int wd = 0; wd < kMaskWords; ++wd) {   // all complete rows (groups of four, then the one to three left over)
            const int r_lo = wd * kRowsPerWord;
            const int r_hi = full_rows < r_lo + kRowsPerWord ? full_rows : r_lo + kRowsPerWord;
            unsigned bit = 1u;
#pragma unroll 4
            for (int r = r_lo; r < r_hi; ++r) {
              const float4 v = sm4[r * kSweepThreads + tid];
              const float2 d0 = shifted(v.x, v.y), d1 = shifted(v.z, v.w);
              const float p0 = fmaf(h1f, d0.y, h0f * d0.x), p1 = fmaf(h1f, d1.y, h0f * d1.x);
              classify_f32(p0, thr_above, thr_keep, d0.x, d0.y, ax, ay, cf, mask[wd], bit);
              classify_f32(p1, thr_above, thr_keep, d1.x, d1.y, ax, ay, cf, mask[wd], bit + bit);
              bit <<= 2;
            }
          }
          if (full_rows < rows_all) {   // the ragged last row: samples beyond N are neither above nor kept
            const float4 v = sm4[full_rows * kSweepThreads + tid];
            const int i0 = full_rows * kRowSamples + 2 * tid;
            const float2 d0 = shifted(v.x, v.y), d1 = shifted(v.z, v.w);
            float p0 = fmaf(h1f, d0.y, h0f * d0.x), p1 = fmaf(h1f, d1.y, h0f * d1.x);
            if (i0 >= N) p0 = __int_as_float(0x7f800000);
            if (i0 + 1 >= N) p1 = __int_as_float(0x7f800000);
            unsigned mk = 0;
            const unsigned bit = 1u << ((2 * full_rows) & 31);
            classify_f32(p0, thr_above, thr_keep, d0.x, d0.y, ax, ay, cf, mk, bit);
            classify_f32(p1, thr_above, thr_keep, d1.x, d1.y, ax, ay, cf, mk, bit + bit);
            const int wg = (2 * full_rows) >> 5;
#pragma unroll
            for (int w2 = 0; w2 < kMaskWords; ++w2) mask[w2] |= (w2 == wg) ? mk : 0u;
          }
        } else {
          mbar_wait(&bars->hdone[par], use & 1);   // fp64 inputs classify with the canonical direction
          const double h0 = win->hand.h0, h1 = win->hand.h1, t_lo = win->t_lo, t_hi = win->hand.t_hi;
#pragma unroll
          for (int wd = 0; wd < kMaskWords; ++wd) {
            const int r_lo = wd * kRowsPerWord;
            const int r_hi = rows_all < r_lo + kRowsPerWord ? rows_all : r_lo + kRowsPerWord;
            unsigned bit = 1u;
            auto one = [&](const V2 v, unsigned bt) {
              const double L = loss_of(h0, h1, v.x, v.y);
              classify_f64(L, t_hi, t_lo, s_gt, c_gt, mask[wd], bt);
            };
            int r = r_lo;
            const int g_hi = r_lo + (((r_hi < full_rows ? r_hi : full_rows) - r_lo) & ~3);
#pragma unroll 1
            for (; r < g_hi; r += 4, bit <<= 4) {
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
        }
        // ---------------------------------------------------------------- phase 2a: compact the masked samples per warp
        int mine_n = 0;
#pragma unroll
        for (int wd = 0; wd < kMaskWords; ++wd)
          if (wd * 32 < rows_all * kPerLoad) mine_n += __popc(mask[wd]);
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
        if (kF32 && n_list <= kWarpList) {
          // fp32 samples: 32-bit shared-space addresses; bit P = 2 r + e of a word <-> byte (P >> 1) * kRowBytes + (P & 1) * 8
          static_assert(!kF32 || kRowBytes == 4096u, "offset arithmetic of the fp32 compaction loop");
          uint32_t dst_s = wlist_s + 8u * static_cast<uint32_t>(excl);
#pragma unroll
          for (int wd = 0; wd < kMaskWords; ++wd) {
            if (wd * 32 < rows_all * kPerLoad) {
              unsigned mm = mask[wd];
              const uint32_t wb = tslot_s + 16u * kRowBytes * wd;
              while (mm) {
                const unsigned bp = 31u - static_cast<unsigned>(__clz(static_cast<int>(mm)));
                mm ^= 1u << bp;
                const uint32_t off = ((bp << 11) & 0xF000u) | ((bp << 3) & 8u);
                DRCVAR_ASSERT(dst_s + 8u <= wlist_s + 8u * kWarpList && (wb - tslot_s + 16u * tid + off) + 8u <= slot_bytes);
                smem_copy8(dst_s, wb + off);
                dst_s += 8u;
              }
            }
          }
          n_list_prev = n_list;
        } else if (n_list <= kWarpList) {
          V2* dst = wlist + excl;
          const unsigned char* tbase = smem_raw + 16u * tid;
#pragma unroll
          for (int wd = 0; wd < kMaskWords; ++wd) {
            if (wd * 32 < rows_all * kPerLoad) {
              unsigned mm = mask[wd];
              const unsigned char* wbase = tbase + (kF32 ? 16u : 32u) * kRowBytes * wd;
              while (mm) {
                const unsigned bp = 31u - static_cast<unsigned>(__clz(static_cast<int>(mm)));
                mm ^= 1u << bp;
                const unsigned off = kF32 ? ((bp >> 1) * kRowBytes + ((bp & 1u) << 3)) : bp * kRowBytes;
                DRCVAR_ASSERT(dst < wlist + kWarpList && (wbase + off) + sizeof(V2) <= smem_raw + slot_bytes);
                *dst++ = *reinterpret_cast<const V2*>(wbase + off);
              }
            }
          }
          n_list_prev = n_list;
        }
        ax_prev = ax;
        ay_prev = ay;
        cf_prev = cf;
        c_gt_prev = c_gt;
        s_gt_prev = s_gt;
      }
      __syncwarp();
      bar_arrive(kBarSlotFree + par, kPBarCount);   // the slot can be refilled
    }
  }
}

}  // namespace drcvar
