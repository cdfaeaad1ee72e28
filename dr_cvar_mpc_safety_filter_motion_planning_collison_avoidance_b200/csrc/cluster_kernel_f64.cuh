// cluster_kernel_f64.cuh — the cluster / DSMEM kernel (cluster_kernel.cuh) for fp64 samples, the reference's native
// dtype: one thread-block cluster of 2 / 4 / 8 CTAs per halfspace, every sample read from HBM once (N = 100 000 fp64 =
// 1.6 MB per halfspace: 8 CTAs x 200 KB).  Same skeleton — producer warp + TMA chunks handed back during sweep B, 16 sweep
// warps (thread t = slot t of the canonical tree), director warp, rotating leader with a finisher warp, st.async
// exchanges — but the arithmetic is the fp64 contract throughout:
//   sweep A   canonical fp64 slot sums per octant (sequential DADD chain per slot); second moments on every 4th row
//   window    needs the CANONICAL direction (director's IEEE div / sqrt chain): t_lo / t_hi in fp64 from the moments
//   sweep B   exact canonical loss L_i = -(h.xi_i) of every sample: L > t_hi -> count + exact sum, t_lo <= L <= t_hi ->
//             per-warp candidate list (ballot prefix); no screening, no phase 2
//   exchange2 / finish as in the fp32 kernel; CVaR from exact partial sums (fixed reduction trees: run-to-run identical)
// Misses / overflow / non-finite data go to the streaming kernel through the redo list.
#pragma once

#include "cluster_kernel.cuh"

namespace drcvar {

__device__ __forceinline__ double2 lds128d(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(kClThreads, 1) cluster_kernel_f64(const KernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const int C = a.cl_ctas, O = kOctants / C;
  const int lgO = 31 - __clz(O);
  const uint32_t rank = cluster_ctarank();
  const long long q = blockIdx.x / C, n_clusters = gridDim.x / C;
  const uint32_t oct_b = static_cast<uint32_t>(octant_bytes(N, 8));
  const uint32_t row_b = static_cast<uint32_t>(N) * 16u;
  const uint32_t part_cap = static_cast<uint32_t>(O) * oct_b;
  const uint32_t part_lo = rank * part_cap;
  const uint32_t part_b = row_b > part_lo ? (row_b - part_lo < part_cap ? row_b - part_lo : part_cap) : 0u;
  const int n_chunks = static_cast<int>((part_b + kBulkChunk - 1) / kBulkChunk);
  const size_t slot_bytes = (static_cast<size_t>(part_cap) + 127) & ~static_cast<size_t>(127);
  ClShared* sh = reinterpret_cast<ClShared*>(smem_raw + slot_bytes);
  const int cap = kClPool / C;

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
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();

  // ============================================================================================ producer warp
  if (warp == kClProducerWarp) {
    if (lane == 0) {
      int it = 0;
      for (long long b = q; b < a.B; b += n_clusters, ++it) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.samples) +
                                   static_cast<size_t>(b) * a.stride_b * sizeof(double) + part_lo;
        for (int j = 0; j < n_chunks; ++j) {
          if (it > 0) mbar_wait(&sh->free_[j], (it - 1) & 1);
          const uint32_t off = static_cast<uint32_t>(j) * kBulkChunk;
          const uint32_t n = part_b - off < kBulkChunk ? part_b - off : kBulkChunk;
          mbar_expect_tx(&sh->full[j], n);
          bulk_g2s(smem_raw + off, src + off, n, &sh->full[j]);
        }
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
      bar_sync(kClBarDirector + par, 64);   // warp 0 of the team has seen exchange 1 complete
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
  if (warp == kClFinisherWarp) {
    int it = 0;
    uint32_t n_lead = 0;
    for (long long b = q; b < a.B; b += n_clusters, ++it) {
      if (it % C != static_cast<int>(rank)) continue;
      while (!mbar_try_wait(&sh->xbar2, n_lead & 1u)) __nanosleep(400);
      ++n_lead;
      Ctl* fc = &sh->fin_ctl;
      bool fast = fc->window_ok != 0 && fc->nonfinite == 0;
      double n_above = 0.0, s_hi = 0.0;
      int ncand = 0;
      if (fast) {
        for (int s = 0; s < C; ++s) {
          n_above += sh->x2[s][0];
          s_hi += sh->x2[s][1];
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
        if (lane == 0)
          write_risk_outputs(a, b, fc, false, s_hi + s4, cnt_hi + c4, T_thr, fc->degenerate ? kStatusDegenerate : 0);
      } else if (lane == 0) {
        a.redo_list[b] = 1;   // redo flag of halfspace b
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->fdone);
    }
    return;
  }

  // ============================================================================================ sweep team
  const uint32_t slot_s = smem_u32(smem_raw);
  const uint32_t toff = 16u * tid;
  const uint32_t woff = 16u * (tid & ~31);
  double* wcand = reinterpret_cast<double*>(sh->list[warp]);   // kClWarpList doubles per warp
  const unsigned lt_mask = (1u << lane) - 1u;
  int it = 0;
  uint32_t n_lead = 0;
  for (long long b = q; b < a.B; b += n_clusters, ++it) {
    const int par = it & 1;
    const int leader = it % C;
    Ctl* ctl = &sh->ctl[par];
    const double2 first = __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const double*>(a.samples) + b * a.stride_b));
    const uint32_t fpar = it & 1;
    int have = 0;
    auto wait_upto = [&](uint32_t byte_off) {
      const int c = static_cast<int>(byte_off >> 15);
      while (have <= c) {
        mbar_wait_spin(&sh->full[have], fpar);
        ++have;
      }
    };

    // ------------------------------------------------------------------ sweep A: canonical fp64 slot sums per octant
    double qxx = 0.0, qyy = 0.0, qxy = 0.0, qdx = 0.0, qdy = 0.0;
    int n_sub = 0;
    for (int k = 0; k < O; ++k) {
      const uint32_t ob = static_cast<uint32_t>(k) * oct_b;
      const uint32_t oe = part_b < ob + oct_b ? part_b : ob + oct_b;
      const uint32_t len = oe > ob ? oe - ob : 0u;
      double s_x = 0.0, s_y = 0.0;
      auto body = [&](const double2 v, uint32_t lr) {
        s_x = __dadd_rn(s_x, v.x);
        s_y = __dadd_rn(s_y, v.y);
        if ((lr & 3u) == 0u) {   // second moments on every 4th row (window placement only)
          const double dx = v.x - first.x, dy = v.y - first.y;
          qdx += dx;
          qdy += dy;
          qxx = fma(dx, dx, qxx);
          qyy = fma(dy, dy, qyy);
          qxy = fma(dx, dy, qxy);
          ++n_sub;
        }
      };
      const uint32_t n_lr = (len + kClLaneRow - 1) / kClLaneRow;
      uint32_t lr = 0;
      for (; (lr + 4) * kClLaneRow <= len; lr += 4) {
        const uint32_t base = ob + lr * kClLaneRow + toff;
        wait_upto(base + 3 * kClLaneRow);
        const double2 v0 = lds128d(slot_s + base), v1 = lds128d(slot_s + base + kClLaneRow),
                      v2 = lds128d(slot_s + base + 2 * kClLaneRow), v3 = lds128d(slot_s + base + 3 * kClLaneRow);
        body(v0, lr);
        body(v1, lr + 1);
        body(v2, lr + 2);
        body(v3, lr + 3);
      }
      for (; lr < n_lr; ++lr) {
        const uint32_t off = ob + lr * kClLaneRow + toff;
        if (off + 16 <= oe) {
          wait_upto(off);
          body(lds128d(slot_s + off), lr);
        }
      }
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
      const double r0 = warp_sum_any(qxx), r1 = warp_sum_any(qyy), r2 = warp_sum_any(qxy), r3 = warp_sum_any(qdx),
                   r4 = warp_sum_any(qdy);
      const int rn = __reduce_add_sync(kFull, n_sub);
      if (lane == 0) {
        double* m = sh->dmom + warp * 6;
        m[0] = r0; m[1] = r1; m[2] = r2; m[3] = r3; m[4] = r4; m[5] = static_cast<double>(rn);
      }
    }
    cl_team_sync();
    // ------------------------------------------------------------------ exchange 1
    if (warp == 0) {
      double mq[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (lane < kClTeamWarps) {
#pragma unroll
        for (int i = 0; i < 6; ++i) mq[i] = sh->dmom[lane * 6 + i];
      }
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1)
#pragma unroll
        for (int i = 0; i < 6; ++i) mq[i] += shfl_xor_d(mq[i], m);
      if (lane < C) {
        if (static_cast<int>(rank) == leader && n_lead > 0) mbar_wait(&sh->fdone, (n_lead - 1) & 1u);
        const uint32_t dst = mapa_u32(smem_u32(&sh->x1[par][rank][0]), static_cast<uint32_t>(lane));
        const uint32_t bar = mapa_u32(smem_u32(&sh->xbar1[par]), static_cast<uint32_t>(lane));
        mbar_arrive_expect_tx_remote(bar, 8u * (2u * O + 6u));
        for (int i = 0; i < 2 * O; ++i) st_async_f64(dst + 8u * i, sh->octtot[i], bar);
#pragma unroll
        for (int i = 0; i < 6; ++i) st_async_f64(dst + 64 + 8u * i, mq[i], bar);
      }
    }
    mbar_wait_cluster(&sh->xbar1[par], (it >> 1) & 1);
    if (warp == 0) bar_arrive(kClBarDirector + par, 64);

    // ------------------------------------------------------------------ window (warp 0, after the canonical direction)
    mbar_wait(&sh->hdone[par], (it >> 1) & 1);
    if (warp == 0) {
      const double h0 = ctl->h0, h1 = ctl->h1, m0 = ctl->m0, m1 = ctl->m1;
      double qq[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int s = 0; s < C; ++s)
#pragma unroll
        for (int i = 0; i < 6; ++i) qq[i] += sh->x1[par][s][8 + i];
      const double ns = qq[5] > 0.0 ? qq[5] : 1.0;
      const double ex = qq[3] / ns, ey = qq[4] / ns;
      const double cxx = qq[0] / ns - ex * ex, cyy = qq[1] / ns - ey * ey, cxy = qq[2] / ns - ex * ey;
      const double var_l = h0 * h0 * cxx + 2.0 * h0 * h1 * cxy + h1 * h1 * cyy;
      const double sigma = sqrt(var_l);
      const double mu_l = -(h0 * m0 + h1 * m1);
      const double t_lo = __dadd_rn(mu_l + a.z_lo * sigma, 0.0);
      const double t_hi = __dadd_rn(mu_l + a.z_hi * sigma, 0.0);
      const int window_ok = a.use_window && !ctl->nonfinite && qq[5] >= 256.0 && var_l > 0.0 &&
                            isfinite(t_lo) && isfinite(t_hi) && t_lo <= t_hi;
      if (lane == 0) {
        ctl->t_lo = t_lo;
        ctl->t_hi = t_hi;
        ctl->window_ok = window_ok;
      }
    }
    cl_team_sync();  // S2
    const bool window = ctl->window_ok != 0;

    int rel = 0;
    auto release_upto = [&](int c) {
      if (c > rel) {
        __syncwarp();
        if (lane == 0)
          for (int j = rel; j < c; ++j) mbar_arrive(&sh->free_[j]);
        rel = c;
      }
    };
    if (!window) {
      release_upto(n_chunks);
      if (lane == 0) {
        if (tid == 0 && static_cast<int>(rank) == leader) {
          sh->fin_ctl = *ctl;
          mbar_arrive(&sh->xbar2);
        } else {
          mbar_arrive_expect_tx_remote(mapa_u32(smem_u32(&sh->xbar2), static_cast<uint32_t>(leader)), 0u);
        }
      }
      if (static_cast<int>(rank) == leader) ++n_lead;
      continue;
    }

    // ------------------------------------------------------------------ sweep B: exact canonical loss of every sample
    const double h0 = ctl->h0, h1 = ctl->h1, t_lo = ctl->t_lo, t_hi = ctl->t_hi;
    int c_gt = 0, nc = 0;
    double s_gt = 0.0;
    auto classify = [&](bool valid, const double2 v) {
      const double L = loss_of(h0, h1, v.x, v.y);
      const bool up = valid && (L > t_hi);
      const bool cd = valid && !up && (L >= t_lo);
      if (up) {
        ++c_gt;
        s_gt += L;
      }
      const unsigned bal = __ballot_sync(kFull, cd);
      if (bal) {
        const int pos = nc + __popc(bal & lt_mask);
        if (cd && pos < kClWarpList) wcand[pos] = L;
        nc += __popc(bal);
      }
    };
    for (int k = 0; k < O; ++k) {
      const uint32_t ob = static_cast<uint32_t>(k) * oct_b;
      const uint32_t oe = part_b < ob + oct_b ? part_b : ob + oct_b;
      const uint32_t len = oe > ob ? oe - ob : 0u;
      const uint32_t n_lr = (len + kClLaneRow - 1) / kClLaneRow;
      uint32_t lr = 0;
      for (; (lr + 4) * kClLaneRow <= len; lr += 4) {
        const uint32_t base = ob + lr * kClLaneRow + toff;
        release_upto(static_cast<int>((ob + lr * kClLaneRow + woff) >> 15));
        const double2 v0 = lds128d(slot_s + base), v1 = lds128d(slot_s + base + kClLaneRow),
                      v2 = lds128d(slot_s + base + 2 * kClLaneRow), v3 = lds128d(slot_s + base + 3 * kClLaneRow);
        classify(true, v0);
        classify(true, v1);
        classify(true, v2);
        classify(true, v3);
      }
      for (; lr < n_lr; ++lr) {   // warp-uniform trip count; the ragged last row masks per thread
        release_upto(static_cast<int>((ob + lr * kClLaneRow + woff) >> 15));
        const uint32_t off = ob + lr * kClLaneRow + toff;
        const bool valid = off + 16 <= oe;
        classify(valid, valid ? lds128d(slot_s + off) : first);
      }
    }
    release_upto(n_chunks);
    const bool overflow = nc > kClWarpList;
    {
      const int wc = __reduce_add_sync(kFull, c_gt);
      const double ws = warp_sum_any(s_gt);
      if (lane == 0) {
        sh->wsum[warp * 4] = static_cast<double>(wc);
        sh->wsum[warp * 4 + 1] = ws;
        sh->wcnt[warp * 2] = nc;
        sh->wcnt[warp * 2 + 1] = overflow ? 1 : 0;
      }
    }
    cl_team_sync();  // S3
    // ------------------------------------------------------------------ exchange 2
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
          sh->fin_ctl = *ctl;
          mbar_expect_tx(&sh->xbar2, tx);
        } else {
          mbar_arrive_expect_tx_remote(bar, tx);
        }
      }
      if (!ovf) {
        const uint32_t dst = mapa_u32(smem_u32(&sh->pool[rank * cap + before]), static_cast<uint32_t>(leader));
        for (int j = lane; j < nc; j += 32) st_async_f64(dst + 8u * j, wcand[j], bar);
      }
      if (warp == 0) {
        double n_above = 0.0, s_hi = 0.0;
        if (lane < kClTeamWarps) {
          n_above = sh->wsum[lane * 4];
          s_hi = sh->wsum[lane * 4 + 1];
        }
#pragma unroll
        for (int m = 8; m >= 1; m >>= 1) {
          n_above += shfl_xor_d(n_above, m);
          s_hi += shfl_xor_d(s_hi, m);
        }
        if (lane == 0) {
          const uint32_t dst = mapa_u32(smem_u32(&sh->x2[rank][0]), static_cast<uint32_t>(leader));
          st_async_f64(dst, n_above, bar);
          st_async_f64(dst + 8, s_hi, bar);
          st_async_f64(dst + 16, 0.0, bar);
          st_async_f64(dst + 24, ovf ? -1.0 : static_cast<double>(total), bar);
        }
      }
      if (static_cast<int>(rank) == leader) ++n_lead;
    }
    // (list / wsum / wcnt are next written behind the team barriers of the next halfspace's octant trees)
  }
  cluster_sync_all();
}

}  // namespace drcvar
