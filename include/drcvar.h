/*
 * drcvar.h — C ABI of libdrcvar.so: B200 (sm_100a) risk-bounded safe-halfspace engine.
 *
 * The reference (RJ-23YP/DR_CVaR_MPC_Safety_Filter_Motion_Planning_Collison_Avoidance) is pure Python and has
 * NO FFI layer; its boundary for this path is the Python API of core/risk_metrics.py and core/halfspaces.py.
 * Each entry point below names the reference interface it replaces (file:line relative to the reference root).
 * The Python binding a maintainer adds is a ctypes stub (see INTEGRATION.md and
 * dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200/_lib.py).
 *
 * Conventions
 *   - plain pointers and sizes only; the library never frees or retains caller memory;
 *   - every function returns DRCVAR_OK (0) or a negative DRCVAR_ERR_* code and never throws;
 *     drcvar_last_error() returns a thread-local message for the last failure;
 *   - `device >= 0`: all pointers are DEVICE pointers on that CUDA device; work is enqueued on `stream`
 *     (a cudaStream_t, NULL = legacy default stream) and the call does not synchronise unless
 *     DRCVAR_FLAG_SYNC is set;
 *   - `device == DRCVAR_HOST (-1)`: all pointers are HOST pointers; the library stages host->device copies
 *     (chunked and overlapped with the kernels), runs on the current device and returns after the
 *     results are back in the caller's host arrays;
 *   - strides are in ELEMENTS of the sample dtype: coordinate c of sample i of halfspace b lives at
 *     samples[b*stride_b + i*stride_n + c*stride_c].  A C-contiguous [B,N,2] array is (2N, 2, 1); the
 *     reference's per-step view traj[:, t, :] of a [N, T+1, 2] array (simulation/environment.py:88) is
 *     (0, 2(T+1), 1) with the base pointer advanced by 2t.
 *
 * Per-halfspace math (closed form of the reference's two ECOS LPs; see DESIGN.md, oracle/closed_form.py):
 *   m = mean(xi); h = unit(m - ego) (fallback [1,0] if |m-ego| < 1e-10)          core/geometry.py:35-53
 *   L_i = -(h.xi_i); CVaR = (sum of the alpha*N largest L_i, fractional last weight) / (alpha*N)
 *   g_cvar   = CVaR + (r_robot+r_obs)|h| - delta                                  core/risk_metrics.py:179-265,305-338
 *   g_star   = CVaR + (r_robot+r_obs)|h| + epsilon/alpha - delta                  core/risk_metrics.py:84-177
 *   g_drcvar = g_star - (r_robot+r_obs)|h|                                        core/risk_metrics.py:267-303
 *   h_mean = unit(m - 0); g_mean = -(h_mean.m - (r_robot+r_obs)|h_mean|)          core/halfspaces.py:70-106
 */
#ifndef DRCVAR_H_
#define DRCVAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRCVAR_ABI_VERSION 1

#define DRCVAR_OK 0
#define DRCVAR_ERR_INVALID (-1)     /* bad argument (alpha not in (0,1], N < 1, null pointer, bad stride) */
#define DRCVAR_ERR_CUDA (-2)        /* a CUDA runtime call failed; see drcvar_last_error() */
#define DRCVAR_ERR_UNSUPPORTED (-3) /* reserved (every N >= 1 is supported: large N uses the streaming kernel) */
#define DRCVAR_ERR_NOMEM (-4)

#define DRCVAR_HOST (-1)

/* flags */
#define DRCVAR_FLAG_SYNC 1u          /* device pointers: synchronise the stream before returning */
#define DRCVAR_FLAG_GENERAL_ONLY 2u  /* disable the statistical candidate window (always use the general select) */
#define DRCVAR_FLAG_NO_BULK 4u       /* disable cp.async.bulk staging (use the generic strided loader) */
#define DRCVAR_FLAG_FORCE_STREAMING 8u /* use the two-pass streaming kernel even when one CTA could hold N samples */
#define DRCVAR_FLAG_NO_CLUSTER 16u   /* large N: do not use the cluster / DSMEM single-read kernel (streaming kernel instead) */
#define DRCVAR_FLAG_FORCE_CLUSTER 32u /* fp64 samples, large N: use the fp64 cluster kernel (opt-in: the streaming kernel is faster there) */
#define DRCVAR_FLAG_NO_PIPELINE 64u  /* resident sizes: use halfspace_kernel (inline general path) instead of the pipelined kernel + redo pass */
#define DRCVAR_FLAG_LARGE_COORDS 128u /* fp32 samples at resident sizes whose coordinates are far from the origin (|xi| beyond ~64 / (alpha N / 256)):
                                         keep the first-sample-relative sums of the pipelined kernel.  Without it such halfspaces are still exact
                                         (handed to the redo pass) but slower; with it small coordinates lose ~3 % of throughput.  Speed only. */

/* per-halfspace status bits written to status_out */
#define DRCVAR_STATUS_NONFINITE 1   /* non-finite input: sentinel 100.0 emitted (core/risk_metrics.py:177,265,303,338) */
#define DRCVAR_STATUS_GENERAL 2     /* the general (multi-sweep) select path was taken */
#define DRCVAR_STATUS_DEGENERATE 4  /* |mean - ego| < 1e-10: fallback direction [1,0] (core/geometry.py:49-51) */

int drcvar_version(void);
const char* drcvar_last_error(void);
int drcvar_device_count(void);

/* Number of summation lanes of the canonical mean reduction (512); part of the arithmetic contract. */
int drcvar_reduction_lanes(void);

/*
 * Tail size for (alpha, N): returns kc = ceil(alpha*N) (alpha*N snapped to an integer when within 1e-9
 * relative), writes the fractional tail mass to *k_f_out when non-NULL.  Negative on invalid input.
 * Replaces the implicit 1/(alpha*N) weighting of core/risk_metrics.py:110,209.
 */
int64_t drcvar_tail_count(double alpha, int64_t n_samples, double* k_f_out);

/* Largest N the single-CTA shared-memory kernel holds for a sample dtype of `elem_bytes` (4 or 8) on `device`.
 * Larger N: fp32 samples with 32768 < N <= ~205000 (contiguous, 16-byte aligned rows) are served by the cluster / DSMEM
 * kernel (one thread-block cluster per halfspace, one read of the samples; halfspaces it cannot finish — window miss,
 * non-finite data — are redone by the streaming kernel in the same stream); everything else by the two-pass streaming
 * kernel (same results, two reads of the samples). */
int64_t drcvar_max_samples(int elem_bytes, int device);

/*
 * Batched safe halfspaces — replaces, for B (scenario, obstacle, step) triples at once,
 *   MeanSafeHalfspace.create / CVaRSafeHalfspace.create / DRCVaRSafeHalfspace.create   core/halfspaces.py:70-194
 *   cvar_halfspace / dr_cvar_halfspace (when h_in != NULL)                                core/risk_metrics.py:267-338
 *   CVaROptimizer.solve / DRCVaROptimizer.solve                                           core/risk_metrics.py:127-177,215-265
 *   compute_safe_halfspaces (loop over obstacles)                                         core/halfspaces.py:196-247
 *
 *   samples      [B] x [N] x [2], dtype float (f32) / double (f64), strides in elements (see above)
 *   ego          [B,2] double ego reference positions, or NULL (= origin)
 *   h_in         [B,2] double explicit normals, or NULL (derive h from the sample mean and ego)
 *   h_out        [B,2] double  normal used by the CVaR / DR-CVaR halfspaces              (required)
 *   h_mean_out   [B,2] double  normal of the mean halfspace (measured from the origin)   (may be NULL)
 *   g_out        [B,3] double  (g_mean, g_cvar, g_drcvar) — the g-tilde of each metric   (required)
 *   cvar_out     [B]   double  CVaR_alpha of the loss                                     (may be NULL)
 *   var_out      [B]   double  ceil(alpha N)-th largest loss (VaR threshold T)           (may be NULL)
 *   gstar_out    [B]   double  DR-CVaR LP optimum g* (before subtracting the radius)     (may be NULL)
 *   status_out   [B]   int32   DRCVAR_STATUS_* bits                                       (may be NULL)
 *   tail_idx_out [B, kc] int32 indices of the kc = drcvar_tail_count() largest losses, ascending,
 *                ties -> lower index (parity mode; may be NULL; slows the kernel)
 */
int drcvar_halfspaces_f32(const float* samples, int64_t B, int64_t N,
                          int64_t stride_b, int64_t stride_n, int64_t stride_c,
                          const double* ego, const double* h_in,
                          double alpha, double delta, double epsilon, double r_robot, double r_obs,
                          uint32_t flags,
                          double* h_out, double* h_mean_out, double* g_out,
                          double* cvar_out, double* var_out, double* gstar_out,
                          int32_t* status_out, int32_t* tail_idx_out,
                          int device, void* stream);

int drcvar_halfspaces_f64(const double* samples, int64_t B, int64_t N,
                          int64_t stride_b, int64_t stride_n, int64_t stride_c,
                          const double* ego, const double* h_in,
                          double alpha, double delta, double epsilon, double r_robot, double r_obs,
                          uint32_t flags,
                          double* h_out, double* h_mean_out, double* g_out,
                          double* cvar_out, double* var_out, double* gstar_out,
                          int32_t* status_out, int32_t* tail_idx_out,
                          int device, void* stream);

/*
 * One-launch trajectory entry — replaces the (t, obstacle) double loop of
 * SafetyFilteringEnvironment.compute_safe_halfspaces_for_trajectory     simulation/environment.py:60-106.
 * HOST pointers only.  traj[i] points at obstacle i's [N, T1, 2] double array (C-contiguous);
 * halfspace (t, i) uses samples traj[i][:, t, :] and ego ego_steps[t].  Outputs are [n_steps, n_obs, ...]
 * with the same meaning as above.
 */
int drcvar_trajectory_f64(const double* const* traj, int64_t n_obs, int64_t N, int64_t T1, int64_t n_steps,
                          const double* ego_steps /* [n_steps,2] */,
                          double alpha, double delta, double epsilon, double r_robot, double r_obs,
                          uint32_t flags,
                          double* h_out, double* h_mean_out, double* g_out, int32_t* status_out);

/*
 * Fused sample generation + halfspaces (fp32 samples, never materialised in HBM) — replaces, for B (obstacle, step)
 * pairs at once, generate_obstacle_sample_trajectories                              simulation/obstacles.py:43-77
 * followed by the halfspace entry points above: halfspace b draws N samples  mean[b] + L[b] z,  z ~ N(0, I2), inside the
 * kernel's staging step and feeds them to the same sweeps.  The reference draws from numpy's sequential MT19937
 * stream (not reproducible by independent threads); here z comes from counter-based Philox4x32-10 + Box-Muller with a
 * fully specified fp32 arithmetic (oracle/sample_gen.py restates it bit for bit), counter = (pair index, index_offset +
 * b), key = seed.
 *   mean [B,2] double nominal positions (nominal_trajectory[t], obstacles.py:75; rounded to fp32 by the kernel)
 *   chol [B,3] double (l00, l10, l11): lower Cholesky factor of noise_cov (obstacles.py:68-72); zeros = no noise (t = 0)
 *   samples_out [B,N,2] float or NULL: dump of the generated samples (parity tests; costs the HBM write)
 * Other arguments and outputs as drcvar_halfspaces_f32.  32768 < N <= ~205000 (even N, no tail indices) runs the cluster
 * kernel, every CTA drawing its part of the samples into shared memory once; other N > drcvar_max_samples(4, device) go
 * through the streaming kernel, which re-draws the samples in each of its passes (same values: the generator is
 * counter-based).
 */
int drcvar_halfspaces_generated_f32(const double* mean, const double* chol, uint64_t seed, int64_t index_offset,
                                    int64_t B, int64_t N, const double* ego, const double* h_in,
                                    double alpha, double delta, double epsilon, double r_robot, double r_obs,
                                    uint32_t flags,
                                    double* h_out, double* h_mean_out, double* g_out,
                                    double* cvar_out, double* var_out, double* gstar_out,
                                    int32_t* status_out, int32_t* tail_idx_out, float* samples_out,
                                    int device, void* stream);

/* Pinned host memory for callers that want full-speed host->device staging through DRCVAR_HOST calls. */
void* drcvar_host_alloc(size_t bytes);
void drcvar_host_free(void* p);

/* CTAs per halfspace (2, 4 or 8; 0 = not served) the cluster / DSMEM kernel uses for N samples of `elem_bytes` each when
 * an SM offers `smem_optin_bytes` of opt-in shared memory (232448 on B200).  N = 100 000: 4 (fp32), 8 (fp64).  Diagnostic;
 * also what the tests pin so that a grown scratch area cannot silently push config 5 onto clusters of 8. */
int drcvar_cluster_ctas(int64_t n_samples, int elem_bytes, int64_t smem_optin_bytes);

/* Telemetry of this process: kernels launched so far; timings (ms) and bytes of the LAST DRCVAR_HOST call. */
int64_t drcvar_launch_count(void);
/* Checked builds only (make -C csrc checked -> libdrcvar_checked.so: every shared-memory list / histogram / pool index and
 * staged byte range of the kernels is asserted in range): number of failed assertions on the current device since the
 * library was loaded, `*first_site` = 100000 * file id + line of the first one.  -1 in the product build.  There is no
 * reference counterpart: the reference has no native code; this stands in for compute-sanitizer where that tool is closed. */
int64_t drcvar_debug_check_failures(int32_t* first_site);
int drcvar_last_host_call_stats(double* stage_ms, double* kernel_ms, int64_t* h2d_bytes, int64_t* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* DRCVAR_H_ */
