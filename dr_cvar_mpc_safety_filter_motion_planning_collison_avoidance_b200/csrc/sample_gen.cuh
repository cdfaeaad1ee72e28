// sample_gen.cuh — on-device obstacle-sample generator, fused into the staging step of halfspace_kernel.
//
// Replaces simulation/obstacles.py:43-77 of the reference (nominal position + N(0, noise_cov) per (obstacle, step)) for
// batches that should never touch HBM or PCIe for their inputs.  The arithmetic is specified in oracle/sample_gen.py
// and reproduced here bit for bit: counter-based Philox4x32-10, Box-Muller in fp32 built from individually rounded
// +,-,*,/ and sqrt only (no FMA: every product and sum below goes through an _rn intrinsic).
#pragma once

#include <cstdint>

namespace drcvar {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kGenStreamTag = 0x44524356u;   // 'DRCV'

struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float gmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float gadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float gsub(float a, float b) { return __fsub_rn(a, b); }

// log(x) for x in [2^-24, 1]: fdlibm-style, same constants and operation order as oracle/sample_gen.py::log_f32
__device__ __forceinline__ float gen_log(float x) {
  const float ln2_hi = 0x1.62e300p-1f, ln2_lo = 0x1.2fefa2p-17f;
  const float lg1 = 0x1.555554p-1f, lg2 = 0x1.999c26p-2f, lg3 = 0x1.23d3dcp-2f, lg4 = 0x1.f13c4cp-3f;
  int ix = __float_as_int(x);
  ix += 0x3f800000 - 0x3f3504f3;
  const int k = (ix >> 23) - 0x7f;
  ix = (ix & 0x007fffff) + 0x3f3504f3;
  const float m = __int_as_float(ix);
  const float f = gsub(m, 1.0f);
  const float s = __fdiv_rn(f, gadd(2.0f, f));
  const float z = gmul(s, s);
  const float w = gmul(z, z);
  const float t1 = gmul(w, gadd(lg2, gmul(w, lg4)));
  const float t2 = gmul(z, gadd(lg1, gmul(w, lg3)));
  const float r = gadd(t2, t1);
  const float hfsq = gmul(gmul(0.5f, f), f);
  const float dk = static_cast<float>(k);
  return gadd(gadd(gsub(gadd(gmul(s, gadd(hfsq, r)), gmul(dk, ln2_lo)), hfsq), f), gmul(dk, ln2_hi));
}

// (cos(2 pi u), sin(2 pi u)) for u in [0, 1): exact quadrant split, reflection, Taylor kernels on [0, pi/4]
__device__ __forceinline__ void gen_sincos(float u, float& c_out, float& s_out) {
  const float half_pi = 0x1.921fb6p+0f;
  const float s1 = static_cast<float>(-1.0 / 6.0), s2 = static_cast<float>(1.0 / 120.0),
              s3 = static_cast<float>(-1.0 / 5040.0), s4 = static_cast<float>(1.0 / 362880.0);
  const float c1 = static_cast<float>(1.0 / 24.0), c2 = static_cast<float>(-1.0 / 720.0),
              c3 = static_cast<float>(1.0 / 40320.0);
  const float t = gmul(4.0f, u);
  const float qf = floorf(t);
  const float r = gsub(t, qf);
  const int q = static_cast<int>(qf);
  const bool flip = r > 0.5f;
  const float rr = flip ? gsub(1.0f, r) : r;
  const float x = gmul(rr, half_pi);
  const float z = gmul(x, x);
  const float sp = gadd(x, gmul(gmul(x, z), gadd(s1, gmul(z, gadd(s2, gmul(z, gadd(s3, gmul(z, s4))))))));
  const float cp = gadd(gsub(1.0f, gmul(0.5f, z)), gmul(gmul(z, z), gadd(c1, gmul(z, gadd(c2, gmul(z, c3))))));
  const float s_ = flip ? cp : sp, c_ = flip ? sp : cp;
  c_out = q == 0 ? c_ : (q == 1 ? -s_ : (q == 2 ? -c_ : s_));
  s_out = q == 0 ? s_ : (q == 1 ? c_ : (q == 2 ? -s_ : -c_));
}

// one sample (x, y) = mean + L z from two random words
__device__ __forceinline__ float2 gen_sample(uint32_t ra, uint32_t rb, float mx, float my, float l00, float l10, float l11) {
  const float u1 = gmul(static_cast<float>((ra >> 8) + 1u), 0x1p-24f);
  const float u2 = gmul(static_cast<float>(rb >> 8), 0x1p-24f);
  const float rad = __fsqrt_rn(gmul(-2.0f, gen_log(u1)));
  float c, s;
  gen_sincos(u2, c, s);
  const float z0 = gmul(rad, c), z1 = gmul(rad, s);
  return make_float2(gadd(mx, gmul(l00, z0)), gadd(my, gadd(gmul(l10, z0), gmul(l11, z1))));
}

}  // namespace drcvar
