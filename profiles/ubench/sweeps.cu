// Isolated sweep-A / sweep-B loops over an 80 KB shared-memory slot (no TMA, no barriers) to find their
// intrinsic rate at 1 or 2 CTAs per SM and 8 or 16 warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ void classify_f32(float p, float ta, float tk, unsigned long long xy, unsigned long long& acc,
                                             int& cnt, unsigned& mask, unsigned bit) {
  asm("{\n\t.reg .pred u, k;\n\tsetp.lt.f32 u, %3, %4;\n\tsetp.le.and.f32 k, %3, %5, !u;\n\t@u add.rn.f32x2 %0, %0, %6;\n\t"
      "@u add.s32 %1, %1, 1;\n\t@k or.b32 %2, %2, %7;\n\t}" : "+l"(acc), "+r"(cnt), "+r"(mask) : "f"(p), "f"(ta), "f"(tk), "l"(xy), "r"(bit));
}
// variant 2: scalar predicated adds
__device__ __forceinline__ void classify_v2(float p, float ta, float tk, float x, float y, float& ax, float& ay,
                                            int& cnt, unsigned& mask, unsigned bit) {
  asm("{\n\t.reg .pred u, k;\n\tsetp.lt.f32 u, %4, %5;\n\tsetp.le.and.f32 k, %4, %6, !u;\n\t@u add.f32 %0, %0, %7;\n\t@u add.f32 %1, %1, %8;\n\t"
      "@u add.s32 %2, %2, 1;\n\t@k add.s32 %3, %3, %9;\n\t}" : "+f"(ax), "+f"(ay), "+r"(cnt), "+r"(mask) : "f"(p), "f"(ta), "f"(tk), "f"(x), "f"(y), "r"(bit));
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int rows, int reps, float h0, float h1, float ta, float tk) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* sm4 = reinterpret_cast<float4*>(smem);
  const int nt = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < rows * nt; i += nt) sm4[i] = make_float4(3.f + 0.001f * (i % 97), -1.f + 0.002f * (i % 89), 3.1f, -1.2f);
  __syncthreads();
  float2 a0 = make_float2(0, 0), a1 = a0; float amax = 0.f;
  unsigned long long ab = 0; int c32 = 0; unsigned mask = 0; float ax = 0, ay = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {
#pragma unroll 4
      for (int m = 0; m < rows; ++m) {
        const float4 v = sm4[m * nt + tid];
        a0 = __fadd2_rn(a0, make_float2(v.x, v.y));
        a1 = __fadd2_rn(a1, make_float2(v.z, v.w));
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else if (MODE == 1) {
      unsigned b0 = 1, b1 = 2;
#pragma unroll 4
      for (int m = 0; m < rows; ++m) {
        const float4 v = sm4[m * nt + tid];
        const float p0 = fmaf(h1, v.y, h0 * v.x), p1 = fmaf(h1, v.w, h0 * v.z);
        classify_f32(p0, ta, tk, pack2(v.x, v.y), ab, c32, mask, b0);
        classify_f32(p1, ta, tk, pack2(v.z, v.w), ab, c32, mask, b1);
        b0 <<= 2; b1 <<= 2;
      }
    } else if (MODE == 2) {
      unsigned b0 = 1, b1 = 2;
#pragma unroll 4
      for (int m = 0; m < rows; ++m) {
        const float4 v = sm4[m * nt + tid];
        const float p0 = fmaf(h1, v.y, h0 * v.x), p1 = fmaf(h1, v.w, h0 * v.z);
        classify_v2(p0, ta, tk, v.x, v.y, ax, ay, c32, mask, b0);
        classify_v2(p1, ta, tk, v.z, v.w, ax, ay, c32, mask, b1);
        b0 <<= 2; b1 <<= 2;
      }
    } else if (MODE == 3) {   // loads only
#pragma unroll 4
      for (int m = 0; m < rows; ++m) {
        const float4 v = sm4[m * nt + tid];
        ax += v.x; ay += v.w;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * nt + tid] = a0.x + a0.y + a1.x + a1.y + amax + (float)(ab & 0xffff) + c32 + mask + ax + ay;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm) {
  const int total_f4 = 5000;               // 10 000 samples = 5000 float4
  const int rows = total_f4 / threads;      // per thread
  const int reps = 64;
  size_t smem = (size_t)rows * threads * 16;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctas_per_sm == 2 ? 110000 : 200000));
  float* out; long long* cyc;
  const int grid = 148 * ctas_per_sm;
  cudaMalloc(&out, (size_t)grid * threads * 4); cudaMalloc(&cyc, grid * 8);
  size_t dyn = ctas_per_sm == 2 ? 110000 : 200000; (void)smem;
  for (int w = 0; w < 2; ++w) k<MODE><<<grid, threads, dyn>>>(out, cyc, rows, reps, 0.8f, 0.6f, 1.0f, 1.1f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[296]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < grid; ++i) c += h[i]; c /= grid;
  printf("%-10s threads=%4d ctas/SM=%d: %8.0f cycles per pass over 10k samples per CTA (%.1f cycles per float4-row per warp)\n", name, threads,
         ctas_per_sm, c / reps, c / reps / rows);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int ctas : {1, 2})
    for (int threads : {256, 512}) {
      run<3>("loads", threads, ctas);
      run<0>("sweepA", threads, ctas);
      run<1>("sweepB", threads, ctas);
      run<2>("sweepB-v2", threads, ctas);
    }
  return 0;
}
