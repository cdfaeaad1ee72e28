// Throughput microbenchmark of the instruction kinds the halfspace kernel leans on (B200, sm_100a).
// Each test: 32 warps/SM (1024 threads, 1 CTA per SM), ILP 8 independent chains per thread, 4096 iterations.
// Reports warp-instructions per cycle per SM.   nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
template <int OP>
__global__ void k(float* out, long long* cyc, float a, float b, int ia) {
  float x[8];
  unsigned u[8];
  double d[8];
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = a + threadIdx.x * 0.001f + i;
    u[i] = ia + threadIdx.x + i;
    d[i] = a + i;
    p[i] = make_float2(a + i, b - i);
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = fmaf(x[i], a, b);                                             // FFMA
      if (OP == 1) { asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %2, %0, q;}" : "+f"(x[i]) : "f"(b), "f"(a)); }  // FSETP+FSEL (dependent)
      if (OP == 2) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(u[i]) : "r"(ia), "r"(it)); }   // LOP3
      if (OP == 3) { asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(it)); }                       // IADD
      if (OP == 4) { asm volatile("max.f32 %0, %0, %1; max.f32 %0, %0, %2;" : "+f"(x[i]) : "f"(p[i].x), "f"(p[i].y)); p[i].x += 1.0f; }  // FMNMX(3) + FADD
      if (OP == 5) p[i] = __fadd2_rn(p[i], make_float2(a, b));                          // FADD2
      if (OP == 6) d[i] = d[i] + (double)a;                                             // DADD
      if (OP == 7) d[i] = d[i] * (double)b;                                             // DMUL
      if (OP == 8) { asm volatile("{.reg .pred q; setp.lt.f64 q, %0, %1; @q add.f64 %0, %0, %2;}" : "+d"(d[i]) : "d"((double)b), "d"((double)a)); }  // DSETP+@DADD
      if (OP == 9) { double t = (double)x[i]; asm volatile("" : "+d"(t)); x[i] = x[i] + 1.0f; d[i] += t; }  // F2F + FADD + DADD
      if (OP == 10) { asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; @q add.f32 %0, %0, %2;}" : "+f"(x[i]) : "f"(b), "f"(a)); }  // FSETP + @FADD
      if (OP == 11) { asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; @q add.s32 %0, %0, 1;}" : "+r"(u[i]) : "f"(x[i]), "f"(b)); x[i] += a; }  // FSETP + @IADD + FADD
      if (OP == 12) u[i] = u[i] * ia + 3;                                              // IMAD
      if (OP == 13) x[i] = x[i] * a;                                                    // FMUL
      if (OP == 14) x[i] = x[i] + a;                                                    // FADD
      if (OP == 15) { asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; @q or.b32 %0, %0, %3;}" : "+r"(u[i]) : "f"(x[i]), "f"(b), "r"(it)); x[i] += a; }  // FSETP + @LOP + FADD
      if (OP == 16) { asm volatile("{.reg .pred q, r; setp.lt.f32 q, %1, %2; setp.le.and.f32 r, %1, %3, !q; @r or.b32 %0, %0, %4;}" : "+r"(u[i]) : "f"(x[i]), "f"(b), "f"(a), "r"(it)); x[i] += a; }  // 2 FSETP + @LOP + FADD
      if (OP == 17) { asm volatile("shf.l.wrap.b32 %0, %0, %0, %1;" : "+r"(u[i]) : "r"(it)); }          // SHF
      if (OP == 18) { asm volatile("popc.b32 %0, %0;" : "+r"(u[i])); u[i] += it; }                        // POPC + IADD
      if (OP == 19) { u[i] = __shfl_xor_sync(0xffffffffu, u[i], 1) + it; }                               // SHFL + IADD
      if (OP == 20) { u[i] = __ballot_sync(0xffffffffu, (u[i] & 1) != 0) + it; }                         // VOTE (+LOP/ISETP/IADD)
      if (OP == 21) { u[i] = __reduce_add_sync(0xffffffffu, u[i]) + it; }                                // REDUX + IADD
      if (OP == 22) { asm volatile("{.reg .pred q; setp.lt.s32 q, %0, %1; selp.s32 %0, %2, %0, q;}" : "+r"(u[i]) : "r"(ia), "r"(it)); }  // ISETP + SEL
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + u[i] + (float)d[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_op, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, threads>>>(out, cyc, 1.0001f, 0.5f, 7);
  k<OP><<<148, threads>>>(out, cyc, 1.0001f, 0.5f, 7);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double winst = (double)ITER * 8 * instr_per_op * (threads / 32);
  printf("%-22s threads=%4d  %7.3f warp-inst/cycle/SM  (cycles %.0f)\n", name, threads, winst / c, c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {1024, 256}) {
    run<0>("FFMA", 1, threads);
    run<14>("FADD", 1, threads);
    run<5>("FADD2", 1, threads);
    run<12>("IMAD", 1, threads);
    run<3>("IADD", 1, threads);
    run<2>("LOP3", 1, threads);
    run<17>("SHF", 1, threads);
    run<1>("FSETP+FSEL", 2, threads);
    run<22>("ISETP+SEL", 2, threads);
    run<10>("FSETP+@FADD", 2, threads);
    run<11>("FSETP+@IADD+FADD", 3, threads);
    run<15>("FSETP+@LOP+FADD", 3, threads);
    run<16>("2FSETP+@LOP+FADD", 4, threads);
    run<4>("2FMNMX+FADD", 3, threads);
    run<18>("POPC+IADD", 2, threads);
    run<19>("SHFL+IADD", 2, threads);
    run<20>("VOTE+3", 4, threads);
    run<21>("REDUX+IADD", 2, threads);
    run<6>("DADD", 1, threads);
    run<8>("DSETP+@DADD", 2, threads);
    run<9>("F2F+FADD+DADD", 3, threads);
  }
  return 0;
}
