// How fast can 2 CTAs per SM pull 80 KB rows into shared memory with 1-D TMA bulk copies and NO compute?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_stream tma_stream.cu && ./tma_stream
// mode 0: one slot per CTA, the next row is requested when the previous one has landed (latency exposed: the resident
//         kernel's worst case);  mode 1: the next row is requested `lead` chunks before the current one has fully landed is
//         not possible with one slot, so mode 1 uses TWO half-size slots per CTA (ping-pong, always one request in flight).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void expect_tx(unsigned long long* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void wait(unsigned long long* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}

// slots per CTA = nslots, each row_bytes; CTA streams rows b = blockIdx.x, += gridDim.x
__global__ void stream_kernel(const unsigned char* src, long long rows, uint32_t row_bytes, int nslots, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bars[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  if (threadIdx.x == 0) {
    long long b = blockIdx.x;
    int it = 0;
    // prologue: fill all slots
    for (int s = 0; s < nslots && b + (long long)s * gridDim.x < rows; ++s) {
      expect_tx(&bars[s], row_bytes);
      const unsigned char* p = src + (size_t)(b + (long long)s * gridDim.x) * row_bytes;
      for (uint32_t off = 0; off < row_bytes; off += 32768) bulk(smem + (size_t)s * row_bytes + off, p + off, min(32768u, row_bytes - off), &bars[s]);
    }
    for (; b < rows; b += gridDim.x, ++it) {
      const int s = it % nslots;
      wait(&bars[s], (it / nslots) & 1);
      acc += reinterpret_cast<float*>(smem + (size_t)s * row_bytes)[it & 15];
      const long long nb = b + (long long)nslots * gridDim.x;
      if (nb < rows) {
        expect_tx(&bars[s], row_bytes);
        const unsigned char* p = src + (size_t)nb * row_bytes;
        for (uint32_t off = 0; off < row_bytes; off += 32768) bulk(smem + (size_t)s * row_bytes + off, p + off, min(32768u, row_bytes - off), &bars[s]);
      }
    }
    sink[blockIdx.x] = acc;
  }
}

int main() {
  const uint32_t row = 80000;
  const long long rows = 262144;   // 21 GB
  unsigned char* d; float* sink;
  cudaMalloc(&d, (size_t)rows * row); cudaMemset(d, 0, (size_t)rows * row); cudaMalloc(&sink, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct { int ctas_per_sm, nslots; uint32_t row_bytes; const char* name; } cfg[] = {
    {2, 1, 80000, "2 CTAs/SM x 1 slot of 80 KB (request after the previous row landed)"},
    {1, 2, 80000, "1 CTA/SM x 2 slots of 80 KB (always one row in flight per SM)"},
    {2, 2, 40000, "2 CTAs/SM x 2 slots of 40 KB (half rows, always one in flight per CTA)"},
    {1, 1, 80000, "1 CTA/SM x 1 slot of 80 KB"},
  };
  for (auto& c : cfg) {
    const size_t smem = (size_t)c.nslots * c.row_bytes;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long nrows = rows * row / c.row_bytes;
    const int grid = 148 * c.ctas_per_sm;
    stream_kernel<<<grid, 32, smem>>>(d, nrows, c.row_bytes, c.nslots, sink);
    cudaEventRecord(e0);
    stream_kernel<<<grid, 32, smem>>>(d, nrows, c.row_bytes, c.nslots, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-72s %8.3f ms  %7.0f GB/s  (%s)\n", c.name, ms, (double)rows * row / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
