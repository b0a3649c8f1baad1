// How fast can one SM sub-partition run the Philox4x32-10 part of the update alone?
// Each thread performs `iters` iterations of two independent Philox calls (as
// update16 does per 16 sites) and xors the outputs into a sink.  Reported:
// cycles per iteration and scheduler at 4 and 5 warps per scheduler, and the same
// with 120 dependent-free LOP3/IADD filler operations per iteration (the size of
// the compare part), to see how the two overlap.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o philox_rate philox_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 philox(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c.x;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c.z;
    uint4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ (k0 + r * 0x9E3779B9u);
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ (k1 + r * 0xBB67AE85u);
    n.w = (uint32_t)p0;
    c = n;
  }
  return c;
}

// the same with the 64-bit products formed as IMAD.HI + IMAD (two instructions)
__device__ __forceinline__ uint4 philox_hilo(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(h0) : "r"(c.x), "r"(0xD2511F53u));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(l0) : "r"(c.x), "r"(0xD2511F53u));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(h1) : "r"(c.z), "r"(0xCD9E8D57u));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(l1) : "r"(c.z), "r"(0xCD9E8D57u));
    uint4 n;
    n.x = h1 ^ c.y ^ (k0 + r * 0x9E3779B9u);
    n.y = l1;
    n.z = h0 ^ c.w ^ (k1 + r * 0xBB67AE85u);
    n.w = l0;
    c = n;
  }
  return c;
}

__global__ void k_hilo(uint32_t *sink, int iters, long long *cyc) {
  uint32_t acc = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint4 a = philox_hilo(make_uint4(i + threadIdx.x, threadIdx.x, blockIdx.x + threadIdx.x, 0u), 1u, 2u);
    const uint4 b = philox_hilo(make_uint4(i + threadIdx.x, threadIdx.x, blockIdx.x + threadIdx.x, 1u), 1u, 2u);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// per-thread counters in every word (nothing for the uniform datapath to take)
__global__ void k_wide(uint32_t *sink, int iters, long long *cyc) {
  uint32_t acc = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint4 a = philox(make_uint4(i + threadIdx.x, threadIdx.x, blockIdx.x + threadIdx.x, 0u), 1u, 2u);
    const uint4 b = philox(make_uint4(i + threadIdx.x, threadIdx.x, blockIdx.x + threadIdx.x, 1u), 1u, 2u);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run2(K kern, int threads, const char *what) {
  uint32_t *sink;
  long long *cyc, h[148];
  cudaMalloc(&sink, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  kern<<<148, threads>>>(sink, 100, cyc);
  kern<<<148, threads>>>(sink, iters, cyc);
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-28s threads %4d: %.1f cycles per iteration and warp, %.1f per warp-iteration and scheduler\n", what,
         threads, avg / iters, avg / iters / (threads / 32 / 4.0));
  cudaFree(sink);
  cudaFree(cyc);
}

template <int FILL>
__global__ void k(uint32_t *sink, int iters, long long *cyc) {
  uint32_t acc = threadIdx.x, f0 = threadIdx.x * 3u, f1 = blockIdx.x, f2 = 7u, f3 = 11u;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint4 a = philox(make_uint4(i, threadIdx.x, blockIdx.x, 0u), 1u, 2u);
    const uint4 b = philox(make_uint4(i, threadIdx.x, blockIdx.x, 1u), 1u, 2u);
    acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
#pragma unroll
    for (int j = 0; j < FILL / 4; ++j) {  // four independent chains of ALU work
      f0 = (f0 ^ acc) + 0x01010101u;
      f1 = (f1 & 0x7f7f7f7fu) ^ f0;
      f2 = (f2 | f1) + 3u;
      f3 = (f3 ^ f2) & 0x0fffffffu;
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ f0 ^ f1 ^ f2 ^ f3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int FILL>
static void run(int threads, const char *what) {
  uint32_t *sink;
  long long *cyc, h[148];
  cudaMalloc(&sink, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  k<FILL><<<148, threads>>>(sink, 100, cyc);
  k<FILL><<<148, threads>>>(sink, iters, cyc);
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double warps_per_sched = threads / 32 / 4.0;
  printf("%-28s threads %4d: %.1f cycles per iteration and warp, %.1f per warp-iteration and scheduler\n", what,
         threads, avg / iters, avg / iters / warps_per_sched);
  cudaFree(sink);
  cudaFree(cyc);
}

int main() {
  run<0>(512, "philox x2");
  run<0>(640, "philox x2");
  run<0>(1024, "philox x2");
  run<120>(512, "philox x2 + 120 ALU ops");
  run<120>(640, "philox x2 + 120 ALU ops");
  run<240>(512, "philox x2 + 240 ALU ops");
  run2(k_wide, 512, "philox x2, 40 IMAD.WIDE");
  run2(k_hilo, 512, "philox x2, IMAD.HI + IMAD");
  return 0;
}
