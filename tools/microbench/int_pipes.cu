// Integer-instruction throughput on one GPU: for each op a kernel with 8
// independent dependency chains per thread; reports warp-instructions per clock
// per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  if (OP == 0) {  // IMAD.WIDE.U32, chain through the high word
    unsigned long long p;
    asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(p) : "r"(a));
    d = (uint32_t)(p >> 32);
  } else if (OP == 1) {
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  } else if (OP == 2) {
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  } else if (OP == 3) {
    asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  } else if (OP == 4) {
    asm volatile("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(d) : "r"(a), "r"(b));
  } else if (OP == 5) {
    asm volatile("shf.l.wrap.b32 %0, %1, %2, 3;" : "=r"(d) : "r"(a), "r"(b));
  } else if (OP == 6) {
    d = __vadd2(a, b);
    asm volatile("" : "+r"(d));
  } else if (OP == 7) {
    asm volatile("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  } else if (OP == 8) {
    asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  } else if (OP == 9) {
    asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(a));
  } else if (OP == 10) {
    asm volatile("max.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  } else if (OP == 11) {
    asm volatile("shl.b32 %0, %1, 2;" : "=r"(d) : "r"(a));
  } else if (OP == 12) {
    asm volatile("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(b), "r"(a));
  } else {
    d = a;
  }
  return d;
}

template <int OP>
__global__ void k(uint32_t *out, uint32_t seed) {
  uint32_t x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = seed + threadIdx.x * 8 + i;
  uint32_t b = seed ^ 0x9E3779B9u, c = seed * 3 + 1;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = op<OP>(x[i], b, c);
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r ^= x[i];
  if (r == 0x12345678u) out[0] = r;
}

template <int OP>
void run(const char *name, int sms, double ghz, double instr_per_op) {
  uint32_t *out;
  cudaMalloc(&out, 4);
  dim3 grid(sms * 2), block(1024);
  k<OP><<<grid, block>>>(out, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<grid, block>>>(out, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)grid.x * 32 * ITERS * CHAINS * instr_per_op;
  double clocks = ms * 1e-3 * ghz * 1e9;
  printf("%-22s %8.3f ms  %6.2f warp-instr/clk/SM (%.1f SASS instr per op)\n", name, ms,
         warp_instr / clocks / sms, instr_per_op);
  cudaFree(out);
}

// two different ops on alternating chains: do their pipes overlap?
template <int OPA, int OPB>
__global__ void kmix(uint32_t *out, uint32_t seed) {
  uint32_t x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = seed + threadIdx.x * 8 + i;
  uint32_t b = seed ^ 0x9E3779B9u, c = seed * 3 + 1;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = (i & 1) ? op<OPB>(x[i], b, c) : op<OPA>(x[i], b, c);
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r ^= x[i];
  if (r == 0x12345678u) out[0] = r;
}
template <int OPA, int OPB>
void runmix(const char *name, int sms, double ghz, int threads = 1024, int ctas_per_sm = 2) {
  uint32_t *out;
  cudaMalloc(&out, 4);
  dim3 grid(sms * ctas_per_sm), block(threads);
  kmix<OPA, OPB><<<grid, block>>>(out, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kmix<OPA, OPB><<<grid, block>>>(out, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)grid.x * (threads / 32) * ITERS * CHAINS;
  double clocks = ms * 1e-3 * ghz * 1e9;
  printf("%-22s %8.3f ms  %6.2f warp-ops/clk/SM (50/50 mix, %d warps/SM)\n", name, ms, ops / clocks / sms,
         threads / 32 * ctas_per_sm);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  int s = p.multiProcessorCount;
  run<0>("IMAD.WIDE.U32 (hi)", s, ghz, 1.0);
  run<1>("LOP3", s, ghz, 1.0);
  run<2>("IMAD (mad.lo)", s, ghz, 1.0);
  run<3>("IADD (add.u32)", s, ghz, 1.0);
  run<4>("PRMT", s, ghz, 1.0);
  run<5>("SHF.L.W", s, ghz, 1.0);
  run<6>("VIADD.16x2", s, ghz, 1.0);
  run<7>("VIMNMX3.S16x2", s, ghz, 0.5);
  run<8>("IDP.4A", s, ghz, 1.0);
  run<9>("POPC", s, ghz, 1.0);
  run<10>("VIMNMX3.U32", s, ghz, 0.5);
  runmix<1, 2>("LOP3 + IMAD", s, ghz);
  runmix<1, 0>("LOP3 + IMAD.WIDE", s, ghz);
  runmix<1, 3>("LOP3 + IADD3", s, ghz);
  runmix<2, 3>("IMAD + IADD3", s, ghz);
  runmix<1, 4>("LOP3 + PRMT", s, ghz);
  runmix<2, 8>("IMAD + IDP.4A", s, ghz);
  runmix<1, 8>("LOP3 + IDP.4A", s, ghz);
  runmix<0, 2>("IMAD.WIDE + IMAD", s, ghz);
  runmix<1, 5>("LOP3 + SHF", s, ghz);
  // occupancy sensitivity of the overlap (the sweep kernels run 16 warps per SM)
  runmix<1, 2>("LOP3 + IMAD", s, ghz, 1024, 1);
  runmix<1, 2>("LOP3 + IMAD", s, ghz, 512, 1);
  runmix<1, 2>("LOP3 + IMAD", s, ghz, 256, 1);
  runmix<1, 0>("LOP3 + IMAD.WIDE", s, ghz, 512, 1);
  runmix<1, 0>("LOP3 + IMAD.WIDE", s, ghz, 256, 1);
  runmix<1, 1>("LOP3 only", s, ghz, 512, 1);
  runmix<1, 1>("LOP3 only", s, ghz, 256, 1);
  runmix<1, 1>("LOP3 only", s, ghz, 128, 1);
  return 0;
}
