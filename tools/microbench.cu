// Instruction-throughput probe for the big-integer inner loops (developer tool, not product).
// Measures ops/clk/SM for: IMAD.WIDE.U32 chains, IMAD (lo), IADD3, 64-bit integer add, DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nova_b200/csrc/field.cuh"

constexpr int ITER = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void probe(uint64_t* out, uint32_t a0, uint32_t b0, double d0) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t acc[ILP];
  uint32_t lo[ILP], hi[ILP], kk[ILP], x = a0 + tid, y = b0 | 1;
  double fd[ILP], da = d0 + tid * 1e-9, db = 1.0000001;
  for (int k = 0; k < ILP; k++) { acc[k] = tid + k; lo[k] = tid * 7 + k; hi[k] = tid + 3 * k; kk[k] = k; fd[k] = d0 + k; }
  for (int i = 0; i < ITER; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      if (OP == 0) {  // IMAD.WIDE.U32: 32x32 + 64
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
      } else if (OP == 1) {  // IMAD lo
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(x), "r"(y));
      } else if (OP == 2) {  // IADD3
        asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[k]) : "r"(x));
      } else if (OP == 3) {  // 64-bit add
        asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[k]) : "l"((uint64_t)x << 20 | y));
      } else if (OP == 4) {  // DFMA
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(fd[k]) : "d"(da), "d"(db));
      } else if (OP == 5) {  // mad.hi
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(x), "r"(y));
      } else if (OP == 7) {  // carry-save product: wide mad with carry-OUT only + addc into a counter
        asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                     : "+r"(lo[k]), "+r"(hi[k]), "+r"(kk[k]) : "r"(x), "r"(y));
      } else if (OP == 8) {  // wide mad with carry-out only, carry discarded into ONE shared counter per 2
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(lo[k]), "+r"(hi[k]) : "r"(x), "r"(y));
      } else if (OP == 9) {  // co-issue probe: one carry-chain pair AND one independent DFMA per slot
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(lo[k]), "+r"(lo[(k + 1) % ILP]) : "r"(x), "r"(y));
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(fd[k]) : "d"(da), "d"(db));
      } else if (OP == 10) {  // co-issue probe: one carry-less wide product AND one DFMA per slot
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(fd[k]) : "d"(da), "d"(db));
      } else if (OP == 11) {  // two DFMAs per carry-chain pair
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(lo[k]), "+r"(lo[(k + 1) % ILP]) : "r"(x), "r"(y));
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(fd[k]) : "d"(da), "d"(db));
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(fd[(k + 3) % ILP]) : "d"(db), "d"(da));
      } else if (OP == 6) {  // carry chain pair (the field.cuh pattern)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(lo[k]), "+r"(lo[(k + 1) % ILP]) : "r"(x), "r"(y));
      }
    }
  }
  uint64_t s = 0;
  for (int k = 0; k < ILP; k++) s += acc[k] + lo[k] + hi[k] + kk[k] + (uint64_t)fd[k];
  out[tid] = s;
}

// the real thing: dependent fe_mul chains, NCH independent chains per thread
template <int NCH, int VAR>
__global__ void __launch_bounds__(128) probe_femul(nova::fe_t* out, int iters) {
  using namespace nova;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  fe_t x[NCH], y;
  for (int k = 0; k < NCH; k++)
    for (int i = 0; i < 8; i++) x[k].l[i] = tid * 31 + i + k;
  for (int i = 0; i < 8; i++) y.l[i] = tid + 77 * i;
  x[0].l[7] &= 0x0fffffff; y.l[7] &= 0x0fffffff;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) x[k] = VAR ? fe_mul_cs<BN254_FQ>(x[k], y) : fe_mul_chain<BN254_FQ>(x[k], y);
  }
  fe_t s = x[0];
  for (int k = 1; k < NCH; k++) s = fe_add<BN254_FQ>(s, x[k]);
  out[tid] = s;
}
// chain_mad block exactly as in field.cuh (4 wide products + carry-out), independent accumulators
__global__ void probe_chain(uint32_t* out, uint32_t a0, uint32_t b0) {
  using namespace nova;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t X[4][9];
  for (int k = 0; k < 4; k++) for (int i = 0; i < 9; i++) X[k][i] = tid + i + k;
  uint32_t x = a0 + tid, y = b0 | 1;
  for (int i = 0; i < ITER / 4; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) chain_mad<0, false>(X[k], x, x + 1, x + 2, x + 3, y);
  }
  uint32_t s = 0;
  for (int k = 0; k < 4; k++) for (int i = 0; i < 9; i++) s += X[k][i];
  out[tid] = s;
}

template <int NCH, int VAR = 0>
void run_femul(int blocks_per_sm) {
  int dev; cudaGetDevice(&dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int blocks = p.multiProcessorCount * blocks_per_sm, threads = 128, iters = 512;
  nova::fe_t* out; cudaMalloc(&out, (size_t)blocks * threads * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe_femul<NCH, VAR><<<blocks, threads>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe_femul<NCH, VAR><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * threads * iters * NCH;
  printf("%s chains=%d warps/SM=%2d  %8.3f ms  %7.2f G mul/s  (%.0f wide-mults/clk/SM of 64)\n", VAR ? "fe_mul_cs" : "fe_mul   ", NCH,
         blocks_per_sm * 4, ms, muls / ms / 1e6, muls * 136 / (ms * 1e-3) / 1.965e9 / p.multiProcessorCount);
  cudaFree(out);
}

void run_chain() {
  int dev; cudaGetDevice(&dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int blocks = p.multiProcessorCount * 4, threads = 256;
  uint32_t* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe_chain<<<blocks, threads>>>(out, 3, 5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe_chain<<<blocks, threads>>>(out, 3, 5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double prods = (double)blocks * threads * (ITER / 4) * 4 * 4;
  printf("chain_mad block (4 wide products + carry)   %8.3f ms  %6.1f wide-products/clk/SM\n", ms,
         prods / (ms * 1e-3) / 1.965e9 / p.multiProcessorCount);
  cudaFree(out);
}

template <int OP>
void run(const char* name, int ops_per_iter) {
  int dev; cudaGetDevice(&dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int blocks = p.multiProcessorCount * 4, threads = 256;
  uint64_t* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<OP><<<blocks, threads>>>(out, 3, 5, 1.5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<OP><<<blocks, threads>>>(out, 3, 5, 1.5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * ITER * ILP * ops_per_iter;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  double per_clk_sm = ops / (ms * 1e-3) / (clk * 1e3) / p.multiProcessorCount;
  printf("%-28s %8.3f ms  %8.2f Gop/s  %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ms,
         ops / ms / 1e6, per_clk_sm, clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("mad.wide.u32 (IMAD.WIDE)", 1);
  run<1>("mad.lo.u32 (IMAD)", 1);
  run<5>("mad.hi.u32", 1);
  run<6>("mad.lo.cc+madc.hi pair", 1);
  run<2>("add.u32 (IADD3)", 1);
  run<3>("add.u64", 1);
  run<4>("fma.rz.f64 (DFMA)", 1);
  run<9>("chain pair + 1 DFMA (slots)", 1);
  run<10>("IMAD.WIDE + 1 DFMA (slots)", 1);
  run<11>("chain pair + 2 DFMA (slots)", 1);
  run_chain();
  run_femul<1>(2); run_femul<1>(4); run_femul<1>(8);
  run_femul<2>(2); run_femul<2>(4);
  run_femul<4>(2);
  run<7>("wide mad carry-out + addc", 1);
  run<8>("wide mad carry-out only", 1);
  run_femul<1, 1>(2); run_femul<1, 1>(4); run_femul<1, 1>(8);
  run_femul<2, 1>(2); run_femul<2, 1>(4);
  run_femul<4, 1>(2);
  return 0;
}
