// Times the batched sum-check round kernel (k_sc_round_batched, nine claims as in ppsnark prove_helper) alone and
// prints where its cycles go (clock64 stamps at the section boundaries).  Build (two variants):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/roundbench.cu -o /tmp/roundbench
//   nvcc ... -DNOVA_ROUND_INLINE_MUL tools/roundbench.cu -o /tmp/roundbench_inl
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
__device__ long long g_stamp[16];
#define SCB_STAMP(i) \
  if ((threadIdx.x & 31u) == 0) g_stamp[i] = clock64();
#include "../nova_b200/csrc/transcript_batched.cuh"
using namespace nova;
using F = BN254_FR;

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));         \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main() {
  scb_desc d;
  memset(&d, 0, sizeof(d));
  d.nclaims = 9;
  d.neq = 2;
  const int kinds[9] = {SCB_LIN2, SCB_LIN2, SCB_EQ_DEG2, SCB_EQ_DEG2, SCB_EQ_DEG2, SCB_EQ_DEG2, SCB_RAW3, SCB_EQ_DEG1, SCB_LIN2};
  const int eqof[9] = {-1, -1, 0, 0, 0, 0, -1, 1, -1};
  std::vector<uint32_t> host(8 * 64);
  uint64_t x = 88172645463325252ull;
  for (auto& w : host) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    w = (uint32_t)x;
  }
  for (size_t i = 7; i < host.size(); i += 8) host[i] &= 0x0fffffffu;  // < p
  void *sums, *taus, *state, *poly, *r;
  CK(cudaMalloc(&sums, 64 * 32));
  CK(cudaMalloc(&taus, 8 * 32));
  CK(cudaMalloc(&state, sizeof(scb_state)));
  CK(cudaMalloc(&poly, 96));
  CK(cudaMalloc(&r, 32));
  CK(cudaMemcpy(sums, host.data(), 64 * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(taus, host.data() + 64, 8 * 32, cudaMemcpyHostToDevice));
  std::vector<unsigned char> st(sizeof(scb_state), 0);
  memcpy(st.data(), host.data(), 32);
  for (size_t i = 144; i + 32 <= st.size(); i += 32) memcpy(st.data() + i, host.data() + 8 * ((i / 32) % 40), 32);
  CK(cudaMemcpy(state, st.data(), st.size(), cudaMemcpyHostToDevice));
  for (int i = 0; i < 9; i++) {
    d.kind[i] = kinds[i];
    d.eq_of[i] = eqof[i];
    d.slot[i] = 3 * i;
    d.slot_m1[i] = -1;
  }
  for (int g = 0; g < 2; g++) {
    d.tau[g] = (char*)taus + 64 * g;
    d.tau_inv[g] = (char*)taus + 64 * g + 32;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int reps = 200;
  for (int pass = 0; pass < 2; pass++) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_sc_round_batched<F><<<1, 32>>>(d, (scb_state*)state, sums, nullptr, 0, 'p', 'c', poly, r);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("pass %d: %.2f us per launch (back-to-back launches)\n", pass, ms * 1e3 / reps);
  }
  long long s[16];
  CK(cudaMemcpyFromSymbol(s, g_stamp, sizeof(s)));
  const char* names[7] = {"load state + claim evals", "combine", "poly + compress", "message", "keccak x4 (2 digests)",
                          "finish (from_uniform, evaluate)", "claim updates + bounds + stores"};
  for (int i = 0; i < 7; i++) printf("  %-34s %8lld cycles\n", names[i], s[i + 1] - s[i]);
  printf("  %-34s %8lld cycles\n", "total", s[7] - s[0]);
  return 0;
}
