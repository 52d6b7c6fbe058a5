// The REAL kernel wrappers k_sc_round / k_sc_round_batched (nova_b200/csrc/transcript*.cuh) run on the CPU
// through the SIMT shim: 32 threads = 32 lanes.  Same C signatures as the sequential hooks in hostcheck.cpp
// (hc_sc_round, hc_sc_round_batched), so the tests can drive complete proofs through either.
#include <cstring>
#include "simt_host.h"
#include "../../nova_b200/csrc/transcript_batched.cuh"
using namespace nova;

extern "C" int hc_simt_sc_round(int fid, int kind, void* state144, const void* res, const void* tau, const void* tau_inv,
                                const void* pending, uint32_t pending_len, int absorb_label, int squeeze_label,
                                void* out_poly, void* out_r) {
  alignas(16) sc_state st;
  memcpy(&st, state144, 144);
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_warp([&] {
      k_sc_round<F>(kind, &st, res, tau, tau_inv, (const uint8_t*)pending, pending_len, (uint8_t)absorb_label,
                    (uint8_t)squeeze_label, out_poly, out_r);
    });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 1: run(BN254_FQ{}); break;
    case 2: run(PALLAS_FP{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  memcpy(state144, &st, 144);
  return 0;
}

extern "C" int hc_simt_sc_round_batched(int fid, const void* desc, void* state, const void* sums, const void* pending,
                                        uint32_t pending_len, int absorb_label, int squeeze_label, void* out_poly,
                                        void* out_r) {
  const scb_desc d = *(const scb_desc*)desc;
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_warp([&] {
      k_sc_round_batched<F>(d, (scb_state*)state, sums, (const uint8_t*)pending, pending_len, (uint8_t)absorb_label,
                            (uint8_t)squeeze_label, out_poly, out_r);
    });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 1: run(BN254_FQ{}); break;
    case 2: run(PALLAS_FP{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}
