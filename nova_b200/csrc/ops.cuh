// Per-field launcher table.  Every kernel in this library is a template over one prime field
// (point kernels over the curve's BASE field, digit/field-vector kernels over the field the
// vector lives in); each ops_<field>.cu instantiates the whole set once and exports it through
// a `field_ops` table so the C ABI (capi.cu) is template-free and the four translation units
// compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace nova {

struct msm_plan;

struct field_ops {
  int field_id;
  // --- MSM stages -------------------------------------------------------------------------
  // scalar-field side
  void (*digits)(cudaStream_t, const void* scalars, const msm_plan&);
  // base-field side
  void (*expand_key)(cudaStream_t, void* tables, size_t n_ck, int ntables, int shift);
  void (*accumulate)(cudaStream_t, const void* tables, const msm_plan&);
  void (*fixup)(cudaStream_t, const msm_plan&);
  void (*reduce)(cudaStream_t, const msm_plan&, void* out_jac);
  // small helpers on points (base field)
  void (*sum_points)(cudaStream_t, const void* tables, const uint32_t* idx_or_null, size_t n,
                     void* scratch_xyzz, void* out_jac);
  void (*jacobian_sum)(cudaStream_t, const void* pts, int k, void* out_jac);
  void (*index_bases)(cudaStream_t, void* bases, size_t n, const void* gen_affine, uint64_t k0);
  // --- field-vector kernels (K4..K8) --------------------------------------------------------
  void (*cross_term)(cudaStream_t, const void* az, const void* bz, const void* cz, const void* e1,
                     const void* e2_or_null, const void* u, size_t n, void* t);
  void (*axpy)(cudaStream_t, const void* a, const void* b, const void* r, size_t n, void* out);
  void (*vec_add)(cudaStream_t, const void* a, const void* b, size_t n, void* out);
  void (*bind_top)(cudaStream_t, void* z, size_t n, const void* r);
};

extern const field_ops OPS_BN254_FR, OPS_BN254_FQ, OPS_PALLAS_FP, OPS_PALLAS_FQ;

inline const field_ops* ops_for_field(int fid) {
  switch (fid) {
    case 0: return &OPS_BN254_FR;
    case 1: return &OPS_BN254_FQ;
    case 2: return &OPS_PALLAS_FP;
    case 3: return &OPS_PALLAS_FQ;
    default: return nullptr;
  }
}

// field-independent MSM stages (msm_common.cu)
void msm_scan(cudaStream_t, const msm_plan&);
void msm_scatter(cudaStream_t, const msm_plan&);
void msm_digits_small(cudaStream_t, const void* scalars, int elem_bytes, const msm_plan&);

}  // namespace nova
