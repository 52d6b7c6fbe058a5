// Kernel instantiations for the field PALLAS_FP (see ops_impl.cuh).
#include "ops_impl.cuh"
namespace nova {
const field_ops OPS_PALLAS_FP = ops_impl<PALLAS_FP>::table();
}
