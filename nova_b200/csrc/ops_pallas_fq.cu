// Kernel instantiations for the field PALLAS_FQ (see ops_impl.cuh).
#include "ops_impl.cuh"
namespace nova {
const field_ops OPS_PALLAS_FQ = ops_impl<PALLAS_FQ>::table();
}
