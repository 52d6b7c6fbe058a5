// Kernel instantiations for the field BN254_FQ (see ops_impl.cuh).
#include "ops_impl.cuh"
namespace nova {
const field_ops OPS_BN254_FQ = ops_impl<BN254_FQ>::table();
}
