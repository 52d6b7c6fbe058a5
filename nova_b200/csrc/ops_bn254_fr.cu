// Kernel instantiations for the field BN254_FR (see ops_impl.cuh).
#include "ops_impl.cuh"
namespace nova {
const field_ops OPS_BN254_FR = ops_impl<BN254_FR>::table();
}
