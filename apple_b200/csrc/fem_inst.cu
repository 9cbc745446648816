// One translation unit per (scalar type, energy): compiled six times by apple_b200/build.py with
//   -DAPL_INST_T=float|double  -DAPL_INST_KIND=0|1|2
// so that the 9 operator combinations x 2 assembly variants of each pair build in parallel.
#include "fem_kernels.cuh"

#ifndef APL_INST_T
#error "APL_INST_T / APL_INST_KIND must be defined"
#endif

namespace apl {
template <>
int launch_fem<APL_INST_T, APL_INST_KIND>(const apl_fem* fem, int ops, const FemArgs<APL_INST_T>& args,
                                           int scatter, cudaStream_t stream) {
    return launch_fem_impl<APL_INST_T, APL_INST_KIND>(fem, ops, args, scatter, stream);
}
}  // namespace apl
