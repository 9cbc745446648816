// One translation unit per (scalar type, energy), compiled by apple_b200/build.py with
//   -DAPL_INST_T=float|double  -DAPL_INST_KIND=0|1|2|3  [-DAPL_INST_SUPERSET]
// so that the operator combinations x assembly variants of each pair build in parallel.  Without APL_INST_SUPERSET
// the unit holds the classic operator sets (the five reference operators and their fused forms); with it, ONLY the
// opt-in supersets (APL_OP_HESS_OFFD / APL_OP_PSD) -- kept apart because their presence in the same unit changes the
// code nvcc generates for the classic kernels (see fem_kernels.cuh).
#include "fem_kernels.cuh"

#ifndef APL_INST_T
#error "APL_INST_T / APL_INST_KIND must be defined"
#endif

namespace apl {
#ifdef APL_INST_SUPERSET
template <>
int launch_fem_superset<APL_INST_T, APL_INST_KIND>(const apl_fem* fem, int ops, const FemArgs<APL_INST_T>& args,
                                                    int scatter, cudaStream_t stream) {
    return launch_fem_superset_impl<APL_INST_T, APL_INST_KIND>(fem, ops, args, scatter, stream);
}
#else
template <>
int launch_fem<APL_INST_T, APL_INST_KIND>(const apl_fem* fem, int ops, const FemArgs<APL_INST_T>& args,
                                           int scatter, cudaStream_t stream) {
    return launch_fem_impl<APL_INST_T, APL_INST_KIND>(fem, ops, args, scatter, stream);
}
#endif
}  // namespace apl
