// Specialised kernels are registered here (none yet: every shape uses the generic path).
#include "p3d_pocs_spec.cuh"

namespace p3d {
SpecKernels select_spec_kernels(int n_iline, int n_xline) {
    (void)n_iline; (void)n_xline;
    return SpecKernels();
}
}  // namespace p3d
