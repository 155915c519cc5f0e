// Specialised (compile-time size, register-resident) POCS kernels.  See DESIGN.md.
#pragma once
#include "p3d_pocs_kernels.cuh"

namespace p3d {

typedef void (*ColsIterLaunch)(const PocsGeom&, const AxisDev<float>&, const BandArgs<float>&, int nslices, int op, cudaStream_t);
typedef void (*RowsIterLaunch)(const PocsGeom&, const AxisDev<float>&, const BandArgs<float>&, int nslices, cudaStream_t);

struct SpecKernels {
    ColsIterLaunch cols_iter = nullptr;
    RowsIterLaunch rows_iter = nullptr;
    const char* cols_name = "generic";
    const char* rows_name = "generic";
};

SpecKernels select_spec_kernels(int n_iline, int n_xline, int variant = 0);

}  // namespace p3d
