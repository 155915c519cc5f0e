// Launch entry points of the generic kernels (defined in p3d_pocs_generic.cu).
#pragma once
#include "p3d_pocs_kernels.cuh"

namespace p3d {

struct GenericCfg {
    PocsGeom geom;
    int col_threads, row_threads;
    size_t col_smem, row_smem;
};

cudaError_t generic_configure(const GenericCfg& c);
void generic_rows_init(const GenericCfg& c, const AxisDev<float>& ax2, const BandArgs<float>& A, int nslices, cudaStream_t st);
void generic_cols_stats(const GenericCfg& c, const AxisDev<float>& ax1, const BandArgs<float>& A, int nslices, cudaStream_t st);
void generic_cols_iter(const GenericCfg& c, const AxisDev<float>& ax1, const BandArgs<float>& A, int nslices, int op, cudaStream_t st);
void generic_rows_iter(const GenericCfg& c, const AxisDev<float>& ax2, const BandArgs<float>& A, int nslices, cudaStream_t st);
void generic_fft2(const GenericCfg& c, const AxisDev<float>& ax1, const AxisDev<float>& ax2, const Cx<float>* in,
                  Cx<float>* out, int nslices, int inverse, cudaStream_t st);

}  // namespace p3d
