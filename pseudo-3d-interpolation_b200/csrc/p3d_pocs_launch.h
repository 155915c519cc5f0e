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
// in-place forward column FFT of slices (the first half of k_cols_generic), used by the percentile operators
void generic_fft_cols(const GenericCfg& c, const AxisDev<float>& ax1, Cx<float>* data, int nslices, cudaStream_t st);
void generic_fft2(const GenericCfg& c, const AxisDev<float>& ax1, const AxisDev<float>& ax2, const Cx<float>* in,
                  Cx<float>* out, int nslices, int inverse, cudaStream_t st);

}  // namespace p3d

// ---- float64 state mode (p3d_pocs_generic64.cu): same kernels instantiated for double ---------
namespace p3d {
cudaError_t generic64_configure(const GenericCfg& c);
void generic64_rows_init(const GenericCfg& c, const AxisDev<double>& ax2, const BandArgs<double>& A, int nslices, cudaStream_t st);
void generic64_cols_stats(const GenericCfg& c, const AxisDev<double>& ax1, const BandArgs<double>& A, int nslices, cudaStream_t st);
void generic64_cols_iter(const GenericCfg& c, const AxisDev<double>& ax1, const BandArgs<double>& A, int nslices, int op, cudaStream_t st);
void generic64_rows_iter(const GenericCfg& c, const AxisDev<double>& ax2, const BandArgs<double>& A, int nslices, cudaStream_t st);
// X0 (complex128, [nslices][ne]) -> stats[s].im64_key = max imaginary part among elements whose real part is the maximum
void generic64_lexmax_imag(const Cx<double>* x0, SliceStats* stats, long long ne, int nslices, cudaStream_t st);
void convert_c64_to_c128(const Cx<float>* in, Cx<double>* out, long long n, cudaStream_t st);
void convert_c128_to_c64(const Cx<double>* in, Cx<float>* out, long long n, cudaStream_t st);
}  // namespace p3d
