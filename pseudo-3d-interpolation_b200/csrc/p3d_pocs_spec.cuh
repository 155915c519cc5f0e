// Specialised (compile-time size, register-resident) POCS kernels.  See DESIGN.md.
#pragma once
#include <vector>
#include "p3d_pocs_kernels.cuh"

namespace p3d {

// tw: the per-pass twiddle tables of the line plan (spec_twiddle_table), device memory
typedef void (*ColsIterLaunch)(const PocsGeom&, const Cx<float>* tw, const BandArgs<float>&, int nslices, int op, cudaStream_t);
typedef void (*RowsIterLaunch)(const PocsGeom&, const Cx<float>* tw, const BandArgs<float>&, int nslices, cudaStream_t);
// packs mask bytes into one word per (row, thread): bit e = mask[row][j + e*T]
typedef void (*PackMaskLaunch)(const uint8_t* mask, uint32_t* bits, int n_masks, int n1, cudaStream_t);

struct SpecKernels {
    ColsIterLaunch cols_iter = nullptr;
    RowsIterLaunch rows_iter = nullptr;
    RowsIterLaunch rows_init = nullptr;   // row FFT of the observed slice (+ nnz, sum|d|, APOCS prologue)
    RowsIterLaunch cols_stats = nullptr;  // column FFT + schedule statistics (same signature)
    PackMaskLaunch pack_mask = nullptr;
    const char* cols_name = "generic";
    const char* rows_name = "generic";
    std::vector<int> cols_radices, rows_radices;   // for the twiddle tables
    int rows_T = 0;                                 // threads per row line (packed-mask layout)
    // optional: builder of the column kernels' whole table buffer (twiddles + algorithm tables, e.g. Rader's)
    std::vector<Cx<float>> (*cols_table)() = nullptr;
    std::vector<Cx<float>> (*rows_table)() = nullptr;
};

// prime iline counts handled by Rader's algorithm (p3d_pocs_rader.cu); no-op for other lengths
void rader_register_cols(SpecKernels& k, int n_iline, int variant);
void rader_register_rows(SpecKernels& k, int n_xline, int variant);
// further lengths with register plans (p3d_pocs_spec_more.cu, p3d_pocs_spec_mix*.cu); false when the length has none
bool more_register_cols(SpecKernels& k, int n_iline);
bool more_register_rows(SpecKernels& k, int n_xline);
bool mix_register_cols(SpecKernels& k, int n_iline);
bool mix_register_rows(SpecKernels& k, int n_xline);
bool mix2_register_cols(SpecKernels& k, int n_iline);
bool mix2_register_rows(SpecKernels& k, int n_xline);
bool mix3_register_cols(SpecKernels& k, int n_iline);
bool mix3_register_rows(SpecKernels& k, int n_xline);

SpecKernels select_spec_kernels(int n_iline, int n_xline, int variant = 0);

// Host: concatenated [k][R] tables of every twiddled pass (see p3d_fft_reg.cuh), computed in double.
std::vector<Cx<float>> spec_twiddle_table(const std::vector<int>& radices);


// ---- float64 state mode ("precision" = 64): iteration kernels only (the once-per-slice kernels stay generic)
typedef void (*ColsIterLaunch64)(const PocsGeom&, const Cx<double>* tw, const BandArgs<double>&, int nslices, int op, cudaStream_t);
typedef void (*RowsIterLaunch64)(const PocsGeom&, const Cx<double>* tw, const BandArgs<double>&, int nslices, cudaStream_t);
struct SpecKernels64 {
    ColsIterLaunch64 cols_iter = nullptr;
    RowsIterLaunch64 rows_iter = nullptr;
    // escalating-precision mode: complex128 state beside complex64 observed data / results (BandArgs::D32 / OUT32)
    RowsIterLaunch64 rows_iter_io32 = nullptr;
    RowsIterLaunch64 rows_init_io32 = nullptr;   // row FFT (complex128) of the complex64 observed slice
    RowsIterLaunch64 cols_stats = nullptr;       // column FFT + exact statistics (per-CTA lexicographic maxima in BandArgs::cand)
    int cols_C = 0;                               // columns per tile of cols_stats (cand entries per slice = ceil(n2 / cols_C))
    PackMaskLaunch pack_mask = nullptr;
    const char* cols_name = "generic64";
    const char* rows_name = "generic64";
    std::vector<int> cols_radices, rows_radices;
    int rows_T = 0;
    std::vector<Cx<double>> (*cols_table)() = nullptr;   // whole table buffer of the column kernels (Rader) instead of plain twiddles
};
void rader_register_cols64(SpecKernels64& k, int n_iline, int variant);
SpecKernels64 select_spec_kernels64(int n_iline, int n_xline, int variant = 0);
std::vector<Cx<double>> spec_twiddle_table64(const std::vector<int>& radices);

}  // namespace p3d
