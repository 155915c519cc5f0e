// Generic-path kernels instantiated for double: the float64 state mode ("precision = 64").
// Same algorithm, complex128 iterate / spectrum / thresholds; used when results must match the
// reference's float64 arithmetic beyond what fp32 can decide (hard-threshold flips, near-tied
// lexicographic maxima).
#include "p3d_pocs_launch.h"
#include <algorithm>

namespace p3d {

template <typename K> static cudaError_t set_smem64(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t generic64_configure(const GenericCfg& cfg) {
    int dev = 0, optin = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    const size_t lim = (size_t)optin - 1024;
    if (cfg.col_smem > lim || cfg.row_smem > lim) return cudaErrorInvalidValue;
#define P3D_SET(k) if ((e = set_smem64(k, lim)) != cudaSuccess) return e
    P3D_SET((k_cols_generic<double, 0, 0>));
    P3D_SET((k_cols_generic<double, 1, P3D_OP_HARD>));
    P3D_SET((k_cols_generic<double, 1, P3D_OP_SOFT>));
    P3D_SET((k_cols_generic<double, 1, P3D_OP_GARROTE>));
    P3D_SET((k_cols_generic<double, 1, P3D_OP_RESTART>));
    P3D_SET((k_rows_generic<double, 0>));
    P3D_SET((k_rows_generic<double, 1>));
#undef P3D_SET
    return cudaSuccess;
}

static inline dim3 row_grid(const GenericCfg& c, int ns) { return dim3((c.geom.n1 + c.geom.RB - 1) / c.geom.RB, ns); }
static inline dim3 col_grid(const GenericCfg& c, int ns) { return dim3((c.geom.n2 + c.geom.C - 1) / c.geom.C, ns); }

void generic64_rows_init(const GenericCfg& c, const AxisDev<double>& ax2, const BandArgs<double>& A, int ns, cudaStream_t st) {
    k_rows_generic<double, 0><<<row_grid(c, ns), c.row_threads, c.row_smem, st>>>(c.geom, ax2, A);
}
void generic64_cols_stats(const GenericCfg& c, const AxisDev<double>& ax1, const BandArgs<double>& A, int ns, cudaStream_t st) {
    k_cols_generic<double, 0, 0><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A);
}
void generic64_cols_iter(const GenericCfg& c, const AxisDev<double>& ax1, const BandArgs<double>& A, int ns, int op, cudaStream_t st) {
    switch (op) {
        case P3D_OP_HARD: k_cols_generic<double, 1, P3D_OP_HARD><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
        case P3D_OP_SOFT: k_cols_generic<double, 1, P3D_OP_SOFT><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
        case P3D_OP_RESTART: k_cols_generic<double, 1, P3D_OP_RESTART><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
        default:          k_cols_generic<double, 1, P3D_OP_GARROTE><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
    }
}
void generic64_rows_iter(const GenericCfg& c, const AxisDev<double>& ax2, const BandArgs<double>& A, int ns, cudaStream_t st) {
    k_rows_generic<double, 1><<<row_grid(c, ns), c.row_threads, c.row_smem, st>>>(c.geom, ax2, A);
}

__global__ void k_lexmax_imag(const Cx<double>* __restrict__ x0, SliceStats* stats, long long ne) {
    const int s = blockIdx.y;
    const unsigned long long rk = stats[s].re64_key;
    const Cx<double>* p = x0 + (long long)s * ne;
    unsigned long long best = 0ull;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ne; i += (long long)gridDim.x * blockDim.x) {
        const Cx<double> v = p[i];
        if (f64_ordered(v.x) == rk) { const unsigned long long k = f64_ordered(v.y); best = k > best ? k : best; }
    }
    best = warp_max_u64(best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&stats[s].im64_key, best);
}
void generic64_lexmax_imag(const Cx<double>* x0, SliceStats* stats, long long ne, int ns, cudaStream_t st) {
    dim3 grid((unsigned)std::min<long long>((ne + 255) / 256, 296), ns);
    k_lexmax_imag<<<grid, 256, 0, st>>>(x0, stats, ne);
}

__global__ void k_c64_to_c128(const Cx<float>* __restrict__ in, Cx<double>* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const Cx<float> v = in[i];
        out[i] = cmake<double>((double)v.x, (double)v.y);
    }
}
__global__ void k_c128_to_c64(const Cx<double>* __restrict__ in, Cx<float>* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const Cx<double> v = in[i];
        out[i] = cmake<float>((float)v.x, (float)v.y);
    }
}
void convert_c64_to_c128(const Cx<float>* in, Cx<double>* out, long long n, cudaStream_t st) {
    k_c64_to_c128<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(in, out, n);
}
void convert_c128_to_c64(const Cx<double>* in, Cx<float>* out, long long n, cudaStream_t st) {
    k_c128_to_c64<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(in, out, n);
}

}  // namespace p3d
