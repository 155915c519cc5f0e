// Generic-path kernels: instantiation + launchers (kept in their own translation unit so the
// host logic can be rebuilt without recompiling them).
#include "p3d_pocs_launch.h"

namespace p3d {

// plain 2-D FFT of slices through the same tiles (numpy fft2 / ifft2 semantics)
template <int DIR>
__global__ void k_fft2_rows(const __grid_constant__ PocsGeom G, const __grid_constant__ AxisDev<float> ax2, const Cx<float>* in, Cx<float>* outp, float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
    const int r0 = blockIdx.x * G.RB, nr = min(G.RB, G.n1 - r0);
    Cx<float>* bufA = reinterpret_cast<Cx<float>*>(smem_raw);
    Cx<float>* bufB = bufA + (size_t)G.pitch2 * G.RB;
    const long long off = (long long)s * G.n1 * G.n2 + (long long)r0 * G.n2;
    const int tot = nr * G.n2;
    for (int w = tid; w < tot; w += nth) { const int rr = w / G.n2, j = w - rr * G.n2; bufA[rr * G.pitch2 + j] = in[off + (long long)rr * G.n2 + j]; }
    __syncthreads();
    TileGeom tg; tg.nlines = nr; tg.line_stride = G.pitch2; tg.elem_stride = 1; tg.line_fastest = 0;
    Cx<float>* X = line_fft<DIR, float>(bufA, bufB, tg, ax2, tid, nth);
    for (int w = tid; w < tot; w += nth) { const int rr = w / G.n2, j = w - rr * G.n2; outp[off + (long long)rr * G.n2 + j] = cscale(X[rr * G.pitch2 + j], scale); }
}
template <int DIR>
__global__ void k_fft2_cols(const __grid_constant__ PocsGeom G, const __grid_constant__ AxisDev<float> ax1, Cx<float>* data) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
    const int c0 = blockIdx.x * G.C, nc = min(G.C, G.n2 - c0);
    Cx<float>* bufA = reinterpret_cast<Cx<float>*>(smem_raw);
    Cx<float>* bufB = bufA + (size_t)ax1.L * G.C;
    Cx<float>* Ws = data + (long long)s * G.n1 * G.n2;
    const int tot = G.n1 * nc;
    for (int w = tid; w < tot; w += nth) { const int i = w / nc, c = w - i * nc; bufA[i * G.C + c] = Ws[(long long)i * G.n2 + c0 + c]; }
    __syncthreads();
    TileGeom tg; tg.nlines = nc; tg.line_stride = 1; tg.elem_stride = G.C; tg.line_fastest = 1;
    Cx<float>* X = line_fft<DIR, float>(bufA, bufB, tg, ax1, tid, nth);
    for (int w = tid; w < tot; w += nth) { const int i = w / nc, c = w - i * nc; Ws[(long long)i * G.n2 + c0 + c] = X[i * G.C + c]; }
}


template <typename K> static cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t generic_configure(const GenericCfg& cfg) {
    // The attribute is per kernel, not per plan: always raise it to the device's opt-in maximum
    // so that plans with different tile sizes can coexist in one process.
    int dev = 0, optin = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    // (kernels with a little static shared memory need headroom below the opt-in limit)
    struct { size_t col_smem, row_smem; } c = {(size_t)optin - 1024, (size_t)optin - 1024};
    if (cfg.col_smem > c.col_smem || cfg.row_smem > c.row_smem) return cudaErrorInvalidValue;
#define P3D_SET(k, b) if ((e = set_smem(k, b)) != cudaSuccess) return e
    P3D_SET((k_cols_generic<float, 0, 0>), c.col_smem);
    P3D_SET((k_cols_generic<float, 1, P3D_OP_HARD>), c.col_smem);
    P3D_SET((k_cols_generic<float, 1, P3D_OP_SOFT>), c.col_smem);
    P3D_SET((k_cols_generic<float, 1, P3D_OP_GARROTE>), c.col_smem);
    P3D_SET((k_cols_generic<float, 1, P3D_OP_FILTER>), c.col_smem);
    P3D_SET((k_rows_generic<float, 0>), c.row_smem);
    P3D_SET((k_rows_generic<float, 1>), c.row_smem);
    P3D_SET(k_fft2_rows<-1>, c.row_smem); P3D_SET(k_fft2_rows<+1>, c.row_smem);
    P3D_SET(k_fft2_cols<-1>, c.col_smem); P3D_SET(k_fft2_cols<+1>, c.col_smem);
#undef P3D_SET
    return cudaSuccess;
}

static inline dim3 row_grid(const GenericCfg& c, int ns) { return dim3((c.geom.n1 + c.geom.RB - 1) / c.geom.RB, ns); }
static inline dim3 col_grid(const GenericCfg& c, int ns) { return dim3((c.geom.n2 + c.geom.C - 1) / c.geom.C, ns); }

void generic_rows_init(const GenericCfg& c, const AxisDev<float>& ax2, const BandArgs<float>& A, int ns, cudaStream_t st) {
    k_rows_generic<float, 0><<<row_grid(c, ns), c.row_threads, c.row_smem, st>>>(c.geom, ax2, A);
}
void generic_cols_stats(const GenericCfg& c, const AxisDev<float>& ax1, const BandArgs<float>& A, int ns, cudaStream_t st) {
    k_cols_generic<float, 0, 0><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A);
}
void generic_cols_iter(const GenericCfg& c, const AxisDev<float>& ax1, const BandArgs<float>& A, int ns, int op, cudaStream_t st) {
    switch (op) {
        case P3D_OP_HARD: k_cols_generic<float, 1, P3D_OP_HARD><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
        case P3D_OP_SOFT: k_cols_generic<float, 1, P3D_OP_SOFT><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
        case P3D_OP_FILTER: k_cols_generic<float, 1, P3D_OP_FILTER><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
        default:          k_cols_generic<float, 1, P3D_OP_GARROTE><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, A); break;
    }
}
void generic_rows_iter(const GenericCfg& c, const AxisDev<float>& ax2, const BandArgs<float>& A, int ns, cudaStream_t st) {
    k_rows_generic<float, 1><<<row_grid(c, ns), c.row_threads, c.row_smem, st>>>(c.geom, ax2, A);
}
void generic_fft_cols(const GenericCfg& c, const AxisDev<float>& ax1, Cx<float>* data, int ns, cudaStream_t st) {
    k_fft2_cols<-1><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, data);
}
void generic_fft2(const GenericCfg& c, const AxisDev<float>& ax1, const AxisDev<float>& ax2, const Cx<float>* in,
                  Cx<float>* out, int ns, int inverse, cudaStream_t st) {
    const float scale = inverse ? (float)(1.0 / ((double)c.geom.n1 * (double)c.geom.n2)) : 1.0f;
    if (!inverse) {
        k_fft2_rows<-1><<<row_grid(c, ns), c.row_threads, c.row_smem, st>>>(c.geom, ax2, in, out, scale);
        k_fft2_cols<-1><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, out);
    } else {
        k_fft2_rows<+1><<<row_grid(c, ns), c.row_threads, c.row_smem, st>>>(c.geom, ax2, in, out, scale);
        k_fft2_cols<+1><<<col_grid(c, ns), c.col_threads, c.col_smem, st>>>(c.geom, ax1, out);
    }
}

}  // namespace p3d
