// Specialised POCS iteration kernels: compile-time line length, register-resident FFTs
// (p3d_fft_reg.cuh), fused epilogues.  Selected per axis length at plan creation; every
// other shape falls back to the generic kernels.
//
//   cols_iter:  thread (c, j) of a tile of C adjacent columns holds rows j + e*T of column c.
//               load (C*8-byte row segments) -> FFT -> threshold(tau_k) -> IFFT -> store.
//   rows_iter:  thread (j, rr) of a tile of RB rows holds columns j + e*T of row rr.
//               load -> IFFT -> x = alpha*d + (1-alpha*m)*y/(N1 N2) -> sum|x| -> [OUT] -> FFT -> store.
#include "p3d_pocs_spec.cuh"
#include "p3d_fft_reg.cuh"

namespace p3d {

// ---- shared-memory accessors ---------------------------------------------------------------------
// column tile: [pos][c], c fastest.  For C = 8 two consecutive positions share a 128-byte
// bank row; the gray-code swizzle keeps every access pattern of the Stockham passes
// (consecutive positions, stride-R positions) conflict free.
template <int C> struct ColAcc {
    Cx<float>* base;   // already offset by c
    __device__ __forceinline__ Cx<float>& at(int pos) const {
        if (C == 8) {
            const int g = (pos ^ (pos >> 1)) & 1;
            return base[((pos >> 1) << 4) + (g << 3)];
        }
        return base[pos * C];
    }
};
// row tile: [row][pos + pad], one padding element every 32 positions
struct RowAcc {
    Cx<float>* base;   // already offset by the row
    __device__ __forceinline__ Cx<float>& at(int pos) const { return base[pos + (pos >> 5)]; }
};
__host__ __device__ constexpr int row_pitch(int n) { return n + (n >> 5) + 1; }

// ---- column kernel ---------------------------------------------------------------------------------
template <typename LP, int C, int MINB>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_cols_spec(const __grid_constant__ PocsGeom G, const Cx<float>* __restrict__ tw, const __grid_constant__ BandArgs<float> A, const int op) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = blockIdx.y;
    if (slice_stopped(A.stop, A.S, s, A.k, A.niter, A.eps)) return;
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const int col = blockIdx.x * C + c;
    const bool ok = col < G.n2;
    Cx<float>* Ws = A.W + (long long)s * N * G.n2 + col;
    ColAcc<C> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + c;

    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ok ? Ws[(long long)(j + e * T) * G.n2] : cmake<float>(0.f, 0.f);

    LP::template fft<-1, float>(v, acc, j, tw);

    const Cx<float> tau = A.tau[(long long)s * A.niter + A.k];
    const float a = tau.x, b = tau.y;
    const float t2re = a * a - b * b, t2im = 2.f * a * b;
    if (op == P3D_OP_HARD) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_HARD, float>(v[e], a, b, t2re, t2im);
    } else if (op == P3D_OP_SOFT) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_SOFT, float>(v[e], a, b, t2re, t2im);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_GARROTE, float>(v[e], a, b, t2re, t2im);
    }

    LP::template fft<+1, float>(v, acc, j, tw);

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Ws[(long long)(j + e * T) * G.n2] = v[e];
    }
}

// ---- row kernel ----------------------------------------------------------------------------------
template <typename LP, int RB, int MINB>
__global__ void __launch_bounds__(LP::T* RB, MINB)
k_rows_spec(const __grid_constant__ PocsGeom G, const Cx<float>* __restrict__ tw, const __grid_constant__ BandArgs<float> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red_s[32];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = blockIdx.y;
    if (A.stop[s] != 0) return;
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const int row = blockIdx.x * RB + rr;
    const bool ok = row < G.n1;
    const long long off = (long long)s * G.n1 * N + (long long)row * N + j;
    const long long moff = ((A.first_slice + s) / G.slices_per_mask) * (long long)G.n1 * N + (long long)row * N + j;
    RowAcc acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + rr * row_pitch(N);

    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ok ? A.W[off + e * T] : cmake<float>(0.f, 0.f);

    LP::template fft<+1, float>(v, acc, j, tw);

    float part = 0.f;
    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const Cx<float> d = A.D[off + e * T];
            const float m = (float)A.mask[moff + e * T];
            const float coef = (1.f - A.alpha * m) * A.inv_n;
            Cx<float> x = cmake<float>(fmaf(coef, v[e].x, A.alpha * d.x), fmaf(coef, v[e].y, A.alpha * d.y));
            part += sqrtf(x.x * x.x + x.y * x.y);
            if (A.write_out) A.OUT[off + e * T] = x;
            if (A.adaptive) {
                const float keep = 1.f - A.alpha * m, om = 1.f - A.alpha;
                const Cx<float> xt = cmake<float>(A.alpha * d.x + keep * x.x, A.alpha * d.y + keep * x.y);
                x = cmake<float>(xt.x + om * (d.x - m * x.x), xt.y + om * (d.y - m * x.y));
            }
            v[e] = x;
        }
    }
    double dp = warp_sum((double)part);
    if ((tid & 31) == 0) red_s[tid >> 5] = dp;
    __syncthreads();
    if (tid < 32) {
        constexpr int NW = (T * RB + 31) / 32;
        double t = tid < NW ? red_s[tid] : 0.0;
        t = warp_sum(t);
        if (tid == 0) atomicAdd(&A.S[(long long)s * (A.niter + 1) + A.k + 1], t);
    }
    if (A.last) return;

    LP::template fft<-1, float>(v, acc, j, tw);

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) A.W[off + e * T] = v[e];
    }
}

// ---- registry ----------------------------------------------------------------------------------------
template <typename LP, int C, int MINB>
static void launch_cols(const PocsGeom& G, const AxisDev<float>& ax, const BandArgs<float>& A, int ns, int op, cudaStream_t st) {
    constexpr size_t smem = (size_t)LP::N * C * sizeof(Cx<float>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_cols_spec<LP, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n2 + C - 1) / C, ns);
    k_cols_spec<LP, C, MINB><<<grid, LP::T * C, smem, st>>>(G, ax.tw, A, op);
}
template <typename LP, int RB, int MINB>
static void launch_rows(const PocsGeom& G, const AxisDev<float>& ax, const BandArgs<float>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)row_pitch(LP::N) * RB * sizeof(Cx<float>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_rows_spec<LP, RB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n1 + RB - 1) / RB, ns);
    k_rows_spec<LP, RB, MINB><<<grid, LP::T * RB, smem, st>>>(G, ax.tw, A);
}

typedef LinePlan<1000, 10, 10, 10, 10> LP1000;
typedef LinePlan<2000, 10, 10, 10, 10, 2> LP2000;
typedef LinePlan<256, 16, 16, 16> LP256;
typedef LinePlan<200, 20, 10, 20> LP200;

SpecKernels select_spec_kernels(int n_iline, int n_xline, int variant) {
    SpecKernels k;
    switch (n_iline) {      // column transforms have the length of the iline axis
        case 1000:
            if (variant == 1) { k.cols_iter = launch_cols<LP1000, 8, 1>; k.cols_name = "spec<1000,E10,10x10x10,C8,1cta>"; }
            else              { k.cols_iter = launch_cols<LP1000, 4, 2>; k.cols_name = "spec<1000,E10,10x10x10,C4,2cta>"; }
            break;
        case 2000: k.cols_iter = launch_cols<LP2000, 4, 1>;  k.cols_name = "spec<2000,E10,10x10x10x2,C4>"; break;
        case 256:  k.cols_iter = launch_cols<LP256, 16, 3>;  k.cols_name = "spec<256,E16,16x16,C16>"; break;
        case 200:  k.cols_iter = launch_cols<LP200, 16, 4>;  k.cols_name = "spec<200,E20,10x20,C16>"; break;
        default: break;
    }
    switch (n_xline) {      // row transforms have the length of the xline axis
        case 1000:
            if (variant == 1) { k.rows_iter = launch_rows<LP1000, 8, 1>; k.rows_name = "spec<1000,E10,10x10x10,RB8,1cta>"; }
            else              { k.rows_iter = launch_rows<LP1000, 4, 2>; k.rows_name = "spec<1000,E10,10x10x10,RB4,2cta>"; }
            break;
        case 2000: k.rows_iter = launch_rows<LP2000, 4, 1>;  k.rows_name = "spec<2000,E10,10x10x10x2,RB4>"; break;
        case 256:  k.rows_iter = launch_rows<LP256, 16, 3>;  k.rows_name = "spec<256,E16,16x16,RB16>"; break;
        case 200:  k.rows_iter = launch_rows<LP200, 16, 4>;  k.rows_name = "spec<200,E20,10x20,RB16>"; break;
        default: break;
    }
    return k;
}

}  // namespace p3d
