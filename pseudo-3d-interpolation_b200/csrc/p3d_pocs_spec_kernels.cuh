// Kernel templates, launchers and registration macros of the specialised (compile-time size, register-resident)
// POCS kernels; included by the translation units that instantiate them (p3d_pocs_spec.cu, p3d_pocs_spec_more.cu).
#pragma once
#include "p3d_pocs_spec.cuh"
#include "p3d_fft_mix.cuh"
#include "p3d_fft_reg.cuh"

#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <type_traits>

namespace p3d {

// L2 prefetch of one 32-byte sector (the data of a tile that a later CTA will load)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// CTAs resident at a time (2 per SM on 148 SMs): the tile that far ahead is the one whose loads
// should already be on their way when its CTA starts
constexpr int P3D_PREFETCH_DISTANCE = 296;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f(double v) { return warp_sum(v); }

// ---- column kernel ---------------------------------------------------------------------------------
// BULK: the tile is loaded with cp.async.cg (global -> shared memory in 16-byte pieces, no register and no L1 staging)
// into exchange buffer 1, which is free until the second exchange, in the [row][c] layout of the accessor, and read
// into registers from there.  A plain load of a 32-byte row segment holds a whole 128-byte L1 line while it is in
// flight, so the L1 capacity left beside the exchange buffers bounded the memory-level parallelism of this kernel
// (1000-point columns, 512 slices: 3181 us with 124 KB of L1, 4418 us with the carveout forced to 100 %); with the
// asynchronous copies a third CTA per SM pays off (3179 -> 2594 us).
// SB: single exchange buffer (ColAcc1)
template <typename F, typename LP, int C, int MINB, bool BULK, bool SB = false>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_cols_spec(const __grid_constant__ PocsGeom G, const Cx<F>* __restrict__ tw, const __grid_constant__ BandArgs<F> A, const int op) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = blockIdx.y;
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const int col = blockIdx.x * C + c;
    const bool ok = col < G.n2;
    Cx<F>* __restrict__ Ws = A.W + (long long)s * N * G.n2 + col;
    typename std::conditional<SB, ColAcc1<F, C, LP::LINE>, ColAcc<F, C, LP::LINE>>::type acc;
    acc.base = reinterpret_cast<Cx<F>*>(smem_raw) + c;

    Cx<F> v[E];
    constexpr int CHUNKS = (C * (int)sizeof(Cx<F>)) / 16;                 // 16-byte pieces per row segment
    // (uniform over the CTA) whole tile inside the slice, rows 16-byte aligned
    const bool bulk = BULK && (CHUNKS >= 1) && (blockIdx.x * C + C <= G.n2) && (((long long)G.n2 * sizeof(Cx<F>)) % 16 == 0) &&
                      ((reinterpret_cast<uintptr_t>(A.W) % 16) == 0);
    if (bulk) {
        const char* src0 = reinterpret_cast<const char*>(A.W + (long long)s * N * G.n2 + blockIdx.x * C);
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(reinterpret_cast<Cx<F>*>(smem_raw) + (SB ? (size_t)0 : (size_t)LP::LINE * C));
        for (int q = tid; q < N * CHUNKS; q += T * C) {
            const int row = q / CHUNKS, part = q - row * CHUNKS;
            const char* src = src0 + (long long)row * G.n2 * sizeof(Cx<F>) + part * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (unsigned)(row * C * sizeof(Cx<F>) + part * 16)), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = ok ? Ws[(long long)(j + e * T) * G.n2] : cmake<F>(F(0), F(0));
    }
    const Cx<F> tau = A.tau[(long long)s * A.niter + A.k];
    // the (rare) early-exit test comes AFTER the loads were issued, so that its own dependent
    // loads (stop flag, two sums) do not delay them
    if (slice_stopped(A.stop, A.S, s, A.k, A.niter, A.eps)) { if (bulk) asm volatile("cp.async.wait_all;"); return; }
    if (bulk) {
        asm volatile("cp.async.wait_all;");
        __syncthreads();
        const Cx<F>* land = acc.line(1);
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = land[(j + e * T) * C];
    }

    LP::template fft<-1, 0, F>(v, acc, j, tw);

    const F a = tau.x, b = tau.y;
    const F t2re = a * a - b * b, t2im = F(2) * a * b;
    if (op == P3D_OP_HARD && !A.exact_tie) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_HARD, F, false>(v[e], a, b, t2re, t2im);
    } else if (op == P3D_OP_HARD) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_HARD, F>(v[e], a, b, t2re, t2im);
    } else if (op == P3D_OP_SOFT) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_SOFT, F>(v[e], a, b, t2re, t2im);
    } else if (op == P3D_OP_GARROTE) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_GARROTE, F>(v[e], a, b, t2re, t2im);
    } else {
        // kx-ky filter mode: real filter plane, same (row, column) position as the coefficient
        const float* __restrict__ H = A.filt + col;
#pragma unroll
        for (int e = 0; e < E; ++e) { const F h = ok ? (F)__ldg(H + (long long)(j + e * T) * G.n2) : F(0); v[e] = cmake<F>(v[e].x * h, v[e].y * h); }
    }

    LP::template fft<+1, (LP::NEXCH & 1), F>(v, acc, j, tw);

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Ws[(long long)(j + e * T) * G.n2] = v[e];
    }
}

// ---- row kernel ----------------------------------------------------------------------------------
template <typename F, typename LP, int RB, int MINB, bool PF>
__global__ void __launch_bounds__(LP::T* RB, MINB)
k_rows_spec(const __grid_constant__ PocsGeom G, const Cx<F>* __restrict__ tw, const __grid_constant__ BandArgs<F> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red_s[32];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    static_assert(E <= 32, "mask bits are packed into one 32-bit word");
    const int s = blockIdx.y;
    const int stopped = A.stop[s];
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const int row = blockIdx.x * RB + rr;
    const bool ok = row < G.n1;
    const long long off = (long long)s * G.n1 * N + (long long)row * N + j;
    RowAcc<F, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<F>*>(smem_raw) + rr * LP::LINE;
    Cx<F>* __restrict__ Wp = A.W + off;
    const Cx<F>* __restrict__ Dp = A.D + off;
    Cx<F>* __restrict__ Op = A.OUT + off;

    Cx<F> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ok ? Wp[e * T] : cmake<F>(F(0), F(0));
    if (PF && ok && (j & 3) == 0) {
        // the observed data of this tile are needed after the inverse transform: start fetching now;
        // and warm L2 with the W rows of the CTA that will run P3D_PREFETCH_DISTANCE blocks later
#pragma unroll
        for (int e = 0; e < E; ++e) prefetch_l2(Dp + e * T);
        const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + P3D_PREFETCH_DISTANCE;
        const long long by = lin / gridDim.x, bx = lin - by * gridDim.x;
        if (by < gridDim.y && bx * RB + rr < G.n1) {
            const Cx<F>* nx = A.W + by * (long long)G.n1 * N + (bx * RB + rr) * (long long)N + j;
#pragma unroll
            for (int e = 0; e < E; ++e) prefetch_l2(nx + e * T);
        }
    }
    // one packed mask word per thread (bit e <-> column j + e*T) rides along with the first loads
    const long long midx = (A.first_slice + s) / G.slices_per_mask;
    const unsigned mbits = ok ? A.mbits[(midx * G.n1 + row) * T + j] : 0u;
    if (stopped != 0) return;

    LP::template fft<+1, 0, F>(v, acc, j, tw);

    F part = F(0);
    if (ok) {
        // all observed-data loads are issued before the first use (one exposed latency, not E)
        Cx<F> d[E];
#pragma unroll
        for (int e = 0; e < E; ++e) d[e] = Dp[e * T];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const F m = ((mbits >> e) & 1u) ? F(1) : F(0);
            const F coef = (F(1) - A.alpha * m) * A.inv_n;
            Cx<F> x = cmake<F>(fma(coef, v[e].x, A.alpha * d[e].x), fma(coef, v[e].y, A.alpha * d[e].y));
            part += sqrt(x.x * x.x + x.y * x.y);
            if (A.write_out) Op[e * T] = x;
            if (A.adaptive) {
                const F keep = F(1) - A.alpha * m, om = F(1) - A.alpha;
                const Cx<F> xt = cmake<F>(A.alpha * d[e].x + keep * x.x, A.alpha * d[e].y + keep * x.y);
                x = cmake<F>(xt.x + om * (d[e].x - m * x.x), xt.y + om * (d[e].y - m * x.y));
            }
            v[e] = x;
        }
    }
    // warp level in the working precision (a warp holds 32 E non-negative terms: float is exact enough for a sum
    // that is compared at 1e-5 relative), rows and slices in double
    double dp = (double)warp_sum_f(part);
    if ((tid & 31) == 0) red_s[tid >> 5] = dp;
    __syncthreads();
    if (tid < 32) {
        constexpr int NW = (T * RB + 31) / 32;
        double t = tid < NW ? red_s[tid] : 0.0;
        t = warp_sum(t);
        if (tid == 0) atomicAdd(&A.S[(long long)s * (A.niter + 1) + A.k + 1], t);
    }
    if (A.last) return;

    LP::template fft<-1, (LP::NEXCH & 1), F>(v, acc, j, tw);

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Wp[e * T] = v[e];
    }
}

// ---- once-per-slice kernels: row FFT of the observed slice, column FFT + schedule statistics -------
template <typename LP, int RB, int MINB>
__global__ void __launch_bounds__(LP::T* RB, MINB)
k_rows_init_spec(const __grid_constant__ PocsGeom G, const Cx<float>* __restrict__ tw, const __grid_constant__ BandArgs<float> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red_s[32];
    __shared__ unsigned long long red_n[32];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = blockIdx.y;
    if (A.adaptive && A.stop[s] != 0) return;
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const int row = blockIdx.x * RB + rr;
    const bool ok = row < G.n1;
    const long long off = (long long)s * G.n1 * N + (long long)row * N + j;
    RowAcc<float, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + rr * LP::LINE;
    const Cx<float>* __restrict__ Dp = A.D + off;
    Cx<float>* __restrict__ Wp = A.W + off;
    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ok ? Dp[e * T] : cmake<float>(0.f, 0.f);
    float part = 0.f;
    unsigned long long nnz = 0ull;
    if (!A.adaptive) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            nnz += (v[e].x != 0.f || v[e].y != 0.f) ? 1ull : 0ull;
            part += sqrtf(v[e].x * v[e].x + v[e].y * v[e].y);
        }
    } else {
        // APOCS prologue with x_old = x (functions/POCS.py:572-575)
        const long long midx = (A.first_slice + s) / G.slices_per_mask;
        const unsigned mbits = ok ? A.mbits[(midx * G.n1 + row) * T + j] : 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float m = ((mbits >> e) & 1u) ? 1.f : 0.f;
            const float keep = 1.f - A.alpha * m, om = 1.f - A.alpha;
            const Cx<float> d = v[e];
            const Cx<float> xt = cmake<float>(A.alpha * d.x + keep * d.x, A.alpha * d.y + keep * d.y);
            v[e] = cmake<float>(xt.x + om * (d.x - m * d.x), xt.y + om * (d.y - m * d.y));
        }
    }
    if (A.accum) {
        double dp = warp_sum((double)part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nnz += __shfl_xor_sync(0xffffffffu, nnz, o);
        if ((tid & 31) == 0) { red_s[tid >> 5] = dp; red_n[tid >> 5] = nnz; }
        __syncthreads();
        if (tid < 32) {
            constexpr int NW = (T * RB + 31) / 32;
            double t = tid < NW ? red_s[tid] : 0.0;
            unsigned long long c = tid < NW ? red_n[tid] : 0ull;
            t = warp_sum(t);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (tid == 0) { atomicAdd(&A.S[(long long)s * (A.niter + 1)], t); atomicAdd(&A.stats[s].nnz, c); }
        }
    }
    LP::template fft<-1, 0, float>(v, acc, j, tw);
    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Wp[e * T] = v[e];
    }
}

template <typename LP, int C, int MINB>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_cols_stats_spec(const __grid_constant__ PocsGeom G, const Cx<float>* __restrict__ tw, const __grid_constant__ BandArgs<float> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = blockIdx.y;
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const int col = blockIdx.x * C + c;
    const bool ok = col < G.n2;
    const Cx<float>* __restrict__ Ws = A.W + (long long)s * N * G.n2 + col;
    ColAcc<float, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + c;
    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ok ? Ws[(long long)(j + e * T) * G.n2] : cmake<float>(0.f, 0.f);
    LP::template fft<-1, 0, float>(v, acc, j, tw);
    unsigned long long kmax = 0ull; float ssf = 0.f; unsigned int amax = 0u, amin = 0xffffffffu;
    if (ok) {
        Cx<float>* __restrict__ X0 = A.OUT + (long long)s * N * G.n2 + col;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const unsigned long long key = lex_key(v[e].x, v[e].y);
            kmax = key > kmax ? key : kmax;
            const float r2 = v[e].x * v[e].x + v[e].y * v[e].y;
            ssf += r2;
            const unsigned int rb = __float_as_uint(sqrtf(r2));
            amax = rb > amax ? rb : amax; amin = rb < amin ? rb : amin;
            if (A.store_x0) X0[(long long)(j + e * T) * G.n2] = v[e];
        }
    }
    kmax = warp_max_u64(kmax); const double ss = warp_sum((double)ssf); amax = warp_max_u32(amax); amin = warp_min_u32(amin);
    if ((tid & 31) == 0) {
        atomicMax(&A.stats[s].lexmax_key, kmax);
        atomicAdd(&A.stats[s].sumsq, ss);
        atomicMax(&A.stats[s].maxabs_bits, amax);
        atomicMin(&A.stats[s].minabs_bits, amin);
    }
}

// mask bytes -> one word per (mask, row, j): bit e = mask[row][j + e*T] != 0
template <int T, int E>
__global__ void k_pack_mask(const uint8_t* __restrict__ mask, uint32_t* __restrict__ bits, long long total_rows) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total_rows * T) return;
    const long long row = i / T;
    const int j = (int)(i - row * T);
    const uint8_t* m = mask + row * (long long)(T * E) + j;
    unsigned w = 0u;
#pragma unroll
    for (int e = 0; e < E; ++e) w |= (m[e * T] != 0) ? (1u << e) : 0u;
    bits[i] = w;
}

// ---- registry ----------------------------------------------------------------------------------------
template <typename LP, int C, int MINB, bool BULK = false, typename F = float, bool SB = false>
static void launch_cols(const PocsGeom& G, const Cx<F>* tw, const BandArgs<F>& A, int ns, int op, cudaStream_t st) {
    constexpr size_t smem = (size_t)(SB ? 1 : 2) * LP::LINE * C * sizeof(Cx<F>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_cols_spec<F, LP, C, MINB, BULK, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n2 + C - 1) / C, ns);
    k_cols_spec<F, LP, C, MINB, BULK, SB><<<grid, LP::T * C, smem, st>>>(G, tw, A, op);
}
template <typename LP, int RB, int MINB, bool PF = false, typename F = float>
static void launch_rows(const PocsGeom& G, const Cx<F>* tw, const BandArgs<F>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * LP::LINE * RB * sizeof(Cx<F>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_rows_spec<F, LP, RB, MINB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n1 + RB - 1) / RB, ns);
    k_rows_spec<F, LP, RB, MINB, PF><<<grid, LP::T * RB, smem, st>>>(G, tw, A);
}
template <typename LP, int RB, int MINB>
static void launch_rows_init(const PocsGeom& G, const Cx<float>* tw, const BandArgs<float>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * LP::LINE * RB * sizeof(Cx<float>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_rows_init_spec<LP, RB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n1 + RB - 1) / RB, ns);
    k_rows_init_spec<LP, RB, MINB><<<grid, LP::T * RB, smem, st>>>(G, tw, A);
}
template <typename LP, int C, int MINB>
static void launch_cols_stats(const PocsGeom& G, const Cx<float>* tw, const BandArgs<float>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * LP::LINE * C * sizeof(Cx<float>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_cols_stats_spec<LP, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n2 + C - 1) / C, ns);
    k_cols_stats_spec<LP, C, MINB><<<grid, LP::T * C, smem, st>>>(G, tw, A);
}
template <typename LP>
static void launch_pack(const uint8_t* mask, uint32_t* bits, int n_masks, int n1, cudaStream_t st) {
    const long long rows = (long long)n_masks * n1;
    const long long total = rows * LP::T;
    k_pack_mask<LP::T, LP::E><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(mask, bits, rows);
}

template <typename LP> static std::vector<int> radices_of() {
    std::vector<int> r(LP::NPASS);
    LP::radices(r.data());
    return r;
}

#define P3D_COLS_BULK(LP, C, MINB, NAME) do { k.cols_iter = launch_cols<LP, C, MINB, true>; k.cols_stats = launch_cols_stats<LP, C, MINB>; \
                                              k.cols_name = NAME; k.cols_radices = radices_of<LP>(); } while (0)
#define P3D_COLS(LP, C, MINB, NAME) do { k.cols_iter = launch_cols<LP, C, MINB>; k.cols_stats = launch_cols_stats<LP, C, MINB>; \
                                         k.cols_name = NAME; k.cols_radices = radices_of<LP>(); } while (0)
#define P3D_ROWS(LP, RB, MINB, NAME) do { k.rows_iter = launch_rows<LP, RB, MINB>; k.rows_init = launch_rows_init<LP, RB, MINB>; \
                                          k.rows_name = NAME; k.rows_radices = radices_of<LP>(); \
                                          k.pack_mask = launch_pack<LP>; k.rows_T = LP::T; } while (0)


}  // namespace p3d
