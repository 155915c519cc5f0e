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

// |x| for the convergence sum.  double: an fp32 reciprocal-square-root seed and one Newton step in double (1e-14
// relative) instead of the library square root (a dozen FP64 instructions and a slow-path call per element); values
// outside the float range take the library path.
__device__ __forceinline__ float norm_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double norm_sqrt(double v) {
    if (!(v > 1e-30 && v < 1e30)) return sqrt(v);
    const double r = (double)rsqrtf((float)v);
    const double y = v * r;
    return fma(fma(-y, y, v), 0.5 * r, y);
}

// plans whose transforms may run on a single exchange buffer (ColAcc1): every LinePlan; a MixPlan3 only when its middle
// pass is one round (operands in registers before the first write)
template <typename LP> struct single_buffer_ok { static constexpr bool value = true; };
template <int N, int Ra, int Rb> struct single_buffer_ok<MixPlan3<N, Ra, Rb>> { static constexpr bool value = MixPlan3<N, Ra, Rb>::SINGLE_BUFFER_OK; };

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers -----------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");       // visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// wait for phase `parity`; traps instead of hanging if the transfer never completes (bad descriptor)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

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
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    static_assert(!SB || single_buffer_ok<LP>::value, "this plan needs two exchange buffers");
    const int s = band_slice(A);
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
    // TMA: one thread fetches the tile as boxes of (C columns x tma_rows rows) of the 3-D tensor (column, row, slice);
    // the transfer never touches the LSU pipe or a register, completion is an mbarrier transaction count
    __shared__ __align__(8) unsigned long long tma_bar;
    const bool tma = bulk && A.use_tma;
    if (tma) {
        if (tid == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&tma_bar, (unsigned)(N * C * sizeof(Cx<F>)));
            Cx<F>* dst = reinterpret_cast<Cx<F>*>(smem_raw) + (SB ? (size_t)0 : (size_t)LP::LINE * C);
            constexpr int PER = (int)(sizeof(Cx<F>) / 8);                   // tensor elements (8 bytes) per complex value
            for (int r0 = 0; r0 < N; r0 += A.tma_rows)
                tma_load_3d(dst + (size_t)r0 * C, &A.tmapW, (int)(blockIdx.x * C * PER), r0, A.tma_slice0 + s, &tma_bar);
        }
    } else if (bulk) {
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
    if ((!A.restart && slice_stopped(A.stop, A.S, s, A.k, A.niter, A.eps)) || slice_escalated(A, s)) {
        if (tma) mbar_wait(&tma_bar, 0); else if (bulk) asm volatile("cp.async.wait_all;");
        return;
    }
    if (tma) {
        mbar_wait(&tma_bar, 0);
    } else if (bulk) {
        asm volatile("cp.async.wait_all;");
        __syncthreads();
    }
    if (bulk) {
        const Cx<F>* land = acc.line(1);
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = land[(j + e * T) * C];
    }

    if (sizeof(F) == 8 && op == P3D_OP_RESTART) {
        // exact restart of the escalating mode: the tile holds a thresholded spectrum, inverse transform only
        LP::template fft<+1, 0, F>(v, acc, j, tw);
        if (ok) {
#pragma unroll
            for (int e = 0; e < E; ++e) Ws[(long long)(j + e * T) * G.n2] = v[e];
        }
        return;
    }

    LP::template fft<-1, 0, F>(v, acc, j, tw);

    const F a = tau.x, b = tau.y;
    const F t2re = a * a - b * b, t2im = F(2) * a * b;
    unsigned nearbits = 0u;
    if (sizeof(F) == 4 && A.guard && op != P3D_OP_FILTER) {
        GuardBand<F> gb(A, s, a, b, op);
#pragma unroll
        for (int e = 0; e < E; ++e) gb.test(v[e], e);
        if (!ok) { gb.hit = false; gb.bits = 0u; }
        if (A.watch) nearbits = gb.bits; else gb.commit(A, s);
    }
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
    if (sizeof(F) == 4 && A.arena && op != P3D_OP_FILTER) {
        // support record of the fp32 pilot (exact restart): which coefficients survived the threshold
        __shared__ int rec_sh[34];
        unsigned idx[E]; unsigned kept = 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            idx[e] = ((unsigned)(j + e * T) << 16) | (unsigned)col;
            if (ok && (v[e].x != F(0) || v[e].y != F(0))) kept |= 1u << e;
        }
        record_support<E>(reinterpret_cast<const BandArgs<float>&>(A), s, idx, kept, rec_sh, nearbits & ~kept);
    }

    LP::template fft<+1, (LP::NEXCH & 1), F>(v, acc, j, tw);

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Ws[(long long)(j + e * T) * G.n2] = v[e];
    }
}

// ---- row kernel ----------------------------------------------------------------------------------
// observed data / result pointers of the row kernels: the state's own type, or complex64 beside a complex128 state (IO32)
template <bool IO32, typename F> struct IoSel {
    typedef F T;
    __device__ __forceinline__ static const Cx<F>* d(const BandArgs<F>& A) { return A.D; }
    __device__ __forceinline__ static Cx<F>* o(const BandArgs<F>& A) { return A.OUT; }
};
template <typename F> struct IoSel<true, F> {
    typedef float T;
    __device__ __forceinline__ static const Cx<float>* d(const BandArgs<F>& A) { return A.D32; }
    __device__ __forceinline__ static Cx<float>* o(const BandArgs<F>& A) { return A.OUT32; }
};

template <typename F, typename LP, int RB, int MINB, bool PF, bool IO32 = false>
__global__ void __launch_bounds__(LP::T* RB, MINB)
k_rows_spec(const __grid_constant__ PocsGeom G, const Cx<F>* __restrict__ tw, const __grid_constant__ BandArgs<F> A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red_s[32];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    static_assert(E <= 32, "mask bits are packed into one 32-bit word");
    typedef typename IoSel<IO32, F>::T FD;
    const int s = band_slice(A);
    const int stopped = A.stop[s];
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const int row = blockIdx.x * RB + rr;
    const bool ok = row < G.n1;
    const long long off = (long long)s * G.n1 * N + (long long)row * N + j;
    RowAcc<F, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<F>*>(smem_raw) + rr * LP::LINE;
    Cx<F>* __restrict__ Wp = A.W + off;
    const Cx<FD>* __restrict__ Dp = IoSel<IO32, F>::d(A) + off;
    Cx<FD>* __restrict__ Op = IoSel<IO32, F>::o(A) + off;
    // complex128 restart launch: the iteration rebuilt is the one before the slice's switch
    const int kk = A.restart ? A.esc[s] - 2 : A.k;

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
        if (!A.list && by < gridDim.y && bx * RB + rr < G.n1) {
            const Cx<F>* nx = A.W + by * (long long)G.n1 * N + (bx * RB + rr) * (long long)N + j;
#pragma unroll
            for (int e = 0; e < E; ++e) prefetch_l2(nx + e * T);
        }
    }
    // one packed mask word per thread (bit e <-> column j + e*T) rides along with the first loads
    const long long midx = (A.first_slice + s) / G.slices_per_mask;
    const unsigned mbits = ok ? A.mbits[(midx * G.n1 + row) * T + j] : 0u;
    if (stopped != 0 || slice_frozen(A, s)) return;
    if (sizeof(F) == 4 && A.astart && blockIdx.x == 0 && tid == 0) close_support_record(A, s);

    LP::template fft<+1, 0, F>(v, acc, j, tw);

    F part = F(0);
    if (ok) {
        // all observed-data loads are issued before the first use (one exposed latency, not E)
        Cx<F> d[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { const Cx<FD> t = Dp[e * T]; d[e] = cmake<F>((F)t.x, (F)t.y); }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const F m = ((mbits >> e) & 1u) ? F(1) : F(0);
            const F coef = (F(1) - A.alpha * m) * A.inv_n;
            Cx<F> x = cmake<F>(fma(coef, v[e].x, A.alpha * d[e].x), fma(coef, v[e].y, A.alpha * d[e].y));
            part += norm_sqrt(x.x * x.x + x.y * x.y);
            if (A.write_out) Op[e * T] = cmake<FD>((FD)x.x, (FD)x.y);
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
        if (tid == 0) atomicAdd(&A.S[(long long)s * (A.niter + 1) + kk + 1], t);
    }
    if (A.last) return;

    LP::template fft<-1, (LP::NEXCH & 1), F>(v, acc, j, tw);

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Wp[e * T] = v[e];
    }
}

// ---- once-per-slice kernels: row FFT of the observed slice, column FFT + schedule statistics -------
// F = float: the observed slice A.D.  F = double (IO32): the complex64 observed slice A.D32.
template <typename F, typename LP, int RB, int MINB>
__global__ void __launch_bounds__(LP::T* RB, MINB)
k_rows_init_spec(const __grid_constant__ PocsGeom G, const Cx<F>* __restrict__ tw, const __grid_constant__ BandArgs<F> A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red_s[32];
    __shared__ unsigned long long red_n[32];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    constexpr bool IO32 = sizeof(F) == 8;
    const int s = band_slice(A);
    if (A.adaptive && A.stop[s] != 0) return;
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const int row = blockIdx.x * RB + rr;
    const bool ok = row < G.n1;
    const long long off = (long long)s * G.n1 * N + (long long)row * N + j;
    RowAcc<F, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<F>*>(smem_raw) + rr * LP::LINE;
    const Cx<float>* __restrict__ Dp = (IO32 ? A.D32 : reinterpret_cast<const Cx<float>*>(A.D)) + off;
    Cx<F>* __restrict__ Wp = A.W + off;
    Cx<F> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { const Cx<float> t = ok ? Dp[e * T] : cmake<float>(0.f, 0.f); v[e] = cmake<F>((F)t.x, (F)t.y); }
    F part = F(0);
    unsigned long long nnz = 0ull;
    if (!A.adaptive) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            nnz += (v[e].x != F(0) || v[e].y != F(0)) ? 1ull : 0ull;
            part += sqrt(v[e].x * v[e].x + v[e].y * v[e].y);
        }
    } else {
        // APOCS prologue with x_old = x (functions/POCS.py:572-575)
        const long long midx = (A.first_slice + s) / G.slices_per_mask;
        const unsigned mbits = ok ? A.mbits[(midx * G.n1 + row) * T + j] : 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const F m = ((mbits >> e) & 1u) ? F(1) : F(0);
            const F keep = F(1) - A.alpha * m, om = F(1) - A.alpha;
            const Cx<F> d = v[e];
            const Cx<F> xt = cmake<F>(A.alpha * d.x + keep * d.x, A.alpha * d.y + keep * d.y);
            v[e] = cmake<F>(xt.x + om * (d.x - m * d.x), xt.y + om * (d.y - m * d.y));
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
    LP::template fft<-1, 0, F>(v, acc, j, tw);
    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Wp[e * T] = v[e];
    }
}

// complex128 statistics of X0 (escalating mode): exact lexicographic maximum per CTA (A.cand), sum |X0|^2, max / min |X0|
template <typename LP, int C, int MINB>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_cols_stats_spec64(const __grid_constant__ PocsGeom G, const Cx<double>* __restrict__ tw, const __grid_constant__ BandArgs<double> A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = band_slice(A);
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const int col = blockIdx.x * C + c;
    const bool ok = col < G.n2;
    const Cx<double>* __restrict__ Ws = A.W + (long long)s * N * G.n2 + col;
    ColAcc<double, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<double>*>(smem_raw) + c;
    Cx<double> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = ok ? Ws[(long long)(j + e * T) * G.n2] : cmake<double>(0.0, 0.0);
    LP::template fft<-1, 0, double>(v, acc, j, tw);
    if (A.store_x0_inplace && ok) {
        Cx<double>* __restrict__ Wo = A.W + (long long)s * N * G.n2 + col;
#pragma unroll
        for (int e = 0; e < E; ++e) Wo[(long long)(j + e * T) * G.n2] = v[e];
    }
    double bre = -INFINITY, bim = -INFINITY, ss = 0.0;
    unsigned long long ak = 0ull, ik = ~0ull;
    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (v[e].x > bre || (v[e].x == bre && v[e].y > bim)) { bre = v[e].x; bim = v[e].y; }
            const double r2 = v[e].x * v[e].x + v[e].y * v[e].y;
            ss += r2;
            const unsigned long long k2 = f64_ordered(sqrt(r2));
            ak = k2 > ak ? k2 : ak; ik = k2 < ik ? k2 : ik;
        }
    }
    ss = warp_sum(ss); ak = warp_max_u64(ak);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long w2 = __shfl_xor_sync(0xffffffffu, ik, o); ik = w2 < ik ? w2 : ik; }
    if ((tid & 31) == 0) {
        atomicAdd(&A.stats[s].sumsq, ss);
        atomicMax(&A.stats[s].maxabs64_key, ak);
        atomicMin(&A.stats[s].minabs64_key, ik);
    }
    __syncthreads();          // the exchange buffers are free now: reuse them for the block reduction
    block_lexmax(bre, bim, reinterpret_cast<double*>(smem_raw));
    if (tid == 0) A.cand[(long long)s * A.cand_stride + blockIdx.x] = make_double2(bre, bim);
}

template <typename LP, int C, int MINB>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_cols_stats_spec(const __grid_constant__ PocsGeom G, const Cx<float>* __restrict__ tw, const __grid_constant__ BandArgs<float> A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int s = band_slice(A);
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
// rows per TMA box: the largest divisor of N that is <= 256 and makes the box a multiple of 128 bytes (0 = none)
constexpr int tma_box_rows(int N, int row_bytes) {
    for (int r = 256; r >= 1; --r)
        if (N % r == 0 && (r * row_bytes) % 128 == 0) return r;
    return 0;
}
template <typename LP, int C, int MINB, bool BULK = false, typename F = float, bool SB = false>
static void launch_cols(const PocsGeom& G, const Cx<F>* tw, const BandArgs<F>& A, int ns, int op, cudaStream_t st) {
    constexpr size_t smem = (size_t)(SB ? 1 : 2) * LP::LINE * C * sizeof(Cx<F>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_cols_spec<F, LP, C, MINB, BULK, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n2 + C - 1) / C, ns);
    // TMA tile load when the caller offers the lane's work array and the tile shape qualifies: boxes and the landing
    // buffer 128-byte aligned, row segments of at least 16 bytes
    constexpr int ROWS = tma_box_rows(LP::N, C * (int)sizeof(Cx<F>));
    constexpr bool TMA_SHAPE = BULK && ROWS > 0 && (C * sizeof(Cx<F>)) % 16 == 0 && ((size_t)LP::LINE * C * sizeof(Cx<F>)) % 128 == 0;
    if (TMA_SHAPE && A.tma_base) {
        // one descriptor per (work array, capacity) and calling thread: lanes launch from their own threads
        thread_local const void* cached_base = nullptr;
        thread_local long long cached_slices = 0;
        thread_local int cached_n2 = 0;
        thread_local bool cached_ok = false;
        thread_local BandArgs<F> B;
        if (cached_base != A.tma_base || cached_slices != A.tma_slices || cached_n2 != G.n2) {
            cached_ok = tma_encode_tile_map(&B.tmapW, A.tma_base, A.tma_slices, LP::N, G.n2, (int)sizeof(Cx<F>), C, ROWS);
            cached_base = A.tma_base; cached_slices = A.tma_slices; cached_n2 = G.n2;
        }
        if (cached_ok) {
            const CUtensorMap keep = B.tmapW;
            B = A;
            B.tmapW = keep; B.use_tma = 1; B.tma_rows = ROWS;
            k_cols_spec<F, LP, C, MINB, BULK, SB><<<grid, LP::T * C, smem, st>>>(G, tw, B, op);
            return;
        }
    }
    k_cols_spec<F, LP, C, MINB, BULK, SB><<<grid, LP::T * C, smem, st>>>(G, tw, A, op);
}
template <typename LP, int RB, int MINB, bool PF = false, typename F = float, bool IO32 = false>
static void launch_rows(const PocsGeom& G, const Cx<F>* tw, const BandArgs<F>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * LP::LINE * RB * sizeof(Cx<F>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_rows_spec<F, LP, RB, MINB, PF, IO32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n1 + RB - 1) / RB, ns);
    k_rows_spec<F, LP, RB, MINB, PF, IO32><<<grid, LP::T * RB, smem, st>>>(G, tw, A);
}
template <typename LP, int RB, int MINB, typename F = float>
static void launch_rows_init(const PocsGeom& G, const Cx<F>* tw, const BandArgs<F>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * LP::LINE * RB * sizeof(Cx<F>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_rows_init_spec<F, LP, RB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n1 + RB - 1) / RB, ns);
    k_rows_init_spec<F, LP, RB, MINB><<<grid, LP::T * RB, smem, st>>>(G, tw, A);
}
template <typename LP, int C, int MINB>
static void launch_cols_stats64(const PocsGeom& G, const Cx<double>* tw, const BandArgs<double>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * LP::LINE * C * sizeof(Cx<double>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_cols_stats_spec64<LP, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n2 + C - 1) / C, ns);
    k_cols_stats_spec64<LP, C, MINB><<<grid, LP::T * C, smem, st>>>(G, tw, A);
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
