// Register-resident line FFT for compile-time lengths (the fast path).
//
// A line of N points is owned by T = N/E threads; thread j holds the E points at positions
// j + e*T (e = 0..E-1) in registers.  That element set is the input set of EVERY Stockham
// pass and also the natural-order output set of the last one, so
//   * global loads/stores are coalesced straight from/to registers (no staging),
//   * only npass-1 shared-memory exchanges are needed per transform,
//   * an inverse transform can be followed by an element-wise step and a forward transform
//     without leaving registers (the fusion the POCS iteration kernels are built on).
// Each pass: per-thread Q = E/R radix-R butterflies (every radix must divide E).
#pragma once
#include "p3d_butterflies.cuh"

namespace p3d {

// Smem accessor concept:  Cx<T>& at(int pos)  -- maps a line position to this thread's line storage.

template <int N, int E, int DIR, int Ns, typename T, typename Acc, int... Rs> struct RegPasses;

template <int N, int E, int DIR, int Ns, typename T, typename Acc>
struct RegPasses<N, E, DIR, Ns, T, Acc> {
    static_assert(Ns == N, "radix sequence does not multiply to N");
    __device__ __forceinline__ static void run(Cx<T> (&)[E], const Acc&, const int, const Cx<T>* __restrict__) {}
};

template <int N, int E, int DIR, int Ns, typename T, typename Acc, int R, int... Rest>
struct RegPasses<N, E, DIR, Ns, T, Acc, R, Rest...> {
    static constexpr int TT = N / E;          // threads per line
    static constexpr int Q = E / R;           // butterflies per thread in this pass
    static_assert(E % R == 0, "every radix must divide the elements per thread");
    static_assert(N % (Ns * R) == 0, "radix sequence does not multiply to N");

    __device__ __forceinline__ static void run(Cx<T> (&v)[E], const Acc& acc, const int j, const Cx<T>* __restrict__ tw) {
        constexpr bool last = (sizeof...(Rest) == 0);
        int kk[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int b = j + q * TT;
            const int k = (Ns == 1) ? 0 : (b % Ns);
            kk[q] = k;
            Cx<T> x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = v[q + r * Q];
            if (Ns > 1) {
                const int t1 = k * (N / (Ns * R));
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const Cx<T> w = tw[r * t1];
                    x[r] = (DIR < 0) ? cmul(x[r], w) : cmulc(x[r], w);
                }
            }
            Bfly<R, DIR, T>::run(x);
#pragma unroll
            for (int r = 0; r < R; ++r) v[q + r * Q] = x[r];
        }
        if constexpr (!last) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int b = j + q * TT;
                const int j0 = (b - kk[q]) * R + kk[q];
#pragma unroll
                for (int r = 0; r < R; ++r) acc.at(j0 + r * Ns) = v[q + r * Q];
            }
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = acc.at(j + e * TT);
            __syncthreads();
            RegPasses<N, E, DIR, Ns * R, T, Acc, Rest...>::run(v, acc, j, tw);
        }
    }
};

// Radix list as a type so kernels can be templated on one "line plan".
template <int N_, int E_, int... Rs> struct LinePlan {
    static constexpr int N = N_;
    static constexpr int E = E_;
    static constexpr int T = N_ / E_;
    template <int DIR, typename TT, typename Acc>
    __device__ __forceinline__ static void fft(Cx<TT> (&v)[E_], const Acc& acc, const int j, const Cx<TT>* __restrict__ tw) {
        RegPasses<N_, E_, DIR, 1, TT, Acc, Rs...>::run(v, acc, j, tw);
    }
};

}  // namespace p3d
