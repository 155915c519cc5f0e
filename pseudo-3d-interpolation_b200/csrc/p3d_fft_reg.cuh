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
//
// Exchanges alternate between two shared-memory buffers, so each exchange needs a single
// __syncthreads(): the writes of exchange i+1 go to the buffer nobody reads any more, and the
// writes of exchange i+2 are ordered after the reads of exchange i by the barrier of i+1.
//
// Exchange layout.  After the pass with (Ns, R) the Stockham output position is
//   pos = hi*(Ns*R) + r*Ns + k        (butterfly b = hi*Ns + k, output r)
// and it is stored at  pos + hi*DELTA : every block of B = Ns*R positions is followed by DELTA
// padding elements.  DELTA is chosen per exchange so that the 16 lanes of a half-warp (8-byte
// accesses) fall into distinct banks for the writes (stride B+DELTA odd for Ns = 1; block
// starts continuing the run of k for Ns > 1), while the reads stay (almost) consecutive.
// Because T is a multiple of B or B a multiple of T, both the write and the read address
// split into a per-thread base plus a compile-time offset: no per-access index arithmetic.
#pragma once
#include "p3d_butterflies.cuh"

namespace p3d {

__host__ __device__ constexpr int exch_delta(int Ns, int R) {
    const int B = Ns * R;
    if (Ns == 1) return (R % 2 == 0) ? 1 : 0;
    int d = (Ns - B) % 16;
    if (d < 0) d += 16;
    return d;
}
// storage needed per line (elements): the largest padded address over all exchanges, + 1
template <int... Rs> struct PaddedLine {
    static constexpr int value() {
        const int r[] = {Rs...};
        const int np = (int)sizeof...(Rs);
        int n = 1;
        for (int i = 0; i < np; ++i) n *= r[i];
        int best = n, ns = 1;
        for (int i = 0; i + 1 < np; ++i) {
            const int b = ns * r[i];
            const int need = n + (n / b - 1) * exch_delta(ns, r[i]);
            best = need > best ? need : best;
            ns = b;
        }
        return (best + 3) & ~3;        // keep rows of the column tile 32-byte aligned
    }
};

// Accessor concept:
//   static constexpr int STRIDE;            element stride of consecutive line positions
//   Cx<T>* line(int buf) const;             this thread's line in buffer 0/1

// Twiddles: one table per twiddled pass with entries W^{r k} (k = b mod Ns < Ns, W = exp(-2 pi i / (Ns R)); r = 0 is 1
// and unused), laid out with k FASTEST: even R as pairs [r/2][k][2] (one 16-byte load fetches W^{2 r2 k}, W^{(2 r2 + 1) k}),
// odd R as [r][k].  A butterfly still reads its factors from ONE computed address (+ compile-time offsets r2 * Ns), and
// the lanes of a warp, whose k are consecutive, now read consecutive 16-byte words: a [k][R] layout (80-byte stride
// between lanes for R = 10) touched five times as many L1 lines per load.  Tables of consecutive passes are
// concatenated (offset TWOFF, each table starting at an even element = 16-byte aligned).  Host side: spec_twiddle_table().
__host__ __device__ constexpr int tw_index(int R, int Ns, int r, int k) {
    return (R % 2 == 0) ? (((r >> 1) * Ns + k) * 2 + (r & 1)) : (r * Ns + k);
}
// complex128 twiddles: W^{r k}, r = 2 .. R-1, from W^k alone (squares and products, depth <= 4).  The L1 / shared-memory
// data pipe is the busiest unit of the complex128 kernels (ncu: 80 % in the row kernel) and R - 1 16-byte table loads per
// butterfly were a quarter of its wavefronts; the FP64 pipe has the headroom for the 3 - 4 extra instructions per factor.
template <int R, typename T>
__device__ __forceinline__ void twiddle_powers(const Cx<T> w1, Cx<T> (&w)[R]) {
    w[0] = cmake<T>(T(1), T(0));
    w[1] = w1;
#pragma unroll
    for (int r = 2; r < R; ++r) {
        if (r % 2 == 0) { const Cx<T> h = w[r / 2]; w[r] = cmake<T>(h.x * h.x - h.y * h.y, T(2) * h.x * h.y); }
        else w[r] = cmul(w[(r + 1) / 2], w[r / 2]);
    }
}

template <int N, int E, int DIR, int Ns, int BUF, int TWOFF, typename T, typename Acc, int... Rs> struct RegPasses;

template <int N, int E, int DIR, int Ns, int BUF, int TWOFF, typename T, typename Acc>
struct RegPasses<N, E, DIR, Ns, BUF, TWOFF, T, Acc> {
    static_assert(Ns == N, "radix sequence does not multiply to N");
    __device__ __forceinline__ static void run(Cx<T> (&)[E], const Acc&, const int, const Cx<T>* __restrict__) {}
};

template <int N, int E, int DIR, int Ns, int BUF, int TWOFF, typename T, typename Acc, int R, int... Rest>
struct RegPasses<N, E, DIR, Ns, BUF, TWOFF, T, Acc, R, Rest...> {
    static constexpr int TT = N / E;          // threads per line
    static constexpr int Q = E / R;           // butterflies per thread in this pass
    static constexpr int B = Ns * R;          // Stockham block of this pass
    // First exchange of a ROW line (unit stride) with a block of B = R positions per thread, B % 4 == 2 (radix 10):
    // no padding and 16-byte stores.  Eight lanes of a 128-bit store then start 2B words apart = 4, 12, 20, 28
    // (mod 32) -> conflict-free, and the reads are consecutive 8-byte words -> conflict-free as well; the padded
    // layout (DELTA = 1, 8-byte stores) costs the reads a 2-way conflict per half-warp (ncu: 832 k instead of
    // 448 k wavefronts per load instruction).
    static constexpr bool VEC1 = (Ns == 1) && (Acc::STRIDE == 1) && (R % 4 == 2) && (sizeof(T) == 4) && (sizeof...(Rest) > 0);
    static constexpr int DELTA = VEC1 ? 0 : exch_delta(Ns, R);
    static_assert(E % R == 0, "every radix must divide the elements per thread");
    static_assert(N % (Ns * R) == 0, "radix sequence does not multiply to N");

    __device__ __forceinline__ static void run(Cx<T> (&v)[E], const Acc& acc, const int j, const Cx<T>* __restrict__ tw) {
        constexpr bool last = (sizeof...(Rest) == 0);
        constexpr int S = Acc::STRIDE;
        int wbase[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int b = j + q * TT;
            const int hi = (Ns == 1) ? b : (b / Ns);
            const int k = (Ns == 1) ? 0 : (b - hi * Ns);
            wbase[q] = hi * (B + DELTA) + k;
            Cx<T> x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = v[q + r * Q];
            if (Ns > 1) {
                if constexpr (R % 2 == 0 && sizeof(T) == 4) {
                    const float4* t4 = reinterpret_cast<const float4*>(tw + TWOFF) + k;
#pragma unroll
                    for (int r2 = 0; r2 < R / 2; ++r2) {
                        const float4 w = __ldg(t4 + r2 * Ns);
                        if (r2 > 0) x[2 * r2] = (DIR < 0) ? cmul(x[2 * r2], cmake<T>(w.x, w.y)) : cmulc(x[2 * r2], cmake<T>(w.x, w.y));
                        x[2 * r2 + 1] = (DIR < 0) ? cmul(x[2 * r2 + 1], cmake<T>(w.z, w.w)) : cmulc(x[2 * r2 + 1], cmake<T>(w.z, w.w));
                    }
                } else if constexpr (sizeof(T) == 8 && (R > 3)) {
                    const Cx<T>* tp = tw + TWOFF;
                    Cx<double> w[R];
                    twiddle_powers<R, double>(tp[tw_index(R, Ns, 1, 0) + k * (R % 2 == 0 ? 2 : 1)], w);
#pragma unroll
                    for (int r = 1; r < R; ++r) x[r] = (DIR < 0) ? cmul(x[r], w[r]) : cmulc(x[r], w[r]);
                } else {
                    const Cx<T>* tp = tw + TWOFF;
#pragma unroll
                    for (int r = 1; r < R; ++r) {
                        const Cx<T> w = tp[tw_index(R, Ns, r, 0) + k * (R % 2 == 0 ? 2 : 1)];
                        x[r] = (DIR < 0) ? cmul(x[r], w) : cmulc(x[r], w);
                    }
                }
            }
            Bfly<R, DIR, T>::run(x);
#pragma unroll
            for (int r = 0; r < R; ++r) v[q + r * Q] = x[r];
        }
        if constexpr (!last) {
            static_assert(TT % B == 0 || B % TT == 0, "threads per line and Stockham block must nest");
            Cx<T>* line = acc.line(BUF);
            acc.pre_sync();          // single-buffered accessors: the previous exchange's reads must be over
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                Cx<T>* w = line + wbase[q] * S;
                if constexpr (VEC1) {
                    float4* w4 = reinterpret_cast<float4*>(w);
#pragma unroll
                    for (int r = 0; r < R; r += 2) {
                        const Cx<T> a = v[q + r * Q], b = v[q + (r + 1) * Q];
                        w4[r / 2] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) w[(r * Ns) * S] = v[q + r * Q];
                }
            }
            acc.sync();
            if constexpr (TT % B == 0) {
                // hi(pos) = j/B + e*(TT/B)
                const Cx<T>* rd = line + (j + DELTA * (j / B)) * S;
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = rd[(e * (TT + DELTA * (TT / B))) * S];
            } else {
                // hi(pos) = e / (B/TT)
                const Cx<T>* rd = line + j * S;
#pragma unroll
                for (int e = 0; e < E; ++e) v[e] = rd[(e * TT + DELTA * (e / (B / TT))) * S];
            }
            RegPasses<N, E, DIR, Ns * R, BUF ^ 1, TWOFF + (Ns > 1 ? ((Ns * R + 1) & ~1) : 0), T, Acc, Rest...>::run(v, acc, j, tw);
        }
    }
};

// Radix list as a type so kernels can be templated on one "line plan".
template <int N_, int E_, int... Rs> struct LinePlan {
    static constexpr int N = N_;
    static constexpr int E = E_;
    static constexpr int T = N_ / E_;
    static constexpr int NEXCH = (int)sizeof...(Rs) - 1;      // shared-memory exchanges per transform
    static constexpr int LINE = PaddedLine<Rs...>::value();   // storage per line and buffer (elements)
    static constexpr int NPASS = (int)sizeof...(Rs);
    static void radices(int* out) { const int r[] = {Rs...}; for (int i = 0; i < NPASS; ++i) out[i] = r[i]; }
    // BUF0: buffer used by the first exchange; the next transform should start with (BUF0 + NEXCH) & 1
    template <int DIR, int BUF0, typename TT, typename Acc>
    __device__ __forceinline__ static void fft(Cx<TT> (&v)[E_], const Acc& acc, const int j, const Cx<TT>* __restrict__ tw) {
        RegPasses<N_, E_, DIR, 1, BUF0, 0, TT, Acc, Rs...>::run(v, acc, j, tw);
    }
};

// ---- shared-memory accessors ---------------------------------------------------------------------
// Two buffers (see p3d_fft_reg.cuh).  Column tile: [addr][c], c fastest; row tile: [row][addr].
template <typename T, int C, int LINE> struct ColAcc {
    static constexpr int STRIDE = C;
    Cx<T>* base;   // already offset by c
    __device__ __forceinline__ Cx<T>* line(int buf) const { return base + buf * (LINE * C); }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void pre_sync() const {}
};
// one buffer instead of two (half the shared memory, one more barrier per exchange)
template <typename T, int C, int LINE> struct ColAcc1 {
    static constexpr int STRIDE = C;
    Cx<T>* base;
    __device__ __forceinline__ Cx<T>* line(int) const { return base; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void pre_sync() const { __syncthreads(); }
};
template <typename T, int RB, int LINE> struct RowAcc {
    static constexpr int STRIDE = 1;
    Cx<T>* base;   // already offset by the row
    __device__ __forceinline__ Cx<T>* line(int buf) const { return base + buf * (RB * LINE); }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void pre_sync() const {}
};
}  // namespace p3d
