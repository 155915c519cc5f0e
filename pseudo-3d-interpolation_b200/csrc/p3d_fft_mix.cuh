// Register-resident line FFT for lengths N = Ra * Rb * Ra whose middle radix does NOT divide the
// elements per thread (847 = 11 * 7 * 11, 1200 = 10 * 12 * 10, ...): the mixed-radix companion of
// p3d_fft_reg.cuh with the same interface (E, T, N, LINE, NEXCH, NPASS, radices(), fft<>()), so the
// POCS iteration kernels can be instantiated with either plan type.
//
// T = Ra * Rb threads own a line; thread j holds the E = Ra points j + e*T.  That set is the input
// set of the first Stockham pass (radix Ra, butterfly b = j) and the natural-order output set of the
// last one (radix Ra, Ns = T), so global loads / stores and the element-wise step between an inverse
// and a forward transform work on registers exactly as in p3d_fft_reg.cuh.  The middle pass (radix
// Rb, Ns = Ra) has Ra*Ra butterflies spread over the T threads (ceil(Ra / Rb) rounds, guarded); its
// operands travel through the two shared-memory exchanges every three-pass transform needs anyway.
//
// Exchange layout: as in p3d_fft_reg.cuh, position pos written by the pass with Stockham block
// B = Ns*R is stored at pos + (pos / B) * DELTA (bank spreading); all addresses split into a
// per-thread base plus compile-time offsets.
// Twiddles: the tables of spec_twiddle_table({Ra, Rb, Ra}) (pass 2 at offset 0, pass 3 at Ra*Rb rounded up to even; layout: tw_index).
#pragma once
#include "p3d_fft_reg.cuh"

namespace p3d {

// factors W^{r k}, r = 1 .. R-1, of the pass (Ns, R) for this thread's k; table layout: tw_index (p3d_fft_reg.cuh)
template <int DIR, int R, int Ns, typename T>
__device__ __forceinline__ void mix_twiddle(Cx<T> (&x)[R], const Cx<T>* __restrict__ tab, const int k) {
    if constexpr (R % 2 == 0 && sizeof(T) == 4) {
        const float4* t4 = reinterpret_cast<const float4*>(tab) + k;
#pragma unroll
        for (int r2 = 0; r2 < R / 2; ++r2) {
            const float4 w = __ldg(t4 + r2 * Ns);
            if (r2 > 0) x[2 * r2] = (DIR < 0) ? cmul(x[2 * r2], cmake<T>(w.x, w.y)) : cmulc(x[2 * r2], cmake<T>(w.x, w.y));
            x[2 * r2 + 1] = (DIR < 0) ? cmul(x[2 * r2 + 1], cmake<T>(w.z, w.w)) : cmulc(x[2 * r2 + 1], cmake<T>(w.z, w.w));
        }
    } else if constexpr (sizeof(T) == 8 && (R > 3)) {
        // complex128: powers of W^k instead of R - 1 table loads (see twiddle_powers)
        Cx<double> w[R];
        twiddle_powers<R, double>(tab[tw_index(R, Ns, 1, 0) + k * (R % 2 == 0 ? 2 : 1)], w);
#pragma unroll
        for (int r = 1; r < R; ++r) x[r] = (DIR < 0) ? cmul(x[r], w[r]) : cmulc(x[r], w[r]);
    } else {
#pragma unroll
        for (int r = 1; r < R; ++r) {
            const Cx<T> w = tab[tw_index(R, Ns, r, 0) + k * (R % 2 == 0 ? 2 : 1)];
            x[r] = (DIR < 0) ? cmul(x[r], w) : cmulc(x[r], w);
        }
    }
}

template <int N_, int Ra, int Rb> struct MixPlan3 {
    static_assert(N_ == Ra * Rb * Ra, "N must be Ra * Rb * Ra");
    static constexpr int N = N_;
    static constexpr int E = Ra;
    static constexpr int T = Ra * Rb;
    static constexpr int NEXCH = 2;
    static constexpr int NPASS = 3;
    static constexpr int D1 = exch_delta(1, Ra);            // padding after every block of Ra positions (exchange 1)
    static constexpr int D2 = exch_delta(Ra, Rb);           // padding after every block of Ra*Rb positions (exchange 2)
    static constexpr int NB2 = Ra * Ra;                     // butterflies of the middle pass
    static constexpr int Q2 = (NB2 + T - 1) / T;            // rounds of the middle pass per thread
    static constexpr bool SINGLE_BUFFER_OK = (Q2 == 1);     // see pass 2
    static constexpr int LINE_RAW = (N + (N / Ra - 1) * D1) > (N + (N / T - 1) * D2) ? (N + (N / Ra - 1) * D1) : (N + (N / T - 1) * D2);
    static constexpr int LINE = (LINE_RAW + 1 + 3) & ~3;
    static void radices(int* out) { out[0] = Ra; out[1] = Rb; out[2] = Ra; }

    template <int DIR, int BUF0, typename F, typename Acc>
    __device__ __forceinline__ static void fft(Cx<F> (&v)[Ra], const Acc& acc, const int j, const Cx<F>* __restrict__ tw) {
        constexpr int S = Acc::STRIDE;
        // ---- pass 1: radix Ra, Ns = 1, butterfly b = j on the register set
        Bfly<Ra, DIR, F>::run(v);
        {
            acc.pre_sync();
            Cx<F>* w = acc.line(BUF0) + (j * (Ra + D1)) * S;
#pragma unroll
            for (int r = 0; r < Ra; ++r) w[r * S] = v[r];
        }
        acc.sync();
        // ---- pass 2: radix Rb, Ns = Ra; inputs b + r*Ra*Ra, outputs hi*(Ra*Rb) + r*Ra + k
        if constexpr (Q2 == 1) {
            // one round: operands in registers before anything is written, so a single exchange buffer works as well
            // (acc.pre_sync() is the barrier between the reads and the writes there, a no-op with two buffers)
            const Cx<F>* in = acc.line(BUF0);
            Cx<F>* out = acc.line(BUF0 ^ 1);
            const int b = j;
            const bool act = (T <= NB2) || (b < NB2);
            const int hi = b / Ra, k = b - hi * Ra;
            Cx<F> x[Rb];
            if (act) {
                const Cx<F>* rd = in + (b + hi * D1) * S;
#pragma unroll
                for (int r = 0; r < Rb; ++r) x[r] = rd[(r * (Ra * Ra + Ra * D1)) * S];
                mix_twiddle<DIR, Rb, Ra, F>(x, tw, k);
                Bfly<Rb, DIR, F>::run(x);
            }
            acc.pre_sync();
            if (act) {
                Cx<F>* wr = out + (hi * (T + D2) + k) * S;
#pragma unroll
                for (int r = 0; r < Rb; ++r) wr[(r * Ra) * S] = x[r];
            }
        } else {
            const Cx<F>* in = acc.line(BUF0);
            Cx<F>* out = acc.line(BUF0 ^ 1);
#pragma unroll
            for (int q = 0; q < Q2; ++q) {
                const int b = j + q * T;
                if ((q + 1) * T <= NB2 || b < NB2) {
                    const int hi = b / Ra, k = b - hi * Ra;
                    Cx<F> x[Rb];
                    const Cx<F>* rd = in + (b + hi * D1) * S;
#pragma unroll
                    for (int r = 0; r < Rb; ++r) x[r] = rd[(r * (Ra * Ra + Ra * D1)) * S];
                    mix_twiddle<DIR, Rb, Ra, F>(x, tw, k);
                    Bfly<Rb, DIR, F>::run(x);
                    Cx<F>* wr = out + (hi * (T + D2) + k) * S;
#pragma unroll
                    for (int r = 0; r < Rb; ++r) wr[(r * Ra) * S] = x[r];
                }
            }
        }
        acc.sync();
        // ---- pass 3: radix Ra, Ns = T; inputs j + r*T, natural-order outputs j + r*T
        {
            const Cx<F>* rd = acc.line(BUF0 ^ 1) + j * S;
#pragma unroll
            for (int r = 0; r < Ra; ++r) v[r] = rd[(r * (T + D2)) * S];
            mix_twiddle<DIR, Ra, T, F>(v, tw + ((Ra * Rb + 1) & ~1), j);
            Bfly<Ra, DIR, F>::run(v);
        }
    }
};

}  // namespace p3d
