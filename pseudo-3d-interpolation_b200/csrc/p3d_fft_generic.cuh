// Generic shared-memory line FFT: any length.
//
//  * lengths whose prime factors are all in {2,3,5,7,11,13} run as Stockham autosort passes
//    (radices 16/10/8/5/4/3/2/7/11/13 chosen on the host), ping-ponging between two
//    shared-memory buffers;
//  * every other length (e.g. the prime 1201 of BASELINE config 3) runs as Bluestein's
//    chirp-z convolution over a smooth length L >= 2n-1 built from the same passes.
//
// A "tile" is `nlines` independent lines.  Element (line, pos) lives at
// buf[line * line_stride + pos * elem_stride]; column tiles use (line_stride=1,
// elem_stride=C) so that adjacent threads touch adjacent lines (bank-conflict free), row
// tiles use (line_stride=pitch, elem_stride=1).
#pragma once
#include "p3d_butterflies.cuh"

namespace p3d {

constexpr int P3D_MAX_PASSES = 12;

// Device-side description of one axis transform (passed by value as a kernel argument).
template <typename T> struct AxisDev {
    int n;                      // logical length
    int L;                      // transform length actually executed (n, or Bluestein length)
    int npass;
    int radix[P3D_MAX_PASSES];
    int bluestein;              // 0/1
    const Cx<T>* tw;            // W_L^t = exp(-2 pi i t / L), t in [0, L)
    const Cx<T>* chirp;         // exp(-i pi j^2 / n), j in [0, n)          (Bluestein only)
    const Cx<T>* bfilt;         // FFT_L(conj-chirp kernel) / L               (Bluestein only)
};

struct TileGeom {
    int nlines;
    int line_stride;
    int elem_stride;
    int line_fastest;           // 1: adjacent threads -> adjacent lines, 0: adjacent butterflies
};

template <int R, int DIR, typename T>
__device__ __forceinline__ void stockham_pass(const Cx<T>* __restrict__ src, Cx<T>* __restrict__ dst,
                                              const TileGeom g, const int L, const int Ns,
                                              const Cx<T>* __restrict__ tw, const int tid, const int nthreads) {
    const int nb = L / R;                    // butterflies per line
    const int total = nb * g.nlines;
    const int tw_step = L / (Ns * R);
    for (int w = tid; w < total; w += nthreads) {
        int line, b;
        if (g.line_fastest) { b = w / g.nlines; line = w - b * g.nlines; }
        else                { line = w / nb;    b = w - line * nb; }
        const int k = b % Ns;
        const Cx<T>* s = src + line * g.line_stride;
        Cx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = s[(b + r * nb) * g.elem_stride];
        if (Ns > 1) {
            const int t1 = k * tw_step;
#pragma unroll
            for (int r = 1; r < R; ++r) {
                Cx<T> wv = tw[r * t1];
                v[r] = (DIR < 0) ? cmul(v[r], wv) : cmulc(v[r], wv);
            }
        }
        Bfly<R, DIR, T>::run(v);
        Cx<T>* d = dst + line * g.line_stride;
        const int j0 = (b - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) d[(j0 + r * Ns) * g.elem_stride] = v[r];
    }
}

// All passes of a length-L transform.  Returns the buffer that holds the result.
// Every thread of the CTA must call this; contains __syncthreads().
template <int DIR, typename T>
__device__ __forceinline__ Cx<T>* fft_passes(Cx<T>* a, Cx<T>* b, const TileGeom g, const AxisDev<T>& ax,
                                             const int tid, const int nthreads) {
    int Ns = 1;
    for (int p = 0; p < ax.npass; ++p) {
        const int R = ax.radix[p];
        switch (R) {
            case 2:  stockham_pass<2, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 3:  stockham_pass<3, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 4:  stockham_pass<4, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 5:  stockham_pass<5, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 7:  stockham_pass<7, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 8:  stockham_pass<8, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 10: stockham_pass<10, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 11: stockham_pass<11, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 13: stockham_pass<13, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            case 16: stockham_pass<16, DIR, T>(a, b, g, ax.L, Ns, ax.tw, tid, nthreads); break;
            default: break;
        }
        Ns *= R;
        __syncthreads();
        Cx<T>* t = a; a = b; b = t;
    }
    return a;
}

// Full line transform of logical length ax.n (unscaled in both directions).
// Input in `a` (positions [0, n) of every line; each line has room for ax.L positions),
// `b` is scratch of the same size.  Returns the buffer holding the result.
// `extra_scale` multiplies the result (used to fold the 1/(N1 N2) of the inverse 2-D FFT).
template <int DIR, typename T>
__device__ __forceinline__ Cx<T>* line_fft(Cx<T>* a, Cx<T>* b, const TileGeom g, const AxisDev<T>& ax,
                                           const int tid, const int nthreads) {
    if (!ax.bluestein) return fft_passes<DIR, T>(a, b, g, ax, tid, nthreads);

    const int n = ax.n, L = ax.L;
    const int total = L * g.nlines;
    // 1. a[j] *= chirp[j] (conj for the inverse), zero-pad to L
    for (int w = tid; w < total; w += nthreads) {
        int line, j;
        if (g.line_fastest) { j = w / g.nlines; line = w - j * g.nlines; }
        else                { line = w / L;     j = w - line * L; }
        Cx<T>* p = a + line * g.line_stride + j * g.elem_stride;
        if (j < n) {
            Cx<T> c = ax.chirp[j];
            *p = (DIR < 0) ? cmul(*p, c) : cmulc(*p, c);
        } else {
            *p = cmake<T>(T(0), T(0));
        }
    }
    __syncthreads();
    // 2. forward FFT_L
    Cx<T>* r = fft_passes<-1, T>(a, b, g, ax, tid, nthreads);
    Cx<T>* o = (r == a) ? b : a;
    // 3. multiply by the filter spectrum (conj for the inverse direction; 1/L is folded in)
    for (int w = tid; w < total; w += nthreads) {
        int line, j;
        if (g.line_fastest) { j = w / g.nlines; line = w - j * g.nlines; }
        else                { line = w / L;     j = w - line * L; }
        Cx<T>* p = r + line * g.line_stride + j * g.elem_stride;
        Cx<T> f = ax.bfilt[j];
        *p = (DIR < 0) ? cmul(*p, f) : cmulc(*p, f);
    }
    __syncthreads();
    // 4. inverse FFT_L
    Cx<T>* q = fft_passes<+1, T>(r, o, g, ax, tid, nthreads);
    // 5. multiply by chirp[k], k < n
    const int totn = n * g.nlines;
    for (int w = tid; w < totn; w += nthreads) {
        int line, j;
        if (g.line_fastest) { j = w / g.nlines; line = w - j * g.nlines; }
        else                { line = w / n;     j = w - line * n; }
        Cx<T>* p = q + line * g.line_stride + j * g.elem_stride;
        Cx<T> c = ax.chirp[j];
        *p = (DIR < 0) ? cmul(*p, c) : cmulc(*p, c);
    }
    __syncthreads();
    return q;
}

}  // namespace p3d
