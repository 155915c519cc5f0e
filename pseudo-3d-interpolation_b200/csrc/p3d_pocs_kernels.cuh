// POCS kernels, generic path (any slice shape): shared-memory tiles + p3d_fft_generic.cuh.
//
// One POCS iteration is two launches over a band of slices:
//   k_cols_iter :  column FFT -> threshold(tau_k) -> column IFFT            (in place on W)
//   k_rows_iter :  row IFFT -> x = alpha*d + (1-alpha*m)*y/(N1 N2) -> sum|x| -> row FFT
// preceded once per slice by k_rows_init (row FFT of the observed slice + count_nonzero +
// sum|d|) and k_cols_stats (column FFT + the statistics the schedule needs).
#pragma once
#include <stdint.h>
#include <cuda.h>          // CUtensorMap (TMA descriptor of the work spectrum; no driver library is linked)
#include "../../include/p3d_b200.h"
#include "p3d_fft_generic.cuh"

namespace p3d {

// Per-slice statistics of the initial spectrum X0 = fft2(x) (functions/POCS.py:286-299,251-260)
struct SliceStats {
    unsigned long long lexmax_key;   // ordered (re, im) key of max_lex X0
    double sumsq;                    // sum |X0|^2
    unsigned int maxabs_bits;        // max |X0| (float bits, non-negative => ordered as uint)
    unsigned int minabs_bits;        // min |X0|
    unsigned long long nnz;          // count_nonzero(x)
    unsigned long long n_cand;       // data-driven: number of candidates inside (tau_min, tau_max)
    // float64 state mode only (ordered keys of doubles)
    unsigned long long re64_key;     // max real part of X0
    unsigned long long im64_key;     // max imaginary part among the elements with that real part
    unsigned long long maxabs64_key; // max |X0|
    unsigned long long minabs64_key; // min |X0|
};

__host__ __device__ __forceinline__ unsigned long long f64_ordered(double f) {
#ifdef __CUDA_ARCH__
    unsigned long long u = (unsigned long long)__double_as_longlong(f);
#else
    unsigned long long u; memcpy(&u, &f, 8);
#endif
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double f64_from_ordered(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double f; memcpy(&f, &u, 8); return f;
#endif
}

__host__ __device__ __forceinline__ unsigned int f32_ordered(float f) {
#ifdef __CUDA_ARCH__
    unsigned int u = __float_as_uint(f);
#else
    unsigned int u; memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(unsigned int u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
__host__ __device__ __forceinline__ unsigned long long lex_key(float re, float im) {
    return ((unsigned long long)f32_ordered(re) << 32) | (unsigned long long)f32_ordered(im);
}

struct PocsGeom {
    int n1, n2;            // rows (iline), columns (xline, contiguous)
    int C;                 // columns per column tile
    int RB;                // rows per row tile
    int pitch2;            // smem pitch of a row line (>= ax2.L)
    int slices_per_mask;   // mask index = global slice / slices_per_mask
};

template <typename T> struct BandArgs {
    Cx<T>* W;                    // work spectrum / iterate, [band][n1][n2]
    const Cx<T>* D;              // observed slices
    Cx<T>* OUT;                  // result slices (also scratch for X0 in the data-driven setup)
    const uint8_t* mask;         // [n_masks][n1][n2]
    const uint32_t* mbits;       // [n_masks][n1][T] packed mask words of the specialised row kernel (or null)
    long long first_slice;       // global index of the band's first slice (for the mask lookup)
    const Cx<T>* tau;            // [band][niter]
    double* S;                   // [band][niter+1]  (S[.][0] = sum|d|, S[.][k+1] = sum|x_k|)
    int* stop;                   // [band] 0 = active, k>0 = stopped after k iterations, -1 = all-zero slice
    SliceStats* stats;           // [band]
    int k, niter;
    double eps;
    T alpha, inv_n;
    int write_out, last, adaptive, store_x0, accum;
    int exact_tie;               // thresholds derived from |X0| values (inverse-proportional): honour exact ties
    const float* filt;           // kx-ky filter mode (P3D_OP_FILTER): real (n1, n2) plane multiplied into the spectrum
    // ---- escalating precision ("precision" = auto) ----
    const int* list;             // optional: blockIdx.y -> slice index inside the band arrays (compacted launches)
    int* esc;                    // optional [band]: 0 = fp32 state, k+1 = a coefficient came within `guard` of Re(tau_k)
                                 // in iteration k: x_k is the last fp32 iterate, complex128 state from k+1 on
    const float* guard;          // optional [band]: half-width of that guard band (fp32 kernels only)
    const Cx<float>* D32;        // complex128 kernels with complex64 observed data / results (no converted copies)
    Cx<float>* OUT32;
    double2* cand;               // complex128 statistics: per-CTA lexicographic maxima [band][cand_stride]
    int cand_stride;
    // support record of the fp32 pilot (exact restart): packed (row << 16 | col) of every coefficient kept in iteration k
    unsigned* arena;             // optional [band][arena_cap]
    int* acnt;                   // [band] entries used
    int* astart;                 // [band][niter + 1] first entry of iteration k (astart[.][0] = 0)
    int arena_cap;
    int scap;                    // freeze the slice after an iteration that kept more than scap coefficients (replay cost ~ |S|^2)
    int watch;                   // 1: a guard-band hit does not freeze the slice; killed coefficients inside the band are recorded
                                 // too (bit 31 set) and the float64 replay verifies every decision inside the band exactly
    int* wflag;                  // [band] set when a slice has recorded such a "watched kill"
    int restart;                 // complex128 kernels: this launch rebuilds x_{k_e - 1} (iteration index = esc[s] - 2 per slice)
    int store_x0_inplace;        // complex128 statistics kernel: leave X0 in W
    // TMA descriptor of the lane's work array as a 3-D tensor (column, row, slice) with a box of C columns x tma_rows rows:
    // the column kernels fetch their tile with cp.async.bulk.tensor instead of per-thread cp.async pieces
    int use_tma;                 // set by the launcher when the tile shape allows it (see launch_cols)
    int tma_rows;                // rows per box (divides n1, <= 256, box bytes a multiple of 128)
    int tma_slice0;              // slice index of W inside the tensor
    const void* tma_base;        // host request: base of the lane's work array and its slice capacity (null = no TMA)
    long long tma_slices;
    alignas(64) CUtensorMap tmapW;
};

// host: encode `map` for a (slices, n1, n2) array of complex values of `elem_bytes` bytes at `base`, box = C columns x
// `rows` rows x 1 slice.  Returns false when the driver entry point is unavailable or rejects the shape.
bool tma_encode_tile_map(CUtensorMap* map, const void* base, long long slices, int n1, int n2, int elem_bytes, int C, int rows, int row_step = 1);

// internal operator of the complex128 column kernel: the tile already holds a thresholded spectrum (exact restart):
// inverse transform only
#define P3D_OP_RESTART 4

// slice handled by this CTA (compacted launches go through the list)
template <typename T> __device__ __forceinline__ int band_slice(const BandArgs<T>& A) {
    return A.list ? A.list[blockIdx.y] : (int)blockIdx.y;
}
// fp32 kernels: the slice left the fp32 path in an EARLIER iteration (the column pass in which the hit happens is
// finished by all of its CTAs; the row pass of that iteration is already skipped: see slice_frozen)
template <typename T> __device__ __forceinline__ bool slice_escalated(const BandArgs<T>& A, const int s) {
    if (sizeof(T) != 4 || !A.esc) return false;
    const int e = A.esc[s];
    return e != 0 && e <= A.k;
}
// fp32 row kernels, one thread per slice and iteration: close this iteration's support record; a support that has
// outgrown the sparse replay freezes the slice from the NEXT iteration on (its own record is complete)
template <typename T> __device__ __forceinline__ void close_support_record(const BandArgs<T>& A, const int s) {
    const int end = A.acnt[s];
    int* as = A.astart + (long long)s * (A.niter + 1) + A.k;
    as[1] = end;
    if (A.scap > 0 && end - as[0] > A.scap && A.esc[s] == 0) A.esc[s] = A.k + 2;
}
// fp32 row kernels: esc = k + 1 was set by the column pass of this iteration k (or earlier): complex128 redoes iteration k
template <typename T> __device__ __forceinline__ bool slice_frozen(const BandArgs<T>& A, const int s) {
    if (sizeof(T) != 4 || !A.esc) return false;
    const int e = A.esc[s];
    return e != 0 && e <= A.k + 1;
}

// Append the packed indices of the coefficients this CTA kept to the slice's support record.  kept: bit e <-> idx[e].
// sh: >= 34 ints of shared memory.  Every thread of the CTA must call this (barriers inside).
// watched: coefficients the threshold killed INSIDE the guard band (watch mode): recorded with bit 31 set.
template <int E>
__device__ __forceinline__ void record_support(const BandArgs<float>& A, const int s, const unsigned (&idx)[E], const unsigned kept, int* sh,
                                               const unsigned watched = 0u) {
    const unsigned both = kept | watched;
    if (!__syncthreads_or(both != 0u)) return;
    if (watched) A.wflag[s] = 1;
    const int cnt = __popc(both);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31 || threadIdx.x == blockDim.x - 1) sh[w] = incl;      // (the last warp of a CTA may be partial)
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < nw; ++i) { const int t = sh[i]; sh[i] = tot; tot += t; }
        sh[32] = atomicAdd(&A.acnt[s], tot);
        sh[33] = tot;
    }
    __syncthreads();
    const int base = sh[32], tot = sh[33];
    if (base + tot > A.arena_cap) {
        // record full: the support has outgrown the sparse replay -> complex128 takes over from this iteration
        if (threadIdx.x == 0) A.esc[s] = A.k + 1;
        return;
    }
    unsigned* dst = A.arena + (long long)s * A.arena_cap + base + sh[w] + incl - cnt;
    int o = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) if ((both >> e) & 1u) dst[o++] = idx[e] | (((watched >> e) & 1u) << 31);
}
// Guard band of the escalating-precision mode: |X| within g of Re(tau_k) means that the decision of this or of a
// later iteration may depend on fp32 rounding.  Tested on |X|^2 (two compares per coefficient).
template <typename T> struct GuardBand {
    T lo2, hi2;
    bool on, hit;
    unsigned bits = 0u;          // bit e: element e is inside the band
    // centre = the modulus at which the operator jumps: Re(tau) for hard and soft, sqrt(Re(tau^2)) for garrote
    __device__ __forceinline__ GuardBand(const BandArgs<T>& A, const int s, const T a, const T b, const int op) {
        on = (sizeof(T) == 4) && (A.guard != nullptr);
        hit = false; lo2 = T(-1); hi2 = T(-1);
        if (on) {
            const T g = (T)A.guard[s];
            T c = a;
            if (op == P3D_OP_GARROTE) { const T c2 = a * a - b * b; c = c2 > T(0) ? sqrt(c2) : T(0); }
            const T lo = c - g, hi = c + g;
            if (hi > T(0)) { hi2 = hi * hi; lo2 = lo > T(0) ? lo * lo : T(-1); }
        }
    }
    __device__ __forceinline__ void test(const Cx<T> v, const int e = 0) {
        const T r2 = v.x * v.x + v.y * v.y;
        const bool in = (r2 > lo2) && (r2 < hi2);
        hit |= in;
        bits |= in ? (1u << e) : 0u;
    }
    // freeze on a hit (kernels without a watch list, or watch mode off)
    __device__ __forceinline__ void commit(const BandArgs<T>& A, const int s) const {
        if (on && hit) A.esc[s] = A.k + 1;       // every writer of this launch stores the same value
    }
};

// lexicographic (re, im) maximum over a CTA; result valid in thread 0.  red: >= 64 doubles of shared memory
__device__ __forceinline__ void block_lexmax(double& re, double& im, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double r2 = __shfl_xor_sync(0xffffffffu, re, o), i2 = __shfl_xor_sync(0xffffffffu, im, o);
        if (r2 > re || (r2 == re && i2 > im)) { re = r2; im = i2; }
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { red[2 * w] = re; red[2 * w + 1] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < nw; ++i) {
            const double r2 = red[2 * i], i2 = red[2 * i + 1];
            if (r2 > re || (r2 == re && i2 > im)) { re = r2; im = i2; }
        }
    }
}

// internal "threshold operator": multiply the spectrum by a real filter plane instead of thresholding
// (cube_postprocessing_3D.py:254,342  ifft2(filter * fft2(slice)))
#define P3D_OP_FILTER 3

// ---------------------------------------------------------------------------------------------
// threshold operators with the reference's complex-tau semantics
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float p3d_rsqrt(float v) { return rsqrtf(v); }
__device__ __forceinline__ double p3d_rsqrt(double v) { return 1.0 / sqrt(v); }

template <int OP, typename T, bool TIE = true>
__device__ __forceinline__ Cx<T> apply_threshold(Cx<T> X, const T a, const T b, const T t2re, const T t2im) {
    const T r2 = X.x * X.x + X.y * X.y;
    if (OP == P3D_OP_HARD && !TIE) {
        // schedules whose tau is p * z (linear / exponential): an exact tie |X| == Re(tau) has
        // measure zero, the squared comparison decides
        const bool kill = (a > T(0)) && (r2 < a * a);
        return kill ? cmake<T>(T(0), T(0)) : X;
    }
    if (OP == P3D_OP_HARD) {
        // |X| < a  <=>  |X|^2 < a^2 for a > 0 (never true for a <= 0): no square root on the hot
        // path.  Only when |X|^2 is within a few ulps of a^2 is the comparison redone on |X|
        // itself, exactly as the statistics kernel formed max|X0| / min|X0| (so that the element
        // that DEFINES a threshold is never killed by it) and with the lexicographic tie rule
        // |X| == a  =>  kill iff Im(tau) > 0.
        const T a2 = a * a;
        bool kill = (a > T(0)) && (r2 < a2);
        if (fabs(r2 - a2) <= T(2e-6) * a2) {
            const T r = sqrt(r2);
            kill = (r < a) || (r == a && b > T(0));
        }
        return kill ? cmake<T>(T(0), T(0)) : X;
    } else if (OP == P3D_OP_SOFT) {
        const T inv = p3d_rsqrt(r2);        // 1/|X| (inf for X == 0, handled below)
        T fre = T(1) - a * inv, fim = -(b * inv);
        const bool zero = (r2 == T(0)) || (fre < T(0)) || (fre == T(0) && fim < T(0));
        if (zero) return cmake<T>(T(0), T(0));
        return cmake<T>(X.x * fre - X.y * fim, X.x * fim + X.y * fre);
    } else {
        const T inv = T(1) / r2;
        T fre = T(1) - t2re * inv, fim = -(t2im * inv);
        const bool zero = (r2 == T(0)) || (fre < T(0)) || (fre == T(0) && fim < T(0));
        if (zero) return cmake<T>(T(0), T(0));
        return cmake<T>(X.x * fre - X.y * fim, X.x * fim + X.y * fre);
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
__device__ __forceinline__ unsigned int warp_max_u32(unsigned int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned int w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
__device__ __forceinline__ unsigned int warp_min_u32(unsigned int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { unsigned int w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}

// device-side replica of the reference's stop test (functions/POCS.py:631): iteration j = k-1
// finished with cost_j; stop when j > 2 and cost_j < eps.  Returns true when the slice is inactive.
__device__ __forceinline__ bool slice_stopped(int* stop, const double* S, int s, int k, int niter, double eps) {
    const int st = stop[s];
    if (st != 0) return true;
    if (eps > 0.0 && k >= 4) {
        const double sk = S[(long long)s * (niter + 1) + k], skm = S[(long long)s * (niter + 1) + k - 1];
        const double d = sk - skm;
        const double cost = (d * d) / (sk * sk);
        if (cost < eps) {
            if (threadIdx.x == 0) stop[s] = k;
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// generic column kernel.  MODE 0: statistics of X0 (optionally store X0), MODE 1: iterate
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int OP>
__global__ void __launch_bounds__(512) k_cols_generic(const __grid_constant__ PocsGeom G, const __grid_constant__ AxisDev<T> ax1,
                               const __grid_constant__ BandArgs<T> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = band_slice(A);
    const int tid = threadIdx.x, nth = blockDim.x;
    if (MODE == 1) { if ((!A.restart && slice_stopped(A.stop, A.S, s, A.k, A.niter, A.eps)) || slice_escalated(A, s)) return; }

    const int c0 = blockIdx.x * G.C;
    const int nc = min(G.C, G.n2 - c0);
    Cx<T>* bufA = reinterpret_cast<Cx<T>*>(smem_raw);
    Cx<T>* bufB = bufA + (size_t)ax1.L * G.C;
    const long long slice_off = (long long)s * G.n1 * G.n2;
    Cx<T>* Ws = A.W + slice_off;

    const int tot = G.n1 * nc;
    for (int w = tid; w < tot; w += nth) {
        const int i = w / nc, c = w - i * nc;
        bufA[i * G.C + c] = Ws[(long long)i * G.n2 + c0 + c];
    }
    __syncthreads();
    TileGeom tg; tg.nlines = nc; tg.line_stride = 1; tg.elem_stride = G.C; tg.line_fastest = 1;
    if (MODE == 1 && OP == P3D_OP_RESTART) {
        // exact restart: the tile holds the thresholded spectrum already, inverse transform only
        Cx<T>* Y = line_fft<+1, T>(bufA, bufB, tg, ax1, tid, nth);
        for (int w = tid; w < tot; w += nth) {
            const int i = w / nc, c = w - i * nc;
            Ws[(long long)i * G.n2 + c0 + c] = Y[i * G.C + c];
        }
        return;
    }
    Cx<T>* X = line_fft<-1, T>(bufA, bufB, tg, ax1, tid, nth);
    Cx<T>* other = (X == bufA) ? bufB : bufA;

    if (MODE == 0) {
        unsigned long long kmax = 0ull; double ss = 0.0; unsigned int amax = 0u, amin = 0xffffffffu;
        Cx<T>* X0s = A.OUT + slice_off;
        for (int w = tid; w < tot; w += nth) {
            const int i = w / nc, c = w - i * nc;
            const Cx<T> v = X[i * G.C + c];
            const unsigned long long key = lex_key((float)v.x, (float)v.y);
            kmax = key > kmax ? key : kmax;
            const float r2 = (float)(v.x * v.x + v.y * v.y);
            ss += (double)r2;
            const unsigned int rb = __float_as_uint(sqrtf(r2));
            amax = rb > amax ? rb : amax; amin = rb < amin ? rb : amin;
            if (A.store_x0) X0s[(long long)i * G.n2 + c0 + c] = v;
        }
        kmax = warp_max_u64(kmax); ss = warp_sum(ss); amax = warp_max_u32(amax); amin = warp_min_u32(amin);
        if ((tid & 31) == 0) {
            atomicMax(&A.stats[s].lexmax_key, kmax);
            atomicAdd(&A.stats[s].sumsq, ss);
            atomicMax(&A.stats[s].maxabs_bits, amax);
            atomicMin(&A.stats[s].minabs_bits, amin);
        }
        if (sizeof(T) == 8) {
            // float64 state: exact keys (the imaginary tie-break runs as a second pass over X0)
            unsigned long long rk = 0ull, ak = 0ull, ik = ~0ull;
            double ss64 = 0.0;
            for (int w = tid; w < tot; w += nth) {
                const int i = w / nc, c = w - i * nc;
                const Cx<T> v = X[i * G.C + c];
                const unsigned long long k1 = f64_ordered((double)v.x);
                rk = k1 > rk ? k1 : rk;
                const double r2 = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
                ss64 += r2;
                const unsigned long long k2 = f64_ordered(sqrt(r2));
                ak = k2 > ak ? k2 : ak; ik = k2 < ik ? k2 : ik;
            }
            rk = warp_max_u64(rk); ak = warp_max_u64(ak);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { unsigned long long w2 = __shfl_xor_sync(0xffffffffu, ik, o); ik = w2 < ik ? w2 : ik; }
            ss64 = warp_sum(ss64) - ss;           // replace the float-accumulated sum by the double one
            if ((tid & 31) == 0) {
                atomicMax(&A.stats[s].re64_key, rk);
                atomicMax(&A.stats[s].maxabs64_key, ak);
                atomicMin(&A.stats[s].minabs64_key, ik);
                atomicAdd(&A.stats[s].sumsq, ss64);
            }
            if (A.store_x0_inplace) {
                for (int w = tid; w < tot; w += nth) {
                    const int i = w / nc, c = w - i * nc;
                    Ws[(long long)i * G.n2 + c0 + c] = X[i * G.C + c];
                }
            }
            if (A.cand) {
                // lexicographic maximum of this tile (escalating mode: no stored X0, no second pass)
                double bre = -INFINITY, bim = -INFINITY;
                for (int w = tid; w < tot; w += nth) {
                    const int i = w / nc, c = w - i * nc;
                    const Cx<T> v = X[i * G.C + c];
                    if ((double)v.x > bre || ((double)v.x == bre && (double)v.y > bim)) { bre = (double)v.x; bim = (double)v.y; }
                }
                __syncthreads();
                block_lexmax(bre, bim, reinterpret_cast<double*>(other));
                if (tid == 0) A.cand[(long long)s * A.cand_stride + blockIdx.x] = make_double2(bre, bim);
            }
        }
        return;
    }

    // threshold in place
    const Cx<T> tau = A.tau[(long long)s * A.niter + A.k];
    const T a = tau.x, b = tau.y;
    const T t2re = a * a - b * b, t2im = T(2) * a * b;
    GuardBand<T> gb(A, s, a, b, OP);
    for (int w = tid; w < tot; w += nth) {
        const int i = w / nc, c = w - i * nc;
        Cx<T>* p = X + i * G.C + c;
        if (OP == P3D_OP_FILTER) { const T h = (T)A.filt[(long long)i * G.n2 + c0 + c]; *p = cmake<T>(p->x * h, p->y * h); }
        else { if (sizeof(T) == 4) gb.test(*p); *p = apply_threshold<OP, T>(*p, a, b, t2re, t2im); }
    }
    gb.commit(A, s);
    if (sizeof(T) == 4 && OP != P3D_OP_FILTER && A.arena) {
        // support record of this tile, 32 coefficients per thread and round (fp32 pilot of the escalating mode)
        __shared__ int rec_sh[34];
        const BandArgs<float>& Af = reinterpret_cast<const BandArgs<float>&>(A);
        for (int w0 = 0; w0 < tot; w0 += nth * 32) {
            unsigned idx[32]; unsigned kept = 0u;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int w = w0 + e * nth + tid;
                idx[e] = 0u;
                if (w < tot) {
                    const int i = w / nc, c = w - i * nc;
                    const Cx<T> v = X[i * G.C + c];
                    idx[e] = ((unsigned)i << 16) | (unsigned)(c0 + c);
                    if (v.x != T(0) || v.y != T(0)) kept |= 1u << e;
                }
            }
            record_support<32>(Af, s, idx, kept, rec_sh);
        }
    }
    __syncthreads();
    Cx<T>* Y = line_fft<+1, T>(X, other, tg, ax1, tid, nth);
    for (int w = tid; w < tot; w += nth) {
        const int i = w / nc, c = w - i * nc;
        Ws[(long long)i * G.n2 + c0 + c] = Y[i * G.C + c];
    }
}

// ---------------------------------------------------------------------------------------------
// generic row kernel.  MODE 0: init (row FFT of d, nnz, sum|d|), MODE 1: iterate
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(512) k_rows_generic(const __grid_constant__ PocsGeom G, const __grid_constant__ AxisDev<T> ax2,
                               const __grid_constant__ BandArgs<T> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red_s[32];
    __shared__ unsigned long long red_n[32];
    const int s = band_slice(A);
    const int tid = threadIdx.x, nth = blockDim.x;
    if (MODE == 1) { if (A.stop[s] != 0 || slice_frozen(A, s)) return; }
    if (MODE == 0 && A.adaptive) { if (A.stop[s] != 0) return; }
    // complex128 restart launch: the iteration rebuilt is the one before the slice's switch
    const int kk = (MODE == 1 && A.restart) ? A.esc[s] - 2 : A.k;
    if (MODE == 1 && sizeof(T) == 4 && A.astart && blockIdx.x == 0 && tid == 0) close_support_record(A, s);

    const int r0 = blockIdx.x * G.RB;
    const int nr = min(G.RB, G.n1 - r0);
    Cx<T>* bufA = reinterpret_cast<Cx<T>*>(smem_raw);
    Cx<T>* bufB = bufA + (size_t)G.pitch2 * G.RB;
    const long long slice_off = (long long)s * G.n1 * G.n2;
    const long long row_off = slice_off + (long long)r0 * G.n2;
    const long long mask_off = ((A.first_slice + s) / G.slices_per_mask) * (long long)G.n1 * G.n2 + (long long)r0 * G.n2;
    const int tot = nr * G.n2;
    TileGeom tg; tg.nlines = nr; tg.line_stride = G.pitch2; tg.elem_stride = 1; tg.line_fastest = 0;

    double part = 0.0;
    unsigned long long nnz = 0ull;
    Cx<T>* cur;

    if (MODE == 0) {
        for (int w = tid; w < tot; w += nth) {
            const int rr = w / G.n2, j = w - rr * G.n2;
            const long long g = row_off + (long long)rr * G.n2 + j;
            Cx<T> d;
            if (A.D32) { const Cx<float> t = A.D32[g]; d = cmake<T>((T)t.x, (T)t.y); }
            else d = A.D[g];
            if (!A.adaptive) {
                nnz += (d.x != T(0) || d.y != T(0)) ? 1ull : 0ull;
                part += (double)sqrt(d.x * d.x + d.y * d.y);
            } else {
                // APOCS prologue with x_old = x (functions/POCS.py:572-575)
                const T m = (T)A.mask[mask_off + (long long)rr * G.n2 + j];
                const T keep = T(1) - A.alpha * m;
                Cx<T> xt = cmake<T>(A.alpha * d.x + keep * d.x, A.alpha * d.y + keep * d.y);
                const T om = T(1) - A.alpha;
                d = cmake<T>(xt.x + om * (d.x - m * d.x), xt.y + om * (d.y - m * d.y));
            }
            bufA[rr * G.pitch2 + j] = d;
        }
        cur = bufA;
    } else {
        for (int w = tid; w < tot; w += nth) {
            const int rr = w / G.n2, j = w - rr * G.n2;
            bufA[rr * G.pitch2 + j] = A.W[row_off + (long long)rr * G.n2 + j];
        }
        __syncthreads();
        cur = line_fft<+1, T>(bufA, bufB, tg, ax2, tid, nth);
        // re-insertion: x = alpha*d + (1 - alpha*m) * y / (N1 N2)      (functions/POCS.py:616-619)
        for (int w = tid; w < tot; w += nth) {
            const int rr = w / G.n2, j = w - rr * G.n2;
            const long long g = (long long)rr * G.n2 + j;
            Cx<T> d;
            if (A.D32) { const Cx<float> t = A.D32[row_off + g]; d = cmake<T>((T)t.x, (T)t.y); }
            else d = A.D[row_off + g];
            const T m = (T)A.mask[mask_off + g];
            const T coef = (T(1) - A.alpha * m) * A.inv_n;
            const Cx<T> y = cur[rr * G.pitch2 + j];
            Cx<T> x = cmake<T>(fma(coef, y.x, A.alpha * d.x), fma(coef, y.y, A.alpha * d.y));
            part += (double)sqrt(x.x * x.x + x.y * x.y);
            if (A.write_out) {
                if (A.OUT32) A.OUT32[row_off + g] = cmake<float>((float)x.x, (float)x.y);
                else A.OUT[row_off + g] = x;
            }
            if (A.adaptive) {
                const T keep = T(1) - A.alpha * m;
                Cx<T> xt = cmake<T>(A.alpha * d.x + keep * x.x, A.alpha * d.y + keep * x.y);
                const T om = T(1) - A.alpha;
                x = cmake<T>(xt.x + om * (d.x - m * x.x), xt.y + om * (d.y - m * x.y));
            }
            cur[rr * G.pitch2 + j] = x;
        }
    }
    // block reduction of sum|.| (and nnz)
    part = warp_sum(part);
    if (MODE == 0) { for (int o = 16; o > 0; o >>= 1) nnz += __shfl_xor_sync(0xffffffffu, nnz, o); }
    if ((tid & 31) == 0) { red_s[tid >> 5] = part; red_n[tid >> 5] = nnz; }
    __syncthreads();
    if (tid < 32) {
        const int nw = (nth + 31) >> 5;
        double v = tid < nw ? red_s[tid] : 0.0;
        unsigned long long c = tid < nw ? red_n[tid] : 0ull;
        v = warp_sum(v);
        if (MODE == 0) { for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o); }
        if (tid == 0) {
            if (MODE == 0) {
                if (A.accum) {
                    atomicAdd(&A.S[(long long)s * (A.niter + 1)], v);
                    atomicAdd(&A.stats[s].nnz, c);
                }
            } else {
                atomicAdd(&A.S[(long long)s * (A.niter + 1) + kk + 1], v);
            }
        }
    }
    if (MODE == 1 && A.last) return;
    Cx<T>* other = (cur == bufA) ? bufB : bufA;
    Cx<T>* X = line_fft<-1, T>(cur, other, tg, ax2, tid, nth);
    for (int w = tid; w < tot; w += nth) {
        const int rr = w / G.n2, j = w - rr * G.n2;
        A.W[row_off + (long long)rr * G.n2 + j] = X[rr * G.pitch2 + j];
    }
}

}  // namespace p3d
