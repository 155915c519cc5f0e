// Column pass of the POCS iteration for a PRIME iline count P (config 3: 1201), register-resident.
//
// Rader's algorithm turns a length-P DFT into a cyclic convolution of length M = P - 1:
//     X[0]      = x[0] + sum_q a[q]                      a[q] = x[g^q mod P]   (g = primitive root mod P)
//     X[g^-m]   = x[0] + (a (*) b)[m]                     b[t] = w^(g^-t),  w = exp(-2 pi i / P)
// and the convolution is two M-point FFTs (M = 1200 = 10 x 12 x 10, p3d_fft_mix.cuh) around a table
// multiply by B = FFT_M(b) / M.  What makes this cheap inside the column kernel:
//   * a column tile is loaded row by row anyway (C adjacent columns per row segment), so the input
//     permutation x[g^q] is just the ROW INDEX of each load - no shuffle, no shared-memory gather;
//   * sum_q a[q] is the DC output of the first FFT, and adding x[0] to every convolution output is
//     adding it to the DC input of the inverse FFT;
//   * the threshold is element-wise, so the spectrum is left in Rader order (position m <-> frequency
//     g^-m): the inverse DFT  y[g^s] = Y[0] + (h (*) b2)[s],  h[m] = Y[g^-m],  b2[t] = conj(w)^(g^t)
//     consumes it in exactly that order and produces y[g^s] in the order of the load permutation.
// One column-iteration therefore costs four 1200-point register FFTs and two table multiplies, against
// four 2560-point shared-memory FFTs for the generic Bluestein path.
// Table buffer layout (Cx<float> slots): [twiddles of 10x12x10 : T + M][B : M][B2 : M][perm : M ints][kperm : M ints]
#include "p3d_pocs_spec.cuh"
#include "p3d_fft_mix.cuh"

#include <cmath>
#include <complex>
#include <cstring>
#include <cstdint>

namespace p3d {

__device__ __forceinline__ Cx<float> ldg_cx(const Cx<float>* p) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    return cmake<float>(t.x, t.y);
}
__device__ __forceinline__ Cx<double> ldg_cx(const Cx<double>* p) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    return cmake<double>(t.x, t.y);
}

// MODE 0: statistics of X0 (+ optional store), MODE 1: iterate (threshold / filter)
// BULK: tile load with cp.async.cg into exchange buffer 1 (see k_cols_spec), the row permutation applied to the SOURCE row
// F = double (complex128 state of the escalating / float64 modes): MODE 0 leaves the exact statistics (per-CTA lexicographic
// maxima in A.cand) and, with A.store_x0_inplace, X0 in natural order in W; op = P3D_OP_RESTART takes a thresholded
// spectrum in natural order and runs the inverse half only.
template <typename F, typename MP, int P, int C, int MINB, int MODE, bool BULK>
__global__ void __launch_bounds__(MP::T* C, MINB)
k_cols_rader(const __grid_constant__ PocsGeom G, const Cx<F>* __restrict__ tab, const __grid_constant__ BandArgs<F> A, const int op) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = MP::E, T = MP::T, M = MP::N;
    static_assert(M == P - 1, "Rader: the convolution length is P - 1");
    const Cx<F>* __restrict__ tw = tab;
    const Cx<F>* __restrict__ Bf = tab + (T + M);
    const Cx<F>* __restrict__ Bi = Bf + M;
    const int* __restrict__ perm = reinterpret_cast<const int*>(Bi + M);
    const int* __restrict__ kperm = perm + M;
    const int s = band_slice(A);
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const int col = blockIdx.x * C + c;
    const bool ok = col < G.n2;
    Cx<F>* __restrict__ Ws = A.W + (long long)s * P * G.n2 + col;
    ColAcc<F, C, MP::LINE> acc; acc.base = reinterpret_cast<Cx<F>*>(smem_raw) + c;

    const bool restart = (MODE == 1) && (sizeof(F) == 8) && (op == P3D_OP_RESTART);
    int row[E];
#pragma unroll
    for (int e = 0; e < E; ++e) row[e] = __ldg(perm + j + e * T);
    Cx<F> v[E];
    constexpr int CHUNKS = (C * (int)sizeof(Cx<F>)) / 16;
    const bool bulk = BULK && (CHUNKS >= 1) && (blockIdx.x * C + C <= G.n2) && (((long long)G.n2 * sizeof(Cx<F>)) % 16 == 0) &&
                      ((reinterpret_cast<uintptr_t>(A.W) % 16) == 0);
    if (bulk) {
        const char* src0 = reinterpret_cast<const char*>(A.W + (long long)s * P * G.n2 + blockIdx.x * C);
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(reinterpret_cast<Cx<F>*>(smem_raw) + (size_t)MP::LINE * C);
        for (int q = tid; q < M * CHUNKS; q += T * C) {
            const int m = q / CHUNKS, part = q - m * CHUNKS;
            const char* src = src0 + (long long)__ldg((restart ? kperm : perm) + m) * G.n2 * sizeof(Cx<F>) + part * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (unsigned)(m * C * sizeof(Cx<F>) + part * 16)), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = ok ? Ws[(long long)(restart ? __ldg(kperm + j + e * T) : row[e]) * G.n2] : cmake<F>(F(0), F(0));
    }
    Cx<F> dc = (ok && j == 0) ? Ws[0] : cmake<F>(F(0), F(0));
    Cx<F> tau = cmake<F>(F(0), F(0));
    if (MODE == 1) {
        tau = A.tau[(long long)s * A.niter + A.k];
        if ((!A.restart && slice_stopped(A.stop, A.S, s, A.k, A.niter, A.eps)) || slice_escalated(A, s)) { if (bulk) asm volatile("cp.async.wait_all;"); return; }
    }
    if (bulk) {
        asm volatile("cp.async.wait_all;");
        __syncthreads();
        const Cx<F>* land = acc.line(1);
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = land[(j + e * T) * C];
    }

    // ---- forward DFT of length P (not for a restart: the tile holds the thresholded spectrum, loaded in Rader order)
    if (!restart) {
        MP::template fft<-1, 0, F>(v, acc, j, tw);
        {
            const Cx<F> x0 = dc;
            if (j == 0) dc = cadd(x0, v[0]);                 // X[0] = x[0] + sum_q a[q]
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = cmul(v[e], ldg_cx(Bf + j + e * T));
            if (j == 0) v[0] = cadd(v[0], x0);               // + x[0] on every convolution output
        }
        MP::template fft<+1, 0, F>(v, acc, j, tw);
    }
    // v[e] = X[kperm[j + e*T]], dc = X[0] (threads j == 0)

    if (MODE == 0 && sizeof(F) == 8) {
        double bre = -INFINITY, bim = -INFINITY, ss = 0.0;
        unsigned long long ak = 0ull, ik = ~0ull;
        Cx<F>* __restrict__ Wo = A.W + (long long)s * P * G.n2 + col;
        auto account = [&](const Cx<F> x, const int frow) {
            if ((double)x.x > bre || ((double)x.x == bre && (double)x.y > bim)) { bre = (double)x.x; bim = (double)x.y; }
            const double r2 = (double)x.x * (double)x.x + (double)x.y * (double)x.y;
            ss += r2;
            const unsigned long long k2 = f64_ordered(sqrt(r2));
            ak = k2 > ak ? k2 : ak; ik = k2 < ik ? k2 : ik;
            if (A.store_x0_inplace) Wo[(long long)frow * G.n2] = x;
        };
        if (ok) {
#pragma unroll
            for (int e = 0; e < E; ++e) account(v[e], __ldg(kperm + j + e * T));
            if (j == 0) account(dc, 0);
        }
        ss = warp_sum(ss); ak = warp_max_u64(ak);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long w2 = __shfl_xor_sync(0xffffffffu, ik, o); ik = w2 < ik ? w2 : ik; }
        if ((tid & 31) == 0) {
            atomicAdd(&A.stats[s].sumsq, ss);
            atomicMax(&A.stats[s].maxabs64_key, ak);
            atomicMin(&A.stats[s].minabs64_key, ik);
        }
        __syncthreads();
        block_lexmax(bre, bim, reinterpret_cast<double*>(smem_raw));
        if (tid == 0 && A.cand) A.cand[(long long)s * A.cand_stride + blockIdx.x] = make_double2(bre, bim);
        return;
    }
    if (MODE == 0) {
        unsigned long long kmax = 0ull; float ssf = 0.f; unsigned int amax = 0u, amin = 0xffffffffu;
        Cx<F>* __restrict__ X0 = A.OUT + (long long)s * P * G.n2 + col;
        auto account = [&](const Cx<F> x, const int frow) {
            const unsigned long long key = lex_key((float)x.x, (float)x.y);
            kmax = key > kmax ? key : kmax;
            const float r2 = (float)(x.x * x.x + x.y * x.y);
            ssf += r2;
            const unsigned int rb = __float_as_uint(sqrtf(r2));
            amax = rb > amax ? rb : amax; amin = rb < amin ? rb : amin;
            if (A.store_x0) X0[(long long)frow * G.n2] = x;
        };
        if (ok) {
#pragma unroll
            for (int e = 0; e < E; ++e) account(v[e], __ldg(kperm + j + e * T));
            if (j == 0) account(dc, 0);
        }
        kmax = warp_max_u64(kmax); const double ss = warp_sum((double)ssf); amax = warp_max_u32(amax); amin = warp_min_u32(amin);
        if ((tid & 31) == 0) {
            atomicMax(&A.stats[s].lexmax_key, kmax);
            atomicAdd(&A.stats[s].sumsq, ss);
            atomicMax(&A.stats[s].maxabs_bits, amax);
            atomicMin(&A.stats[s].minabs_bits, amin);
        }
        return;
    }

    // ---- threshold (or kx-ky filter) in Rader order
    const F a = tau.x, b = tau.y;
    const F t2re = a * a - b * b, t2im = F(2) * a * b;
    unsigned nearbits = 0u;
    if (sizeof(F) == 4 && A.guard && op != P3D_OP_FILTER) {
        GuardBand<F> gb(A, s, a, b, op);
#pragma unroll
        for (int e = 0; e < E; ++e) gb.test(v[e], e);
        if (j == 0) gb.test(dc, E);
        if (!ok) { gb.hit = false; gb.bits = 0u; }
        if (A.watch) nearbits = gb.bits; else gb.commit(A, s);
    }
    if (op == P3D_OP_HARD && !A.exact_tie) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_HARD, F, false>(v[e], a, b, t2re, t2im);
        dc = apply_threshold<P3D_OP_HARD, F, false>(dc, a, b, t2re, t2im);
    } else if (op == P3D_OP_HARD) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_HARD, F>(v[e], a, b, t2re, t2im);
        dc = apply_threshold<P3D_OP_HARD, F>(dc, a, b, t2re, t2im);
    } else if (op == P3D_OP_SOFT) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_SOFT, F>(v[e], a, b, t2re, t2im);
        dc = apply_threshold<P3D_OP_SOFT, F>(dc, a, b, t2re, t2im);
    } else if (op == P3D_OP_GARROTE) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = apply_threshold<P3D_OP_GARROTE, F>(v[e], a, b, t2re, t2im);
        dc = apply_threshold<P3D_OP_GARROTE, F>(dc, a, b, t2re, t2im);
    } else if (op == P3D_OP_FILTER) {
        const float* __restrict__ H = A.filt + col;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const F h = ok ? (F)__ldg(H + (long long)__ldg(kperm + j + e * T) * G.n2) : F(0);
            v[e] = cmake<F>(v[e].x * h, v[e].y * h);
        }
        const F h0 = (ok && j == 0) ? (F)__ldg(H) : F(0);
        dc = cmake<F>(dc.x * h0, dc.y * h0);
    }

    if (sizeof(F) == 4 && A.arena && op != P3D_OP_FILTER) {
        // support record of the fp32 pilot (exact restart): natural frequency row of every surviving coefficient
        __shared__ int rec_sh[34];
        unsigned idx[E + 1]; unsigned kept = 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            idx[e] = ((unsigned)__ldg(kperm + j + e * T) << 16) | (unsigned)col;
            if (ok && (v[e].x != F(0) || v[e].y != F(0))) kept |= 1u << e;
        }
        idx[E] = (unsigned)col;
        if (ok && j == 0 && (dc.x != F(0) || dc.y != F(0))) kept |= 1u << E;
        record_support<E + 1>(reinterpret_cast<const BandArgs<float>&>(A), s, idx, kept, rec_sh, nearbits & ~kept);
    }

    // ---- inverse DFT of length P (unscaled): h[m] = Y[g^-m] is already in natural order of m
    MP::template fft<-1, 0, F>(v, acc, j, tw);
    {
        const Cx<F> y0 = dc;
        if (j == 0) dc = cadd(y0, v[0]);                 // y[0] = sum_k Y[k]
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], ldg_cx(Bi + j + e * T));
        if (j == 0) v[0] = cadd(v[0], y0);
    }
    MP::template fft<+1, 0, F>(v, acc, j, tw);
    // v[e] = y[g^(j + e*T)] = y[row[e]]

    if (ok) {
#pragma unroll
        for (int e = 0; e < E; ++e) Ws[(long long)row[e] * G.n2] = v[e];
        if (j == 0) Ws[0] = dc;
    }
}

// ---- host: tables ----------------------------------------------------------------------------------------
static long powmod(long b, long e, long m) { long r = 1; b %= m; while (e > 0) { if (e & 1) r = r * b % m; b = b * b % m; e >>= 1; } return r; }
static int primitive_root(int p) {
    std::vector<int> factors;
    int phi = p - 1, n = phi;
    for (int f = 2; f * f <= n; ++f) if (n % f == 0) { factors.push_back(f); while (n % f == 0) n /= f; }
    if (n > 1) factors.push_back(n);
    for (int g = 2; g < p; ++g) {
        bool ok = true;
        for (int f : factors) if (powmod(g, phi / f, p) == 1) { ok = false; break; }
        if (ok) return g;
    }
    return -1;
}

// ROWS = false: kernels of the column pass (forward DFT first: B = FFT(w^(g^-t)), B2 = FFT(conj(w)^(g^t)));
// ROWS = true: kernels of the row pass (inverse DFT first: B = FFT(conj(w)^(g^-t)), B2 = FFT(w^(g^t)))
static std::vector<Cx<float>> twiddles_of(const std::vector<int>& r, float) { return spec_twiddle_table(r); }
static std::vector<Cx<double>> twiddles_of(const std::vector<int>& r, double) { return spec_twiddle_table64(r); }

template <typename MP, int P, bool ROWS = false, typename F = float> static std::vector<Cx<F>> rader_tables() {
    constexpr int M = P - 1;
    int rad[3]; MP::radices(rad);
    std::vector<Cx<F>> t = twiddles_of(std::vector<int>(rad, rad + 3), F(0));
    static_assert(MP::T % 2 == 0, "Rader table offsets assume an even pass-2 table");
    t.resize((size_t)MP::T + M);                                    // exact size of the twiddle section
    const int g = primitive_root(P);
    const long ginv = powmod(g, P - 2, P);
    std::vector<int> perm(M), kperm(M);
    for (int q = 0; q < M; ++q) { perm[q] = (int)powmod(g, q, P); kperm[q] = (int)powmod(ginv, q, P); }
    typedef std::complex<double> cd;
    std::vector<cd> bf(M), bi(M);
    for (int q = 0; q < M; ++q) {
        const double sg = ROWS ? -1.0 : 1.0;
        bf[q] = std::polar(1.0, -sg * 2.0 * M_PI * (double)kperm[q] / (double)P);      // w^(g^-q)       (rows: conjugate)
        bi[q] = std::polar(1.0, +sg * 2.0 * M_PI * (double)perm[q] / (double)P);       // conj(w)^(g^q)  (rows: conjugate)
    }
    // M-point DFTs of the two kernels (once per plan: M^2 = 1.44 M complex multiplies), scaled by 1/M
    std::vector<cd> wM(M);
    for (int k = 0; k < M; ++k) wM[k] = std::polar(1.0, -2.0 * M_PI * (double)k / (double)M);
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<cd>& src = pass == 0 ? bf : bi;
        for (int k = 0; k < M; ++k) {
            cd acc(0, 0);
            for (int q = 0; q < M; ++q) acc += src[q] * wM[(int)(((long)k * q) % M)];
            acc /= (double)M;
            t.push_back(cmake<F>((F)acc.real(), (F)acc.imag()));
        }
    }
    const size_t int_slots = ((size_t)2 * M * sizeof(int) + sizeof(Cx<F>) - 1) / sizeof(Cx<F>);      // 2 * M ints
    const size_t base = t.size();
    t.resize(base + int_slots);
    int* ip = reinterpret_cast<int*>(t.data() + base);
    memcpy(ip, perm.data(), sizeof(int) * M);
    memcpy(ip + M, kperm.data(), sizeof(int) * M);
    return t;
}

template <typename MP, int P, int C, int MINB, int MODE, bool BULK = false, typename F = float>
static void launch_rader(const PocsGeom& G, const Cx<F>* tab, const BandArgs<F>& A, int ns, int op, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * MP::LINE * C * sizeof(Cx<F>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_cols_rader<F, MP, P, C, MINB, MODE, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid((G.n2 + C - 1) / C, ns);
    k_cols_rader<F, MP, P, C, MINB, MODE, BULK><<<grid, MP::T * C, smem, st>>>(G, tab, A, op);
}
template <typename MP, int P, int C, int MINB>
static void launch_rader_stats64(const PocsGeom& G, const Cx<double>* tab, const BandArgs<double>& A, int ns, cudaStream_t st) {
    launch_rader<MP, P, C, MINB, 0, false, double>(G, tab, A, ns, 0, st);
}
template <typename MP, int P, int C, int MINB>
static void launch_rader_stats(const PocsGeom& G, const Cx<float>* tab, const BandArgs<float>& A, int ns, cudaStream_t st) {
    launch_rader<MP, P, C, MINB, 0>(G, tab, A, ns, 0, st);
}

typedef MixPlan3<1200, 10, 12> MP1200;
typedef MixPlan3<1200, 20, 3> MP1200E20;

// ---- row pass for a prime xline count -------------------------------------------------------------------------
// One row per CTA.  The row of W (and of the observed data) is staged in shared memory in natural order by coalesced
// loads; the Rader permutation is the gather from there.  Inverse DFT first (kernel conj(w)): slot m of the result
// holds y[g^-m], so the re-insertion reads d and the mask at n = kperm[m]; the forward DFT takes h[m] = x[g^-m] as it
// is and leaves X[g^s] in slot s, which goes back through shared memory to a coalesced store.
template <typename MP, int P, int MINB>
__global__ void __launch_bounds__(MP::T, MINB)
k_rows_rader(const __grid_constant__ PocsGeom G, const Cx<float>* __restrict__ tab, const __grid_constant__ BandArgs<float> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red_s[32];
    constexpr int E = MP::E, T = MP::T, M = MP::N;
    static_assert(M == P - 1 && MP::LINE >= P, "Rader rows: convolution length P - 1, staging fits an exchange buffer");
    const Cx<float>* __restrict__ tw = tab;
    const Cx<float>* __restrict__ B1 = tab + (T + M);
    const Cx<float>* __restrict__ B2 = B1 + M;
    const int* __restrict__ perm = reinterpret_cast<const int*>(B2 + M);
    const int* __restrict__ kperm = perm + M;
    const int s = band_slice(A), row = blockIdx.x, j = threadIdx.x;
    const int stopped = A.stop[s];

    const long long base = ((long long)s * G.n1 + row) * P;
    Cx<float>* __restrict__ Wp = A.W + base;
    const Cx<float>* __restrict__ Dp = A.D + base;
    Cx<float>* __restrict__ Op = A.OUT + base;
    RowAcc<float, 1, MP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw);
    Cx<float>* land = acc.line(1);                                       // W row, natural order (free until exchange 2)
    Cx<float>* dst = reinterpret_cast<Cx<float>*>(smem_raw) + 2 * MP::LINE;   // observed row, natural order
    for (int i = j; i < P; i += T) { land[i] = Wp[i]; dst[i] = Dp[i]; }
    if (stopped != 0 || slice_frozen(A, s)) return;
    if (A.astart && row == 0 && j == 0) close_support_record(A, s);
    __syncthreads();

    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = land[__ldg(perm + j + e * T)];
    Cx<float> dc = land[0];

    // ---- inverse DFT of length P (unscaled)
    MP::template fft<-1, 0, float>(v, acc, j, tw);
    {
        const Cx<float> w0 = dc;
        if (j == 0) dc = cadd(w0, v[0]);                 // y[0]
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], ldg_cx(B1 + j + e * T));
        if (j == 0) v[0] = cadd(v[0], w0);
    }
    MP::template fft<+1, 0, float>(v, acc, j, tw);
    // v[e] = y[kperm[j + e*T]], dc = y[0] (thread 0)

    // ---- re-insertion at the natural positions, norm
    const long long mrow = (((A.first_slice + s) / G.slices_per_mask) * G.n1 + row) * (long long)P;
    float part = 0.f;
    auto reinsert = [&](Cx<float> y, const int n) {
        const Cx<float> d = dst[n];
        const float m = (float)A.mask[mrow + n];
        const float coef = (1.f - A.alpha * m) * A.inv_n;
        Cx<float> x = cmake<float>(fmaf(coef, y.x, A.alpha * d.x), fmaf(coef, y.y, A.alpha * d.y));
        part += sqrtf(x.x * x.x + x.y * x.y);
        if (A.write_out) Op[n] = x;
        if (A.adaptive) {
            const float keep = 1.f - A.alpha * m, om = 1.f - A.alpha;
            const Cx<float> xt = cmake<float>(A.alpha * d.x + keep * x.x, A.alpha * d.y + keep * x.y);
            x = cmake<float>(xt.x + om * (d.x - m * x.x), xt.y + om * (d.y - m * x.y));
        }
        return x;
    };
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = reinsert(v[e], __ldg(kperm + j + e * T));
    if (j == 0) dc = reinsert(dc, 0);
    double dp = warp_sum((double)part);
    if ((j & 31) == 0) red_s[j >> 5] = dp;
    __syncthreads();
    if (j < 32) {
        constexpr int NW = (T + 31) / 32;
        double t = j < NW ? red_s[j] : 0.0;
        t = warp_sum(t);
        if (j == 0) atomicAdd(&A.S[(long long)s * (A.niter + 1) + A.k + 1], t);
    }
    if (A.last) return;

    // ---- forward DFT of length P: h[m] = x[g^-m] is the register order
    MP::template fft<-1, 0, float>(v, acc, j, tw);
    {
        const Cx<float> x0 = dc;
        if (j == 0) dc = cadd(x0, v[0]);                 // X[0]
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], ldg_cx(B2 + j + e * T));
        if (j == 0) v[0] = cadd(v[0], x0);
    }
    MP::template fft<+1, 0, float>(v, acc, j, tw);
    // v[e] = X[perm[j + e*T]]: back to natural order through exchange buffer 0 (last read before the second barrier
    // of the transform above), coalesced store
    Cx<float>* st = acc.line(0);
#pragma unroll
    for (int e = 0; e < E; ++e) st[__ldg(perm + j + e * T)] = v[e];
    if (j == 0) st[0] = dc;
    __syncthreads();
    for (int i = j; i < P; i += T) Wp[i] = st[i];
}

template <typename MP, int P, int MINB>
static void launch_rows_rader(const PocsGeom& G, const Cx<float>* tab, const BandArgs<float>& A, int ns, cudaStream_t st) {
    constexpr size_t smem = (size_t)(2 * MP::LINE + P) * sizeof(Cx<float>);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_rows_rader<MP, P, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid(G.n1, ns);
    k_rows_rader<MP, P, MINB><<<grid, MP::T, smem, st>>>(G, tab, A);
}

void rader_register_rows(SpecKernels& k, int n_xline, int variant) {
    if (n_xline != 1201) return;
    if (variant == 1) {
        k.rows_iter = launch_rows_rader<MP1200E20, 1201, 6>;
        k.rows_name = "rader<1201,20x3x20,RB1>";
        k.rows_radices = {20, 3, 20};
        k.rows_table = rader_tables<MP1200E20, 1201, true>;
    } else {
        k.rows_iter = launch_rows_rader<MP1200, 1201, 6>;
        k.rows_name = "rader<1201,10x12x10,RB1>";
        k.rows_radices = {10, 12, 10};
        k.rows_table = rader_tables<MP1200, 1201, true>;
    }
    k.rows_init = nullptr; k.pack_mask = nullptr; k.rows_T = 0;       // generic row FFT of the observed slice, byte mask
}

// complex128 state (escalating / float64 modes): the same algorithm over 10 x 12 x 10 (E = 10: 40 data registers)
void rader_register_cols64(SpecKernels64& k, int n_iline, int variant) {
    if (n_iline != 1201) return;
    (void)variant;
    k.cols_iter = launch_rader<MP1200, 1201, 2, 2, 1, false, double>;
    k.cols_stats = launch_rader_stats64<MP1200, 1201, 2, 2>;
    k.cols_C = 2;
    k.cols_name = "rader64<1201,10x12x10,C2,2cta>";
    k.cols_radices = {10, 12, 10};
    k.cols_table = rader_tables<MP1200, 1201, false, double>;
}

void rader_register_cols(SpecKernels& k, int n_iline, int variant) {
    if (n_iline != 1201) return;
    k.cols_radices = {20, 3, 20};
    k.cols_table = rader_tables<MP1200E20, 1201>;
    if (variant == 1) {
        k.cols_iter = launch_rader<MP1200E20, 1201, 4, 2, 1, true>; k.cols_stats = launch_rader_stats<MP1200E20, 1201, 4, 2>;
        k.cols_name = "rader<1201,20x3x20,C4,2cta,cp.async>";
    } else if (variant == 2) {
        k.cols_iter = launch_rader<MP1200E20, 1201, 2, 4, 1, true>; k.cols_stats = launch_rader_stats<MP1200E20, 1201, 4, 2>;
        k.cols_name = "rader<1201,20x3x20,C2,4cta,cp.async>";
    } else if (variant == 3) {
        k.cols_iter = launch_rader<MP1200E20, 1201, 2, 5, 1, true>; k.cols_stats = launch_rader_stats<MP1200E20, 1201, 4, 2>;
        k.cols_name = "rader<1201,20x3x20,C2,5cta,cp.async>";
    } else if (variant == 4) {
        k.cols_iter = launch_rader<MP1200, 1201, 4, 2, 1, true>; k.cols_stats = launch_rader_stats<MP1200, 1201, 4, 2>;
        k.cols_name = "rader<1201,10x12x10,C4,2cta,cp.async>";
        k.cols_radices = {10, 12, 10};
        k.cols_table = rader_tables<MP1200, 1201>;
    } else {       // the Rader kernel is bound by its four transforms, not by the tile load: cp.async changes nothing (1500 vs 1485 us)
        k.cols_iter = launch_rader<MP1200E20, 1201, 4, 2, 1>; k.cols_stats = launch_rader_stats<MP1200E20, 1201, 4, 2>;
        k.cols_name = "rader<1201,20x3x20,C4,2cta>";
    }
}

}  // namespace p3d
