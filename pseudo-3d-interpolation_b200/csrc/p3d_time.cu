// Kernel 1: batched trace rfft / irfft / envelope along the time axis of a (nt, n_traces) cube.
//
// The time axis is the SLOWEST axis of the (twt, iline, xline) cube, so a trace is a strided
// column.  A CTA takes a tile of 2C adjacent traces (contiguous 8C bytes per time sample),
// packs two real traces into one complex line (z = x_a + i x_b), runs one complex FFT per
// pair and separates the two spectra with the Hermitian identities; the epilogue applies
// dt * exp(-2 pi i f_k t0) * window[k] and writes the slice-major (nf, n_traces) complex64
// layout the POCS kernels consume.  The inverse does the mirror image (phase, Hermitian
// symmetrisation = "take the real part", packed complex IFFT).
//
// Implementations, in the order try_time_spec() prefers them:
//   k_time_{fwd,inv,env}_tma    one pass, persistent CTAs, TMA-staged tiles, register FFT (512 / 1000 / 1024 / 2500 samples)
//   k_time_{fwd,inv,env}_split  the same with N = 2 H: two H-point register FFTs and a radix-2 step, tiles twice as wide
//                               (2048 / 4096 / 2000 / 4000 samples)
//   k_time_{fwd,inv}_spec       direct register kernels, plain loads, one tile per CTA
//   k_transpose + k_time_*_rows transposing pipeline through L2 (any trace count / alignment)
//   k_time_fwd / _inv / _env    generic shared-memory kernels (any record length, Bluestein included)
#include "p3d_host.h"
#include "p3d_fft_reg.cuh"
#include "p3d_pocs_spec.cuh"
#include "p3d_pocs_spec_kernels.cuh"      // TMA / mbarrier helpers

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace p3d;

namespace {

struct TimeGeom {
    long long nt;          // samples present in the time cube (forward: input; inverse: output rows)
    long long nf;          // frequency rows
    long long ntr;         // traces
    int nfft;
    int C;                 // complex lines (trace pairs) per tile
    int compute_real;
    int ascending;
};

// forward: x (nt, ntr) float32 -> F (nf, ntr) complex64
__global__ void k_time_fwd(const __grid_constant__ TimeGeom G, const __grid_constant__ AxisDev<float> ax,
                           const float* __restrict__ x, Cx<float>* __restrict__ F, const Cx<float>* __restrict__ phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nth = blockDim.x;
    Cx<float>* bufA = reinterpret_cast<Cx<float>*>(smem_raw);
    Cx<float>* bufB = bufA + (size_t)ax.L * G.C;
    const long long tr0 = (long long)blockIdx.x * 2 * G.C;         // first trace of the tile
    const int ntr_tile = (int)min((long long)2 * G.C, G.ntr - tr0);
    const int nc = (ntr_tile + 1) / 2;
    const int N = G.nfft;

    // load + pack (zero beyond nt and beyond the last trace)
    const int tot = N * nc;
    for (int w = tid; w < tot; w += nth) {
        const int n = w / nc, c = w - n * nc;
        float a = 0.f, b = 0.f;
        if (n < G.nt) {
            const long long base = (long long)n * G.ntr + tr0 + 2 * c;
            a = x[base];
            if (2 * c + 1 < ntr_tile) b = x[base + 1];
        }
        bufA[n * G.C + c] = cmake<float>(a, b);
    }
    __syncthreads();
    TileGeom tg; tg.nlines = nc; tg.line_stride = 1; tg.elem_stride = G.C; tg.line_fastest = 1;
    Cx<float>* Z = line_fft<-1, float>(bufA, bufB, tg, ax, tid, nth);

    // separate the two real traces: Xa[k] = (Z[k] + conj Z[N-k]) / 2, Xb[k] = (Z[k] - conj Z[N-k]) / (2i)
    const int totf = (int)G.nf * nc;
    for (int w = tid; w < totf; w += nth) {
        const int k = w / nc, c = w - k * nc;
        const int km = (k == 0) ? 0 : N - k;
        const Cx<float> z1 = Z[k * G.C + c], z2 = Z[km * G.C + c];
        const Cx<float> xa = cmake<float>(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
        const Cx<float> xb = cmake<float>(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
        const Cx<float> ph = phase[k];
        const long long o = (long long)k * G.ntr + tr0 + 2 * c;
        F[o] = cmul(xa, ph);
        if (2 * c + 1 < ntr_tile) F[o + 1] = cmul(xb, ph);
    }
}

// inverse: F (nf, ntr) complex64 -> x (nt_out, ntr) float32
__global__ void k_time_inv(const __grid_constant__ TimeGeom G, const __grid_constant__ AxisDev<float> ax,
                           const Cx<float>* __restrict__ F, float* __restrict__ x, const Cx<float>* __restrict__ phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nth = blockDim.x;
    Cx<float>* bufA = reinterpret_cast<Cx<float>*>(smem_raw);
    Cx<float>* bufB = bufA + (size_t)ax.L * G.C;
    const long long tr0 = (long long)blockIdx.x * 2 * G.C;
    const int ntr_tile = (int)min((long long)2 * G.C, G.ntr - tr0);
    const int nc = (ntr_tile + 1) / 2;
    const int N = G.nfft;
    const int half = N / 2;

    // stage G_a[k] * phase[k] into bufA and G_b[k] * phase[k] into bufB, indexed by FFT bin k
    const int nbins = G.compute_real ? half + 1 : N;
    const int tot = nbins * nc;
    for (int w = tid; w < tot; w += nth) {
        const int k = w / nc, c = w - k * nc;
        // row of the input that holds bin k
        long long row = k;
        if (!G.compute_real && G.ascending) row = (k + half) % N;
        const Cx<float> ph = phase[k];
        const long long o = row * G.ntr + tr0 + 2 * c;
        Cx<float> ga = cmul(F[o], ph);
        Cx<float> gb = cmake<float>(0.f, 0.f);
        if (2 * c + 1 < ntr_tile) gb = cmul(F[o + 1], ph);
        bufA[k * G.C + c] = ga;
        bufB[k * G.C + c] = gb;
    }
    __syncthreads();
    // Hermitian symmetrisation (real part of the inverse transform) + packing H = Ha + i Hb;
    // one thread owns the pair (k, N-k)
    const int totp = (half + 1) * nc;
    for (int w = tid; w < totp; w += nth) {
        const int k = w / nc, c = w - k * nc;
        const int km = (k == 0) ? 0 : N - k;
        Cx<float> a1 = bufA[k * G.C + c], b1 = bufB[k * G.C + c];
        Cx<float> ha, hb;        // symmetrised spectra at bin k
        if (G.compute_real) {
            // irfft semantics: bins 0..N/2 given, imaginary part of DC / Nyquist ignored
            ha = a1; hb = b1;
            if (k == 0 || k == half) { ha.y = 0.f; hb.y = 0.f; }
        } else {
            const Cx<float> a2 = bufA[km * G.C + c], b2 = bufB[km * G.C + c];
            ha = cmake<float>(0.5f * (a1.x + a2.x), 0.5f * (a1.y - a2.y));
            hb = cmake<float>(0.5f * (b1.x + b2.x), 0.5f * (b1.y - b2.y));
        }
        // H[k] = ha + i hb ; H[N-k] = conj(ha) + i conj(hb)
        bufA[k * G.C + c] = cmake<float>(ha.x - hb.y, ha.y + hb.x);
        if (km != k) bufA[km * G.C + c] = cmake<float>(ha.x + hb.y, hb.x - ha.y);
    }
    __syncthreads();
    TileGeom tg; tg.nlines = nc; tg.line_stride = 1; tg.elem_stride = G.C; tg.line_fastest = 1;
    Cx<float>* Z = line_fft<+1, float>(bufA, bufB, tg, ax, tid, nth);
    const int toto = (int)G.nt * nc;
    for (int w = tid; w < toto; w += nth) {
        const int n = w / nc, c = w - n * nc;
        const Cx<float> z = Z[n * G.C + c];
        const long long o = (long long)n * G.ntr + tr0 + 2 * c;
        x[o] = z.x;
        if (2 * c + 1 < ntr_tile) x[o + 1] = z.y;
    }
}


// ------------------------------------------------------------------------------------------------
// register-resident versions for the usual record lengths (same plans as the POCS column kernel):
// thread (c, j) holds samples j + e*T of the packed trace pair c.  Requires an even number of
// traces (8-byte loads of a trace pair, 16-byte stores of its two spectra).
// ------------------------------------------------------------------------------------------------
template <typename LP, int C, int MINB>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_time_fwd_spec(const __grid_constant__ TimeGeom G, const Cx<float>* __restrict__ tw, const float* __restrict__ x,
                Cx<float>* __restrict__ F, const Cx<float>* __restrict__ phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const long long tr = (long long)blockIdx.x * 2 * C + 2 * c;          // first trace of this pair
    const bool ok = tr < G.ntr;                                          // ntr is even: the pair is complete
    ColAcc<float, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + c;

    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int n = j + e * T;
        v[e] = cmake<float>(0.f, 0.f);
        if (ok && n < G.nt) {
            const float2 p = *reinterpret_cast<const float2*>(x + (long long)n * G.ntr + tr);
            v[e] = cmake<float>(p.x, p.y);
        }
    }
    LP::template fft<-1, 0, float>(v, acc, j, tw);

    // one more exchange: Z[N-k] lives in another thread
    Cx<float>* buf = acc.line(LP::NEXCH & 1);
#pragma unroll
    for (int e = 0; e < E; ++e) buf[(j + e * T) * C] = v[e];
    __syncthreads();
    if (!ok) return;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = j + e * T;
        if (k < G.nf) {
            const int km = (k == 0) ? 0 : N - k;
            const Cx<float> z1 = v[e], z2 = buf[km * C];
            const Cx<float> xa = cmake<float>(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
            const Cx<float> xb = cmake<float>(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
            const Cx<float> ph = phase[k];
            const Cx<float> fa = cmul(xa, ph), fb = cmul(xb, ph);
            *reinterpret_cast<float4*>(F + (long long)k * G.ntr + tr) = make_float4(fa.x, fa.y, fb.x, fb.y);
        }
    }
}

template <typename LP, int C, int MINB>
__global__ void __launch_bounds__(LP::T* C, MINB)
k_time_inv_spec(const __grid_constant__ TimeGeom G, const Cx<float>* __restrict__ tw, const Cx<float>* __restrict__ F,
                float* __restrict__ x, const Cx<float>* __restrict__ phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    constexpr int half = N / 2;
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    const long long tr = (long long)blockIdx.x * 2 * C + 2 * c;
    const bool ok = tr < G.ntr;
    ColAcc<float, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + c;
    Cx<float>* bufA = acc.line(0);
    Cx<float>* bufB = acc.line(1);

    // stage G_a * phase and G_b * phase by FFT bin (both buffers, natural order)
    Cx<float> ga[E], gb[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = j + e * T;
        ga[e] = cmake<float>(0.f, 0.f); gb[e] = cmake<float>(0.f, 0.f);
        const bool have = G.compute_real ? (k <= half) : true;
        if (ok && have) {
            long long row = k;
            if (!G.compute_real && G.ascending) row = (k + half) % N;
            const float4 p = *reinterpret_cast<const float4*>(F + row * G.ntr + tr);
            const Cx<float> ph = phase[k];
            ga[e] = cmul(cmake<float>(p.x, p.y), ph);
            gb[e] = cmul(cmake<float>(p.z, p.w), ph);
        }
        bufA[k * C] = ga[e];
        bufB[k * C] = gb[e];
    }
    __syncthreads();
    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = j + e * T;
        const int km = (k == 0) ? 0 : N - k;
        const Cx<float> a2 = bufA[km * C], b2 = bufB[km * C];
        Cx<float> ha, hb;
        if (G.compute_real) {
            if (k <= half) { ha = ga[e]; hb = gb[e]; if (k == 0 || k == half) { ha.y = 0.f; hb.y = 0.f; } }
            else           { ha = cmake<float>(a2.x, -a2.y); hb = cmake<float>(b2.x, -b2.y); }
        } else {
            ha = cmake<float>(0.5f * (ga[e].x + a2.x), 0.5f * (ga[e].y - a2.y));
            hb = cmake<float>(0.5f * (gb[e].x + b2.x), 0.5f * (gb[e].y - b2.y));
        }
        v[e] = cmake<float>(ha.x - hb.y, ha.y + hb.x);          // H[k] = ha + i hb
    }
    __syncthreads();                                             // the transform reuses both buffers
    LP::template fft<+1, 0, float>(v, acc, j, tw);
    if (!ok) return;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int n = j + e * T;
        if (n < G.nt) *reinterpret_cast<float2*>(x + (long long)n * G.ntr + tr) = make_float2(v[e].x, v[e].y);
    }
}


// ------------------------------------------------------------------------------------------------
// One-pass kernels with TMA-staged tiles (the default for the register-plan record lengths).
// A persistent CTA walks over tiles of 2C adjacent traces.  The tile is copied into a shared-memory
// stage by the TMA unit ([rows x 8C bytes] boxes, cp.async.bulk.tensor, completion on an mbarrier);
// as soon as every thread has picked its samples out of the stage, one thread asks for the NEXT
// tile, so the strided column load - thousands of 32-byte row segments a few MB apart, each of
// which would hold a whole 128-byte L1 line while in flight if it were a plain load - neither
// passes through L1 nor stalls the transform; rows beyond nt (zero padding to nfft) and traces
// beyond the cube are zero-filled by the unit.  One stage and ONE exchange buffer (ColAcc1): what
// is left of the 256 KB beside them is the L1 that has to hold the twiddle and phase tables (with
// two stages it did not: 13 % L1 hit rate, the warps waiting on table loads from L2).
// ------------------------------------------------------------------------------------------------
// TMA = false: the same pipeline with per-thread 16-byte cp.async.cg copies (zero-filled outside the cube) instead of
// the tensor unit, which spends a fixed time per box ROW - for 2048-sample records and 32-byte rows that is longer than
// the transform takes.
__device__ __forceinline__ void cp_async_16(unsigned dst, const void* src, bool valid) {
    const unsigned sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

template <typename LP, int C, bool TMA>
__global__ void __launch_bounds__(LP::T* C, 1)
k_time_fwd_tma(const __grid_constant__ TimeGeom G, const __grid_constant__ CUtensorMap tmap, const Cx<float>* __restrict__ tw,
               const float* __restrict__ x, Cx<float>* __restrict__ F, const Cx<float>* __restrict__ phase, const int ntiles,
               const int box_rows, const int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned long long bar;
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);      // TMA destinations: 128-byte aligned
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int STAGE = stage_rows * C;                                       // trace pairs (8 bytes) in the stage (rows >= N: whole boxes)
    float2* stage = reinterpret_cast<float2*>(smem_raw);                    // [stage_rows][C]
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    ColAcc1<float, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw + (size_t)STAGE * sizeof(float2)) + c;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int tile) {
        if constexpr (TMA) {
            if (tid == 0 && tile < ntiles) {
                mbar_expect_tx(&bar, (unsigned)(STAGE * sizeof(float2)));
                for (int r0 = 0; r0 < stage_rows; r0 += box_rows) tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C, r0, 0, &bar);
            }
        } else {
            if (tile < ntiles) {
                constexpr int CH = C / 2;                                   // 16-byte chunks (two trace pairs) per row
                const long long t0 = (long long)tile * 2 * C;
                const unsigned dst0 = smem_u32(stage);
                for (int q = tid; q < N * CH; q += T * C) {
                    const int row = q / CH, part = q - row * CH;
                    const bool valid = row < G.nt && t0 + 4 * part < G.ntr;
                    cp_async_16(dst0 + (unsigned)(row * C * 8 + part * 16), valid ? (const void*)(x + (long long)row * G.ntr + t0 + 4 * part) : (const void*)x, valid);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    int tile = blockIdx.x;
    issue(tile);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        // opaque copy of the row pitch: the 2 E output row addresses are recomputed per tile instead of being hoisted out of
        // the tile loop, where they would occupy 4 E registers for the whole kernel (spills)
        long long ntr = G.ntr;
        asm volatile("" : "+l"(ntr));
        const Cx<float>* ph_t = phase;                    // likewise the E phase factors (loop-invariant loads)
        asm volatile("" : "+l"(ph_t));
        if constexpr (TMA) mbar_wait(&bar, (unsigned)it & 1u);
        else { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncthreads(); }
        const float2* st = stage + c;
        Cx<float> v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { const float2 p = st[(j + e * T) * C]; v[e] = cmake<float>(p.x, p.y); }
        __syncthreads();                                  // the stage is free: the next tile arrives while this one is transformed
        issue(tile + gridDim.x);
        LP::template fft<-1, 0, float>(v, acc, j, tw);
        // one more exchange: Z[N-k] lives in another thread
        Cx<float>* buf = acc.line(0);
        acc.pre_sync();
#pragma unroll
        for (int e = 0; e < E; ++e) buf[(j + e * T) * C] = v[e];
        __syncthreads();
        const long long tr = (long long)tile * 2 * C + 2 * c;
        if (tr < G.ntr) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int k = j + e * T;
                if (k < G.nf) {
                    const int km = (k == 0) ? 0 : N - k;
                    const Cx<float> z1 = v[e], z2 = buf[km * C];
                    const Cx<float> xa = cmake<float>(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
                    const Cx<float> xb = cmake<float>(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
                    const Cx<float> ph = ph_t[k];
                    const Cx<float> fa = cmul(xa, ph), fb = cmul(xb, ph);
                    __stcs(reinterpret_cast<float4*>(F + (long long)k * ntr + tr), make_float4(fa.x, fa.y, fb.x, fb.y));
                }
            }
        }
    }
}

// inverse: the stage holds the spectra of the tile, [rows][C] pairs of complex64 (16 bytes), rows = nf (padded to whole
// boxes); the Hermitian partner of a bin is read from the stage as well, so no exchange buffer is needed before the transform
template <typename LP, int C, bool TMA>
__global__ void __launch_bounds__(LP::T* C, 1)
k_time_inv_tma(const __grid_constant__ TimeGeom G, const __grid_constant__ CUtensorMap tmap, const Cx<float>* __restrict__ tw,
               const Cx<float>* __restrict__ Fin, float* __restrict__ x, const Cx<float>* __restrict__ phase, const int ntiles,
               const int box_rows, const int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned long long bar;
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    constexpr int half = N / 2;
    const int STAGE = stage_rows * C;                                       // pairs of complex64 (16 bytes) in the stage
    float4* stage = reinterpret_cast<float4*>(smem_raw);                    // [stage_rows][C]
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    ColAcc1<float, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw + (size_t)STAGE * sizeof(float4)) + c;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int tile) {
        if constexpr (TMA) {
            if (tid == 0 && tile < ntiles) {
                mbar_expect_tx(&bar, (unsigned)(STAGE * sizeof(float4)));
                for (int r0 = 0; r0 < stage_rows; r0 += box_rows) tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C * 2, r0, 0, &bar);
            }
        } else {
            if (tile < ntiles) {
                const long long t0 = (long long)tile * 2 * C;
                const unsigned dst0 = smem_u32(stage);
                const int nrows = (int)G.nf;
                for (int q = tid; q < nrows * C; q += T * C) {              // one 16-byte chunk = the two spectra of a trace pair
                    const int row = q / C, part = q - row * C;
                    const bool valid = t0 + 2 * part < G.ntr;
                    cp_async_16(dst0 + (unsigned)(q * 16), valid ? (const void*)(Fin + (long long)row * G.ntr + t0 + 2 * part) : (const void*)Fin, valid);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    const bool real = G.compute_real != 0;
    const bool shifted = !real && G.ascending;
    int tile = blockIdx.x;
    issue(tile);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        // opaque copy of the row pitch: the 2 E output row addresses are recomputed per tile instead of being hoisted out of
        // the tile loop, where they would occupy 4 E registers for the whole kernel (spills)
        long long ntr = G.ntr;
        asm volatile("" : "+l"(ntr));
        const Cx<float>* ph_t = phase;                    // likewise the E phase factors (loop-invariant loads)
        asm volatile("" : "+l"(ph_t));
        if constexpr (TMA) mbar_wait(&bar, (unsigned)it & 1u);
        else { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncthreads(); }
        const float4* st = stage + c;
        Cx<float> v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = j + e * T;
            const int km = (k == 0) ? 0 : N - k;
            Cx<float> ha, hb;
            if (real) {
                // irfft semantics: bins 0..N/2 given, the rest is their conjugate mirror; imaginary part of DC / Nyquist ignored
                const int kk = (k <= half) ? k : km;
                const float4 p = st[kk * C];
                const Cx<float> ph = ph_t[kk];
                ha = cmul(cmake<float>(p.x, p.y), ph); hb = cmul(cmake<float>(p.z, p.w), ph);
                if (k > half) { ha.y = -ha.y; hb.y = -hb.y; }
                if (k == 0 || k == half) { ha.y = 0.f; hb.y = 0.f; }
            } else {
                const int r1 = shifted ? (k + half) % N : k, r2 = shifted ? (km + half) % N : km;
                const float4 p = st[r1 * C], q = st[r2 * C];
                const Cx<float> ph1 = ph_t[k], ph2 = ph_t[km];
                const Cx<float> ga = cmul(cmake<float>(p.x, p.y), ph1), gb = cmul(cmake<float>(p.z, p.w), ph1);
                const Cx<float> a2 = cmul(cmake<float>(q.x, q.y), ph2), b2 = cmul(cmake<float>(q.z, q.w), ph2);
                ha = cmake<float>(0.5f * (ga.x + a2.x), 0.5f * (ga.y - a2.y));
                hb = cmake<float>(0.5f * (gb.x + b2.x), 0.5f * (gb.y - b2.y));
            }
            v[e] = cmake<float>(ha.x - hb.y, ha.y + hb.x);          // H[k] = ha + i hb
        }
        __syncthreads();                                  // the stage is free: the next tile arrives while this one is transformed
        issue(tile + gridDim.x);
        LP::template fft<+1, 0, float>(v, acc, j, tw);
        const long long tr = (long long)tile * 2 * C + 2 * c;
        if (tr < G.ntr) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int n = j + e * T;
                if (n < G.nt) __stcs(reinterpret_cast<float2*>(x + (long long)n * ntr + tr), make_float2(v[e].x, v[e].y));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One-pass kernels for the long records (2048 / 4096 / 2000 / 4000 samples): N = 2 H, two H-point
// register transforms per line and one radix-2 step between them.  A thread then holds only H / T
// values at a time, so a CTA of the same size covers TWICE as many traces as the N-point plan
// could - 16 traces of 2048 samples: 64-byte rows on the way in, 128-byte rows on the way out -,
// and the row width of the strided tile is what the time-axis kernels are bound by (8 / 16 / 32 /
// 64 / 128 bytes per row: 0.7 / 1.2 / 2.0 / 3.2 / 4.5 TB/s, profiles/r2_time_axis_tile_width.txt).
//   forward (decimation in time):  Ze = FFT_H(z[2m]), Zo = FFT_H(z[2m+1]);  Z[k] = Ze[k] + w^k Zo[k],
//                                  Z[k+H] = Ze[k] - w^k Zo[k],  w = exp(-2 pi i / N)
//   inverse (decimation in frequency):  x[2m] = IFFT_H(P[k] + P[k+H]),  x[2m+1] = IFFT_H((P[k] - P[k+H]) conj(w^k))
// wsplit: w^k, k < H.
// ------------------------------------------------------------------------------------------------
template <typename LPH, int C, bool PAR>
__global__ void __launch_bounds__(LPH::T* C, 1)
k_time_fwd_split(const __grid_constant__ TimeGeom G, const __grid_constant__ CUtensorMap tmap, const Cx<float>* __restrict__ tw,
                 const Cx<float>* __restrict__ wsplit, Cx<float>* __restrict__ F, const Cx<float>* __restrict__ phase,
                 const int ntiles, const int box_rows, const int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned long long bar;
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    constexpr int E = LPH::E, T = LPH::T, H = LPH::N;
    // PAR (tiles of 64-byte rows): the tensor unit walks the box with a row step of 2, so the even samples land in rows
    // [0, stage_rows) and the odd samples in rows [stage_rows, 2 stage_rows) (stage_rows >= H: whole boxes) and a warp
    // reads 256 consecutive bytes of either half - with the samples interleaved as in the cube, the two rows a half-warp
    // touches lie 128 bytes apart, a 2-way bank conflict on a quarter of the kernel's wavefronts (2.86 -> 3.04 TB/s at
    // 2048 samples).  Tiles of 32-byte rows (4096 / 4000 samples) lose more to the doubled number of boxes: interleaved.
    const int STAGE = (PAR ? 2 : 1) * stage_rows * C;
    constexpr int ev_mul = PAR ? 1 : 2;                                     // even sample m: row ev_mul m, odd sample: od_off rows on
    const int od_off = PAR ? stage_rows : 1;
    float2* stage = reinterpret_cast<float2*>(smem_raw);                    // trace pairs
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    ColAcc1<float, C, LPH::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw + (size_t)STAGE * sizeof(float2)) + c;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int tile) {
        if (tid == 0 && tile < ntiles) {
            mbar_expect_tx(&bar, (unsigned)(STAGE * sizeof(float2)));
            for (int r0 = 0; r0 < stage_rows; r0 += box_rows) {
                if (PAR) {
                    tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C, 2 * r0, 0, &bar);
                    tma_load_3d(stage + (size_t)(stage_rows + r0) * C, &tmap, tile * C, 2 * r0 + 1, 0, &bar);
                } else {
                    tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C, r0, 0, &bar);
                }
            }
        }
    };
    const bool two_sided = G.nf > H + 1;
    int tile = blockIdx.x;
    issue(tile);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        // opaque copy of the row pitch: the 2 E output row addresses are recomputed per tile instead of being hoisted out of
        // the tile loop, where they would occupy 4 E registers for the whole kernel (spills)
        long long ntr = G.ntr;
        asm volatile("" : "+l"(ntr));
        const Cx<float>* ph_t = phase;                    // likewise the E phase factors (loop-invariant loads)
        asm volatile("" : "+l"(ph_t));
        mbar_wait(&bar, (unsigned)it & 1u);
        float2* st = stage + c;
        const Cx<float>* ws_t = wsplit;
        asm volatile("" : "+l"(ws_t));
        Cx<float> lo[E], hi[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { const float2 p = st[(ev_mul * (j + e * T)) * C]; lo[e] = cmake<float>(p.x, p.y); }
        LPH::template fft<-1, 0, float>(lo, acc, j, tw);
        // park Ze[k] in the slots this thread has just emptied (the even sample k of its own column)
#pragma unroll
        for (int e = 0; e < E; ++e) st[(ev_mul * (j + e * T)) * C] = make_float2(lo[e].x, lo[e].y);
#pragma unroll
        for (int e = 0; e < E; ++e) { const float2 p = st[(ev_mul * (j + e * T) + od_off) * C]; hi[e] = cmake<float>(p.x, p.y); }
        LPH::template fft<-1, 0, float>(hi, acc, j, tw);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = j + e * T;
            const float2 p = st[(ev_mul * k) * C];
            const Cx<float> t = cmul(hi[e], ws_t[k]);
            lo[e] = cmake<float>(p.x + t.x, p.y + t.y);          // Z[k]
            hi[e] = cmake<float>(p.x - t.x, p.y - t.y);          // Z[k + H]
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the parked values were written through the generic proxy; the refill is not
        __syncthreads();                                  // the stage is free: the next tile arrives during the epilogue
        issue(tile + gridDim.x);
        // separate the two real traces: rows k <= H need Z[N - k] = Z[(H - k) + H], the upper value of another thread
        Cx<float>* buf = acc.line(0);
#pragma unroll
        for (int e = 0; e < E; ++e) buf[(j + e * T) * C] = hi[e];
        __syncthreads();
        const long long tr = (long long)tile * 2 * C + 2 * c;
        const bool ok = tr < G.ntr;
        auto emit = [&](int row, Cx<float> z1, Cx<float> z2) {
            const Cx<float> xa = cmake<float>(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
            const Cx<float> xb = cmake<float>(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
            const Cx<float> ph = ph_t[row];
            const Cx<float> fa = cmul(xa, ph), fb = cmul(xb, ph);
            __stcs(reinterpret_cast<float4*>(F + (long long)row * ntr + tr), make_float4(fa.x, fa.y, fb.x, fb.y));
        };
        if (ok) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int k = j + e * T;
                if (k == 0) { emit(0, lo[e], lo[e]); emit(H, hi[e], hi[e]); }
                else emit(k, lo[e], buf[(H - k) * C]);
            }
        }
        if (two_sided) {
            // rows k + H (k >= 1) need Z[N - k - H] = Z[H - k], the lower value of another thread
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; ++e) buf[(j + e * T) * C] = lo[e];
            __syncthreads();
            if (ok) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int k = j + e * T;
                    if (k > 0) emit(k + H, hi[e], buf[(H - k) * C]);
                }
            }
        }
    }
}

// inverse of a one-sided spectrum (compute_real): the stage holds rows 0 .. H of the tile, [row][C] pairs of complex64
template <typename LPH, int C>
__global__ void __launch_bounds__(LPH::T* C, 1)
k_time_inv_split(const __grid_constant__ TimeGeom G, const __grid_constant__ CUtensorMap tmap, const Cx<float>* __restrict__ tw,
                 const Cx<float>* __restrict__ wsplit, float* __restrict__ x, const Cx<float>* __restrict__ phase,
                 const int ntiles, const int box_rows, const int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned long long bar;
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    constexpr int E = LPH::E, T = LPH::T, H = LPH::N;
    const int STAGE = stage_rows * C;
    float4* stage = reinterpret_cast<float4*>(smem_raw);                    // [stage_rows >= H + 1][C]
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    ColAcc1<float, C, LPH::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw + (size_t)STAGE * sizeof(float4)) + c;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int tile) {
        if (tid == 0 && tile < ntiles) {
            mbar_expect_tx(&bar, (unsigned)(STAGE * sizeof(float4)));
            for (int r0 = 0; r0 < stage_rows; r0 += box_rows) tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C * 2, r0, 0, &bar);
        }
    };
    int tile = blockIdx.x;
    issue(tile);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        // opaque copy of the row pitch: the 2 E output row addresses are recomputed per tile instead of being hoisted out of
        // the tile loop, where they would occupy 4 E registers for the whole kernel (spills)
        long long ntr = G.ntr;
        asm volatile("" : "+l"(ntr));
        const Cx<float>* ph_t = phase;                    // likewise the E phase factors (loop-invariant loads)
        asm volatile("" : "+l"(ph_t));
        mbar_wait(&bar, (unsigned)it & 1u);
        const float4* st = stage + c;
        const Cx<float>* ws_t = wsplit;
        asm volatile("" : "+l"(ws_t));
        // packed Hermitian spectrum P = ha + i hb of the trace pair at bins k and k + H (k < H):
        //   P[k]     from row k (imaginary parts of DC ignored),
        //   P[k + H] from row H (k = 0, imaginary parts of Nyquist ignored) or the conjugate mirror of row H - k
        auto bins = [&](int k, Cx<float>& p0, Cx<float>& p1) {
            const float4 a = st[k * C];
            const Cx<float> ph = ph_t[k];
            Cx<float> ha = cmul(cmake<float>(a.x, a.y), ph), hb = cmul(cmake<float>(a.z, a.w), ph);
            if (k == 0) { ha.y = 0.f; hb.y = 0.f; }
            p0 = cmake<float>(ha.x - hb.y, ha.y + hb.x);
            const int km = H - k;                                   // k = 0: row H itself
            const float4 b = st[km * C];
            const Cx<float> pm = ph_t[km];
            Cx<float> ma = cmul(cmake<float>(b.x, b.y), pm), mb = cmul(cmake<float>(b.z, b.w), pm);
            if (k == 0) { ma.y = 0.f; mb.y = 0.f; } else { ma.y = -ma.y; mb.y = -mb.y; }
            p1 = cmake<float>(ma.x - mb.y, ma.y + mb.x);
        };
        const long long tr = (long long)tile * 2 * C + 2 * c;
        const bool ok = tr < G.ntr;
        Cx<float> v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            Cx<float> p0, p1; bins(j + e * T, p0, p1); v[e] = cmake<float>(p0.x + p1.x, p0.y + p1.y);
            if (e % 4 == 3) __syncwarp();          // scheduling fence: the loads of at most four bins in flight (registers)
        }
        LPH::template fft<+1, 0, float>(v, acc, j, tw);
        if (ok) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int n = 2 * (j + e * T);
                if (n < G.nt) __stcs(reinterpret_cast<float2*>(x + (long long)n * ntr + tr), make_float2(v[e].x, v[e].y));
            }
        }
        asm volatile("" : "+l"(ph_t));                    // second pass: reload instead of keeping 2 E factors across the transform
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = j + e * T;
            Cx<float> p0, p1; bins(k, p0, p1);
            v[e] = cmulc(cmake<float>(p0.x - p1.x, p0.y - p1.y), ws_t[k]);
            if (e % 4 == 3) __syncwarp();
        }
        __syncthreads();                                  // the stage is free: the next tile arrives during the second transform
        issue(tile + gridDim.x);
        LPH::template fft<+1, 0, float>(v, acc, j, tw);
        if (ok) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int n = 2 * (j + e * T) + 1;
                if (n < G.nt) __stcs(reinterpret_cast<float2*>(x + (long long)n * ntr + tr), make_float2(v[e].x, v[e].y));
            }
        }
    }
}

typedef LinePlan<512, 16, 16, 16, 2> TP512;
typedef LinePlan<1024, 16, 16, 16, 4> TP1024;
typedef LinePlan<2048, 16, 16, 16, 8> TP2048;
typedef LinePlan<4096, 16, 16, 16, 16> TP4096;
// 10-smooth record lengths (1 / 2 / 4 ms sampling of 1 - 5 s records)
typedef LinePlan<1000, 10, 10, 10, 10> TP1000;
typedef LinePlan<2000, 20, 20, 10, 10> TP2000;
typedef LinePlan<2500, 10, 5, 5, 10, 10> TP2500;
typedef LinePlan<3000, 30, 10, 10, 30> TP3000;
typedef LinePlan<4000, 20, 20, 20, 10> TP4000;
typedef LinePlan<5000, 10, 10, 10, 10, 5> TP5000;

// ------------------------------------------------------------------------------------------------
// Transposing pipeline (takes what the one-pass and direct kernels decline: odd trace counts,
// unaligned buffers, 3000 / 5000 samples).  A chunk of traces is (1) transposed to trace-major
// with a tiled transpose that streams a few dozen rows at a time, (2) transformed along the now
// contiguous axis with the register-resident FFT (two real traces per complex line), (3)
// transposed back into the slice-major spectrum.  The chunk is sized so that both intermediates
// stay in the 126 MB L2.  1.3 TB/s algorithmic for every record length (three kernels, 36 bytes
// of L2 traffic per sample against 12 for one pass).
// ------------------------------------------------------------------------------------------------
// out[c][r] = in[r][c]; 32 x 32 tiles, 32 x 8 threads.  ROWS_FAST: consecutive CTAs walk along the
// rows of `in` (else along its columns) -- always chosen so that consecutive CTAs walk along the
// TRACE axis and the GPU streams only a few dozen long time/frequency rows of the cube at a time.
template <typename T, bool ROWS_FAST>
__global__ void k_transpose(const T* __restrict__ in, T* __restrict__ out, long long rows, long long cols,
                            long long in_ld, long long out_ld) {
    __shared__ T tile[32][33];
    const long long c0 = (long long)(ROWS_FAST ? blockIdx.y : blockIdx.x) * 32;
    const long long r0 = (long long)(ROWS_FAST ? blockIdx.x : blockIdx.y) * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const long long r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + i][tx] = in[r * in_ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const long long c = c0 + ty + i, r = r0 + tx;
        if (c < cols && r < rows) out[c * out_ld + r] = tile[tx][ty + i];
    }
}

// trace-major forward: xt (ntr_chunk, nt) float -> Ft (ntr_chunk, nf) complex; RB trace PAIRS per CTA
template <typename LP, int RB>
__global__ void __launch_bounds__(LP::T* RB, 2)
k_time_fwd_rows(const __grid_constant__ TimeGeom G, const Cx<float>* __restrict__ tw, const float* __restrict__ xt,
                Cx<float>* __restrict__ Ft, const Cx<float>* __restrict__ phase, const long long ntr_chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const long long ta = ((long long)blockIdx.x * RB + rr) * 2, tb = ta + 1;
    const bool oka = ta < ntr_chunk, okb = tb < ntr_chunk;
    RowAcc<float, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + rr * LP::LINE;
    const float* __restrict__ pa = xt + ta * G.nt + j;
    const float* __restrict__ pb = xt + tb * G.nt + j;
    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int n = e * T;
        const bool in = j + n < G.nt;
        v[e] = cmake<float>((oka && in) ? pa[n] : 0.f, (okb && in) ? pb[n] : 0.f);
    }
    LP::template fft<-1, 0, float>(v, acc, j, tw);
    Cx<float>* buf = acc.line(LP::NEXCH & 1);
#pragma unroll
    for (int e = 0; e < E; ++e) buf[j + e * T] = v[e];
    __syncthreads();
    Cx<float>* __restrict__ oa = Ft + ta * G.nf;
    Cx<float>* __restrict__ ob = Ft + tb * G.nf;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = j + e * T;
        if (k < G.nf) {
            const Cx<float> z1 = v[e], z2 = buf[(k == 0) ? 0 : N - k];
            const Cx<float> xa = cmake<float>(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
            const Cx<float> xb = cmake<float>(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
            const Cx<float> ph = phase[k];
            if (oka) oa[k] = cmul(xa, ph);
            if (okb) ob[k] = cmul(xb, ph);
        }
    }
}

// trace-major inverse: Ft (ntr_chunk, nf) complex (rows already in FFT-bin order k = 0..nbins-1) -> xt (ntr_chunk, nt)
template <typename LP, int RB>
__global__ void __launch_bounds__(LP::T* RB, 2)
k_time_inv_rows(const __grid_constant__ TimeGeom G, const Cx<float>* __restrict__ tw, const Cx<float>* __restrict__ Ft,
                float* __restrict__ xt, const Cx<float>* __restrict__ phase, const long long ntr_chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    constexpr int half = N / 2;
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const long long ta = ((long long)blockIdx.x * RB + rr) * 2, tb = ta + 1;
    const bool oka = ta < ntr_chunk, okb = tb < ntr_chunk;
    RowAcc<float, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + rr * LP::LINE;
    Cx<float>* bufA = acc.line(0);
    Cx<float>* bufB = acc.line(1);
    const Cx<float>* __restrict__ ia = Ft + ta * G.nf;
    const Cx<float>* __restrict__ ib = Ft + tb * G.nf;
    Cx<float> ga[E], gb[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = j + e * T;
        ga[e] = cmake<float>(0.f, 0.f); gb[e] = cmake<float>(0.f, 0.f);
        const bool have = G.compute_real ? (k <= half) : true;
        if (have) {
            int row = k;
            if (!G.compute_real && G.ascending) row = (k + half) % N;
            const Cx<float> ph = phase[k];
            if (oka) ga[e] = cmul(ia[row], ph);
            if (okb) gb[e] = cmul(ib[row], ph);
        }
        bufA[k] = ga[e];
        bufB[k] = gb[e];
    }
    __syncthreads();
    Cx<float> v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int k = j + e * T;
        const int km = (k == 0) ? 0 : N - k;
        const Cx<float> a2 = bufA[km], b2 = bufB[km];
        Cx<float> ha, hb;
        if (G.compute_real) {
            if (k <= half) { ha = ga[e]; hb = gb[e]; if (k == 0 || k == half) { ha.y = 0.f; hb.y = 0.f; } }
            else           { ha = cmake<float>(a2.x, -a2.y); hb = cmake<float>(b2.x, -b2.y); }
        } else {
            ha = cmake<float>(0.5f * (ga[e].x + a2.x), 0.5f * (ga[e].y - a2.y));
            hb = cmake<float>(0.5f * (gb[e].x + b2.x), 0.5f * (gb[e].y - b2.y));
        }
        v[e] = cmake<float>(ha.x - hb.y, ha.y + hb.x);
    }
    __syncthreads();
    LP::template fft<+1, 0, float>(v, acc, j, tw);
    float* __restrict__ oa = xt + ta * G.nt + j;
    float* __restrict__ ob = xt + tb * G.nt + j;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int n = e * T;
        if (j + n < G.nt) { if (oka) oa[n] = v[e].x; if (okb) ob[n] = v[e].y; }
    }
}

// ------------------------------------------------------------------------------------------------
// Amplitude envelope |x + i H[x]| along the time axis (functions/signal.py:672-690: scipy.signal.hilbert,
// abs, cast to the input dtype).  Two real traces are packed into one complex line z = a + i b; since
// H[a], H[b] are real, ONE inverse transform of  -i sgn(k) FFT(z)[k] / N  returns H[a] + i H[b]
// (sgn = +1 below Nyquist, -1 above, 0 at DC and Nyquist: scipy's one-sided weights 1, 2, ..., 2, 1, 0, ...).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Cx<float> hilbert_weight(Cx<float> z, int k, int N, float inv_n) {
    // -i * sgn(k) * z / N
    const bool zero = (k == 0) || (2 * k == N);
    const float sg = zero ? 0.f : ((2 * k < N) ? inv_n : -inv_n);
    return cmake<float>(sg * z.y, -sg * z.x);
}

// generic direct kernel (any record length): x (nt, ntr) -> env (nt, ntr)
__global__ void k_time_env(const __grid_constant__ TimeGeom G, const __grid_constant__ AxisDev<float> ax,
                           const float* __restrict__ x, float* __restrict__ env) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nth = blockDim.x;
    Cx<float>* bufA = reinterpret_cast<Cx<float>*>(smem_raw);
    Cx<float>* bufB = bufA + (size_t)ax.L * G.C;
    const long long tr0 = (long long)blockIdx.x * 2 * G.C;
    const int ntr_tile = (int)min((long long)2 * G.C, G.ntr - tr0);
    const int nc = (ntr_tile + 1) / 2;
    const int N = G.nfft;
    const int tot = N * nc;
    for (int w = tid; w < tot; w += nth) {
        const int n = w / nc, c = w - n * nc;
        const long long base = (long long)n * G.ntr + tr0 + 2 * c;
        const float a = x[base];
        const float b = (2 * c + 1 < ntr_tile) ? x[base + 1] : 0.f;
        bufA[n * G.C + c] = cmake<float>(a, b);
    }
    __syncthreads();
    TileGeom tg; tg.nlines = nc; tg.line_stride = 1; tg.elem_stride = G.C; tg.line_fastest = 1;
    Cx<float>* Z = line_fft<-1, float>(bufA, bufB, tg, ax, tid, nth);
    Cx<float>* other = (Z == bufA) ? bufB : bufA;
    const float inv_n = 1.f / (float)N;
    for (int w = tid; w < tot; w += nth) {
        const int k = w / nc, c = w - k * nc;
        Z[k * G.C + c] = hilbert_weight(Z[k * G.C + c], k, N, inv_n);
    }
    __syncthreads();
    Cx<float>* Hx = line_fft<+1, float>(Z, other, tg, ax, tid, nth);
    for (int w = tid; w < tot; w += nth) {
        const int n = w / nc, c = w - n * nc;
        const long long base = (long long)n * G.ntr + tr0 + 2 * c;
        const Cx<float> h = Hx[n * G.C + c];
        const float a = x[base];
        env[base] = sqrtf(a * a + h.x * h.x);
        if (2 * c + 1 < ntr_tile) { const float b = x[base + 1]; env[base + 1] = sqrtf(b * b + h.y * h.y); }
    }
}

// trace-major register-resident version: xt (ntr_chunk, nt) -> et (ntr_chunk, nt); RB trace PAIRS per CTA
template <typename LP, int RB>
__global__ void __launch_bounds__(LP::T* RB, 2)
k_time_env_rows(const __grid_constant__ TimeGeom G, const Cx<float>* __restrict__ tw, const float* __restrict__ xt,
                float* __restrict__ et, const long long ntr_chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int tid = threadIdx.x;
    const int j = tid % T, rr = tid / T;
    const long long ta = ((long long)blockIdx.x * RB + rr) * 2, tb = ta + 1;
    const bool oka = ta < ntr_chunk, okb = tb < ntr_chunk;
    RowAcc<float, RB, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw) + rr * LP::LINE;
    const float* __restrict__ pa = xt + ta * N + j;
    const float* __restrict__ pb = xt + tb * N + j;
    Cx<float> v[E], x0[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { x0[e] = cmake<float>(oka ? pa[e * T] : 0.f, okb ? pb[e * T] : 0.f); v[e] = x0[e]; }
    LP::template fft<-1, 0, float>(v, acc, j, tw);
    const float inv_n = 1.f / (float)N;
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = hilbert_weight(v[e], j + e * T, N, inv_n);
    LP::template fft<+1, (LP::NEXCH & 1), float>(v, acc, j, tw);
    float* __restrict__ oa = et + ta * N + j;
    float* __restrict__ ob = et + tb * N + j;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if (oka) oa[e * T] = sqrtf(x0[e].x * x0[e].x + v[e].x * v[e].x);
        if (okb) ob[e * T] = sqrtf(x0[e].y * x0[e].y + v[e].y * v[e].y);
    }
}

// one-pass envelope: the same TMA-staged tiles; forward transform, Hilbert weights, inverse transform, and the magnitude
// against the samples still sitting in the stage (so the next tile is requested only after the tile is finished)
template <typename LP, int C>
__global__ void __launch_bounds__(LP::T* C, 1)
k_time_env_tma(const __grid_constant__ TimeGeom G, const __grid_constant__ CUtensorMap tmap, const Cx<float>* __restrict__ tw,
               float* __restrict__ env, const int ntiles, const int box_rows, const int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned long long bar;
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    constexpr int E = LP::E, T = LP::T, N = LP::N;
    const int STAGE = stage_rows * C;
    float2* stage = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    ColAcc1<float, C, LP::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw + (size_t)STAGE * sizeof(float2)) + c;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int tile) {
        if (tid == 0 && tile < ntiles) {
            mbar_expect_tx(&bar, (unsigned)(STAGE * sizeof(float2)));
            for (int r0 = 0; r0 < stage_rows; r0 += box_rows) tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C, r0, 0, &bar);
        }
    };
    const float inv_n = 1.f / (float)N;
    int tile = blockIdx.x;
    issue(tile);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        mbar_wait(&bar, (unsigned)it & 1u);
        const float2* st = stage + c;
        Cx<float> v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { const float2 p = st[(j + e * T) * C]; v[e] = cmake<float>(p.x, p.y); }
        LP::template fft<-1, 0, float>(v, acc, j, tw);
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = hilbert_weight(v[e], j + e * T, N, inv_n);
        LP::template fft<+1, 0, float>(v, acc, j, tw);
        const long long tr = (long long)tile * 2 * C + 2 * c;
        if (tr < G.ntr) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int n = j + e * T;
                const float2 p = st[n * C];
                __stcs(reinterpret_cast<float2*>(env + (long long)n * G.ntr + tr),
                       make_float2(sqrtf(p.x * p.x + v[e].x * v[e].x), sqrtf(p.y * p.y + v[e].y * v[e].y)));
            }
        }
        __syncthreads();
        issue(tile + gridDim.x);
    }
}

// envelope of the long records with the same radix-2 split: forward as in k_time_fwd_split, the Hilbert weights on
// Z[k] and Z[k + H] (both held by the same thread), then the decimation-in-frequency inverse - no exchange between the
// two directions.  The samples stay in the stage for the magnitude, so the half-spectra wait in registers (and partly in
// local memory) instead of being parked there, and the next tile is requested only when this one is finished.
template <typename LPH, int C>
__global__ void __launch_bounds__(LPH::T* C, 1)
k_time_env_split(const __grid_constant__ TimeGeom G, const __grid_constant__ CUtensorMap tmap, const Cx<float>* __restrict__ tw,
                 const Cx<float>* __restrict__ wsplit, float* __restrict__ env, const int ntiles, const int box_rows, const int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    __shared__ unsigned long long bar;
    unsigned char* smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    constexpr int E = LPH::E, T = LPH::T, H = LPH::N;
    const int STAGE = stage_rows * C;
    float2* stage = reinterpret_cast<float2*>(smem_raw);
    const int tid = threadIdx.x;
    const int c = tid % C, j = tid / C;
    ColAcc1<float, C, LPH::LINE> acc; acc.base = reinterpret_cast<Cx<float>*>(smem_raw + (size_t)STAGE * sizeof(float2)) + c;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto issue = [&](int tile) {
        if (tid == 0 && tile < ntiles) {
            mbar_expect_tx(&bar, (unsigned)(STAGE * sizeof(float2)));
            for (int r0 = 0; r0 < stage_rows; r0 += box_rows) tma_load_3d(stage + (size_t)r0 * C, &tmap, tile * C, r0, 0, &bar);
        }
    };
    const float inv_n = 1.f / (float)(2 * H);
    int tile = blockIdx.x;
    issue(tile);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        long long ntr = G.ntr;
        asm volatile("" : "+l"(ntr));
        const Cx<float>* ws_t = wsplit;
        asm volatile("" : "+l"(ws_t));
        mbar_wait(&bar, (unsigned)it & 1u);
        const float2* st = stage + c;
        Cx<float> lo[E], hi[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { const float2 p = st[(2 * (j + e * T)) * C]; lo[e] = cmake<float>(p.x, p.y); }
        LPH::template fft<-1, 0, float>(lo, acc, j, tw);
#pragma unroll
        for (int e = 0; e < E; ++e) { const float2 p = st[(2 * (j + e * T) + 1) * C]; hi[e] = cmake<float>(p.x, p.y); }
        LPH::template fft<-1, 0, float>(hi, acc, j, tw);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int k = j + e * T;
            const Cx<float> w = ws_t[k];
            const Cx<float> t = cmul(hi[e], w);
            const Cx<float> z0 = cmake<float>(lo[e].x + t.x, lo[e].y + t.y);          // Z[k]
            const Cx<float> z1 = cmake<float>(lo[e].x - t.x, lo[e].y - t.y);          // Z[k + H]
            // -i sgn z / N: + below Nyquist (bin k), - above (bin k + H), 0 at DC and Nyquist (k = 0)
            const float sg = (k == 0) ? 0.f : inv_n;
            const Cx<float> p0 = cmake<float>(sg * z0.y, -sg * z0.x);
            const Cx<float> p1 = cmake<float>(-sg * z1.y, sg * z1.x);
            lo[e] = cmake<float>(p0.x + p1.x, p0.y + p1.y);
            hi[e] = cmulc(cmake<float>(p0.x - p1.x, p0.y - p1.y), w);
        }
        LPH::template fft<+1, 0, float>(lo, acc, j, tw);
        LPH::template fft<+1, 0, float>(hi, acc, j, tw);
        const long long tr = (long long)tile * 2 * C + 2 * c;
        if (tr < G.ntr) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int n = 2 * (j + e * T);
                const float2 p = st[n * C], q = st[(n + 1) * C];
                __stcs(reinterpret_cast<float2*>(env + (long long)n * ntr + tr),
                       make_float2(sqrtf(p.x * p.x + lo[e].x * lo[e].x), sqrtf(p.y * p.y + lo[e].y * lo[e].y)));
                __stcs(reinterpret_cast<float2*>(env + (long long)(n + 1) * ntr + tr),
                       make_float2(sqrtf(q.x * q.x + hi[e].x * hi[e].x), sqrtf(q.y * q.y + hi[e].y * hi[e].y)));
            }
        }
        __syncthreads();
        issue(tile + gridDim.x);
    }
}

// device time of the kernels of the last p3d_time_fft / p3d_time_ifft call of this thread
static thread_local double g_last_kernel_ms = 0.0;
// which kernels served it: "tma" (one pass, TMA-staged), "pipeline" (transpose / FFT / transpose), "direct", "generic"
static thread_local const char* g_last_path = "none";

// grow-only scratch kept between calls (per device): the two L2-sized intermediates
struct TimeScratch { int device = -1; float* tmp_t = nullptr; size_t n_t = 0; Cx<float>* tmp_f = nullptr; size_t n_f = 0; };
static TimeScratch g_scratch[16];
static void scratch_get(int device, size_t n_t, size_t n_f, float** tt, Cx<float>** tf) {
    TimeScratch& S = g_scratch[device & 15];
    if (S.n_t < n_t) { if (S.tmp_t) cudaFree(S.tmp_t); S.tmp_t = nullptr; S.n_t = 0; P3D_CUDA(cudaMalloc(&S.tmp_t, sizeof(float) * n_t)); S.n_t = n_t; }
    if (S.n_f < n_f) { if (S.tmp_f) cudaFree(S.tmp_f); S.tmp_f = nullptr; S.n_f = 0; P3D_CUDA(cudaMalloc(&S.tmp_f, sizeof(Cx<float>) * n_f)); S.n_f = n_f; }
    *tt = S.tmp_t; *tf = S.tmp_f;
}

template <typename LP, int RB>
bool launch_time_pipeline(const TimeGeom& G, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    constexpr size_t smem = (size_t)2 * LP::LINE * RB * sizeof(Cx<float>);
    if (smem > smem_optin - 1024) return false;
    std::vector<int> rad(LP::NPASS);
    LP::radices(rad.data());
    std::vector<Cx<float>> t = spec_twiddle_table(rad);
    // chunk of traces whose two intermediates (nt floats + nf complex per trace) stay L2 resident
    const size_t per_trace = (size_t)G.nt * sizeof(float) + (size_t)G.nf * sizeof(Cx<float>);
    long long chunk = (long long)((size_t)72 * 1024 * 1024 / per_trace);
    chunk = std::max<long long>(256, (chunk / 64) * 64);
    chunk = std::min<long long>(chunk, ((G.ntr + 1) / 2) * 2);
    Cx<float>* d_tw = nullptr; float* tmp_t = nullptr; Cx<float>* tmp_f = nullptr;
    struct Free { void* p; cudaEvent_t e0, e1; ~Free() { if (p) cudaFree(p); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); } } fr{nullptr, nullptr, nullptr};
    P3D_CUDA(cudaMalloc(&d_tw, sizeof(Cx<float>) * t.size())); fr.p = d_tw;
    int dev = 0; P3D_CUDA(cudaGetDevice(&dev));
    scratch_get(dev, (size_t)G.nt * chunk, (size_t)G.nf * chunk, &tmp_t, &tmp_f);
    P3D_CUDA(cudaMemcpy(d_tw, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
    P3D_CUDA(cudaEventCreate(&fr.e0)); P3D_CUDA(cudaEventCreate(&fr.e1));
    P3D_CUDA(cudaEventRecord(fr.e0, 0));
    P3D_CUDA(cudaFuncSetAttribute(k_time_fwd_rows<LP, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P3D_CUDA(cudaFuncSetAttribute(k_time_inv_rows<LP, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 tb(32, 8);
    for (long long c0 = 0; c0 < G.ntr; c0 += chunk) {
        const long long nc = std::min<long long>(chunk, G.ntr - c0);
        const unsigned ctas = (unsigned)((nc + 2 * RB - 1) / (2 * RB));
        if (!inverse) {
            const float* x = (const float*)din + c0;
            Cx<float>* F = (Cx<float>*)dout + c0;
            // (nt x nc) -> (nc x nt): columns of `in` are traces -> columns fast
            k_transpose<float, false><<<dim3((unsigned)((nc + 31) / 32), (unsigned)((G.nt + 31) / 32)), tb>>>(x, tmp_t, G.nt, nc, G.ntr, G.nt);
            k_time_fwd_rows<LP, RB><<<ctas, LP::T * RB, smem>>>(G, d_tw, tmp_t, tmp_f, d_ph, nc);
            // (nc x nf) -> (nf x nc): rows of `in` are traces -> rows fast
            k_transpose<Cx<float>, true><<<dim3((unsigned)((nc + 31) / 32), (unsigned)((G.nf + 31) / 32)), tb>>>(tmp_f, F, nc, G.nf, G.nf, G.ntr);
        } else {
            const Cx<float>* F = (const Cx<float>*)din + c0;
            float* x = (float*)dout + c0;
            k_transpose<Cx<float>, false><<<dim3((unsigned)((nc + 31) / 32), (unsigned)((G.nf + 31) / 32)), tb>>>(F, tmp_f, G.nf, nc, G.ntr, G.nf);
            k_time_inv_rows<LP, RB><<<ctas, LP::T * RB, smem>>>(G, d_tw, tmp_f, tmp_t, d_ph, nc);
            k_transpose<float, true><<<dim3((unsigned)((nc + 31) / 32), (unsigned)((G.nt + 31) / 32)), tb>>>(tmp_t, x, nc, G.nt, G.nt, G.ntr);
        }
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaEventRecord(fr.e1, 0));
    P3D_CUDA(cudaDeviceSynchronize());
    { float ms = 0.f; cudaEventElapsedTime(&ms, fr.e0, fr.e1); g_last_kernel_ms = ms; g_last_path = "pipeline"; }
    return true;
}

template <typename LP, int C, int MINB = 1>
bool launch_time_spec(const TimeGeom& G0, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    constexpr size_t smem = (size_t)2 * LP::LINE * C * sizeof(Cx<float>);
    if (smem > smem_optin - 1024) return false;
    std::vector<int> rad(LP::NPASS);
    LP::radices(rad.data());
    std::vector<Cx<float>> t = spec_twiddle_table(rad);
    Cx<float>* d_tw = nullptr;
    P3D_CUDA(cudaMalloc(&d_tw, sizeof(Cx<float>) * t.size()));
    struct Free { Cx<float>* p; ~Free() { cudaFree(p); } } fr{d_tw};
    P3D_CUDA(cudaMemcpy(d_tw, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
    TimeGeom G = G0; G.C = C;
    const long long tiles = (G.ntr + 2 * C - 1) / (2 * C);
    cudaEvent_t e0, e1; P3D_CUDA(cudaEventCreate(&e0)); P3D_CUDA(cudaEventCreate(&e1));
    struct FreeEv { cudaEvent_t a, b; ~FreeEv() { cudaEventDestroy(a); cudaEventDestroy(b); } } fe{e0, e1};
    P3D_CUDA(cudaEventRecord(e0, 0));
    if (!inverse) {
        P3D_CUDA(cudaFuncSetAttribute(k_time_fwd_spec<LP, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_time_fwd_spec<LP, C, MINB><<<(unsigned)tiles, LP::T * C, smem>>>(G, d_tw, (const float*)din, (Cx<float>*)dout, d_ph);
    } else {
        P3D_CUDA(cudaFuncSetAttribute(k_time_inv_spec<LP, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_time_inv_spec<LP, C, MINB><<<(unsigned)tiles, LP::T * C, smem>>>(G, d_tw, (const Cx<float>*)din, (float*)dout, d_ph);
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaEventRecord(e1, 0));
    P3D_CUDA(cudaDeviceSynchronize());
    { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); g_last_kernel_ms = ms; g_last_path = "direct"; }
    return true;
}

// TMA boxes of a [rows x row_bytes] stage: at most 256 rows each and a multiple of 128 bytes (destination alignment);
// the last box may reach beyond `rows` (zero-filled by the unit), so the stage holds stage_rows >= rows
static void pick_box(int rows, int row_bytes, int* box_rows, int* stage_rows, int cap = 256) {
    int align = 1;
    while ((align * row_bytes) % 128) align *= 2;
    for (int nbox = (rows + cap - 1) / cap;; ++nbox) {
        int br = (rows + nbox - 1) / nbox;
        br = (br + align - 1) / align * align;
        if (br <= cap) { *box_rows = br; *stage_rows = nbox * br; return; }
    }
}

// What the one-pass launchers share: the checks, the stage and its TMA boxes, the tensor map, the tables on the device,
// the grid (one persistent CTA per SM) and the CUDA events around the launch.
struct OnePass {
    TimeGeom G;
    long long tiles = 0;
    int box_rows = 0, stage_rows = 0;
    size_t smem = 0;
    unsigned grid = 0;
    CUtensorMap map;
    Cx<float>* d_tw = nullptr;            // twiddle tables of the plan [+ w^k, k < H, for the radix-2 split]
    const Cx<float>* d_wsplit = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    OnePass() { memset(&map, 0, sizeof(map)); }
    ~OnePass() { if (d_tw) cudaFree(d_tw); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }

    // stage of `rows` rows of C elements (elem_bytes: 8 = a pair of samples, 16 = a pair of spectral values) cut out of a
    // (map_rows, n_traces / 2) tensor at `src`; `line` = exchange-buffer elements per line.  false: this path declines.
    // parity: the even rows and the odd rows of the tile land in two halves of the stage (rows = rows per half)
    bool setup(const TimeGeom& G0, const void* src, const void* dst, int C, int rows, long long map_rows, int elem_bytes, int line,
               bool use_tma, size_t smem_optin, bool parity = false) {
        if (G0.ntr % 4 != 0) return false;                                   // 16-byte row pitch and whole trace pairs
        if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) return false;
        G = G0; G.C = C;
        tiles = (G.ntr + 2 * C - 1) / (2 * C);
        if (tiles > 2147483647LL / (4 * C)) return false;
        pick_box(rows, C * elem_bytes, &box_rows, &stage_rows, parity ? 128 : 256);
        if (!use_tma) stage_rows = rows;
        smem = (size_t)stage_rows * (parity ? 2 : 1) * C * elem_bytes + (size_t)line * C * sizeof(Cx<float>) + 128;
        if (smem > smem_optin - 1024) return false;
        if (use_tma && !tma_encode_tile_map(&map, src, 1, (int)map_rows, (int)(G.ntr / 2), elem_bytes, C, box_rows, parity ? 2 : 1)) return false;
        int dev = 0, sms = 0;
        P3D_CUDA(cudaGetDevice(&dev));
        P3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (const char* g = getenv("P3D_TIME_GRID")) sms = std::max(1, atoi(g));          // tests: many tiles per CTA on small inputs
        grid = (unsigned)std::min<long long>(tiles, sms);
        return true;
    }
    void tables(const std::vector<int>& radices, int h_split) {
        std::vector<Cx<float>> t = spec_twiddle_table(radices);
        const size_t woff = (t.size() + 1) & ~(size_t)1;
        if (h_split > 0) {
            t.resize(woff + (size_t)h_split);
            for (int k = 0; k < h_split; ++k) {
                const double a = -2.0 * M_PI * (double)k / (double)(2 * h_split);
                t[woff + k] = cmake<float>((float)cos(a), (float)sin(a));
            }
        }
        P3D_CUDA(cudaMalloc(&d_tw, sizeof(Cx<float>) * t.size()));
        P3D_CUDA(cudaMemcpy(d_tw, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
        d_wsplit = d_tw + woff;
    }
    template <typename K> void begin(K kernel) {
        P3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        P3D_CUDA(cudaEventCreate(&e0)); P3D_CUDA(cudaEventCreate(&e1));
        P3D_CUDA(cudaEventRecord(e0, 0));
    }
    void end() {
        P3D_CUDA(cudaGetLastError());
        P3D_CUDA(cudaEventRecord(e1, 0));
        P3D_CUDA(cudaDeviceSynchronize());
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); g_last_kernel_ms = ms; g_last_path = "tma";
    }
};

template <typename LP> std::vector<int> radices_of_plan() { std::vector<int> r(LP::NPASS); LP::radices(r.data()); return r; }

template <typename LP, int C, bool TMA = true>
bool launch_time_tma(const TimeGeom& G0, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    if (const char* a = getenv("P3D_TIME_ASYNC")) {                     // experiments: 1 = cp.async copies, 0 = TMA
        if constexpr (TMA) { if (atoi(a) == 1) return launch_time_tma<LP, C, false>(G0, din, dout, d_ph, inverse, smem_optin); }
        else               { if (atoi(a) == 0) return launch_time_tma<LP, C, true>(G0, din, dout, d_ph, inverse, smem_optin); }
    }
    OnePass op;
    if (!inverse) { if (!op.setup(G0, din, dout, C, LP::N, G0.nt, (int)sizeof(float2), LP::LINE, TMA, smem_optin)) return false; }
    else          { if (!op.setup(G0, din, dout, C, (int)G0.nf, G0.nf, (int)sizeof(float4), LP::LINE, TMA, smem_optin)) return false; }
    op.tables(radices_of_plan<LP>(), 0);
    if (!inverse) {
        op.begin(k_time_fwd_tma<LP, C, TMA>);
        k_time_fwd_tma<LP, C, TMA><<<op.grid, LP::T * C, op.smem>>>(op.G, op.map, op.d_tw, (const float*)din, (Cx<float>*)dout, d_ph, (int)op.tiles, op.box_rows, op.stage_rows);
    } else {
        op.begin(k_time_inv_tma<LP, C, TMA>);
        k_time_inv_tma<LP, C, TMA><<<op.grid, LP::T * C, op.smem>>>(op.G, op.map, op.d_tw, (const Cx<float>*)din, (float*)dout, d_ph, (int)op.tiles, op.box_rows, op.stage_rows);
    }
    op.end();
    return true;
}

template <typename LPH, int C>
bool launch_time_split(const TimeGeom& G0, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    constexpr int H = LPH::N;
    if (G0.nfft != 2 * H) return false;
    if (inverse && !G0.compute_real) return false;                       // a two-sided spectrum does not fit the stage
    OnePass op;
    constexpr bool PAR = C >= 8;                                         // even / odd halves of the stage (k_time_fwd_split)
    if (!inverse) { if (!op.setup(G0, din, dout, C, PAR ? H : 2 * H, G0.nt, (int)sizeof(float2), LPH::LINE, true, smem_optin, PAR)) return false; }
    else          { if (!op.setup(G0, din, dout, C, H + 1, G0.nf, (int)sizeof(float4), LPH::LINE, true, smem_optin)) return false; }
    op.tables(radices_of_plan<LPH>(), H);
    if (!inverse) {
        op.begin(k_time_fwd_split<LPH, C, PAR>);
        k_time_fwd_split<LPH, C, PAR><<<op.grid, LPH::T * C, op.smem>>>(op.G, op.map, op.d_tw, op.d_wsplit, (Cx<float>*)dout, d_ph, (int)op.tiles, op.box_rows, op.stage_rows);
    } else {
        op.begin(k_time_inv_split<LPH, C>);
        k_time_inv_split<LPH, C><<<op.grid, LPH::T * C, op.smem>>>(op.G, op.map, op.d_tw, op.d_wsplit, (float*)dout, d_ph, (int)op.tiles, op.box_rows, op.stage_rows);
    }
    op.end();
    return true;
}

bool try_time_split(const TimeGeom& G, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    switch (G.nfft) {
        case 2048: return launch_time_split<TP1024, 8>(G, din, dout, d_ph, inverse, smem_optin);
        case 4096: return launch_time_split<TP2048, 4>(G, din, dout, d_ph, inverse, smem_optin);
        case 2000: return launch_time_split<TP1000, 8>(G, din, dout, d_ph, inverse, smem_optin);
        case 4000: return launch_time_split<TP2000, 4>(G, din, dout, d_ph, inverse, smem_optin);
        default: return false;
    }
}

bool try_time_tma(const TimeGeom& G, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    {
        const char* sp = getenv("P3D_TIME_SPLIT");                      // experiments: 0 = N-point plans for the long records too
        if (!(sp && atoi(sp) == 0) && try_time_split(G, din, dout, d_ph, inverse, smem_optin)) return true;
    }
    // the inverse of a two-sided spectrum stages nfft rows of 16 bytes per trace pair: half the tile width if that is too much
    const bool narrow = inverse && !G.compute_real;
#define P3D_TIME_TMA(LP, C, TMA) \
    (launch_time_tma<LP, C, TMA>(G, din, dout, d_ph, inverse, smem_optin) || \
     (narrow && (C) >= 4 && launch_time_tma<LP, ((C) >= 4 ? (C) / 2 : (C)), TMA>(G, din, dout, d_ph, inverse, smem_optin)))
    // copies: TMA boxes, except where a box row is 32 bytes or less and the tensor unit's time per row exceeds the transform's
    switch (G.nfft) {
        case 512:  return P3D_TIME_TMA(TP512, 16, true);
        case 1024: return P3D_TIME_TMA(TP1024, 8, true);
        case 1000: return P3D_TIME_TMA(TP1000, 8, true);
        case 2500: return P3D_TIME_TMA(TP2500, 2, true);
        case 2048: return inverse ? P3D_TIME_TMA(TP2048, 4, true) : P3D_TIME_TMA(TP2048, 4, false);
        case 2000: return inverse ? P3D_TIME_TMA(TP2000, 4, true) : P3D_TIME_TMA(TP2000, 4, false);
        case 4096: return inverse ? P3D_TIME_TMA(TP4096, 2, false) : P3D_TIME_TMA(TP4096, 2, true);
        case 4000: return inverse ? P3D_TIME_TMA(TP4000, 2, false) : P3D_TIME_TMA(TP4000, 2, true);
        default: return false;
    }
#undef P3D_TIME_TMA
}

bool try_time_pipeline(const TimeGeom& G, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    switch (G.nfft) {
        case 512:  return launch_time_pipeline<TP512, 4>(G, din, dout, d_ph, inverse, smem_optin);
        case 1024: return launch_time_pipeline<TP1024, 4>(G, din, dout, d_ph, inverse, smem_optin);
        case 2048: return launch_time_pipeline<TP2048, 2>(G, din, dout, d_ph, inverse, smem_optin);
        case 4096: return launch_time_pipeline<TP4096, 1>(G, din, dout, d_ph, inverse, smem_optin);
        case 1000: return launch_time_pipeline<TP1000, 4>(G, din, dout, d_ph, inverse, smem_optin);
        case 2000: return launch_time_pipeline<TP2000, 2>(G, din, dout, d_ph, inverse, smem_optin);
        case 2500: return launch_time_pipeline<TP2500, 2>(G, din, dout, d_ph, inverse, smem_optin);
        case 3000: return launch_time_pipeline<TP3000, 1>(G, din, dout, d_ph, inverse, smem_optin);
        case 4000: return launch_time_pipeline<TP4000, 1>(G, din, dout, d_ph, inverse, smem_optin);
        case 5000: return launch_time_pipeline<TP5000, 1>(G, din, dout, d_ph, inverse, smem_optin);
        default: return false;
    }
}

bool try_time_direct(const TimeGeom& G, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    if (G.ntr % 2 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(din) | reinterpret_cast<uintptr_t>(dout)) & 15) return false;
    switch (G.nfft) {
        case 512:  return launch_time_spec<TP512, 16>(G, din, dout, d_ph, inverse, smem_optin);
        case 1024: return launch_time_spec<TP1024, 8>(G, din, dout, d_ph, inverse, smem_optin);
        case 2048: return launch_time_spec<TP2048, 4>(G, din, dout, d_ph, inverse, smem_optin);
        case 4096: return launch_time_spec<TP4096, 2>(G, din, dout, d_ph, inverse, smem_optin);
        case 2000: return launch_time_spec<TP2000, 4>(G, din, dout, d_ph, inverse, smem_optin);
        case 4000: return launch_time_spec<TP4000, 2>(G, din, dout, d_ph, inverse, smem_optin);
        default: return false;
    }
}

// Which implementation first (B200, 10^6 traces, profiles/r2_time_axis_paths_v2.txt): the one-pass TMA kernels - N-point
// plans up to 1024 samples (3.1 - 4.5 TB/s algorithmic), the radix-2 split for 2048 / 4096 / 2000 / 4000 (2.5 - 2.9 TB/s
// forward, 1.7 - 2.8 inverse); the direct register kernels (2.0 TB/s; tiles half as wide) for the inverse of a two-sided
// spectrum of a long record, which does not fit the stage, and for trace counts that are even but not divisible by 4; the
// transposing pipeline (1.3 TB/s for every length) takes what the others decline (odd trace counts, unaligned buffers).
// P3D_TIME_PATH = tma | direct | pipeline overrides the first choice (read at every call; used by the tests).
bool try_time_spec(const TimeGeom& G, const void* din, void* dout, const Cx<float>* d_ph, bool inverse, size_t smem_optin) {
    const char* path = getenv("P3D_TIME_PATH");
    const bool long_record = G.nfft == 2048 || G.nfft == 4096 || G.nfft == 2000 || G.nfft == 4000;
    int first = (long_record && inverse && !G.compute_real) ? 1 : 0;          // 0 one pass, 1 direct, 2 pipeline
    if (getenv("P3D_TIME_DIRECT") || (path && !strcmp(path, "direct"))) first = 1;
    else if (path && !strcmp(path, "tma")) first = 0;
    else if (path && !strcmp(path, "pipeline")) first = 2;
    if (first == 0) {
        if (try_time_tma(G, din, dout, d_ph, inverse, smem_optin)) return true;
        if (!path && try_time_direct(G, din, dout, d_ph, inverse, smem_optin)) return true;
    } else if (first == 1) {
        if (try_time_direct(G, din, dout, d_ph, inverse, smem_optin)) return true;
        if (path && !strcmp(path, "direct")) return false;                         // forced: the generic kernels take over
        if (try_time_tma(G, din, dout, d_ph, inverse, smem_optin)) return true;
    }
    return try_time_pipeline(G, din, dout, d_ph, inverse, smem_optin);
}

template <typename LP, int RB>
bool launch_env_pipeline(const TimeGeom& G, const float* din, float* dout, size_t smem_optin) {
    constexpr size_t smem = (size_t)2 * LP::LINE * RB * sizeof(Cx<float>);
    if (smem > smem_optin - 1024) return false;
    std::vector<int> rad(LP::NPASS);
    LP::radices(rad.data());
    std::vector<Cx<float>> t = spec_twiddle_table(rad);
    // the two trace-major intermediates (input and envelope, nt floats each) of a chunk stay L2 resident
    long long chunk = (long long)((size_t)72 * 1024 * 1024 / ((size_t)G.nt * 2 * sizeof(float)));
    chunk = std::max<long long>(256, (chunk / 64) * 64);
    chunk = std::min<long long>(chunk, ((G.ntr + 1) / 2) * 2);
    Cx<float>* d_tw = nullptr; float* tmp_a = nullptr; Cx<float>* tmp_b = nullptr;
    struct Free { void* p; cudaEvent_t e0, e1; ~Free() { if (p) cudaFree(p); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); } } fr{nullptr, nullptr, nullptr};
    P3D_CUDA(cudaMalloc(&d_tw, sizeof(Cx<float>) * t.size())); fr.p = d_tw;
    int dev = 0; P3D_CUDA(cudaGetDevice(&dev));
    scratch_get(dev, (size_t)G.nt * chunk, ((size_t)G.nt * chunk + 1) / 2, &tmp_a, &tmp_b);
    float* tmp_e = reinterpret_cast<float*>(tmp_b);
    P3D_CUDA(cudaMemcpy(d_tw, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
    P3D_CUDA(cudaEventCreate(&fr.e0)); P3D_CUDA(cudaEventCreate(&fr.e1));
    P3D_CUDA(cudaEventRecord(fr.e0, 0));
    P3D_CUDA(cudaFuncSetAttribute(k_time_env_rows<LP, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 tb(32, 8);
    for (long long c0 = 0; c0 < G.ntr; c0 += chunk) {
        const long long nc = std::min<long long>(chunk, G.ntr - c0);
        const unsigned ctas = (unsigned)((nc + 2 * RB - 1) / (2 * RB));
        k_transpose<float, false><<<dim3((unsigned)((nc + 31) / 32), (unsigned)((G.nt + 31) / 32)), tb>>>(din + c0, tmp_a, G.nt, nc, G.ntr, G.nt);
        k_time_env_rows<LP, RB><<<ctas, LP::T * RB, smem>>>(G, d_tw, tmp_a, tmp_e, nc);
        k_transpose<float, true><<<dim3((unsigned)((nc + 31) / 32), (unsigned)((G.nt + 31) / 32)), tb>>>(tmp_e, dout + c0, nc, G.nt, G.nt, G.ntr);
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaEventRecord(fr.e1, 0));
    P3D_CUDA(cudaDeviceSynchronize());
    { float ms = 0.f; cudaEventElapsedTime(&ms, fr.e0, fr.e1); g_last_kernel_ms = ms; g_last_path = "pipeline"; }
    return true;
}

template <typename LP, int C>
bool launch_env_tma(const TimeGeom& G0, const float* din, float* dout, size_t smem_optin) {
    OnePass op;
    if (!op.setup(G0, din, dout, C, LP::N, G0.nt, (int)sizeof(float2), LP::LINE, true, smem_optin)) return false;
    op.tables(radices_of_plan<LP>(), 0);
    op.begin(k_time_env_tma<LP, C>);
    k_time_env_tma<LP, C><<<op.grid, LP::T * C, op.smem>>>(op.G, op.map, op.d_tw, dout, (int)op.tiles, op.box_rows, op.stage_rows);
    op.end();
    return true;
}

template <typename LPH, int C>
bool launch_env_split(const TimeGeom& G0, const float* din, float* dout, size_t smem_optin) {
    constexpr int H = LPH::N;
    if (G0.nfft != 2 * H) return false;
    OnePass op;
    if (!op.setup(G0, din, dout, C, 2 * H, G0.nt, (int)sizeof(float2), LPH::LINE, true, smem_optin)) return false;
    op.tables(radices_of_plan<LPH>(), H);
    op.begin(k_time_env_split<LPH, C>);
    k_time_env_split<LPH, C><<<op.grid, LPH::T * C, op.smem>>>(op.G, op.map, op.d_tw, op.d_wsplit, dout, (int)op.tiles, op.box_rows, op.stage_rows);
    op.end();
    return true;
}

bool try_env_tma(const TimeGeom& G, const float* din, float* dout, size_t smem_optin) {
    const char* path = getenv("P3D_TIME_PATH");
    if (path && strcmp(path, "tma")) return false;
    {
        const char* sp = getenv("P3D_TIME_SPLIT");
        bool done = false;
        if (!(sp && atoi(sp) == 0)) {
            switch (G.nfft) {
                case 2048: done = launch_env_split<TP1024, 8>(G, din, dout, smem_optin); break;
                case 4096: done = launch_env_split<TP2048, 4>(G, din, dout, smem_optin); break;
                case 2000: done = launch_env_split<TP1000, 8>(G, din, dout, smem_optin); break;
                case 4000: done = launch_env_split<TP2000, 4>(G, din, dout, smem_optin); break;
                default: break;
            }
        }
        if (done) return true;
    }
    if (!path && G.nfft == 2048) return false;          // N-point plan at 2048 samples, 32-byte box rows: the pipeline is as fast
    switch (G.nfft) {
        case 512:  return launch_env_tma<TP512, 16>(G, din, dout, smem_optin);
        case 1024: return launch_env_tma<TP1024, 8>(G, din, dout, smem_optin);
        case 2048: return launch_env_tma<TP2048, 4>(G, din, dout, smem_optin);
        case 4096: return launch_env_tma<TP4096, 2>(G, din, dout, smem_optin);
        case 1000: return launch_env_tma<TP1000, 8>(G, din, dout, smem_optin);
        case 2000: return launch_env_tma<TP2000, 4>(G, din, dout, smem_optin);
        case 2500: return launch_env_tma<TP2500, 2>(G, din, dout, smem_optin);
        case 4000: return launch_env_tma<TP4000, 2>(G, din, dout, smem_optin);
        default: return false;
    }
}

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

int choose_cols(int L, size_t smem_optin) {
    int C = 8;
    while (C > 1 && (size_t)2 * L * C * sizeof(Cx<float>) > smem_optin - 2048) C >>= 1;
    return C;
}

int time_impl(int device, const void* x, int x_mem, void* out, int out_mem, int64_t nt, int64_t nfft,
              int64_t ntr, double dt, double t0, int compute_real, int ascending, const double* window, bool inverse) {
    P3D_REQUIRE(x && out, P3D_ERR_BAD_ARG, "null argument");
    P3D_REQUIRE(nfft >= 2 && nfft % 2 == 0, P3D_ERR_BAD_ARG, "nfft must be even and >= 2 (got %lld)", (long long)nfft);
    P3D_REQUIRE(nt >= 1 && nt <= nfft, P3D_ERR_BAD_ARG, "need 1 <= nt <= nfft (nt=%lld nfft=%lld)", (long long)nt, (long long)nfft);
    P3D_REQUIRE(ntr >= 1, P3D_ERR_BAD_ARG, "n_traces must be >= 1");
    P3D_REQUIRE(dt != 0.0, P3D_ERR_BAD_ARG, "dt must be non-zero");
    int ndev = 0; P3D_CUDA(cudaGetDeviceCount(&ndev));
    P3D_REQUIRE(device >= 0 && device < ndev, P3D_ERR_BAD_ARG, "device %d out of range", device);
    DeviceGuard guard(device);
    struct { size_t sharedMemPerBlockOptin; } prop;
    { int v = 0; P3D_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); prop.sharedMemPerBlockOptin = (size_t)v; }
    g_last_kernel_ms = 0.0; g_last_path = "generic";

    // the generic direct kernels need an axis plan; the register-resident pipeline (record lengths
    // 512 .. 4096) does not, so it is only built when that path declines
    static const bool no_spec = getenv("P3D_TIME_GENERIC") != nullptr;
    const bool pow2_len = nfft == 512 || nfft == 1024 || nfft == 2048 || nfft == 4096;
    const bool smooth_len = nfft == 1000 || nfft == 2000 || nfft == 2500 || nfft == 3000 || nfft == 4000 || nfft == 5000;
    const bool spec_len = !no_spec && (pow2_len || smooth_len);
    AxisPlan ax;
    struct Cleanup { AxisPlan* a; void* p[3]; ~Cleanup() { a->release(); for (void* q : p) if (q) cudaFree(q); } } cl{&ax, {nullptr, nullptr, nullptr}};
    int C = 1; size_t smem = 0;
    if (!spec_len) {
        ax.build((int)nfft);
        C = choose_cols(ax.L, prop.sharedMemPerBlockOptin);
        smem = (size_t)2 * ax.L * C * sizeof(Cx<float>);
        P3D_REQUIRE(smem <= prop.sharedMemPerBlockOptin, P3D_ERR_NOT_IMPLEMENTED, "time axis of %lld samples does not fit in shared memory", (long long)nfft);
    }

    const int64_t nf = compute_real ? nfft / 2 + 1 : nfft;
    // phase table in double: forward dt*exp(-2 pi i f t0)*window ; inverse exp(+2 pi i f t0)/(dt*nfft)
    const int64_t nbins = inverse ? (compute_real ? nfft / 2 + 1 : nfft) : nf;
    std::vector<Cx<float>> ph((size_t)nbins);
    for (int64_t k = 0; k < nbins; ++k) {
        const int64_t ks = (!compute_real && k >= nfft / 2) ? k - nfft : k;     // fftfreq: bin nfft/2 is -Nyquist
        const double f = (double)ks / ((double)nfft * dt);
        double cyc = f * t0; cyc -= std::floor(cyc);
        const double ang = 2.0 * M_PI * cyc;
        double amp = inverse ? 1.0 / (dt * (double)nfft) : dt;
        if (!inverse && window) amp *= window[k];
        ph[k] = inverse ? cmake<float>((float)(amp * cos(ang)), (float)(amp * sin(ang)))
                        : cmake<float>((float)(amp * cos(ang)), (float)(-amp * sin(ang)));
    }
    Cx<float>* d_ph = nullptr;
    P3D_CUDA(cudaMalloc(&d_ph, sizeof(Cx<float>) * nbins)); cl.p[0] = d_ph;
    P3D_CUDA(cudaMemcpy(d_ph, ph.data(), sizeof(Cx<float>) * nbins, cudaMemcpyHostToDevice));

    const size_t in_bytes = inverse ? sizeof(Cx<float>) * nf * ntr : sizeof(float) * nt * ntr;
    const size_t out_bytes = inverse ? sizeof(float) * nt * ntr : sizeof(Cx<float>) * nf * ntr;
    const void* din = x; void* dout = out;
    if (x_mem == P3D_MEM_HOST) { void* p = nullptr; P3D_CUDA(cudaMalloc(&p, in_bytes)); cl.p[1] = p; P3D_CUDA(cudaMemcpy(p, x, in_bytes, cudaMemcpyHostToDevice)); din = p; }
    if (out_mem == P3D_MEM_HOST) { void* p = nullptr; P3D_CUDA(cudaMalloc(&p, out_bytes)); cl.p[2] = p; dout = p; }

    TimeGeom G; G.nt = nt; G.nf = nf; G.ntr = ntr; G.nfft = (int)nfft; G.C = C; G.compute_real = compute_real; G.ascending = ascending;
    const long long tiles = (ntr + 2 * C - 1) / (2 * C);
    P3D_REQUIRE(tiles < 2147483647LL, P3D_ERR_BAD_ARG, "too many traces");
    const int threads = 512;
    bool done = spec_len && try_time_spec(G, din, dout, d_ph, inverse, prop.sharedMemPerBlockOptin);
    if (!done && spec_len) {          // the pipeline declined (e.g. shared memory): fall back to the direct kernels
        ax.build((int)nfft);
        C = choose_cols(ax.L, prop.sharedMemPerBlockOptin);
        smem = (size_t)2 * ax.L * C * sizeof(Cx<float>);
        G.C = C;
    }
    const long long tiles2 = (ntr + 2 * C - 1) / (2 * C);
    if (done) {
        // done by the register-resident kernels
    } else if (!inverse) {
        P3D_CUDA(cudaFuncSetAttribute(k_time_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));
        k_time_fwd<<<(unsigned)tiles2, threads, smem>>>(G, ax.dev(), (const float*)din, (Cx<float>*)dout, d_ph);
    } else {
        P3D_CUDA(cudaFuncSetAttribute(k_time_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - 1024));
        k_time_inv<<<(unsigned)tiles2, threads, smem>>>(G, ax.dev(), (const Cx<float>*)din, (float*)dout, d_ph);
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaDeviceSynchronize());
    if (out_mem == P3D_MEM_HOST) P3D_CUDA(cudaMemcpy(out, dout, out_bytes, cudaMemcpyDeviceToHost));
    return P3D_OK;
}

int envelope_impl(int device, const float* x, int x_mem, float* out, int out_mem, int64_t nt, int64_t ntr) {
    P3D_REQUIRE(x && out, P3D_ERR_BAD_ARG, "null argument");
    P3D_REQUIRE(nt >= 1 && ntr >= 1, P3D_ERR_BAD_ARG, "need nt >= 1 and n_traces >= 1");
    int ndev = 0; P3D_CUDA(cudaGetDeviceCount(&ndev));
    P3D_REQUIRE(device >= 0 && device < ndev, P3D_ERR_BAD_ARG, "device %d out of range", device);
    DeviceGuard guard(device);
    int optin = 0; P3D_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    g_last_kernel_ms = 0.0; g_last_path = "generic";
    AxisPlan ax;
    struct Cleanup { AxisPlan* a; void* p[2]; ~Cleanup() { a->release(); for (void* q : p) if (q) cudaFree(q); } } cl{&ax, {nullptr, nullptr}};
    const size_t bytes = sizeof(float) * nt * ntr;
    const float* din = x; float* dout = out;
    if (x_mem == P3D_MEM_HOST) { void* p = nullptr; P3D_CUDA(cudaMalloc(&p, bytes)); cl.p[0] = p; P3D_CUDA(cudaMemcpy(p, x, bytes, cudaMemcpyHostToDevice)); din = (const float*)p; }
    if (out_mem == P3D_MEM_HOST) { void* p = nullptr; P3D_CUDA(cudaMalloc(&p, bytes)); cl.p[1] = p; dout = (float*)p; }
    TimeGeom G; G.nt = nt; G.nf = nt; G.ntr = ntr; G.nfft = (int)nt; G.C = 1; G.compute_real = 0; G.ascending = 0;
    static const bool no_spec = getenv("P3D_TIME_GENERIC") != nullptr;
    bool done = !no_spec && try_env_tma(G, din, dout, (size_t)optin);
    if (!no_spec && !done) {
        switch (nt) {
            case 512:  done = launch_env_pipeline<TP512, 4>(G, din, dout, (size_t)optin); break;
            case 1024: done = launch_env_pipeline<TP1024, 4>(G, din, dout, (size_t)optin); break;
            case 2048: done = launch_env_pipeline<TP2048, 2>(G, din, dout, (size_t)optin); break;
            case 4096: done = launch_env_pipeline<TP4096, 1>(G, din, dout, (size_t)optin); break;
            case 1000: done = launch_env_pipeline<TP1000, 4>(G, din, dout, (size_t)optin); break;
            case 2000: done = launch_env_pipeline<TP2000, 2>(G, din, dout, (size_t)optin); break;
            case 2500: done = launch_env_pipeline<TP2500, 2>(G, din, dout, (size_t)optin); break;
            case 3000: done = launch_env_pipeline<TP3000, 1>(G, din, dout, (size_t)optin); break;
            case 4000: done = launch_env_pipeline<TP4000, 1>(G, din, dout, (size_t)optin); break;
            case 5000: done = launch_env_pipeline<TP5000, 1>(G, din, dout, (size_t)optin); break;
            default: break;
        }
    }
    if (!done) {
        ax.build((int)nt);
        const int C = choose_cols(ax.L, (size_t)optin);
        const size_t smem = (size_t)2 * ax.L * C * sizeof(Cx<float>);
        P3D_REQUIRE(smem <= (size_t)optin, P3D_ERR_NOT_IMPLEMENTED, "time axis of %lld samples does not fit in shared memory", (long long)nt);
        G.C = C;
        const long long tiles = (ntr + 2 * C - 1) / (2 * C);
        P3D_REQUIRE(tiles < 2147483647LL, P3D_ERR_BAD_ARG, "too many traces");
        P3D_CUDA(cudaFuncSetAttribute(k_time_env, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024));
        k_time_env<<<(unsigned)tiles, 512, smem>>>(G, ax.dev(), din, dout);
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaDeviceSynchronize());
    if (out_mem == P3D_MEM_HOST) P3D_CUDA(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    return P3D_OK;
}

}  // namespace

extern "C" {

int p3d_time_last_kernel_ms(double* ms) { if (!ms) return P3D_ERR_BAD_ARG; *ms = g_last_kernel_ms; return P3D_OK; }
const char* p3d_time_last_path(void) { return g_last_path; }

int p3d_time_fft(int device, const float* x, int x_mem, void* out, int out_mem, int64_t nt, int64_t nfft,
                 int64_t n_traces, double dt, double t0, int compute_real, const double* window) {
    try { return time_impl(device, x, x_mem, out, out_mem, nt, nfft, n_traces, dt, t0, compute_real, 0, window, false); }
    catch (const P3dFail& f) { return f.code; }
}

int p3d_time_ifft(int device, const void* x, int x_mem, float* out, int out_mem, int64_t nfft, int64_t nt_out,
                  int64_t n_traces, double dt, double t0, int compute_real, int ascending) {
    try { return time_impl(device, x, x_mem, out, out_mem, nt_out, nfft, n_traces, dt, t0, compute_real, ascending, nullptr, true); }
    catch (const P3dFail& f) { return f.code; }
}

int p3d_time_envelope(int device, const float* x, int x_mem, float* out, int out_mem, int64_t nt, int64_t n_traces) {
    try { return envelope_impl(device, x, x_mem, out, out_mem, nt, n_traces); }
    catch (const P3dFail& f) { return f.code; }
}

}  // extern "C"
