// Plan, band scheduler and C ABI of the POCS path.
#include "p3d_host.h"
#include "p3d_pocs_kernels.cuh"
#include "p3d_pocs_spec.cuh"
#include "p3d_pocs_launch.h"
#include "p3d_pocs_f64.h"
#include "p3d_schedule.h"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

using namespace p3d;

namespace {

struct EventPair { cudaEvent_t a, b; int kind; };

struct Lane {
    cudaStream_t stream = nullptr;
    int64_t cap = 0;                 // slice capacity of the buffers below
    bool own_d = false, own_out = false;
    Cx<float>* W = nullptr;
    Cx<float>* D = nullptr;          // staging of x when x is host memory
    Cx<float>* OUT = nullptr;        // staging of out when out is host memory
    Cx<float>* tau = nullptr;        // [cap][niter]
    double* S = nullptr;             // [cap][niter+1]
    int* stop = nullptr;             // [cap]
    SliceStats* stats = nullptr;     // [cap]
    int niter_cap = 0;
    // pinned host mirrors
    Cx<float>* h_tau = nullptr; double* h_S = nullptr; int* h_stop = nullptr; SliceStats* h_stats = nullptr;
    // pending chunk (results not yet collected)
    bool pending = false; int64_t p_first = 0, p_count = 0;
};

}  // namespace

struct p3d_plan {
    int device = 0;
    int n1 = 0, n2 = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    AxisPlan ax1, ax2;
    PocsGeom geom{};
    int col_threads = 256, row_threads = 256;
    size_t col_smem = 0, row_smem = 0;
    SpecKernels spec{};              // specialised register-resident kernels when available
    bool force_generic = false;
    int spec_variant64 = 0;
    int precision = 32;              // 32: fp32 fast path, 64: float64 state mode (p3d_pocs_f64.cu)
    F64Runner* f64 = nullptr;
    int64_t max_slices = 0;
    int band_slices = 0;
    int n_lanes = 0;                 // 0 = auto
    std::vector<Lane> lanes;
    uint8_t* d_mask = nullptr; int64_t d_mask_bytes = 0;
    uint8_t* d_zero_mask = nullptr;                             // kx-ky filter mode: all-zero mask plane
    float* d_filt = nullptr;                                    // kx-ky filter mode: staged filter plane
    uint32_t* d_mbits = nullptr; int64_t d_mbits_words = 0;    // packed mask of the specialised row kernel
    Cx<float>* spec_tw_cols = nullptr;                          // per-pass twiddle tables (p3d_fft_reg.cuh)
    Cx<float>* spec_tw_rows = nullptr;
    void* cub_temp = nullptr; size_t cub_temp_bytes = 0;
    // percentile operators: spectrum scratch + |X| keys of a few slices
    Cx<float>* pct_scr = nullptr; unsigned int* pct_keys = nullptr; unsigned int* pct_sorted = nullptr; int64_t pct_slices = 0;
    cudaEvent_t ev[8] = {nullptr};
    // profiling
    bool profiling = false;
    std::vector<EventPair> events;
    double prof_ms[P3D_PROFILE_KINDS] = {0};
    int64_t prof_n[P3D_PROFILE_KINDS] = {0};
};

namespace {

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

void prof_begin(p3d_plan* P, cudaStream_t st, int kind) {
    if (!P->profiling) return;
    EventPair e; e.kind = kind;
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
    cudaEventRecord(e.a, st);
    P->events.push_back(e);
}
void prof_end(p3d_plan* P, cudaStream_t st) {
    if (!P->profiling) return;
    cudaEventRecord(P->events.back().b, st);
}
void prof_collect(p3d_plan* P) {
    for (auto& e : P->events) {
        float ms = 0.f;
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            P->prof_ms[e.kind] += ms; P->prof_n[e.kind] += 1;
        }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    P->events.clear();
}

void free_lane(Lane& L) {
    if (L.W) cudaFree(L.W);
    if (L.own_d && L.D) cudaFree(L.D);
    if (L.own_out && L.OUT) cudaFree(L.OUT);
    if (L.tau) cudaFree(L.tau);
    if (L.S) cudaFree(L.S);
    if (L.stop) cudaFree(L.stop);
    if (L.stats) cudaFree(L.stats);
    if (L.h_tau) cudaFreeHost(L.h_tau);
    if (L.h_S) cudaFreeHost(L.h_S);
    if (L.h_stop) cudaFreeHost(L.h_stop);
    if (L.h_stats) cudaFreeHost(L.h_stats);
    L.W = L.D = L.OUT = L.tau = nullptr; L.S = nullptr; L.stop = nullptr; L.stats = nullptr;
    L.h_tau = nullptr; L.h_S = nullptr; L.h_stop = nullptr; L.h_stats = nullptr;
    L.cap = 0; L.niter_cap = 0; L.own_d = L.own_out = false;
}

void ensure_lane(p3d_plan* P, Lane& L, int64_t cap, int niter, bool need_d, bool need_out) {
    const int64_t ne = (int64_t)P->n1 * P->n2;
    if (!L.stream) P3D_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    const bool regrow = cap > L.cap || niter > L.niter_cap || (need_d && !L.own_d) || (need_out && !L.own_out);
    if (!regrow) return;
    cudaStream_t st = L.stream;
    cap = std::max(cap, L.cap); niter = std::max(niter, L.niter_cap);
    need_d = need_d || L.own_d; need_out = need_out || L.own_out;
    L.stream = nullptr; free_lane(L); L.stream = st;
    P3D_CUDA(cudaMalloc(&L.W, sizeof(Cx<float>) * ne * cap));
    if (need_d) { P3D_CUDA(cudaMalloc(&L.D, sizeof(Cx<float>) * ne * cap)); L.own_d = true; }
    if (need_out) { P3D_CUDA(cudaMalloc(&L.OUT, sizeof(Cx<float>) * ne * cap)); L.own_out = true; }
    P3D_CUDA(cudaMalloc(&L.tau, sizeof(Cx<float>) * cap * niter));
    P3D_CUDA(cudaMalloc(&L.S, sizeof(double) * cap * (niter + 1)));
    P3D_CUDA(cudaMalloc(&L.stop, sizeof(int) * cap));
    P3D_CUDA(cudaMalloc(&L.stats, sizeof(SliceStats) * cap));
    P3D_CUDA(cudaMallocHost(&L.h_tau, sizeof(Cx<float>) * cap * niter));
    P3D_CUDA(cudaMallocHost(&L.h_S, sizeof(double) * cap * (niter + 1)));
    P3D_CUDA(cudaMallocHost(&L.h_stop, sizeof(int) * cap));
    P3D_CUDA(cudaMallocHost(&L.h_stats, sizeof(SliceStats) * cap));
    L.cap = cap; L.niter_cap = niter;
}

// ---- tile geometry of the generic kernels -------------------------------------------------------
void choose_geometry(p3d_plan* P) {
    const size_t budget_one = P->smem_optin - 2048;       // one CTA per SM
    const size_t budget_two = (228 * 1024) / 2 - 2048;    // two CTAs per SM
    const int L1 = P->ax1.L, L2 = P->ax2.L;
    // columns per tile: prefer the widest tile that still lets two CTAs share an SM, but
    // never narrower than 4 columns (32 B segments) if a single CTA can hold it
    int C = 16;
    while (C > 1 && (size_t)2 * L1 * C * sizeof(Cx<float>) > budget_two) C >>= 1;
    if (C < 4) { C = 4; while (C > 1 && (size_t)2 * L1 * C * sizeof(Cx<float>) > budget_one) C >>= 1; }
    P3D_REQUIRE((size_t)2 * L1 * C * sizeof(Cx<float>) <= budget_one, P3D_ERR_NOT_IMPLEMENTED,
                "iline axis of length %d (transform length %d) does not fit in shared memory", P->n1, L1);
    C = std::min(C, std::max(1, P->n2));
    const int pitch2 = L2 + 1;
    int RB = 16;
    while (RB > 1 && (size_t)2 * pitch2 * RB * sizeof(Cx<float>) > budget_two) RB >>= 1;
    P3D_REQUIRE((size_t)2 * pitch2 * RB * sizeof(Cx<float>) <= budget_one, P3D_ERR_NOT_IMPLEMENTED,
                "xline axis of length %d (transform length %d) does not fit in shared memory", P->n2, L2);
    RB = std::min(RB, std::max(1, P->n1));
    P->geom.n1 = P->n1; P->geom.n2 = P->n2; P->geom.C = C; P->geom.RB = RB; P->geom.pitch2 = pitch2;
    P->geom.slices_per_mask = 1;
    P->col_smem = (size_t)2 * L1 * C * sizeof(Cx<float>);
    P->row_smem = (size_t)2 * pitch2 * RB * sizeof(Cx<float>);
    auto pick_threads = [](long elems) { long t = ((elems / 8 + 31) / 32) * 32; return (int)std::min<long>(512, std::max<long>(128, t)); };
    P->col_threads = pick_threads((long)L1 * C);
    P->row_threads = pick_threads((long)L2 * RB);
}

GenericCfg generic_cfg(p3d_plan* P) {
    GenericCfg c; c.geom = P->geom; c.col_threads = P->col_threads; c.row_threads = P->row_threads;
    c.col_smem = P->col_smem; c.row_smem = P->row_smem;
    return c;
}

// (re)select the specialised kernels and upload their per-pass twiddle tables
void install_spec(p3d_plan* P, int variant) {
    P->spec = select_spec_kernels(P->n1, P->n2, variant);
    auto upload = [](const std::vector<int>& radices, Cx<float>** dst) {
        if (*dst) { cudaFree(*dst); *dst = nullptr; }
        if (radices.empty()) return;
        std::vector<Cx<float>> t = spec_twiddle_table(radices);
        P3D_CUDA(cudaMalloc(dst, sizeof(Cx<float>) * t.size()));
        P3D_CUDA(cudaMemcpy(*dst, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
    };
    if (P->spec.cols_table) {
        if (P->spec_tw_cols) { cudaFree(P->spec_tw_cols); P->spec_tw_cols = nullptr; }
        std::vector<Cx<float>> t = P->spec.cols_table();
        P3D_CUDA(cudaMalloc(&P->spec_tw_cols, sizeof(Cx<float>) * t.size()));
        P3D_CUDA(cudaMemcpy(P->spec_tw_cols, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
    } else upload(P->spec.cols_radices, &P->spec_tw_cols);
    if (P->spec.rows_table) {
        if (P->spec_tw_rows) { cudaFree(P->spec_tw_rows); P->spec_tw_rows = nullptr; }
        std::vector<Cx<float>> t = P->spec.rows_table();
        P3D_CUDA(cudaMalloc(&P->spec_tw_rows, sizeof(Cx<float>) * t.size()));
        P3D_CUDA(cudaMemcpy(P->spec_tw_rows, t.data(), sizeof(Cx<float>) * t.size(), cudaMemcpyHostToDevice));
    } else
    upload(P->spec.rows_radices, &P->spec_tw_rows);
    P->d_mbits_words = 0;           // layout may have changed: repack on the next run
}

// ---- launches -----------------------------------------------------------------------------------
void launch_rows_init(p3d_plan* P, cudaStream_t st, const BandArgs<float>& A, int nslices) {
    prof_begin(P, st, 0);
    if (P->spec.rows_init && !P->force_generic) P->spec.rows_init(P->geom, P->spec_tw_rows, A, nslices, st);
    else generic_rows_init(generic_cfg(P), P->ax2.dev(), A, nslices, st);
    prof_end(P, st);
    P3D_CUDA(cudaGetLastError());
}
void launch_cols_stats(p3d_plan* P, cudaStream_t st, const BandArgs<float>& A, int nslices) {
    prof_begin(P, st, 1);
    if (P->spec.cols_stats && !P->force_generic) P->spec.cols_stats(P->geom, P->spec_tw_cols, A, nslices, st);
    else generic_cols_stats(generic_cfg(P), P->ax1.dev(), A, nslices, st);
    prof_end(P, st);
    P3D_CUDA(cudaGetLastError());
}
void launch_cols_iter(p3d_plan* P, cudaStream_t st, const BandArgs<float>& A, int nslices, int op) {
    prof_begin(P, st, 2);
    if (P->spec.cols_iter && !P->force_generic) P->spec.cols_iter(P->geom, P->spec_tw_cols, A, nslices, op, st);
    else generic_cols_iter(generic_cfg(P), P->ax1.dev(), A, nslices, op, st);
    prof_end(P, st);
    P3D_CUDA(cudaGetLastError());
}
void launch_rows_iter(p3d_plan* P, cudaStream_t st, const BandArgs<float>& A, int nslices) {
    prof_begin(P, st, 3);
    if (P->spec.rows_iter && !P->force_generic) P->spec.rows_iter(P->geom, P->spec_tw_rows, A, nslices, st);
    else generic_rows_iter(generic_cfg(P), P->ax2.dev(), A, nslices, st);
    prof_end(P, st);
    P3D_CUDA(cudaGetLastError());
}

// ---- data-driven schedule helpers -----------------------------------------------------------------
__global__ void k_make_keys(unsigned long long* keys, long long n, unsigned long long lo, unsigned long long hi,
                            SliceStats* st) {
    // keys[] holds X0 as complex64 on entry; replace by its ordered key if lo < key < hi, else 0
    unsigned long long cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = reinterpret_cast<const float2*>(keys)[i];
        const unsigned long long key = lex_key(v.x, v.y);
        const bool in = key > lo && key < hi;
        keys[i] = in ? key : 0ull;
        cnt += in ? 1ull : 0ull;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&st->n_cand, cnt);
}
__global__ void k_pick_tau(const unsigned long long* sorted_desc, const SliceStats* st, Cx<float>* tau, int niter) {
    const unsigned long long nv = st->n_cand;
    for (int k = threadIdx.x; k < niter; k += blockDim.x) {
        if (nv == 0) { tau[k] = cmake<float>(nanf(""), nanf("")); continue; }
        long long idx = 0;
        if (k > 0) idx = (long long)ceil((double)((long long)k * (long long)(nv - 1)) / (double)(niter - 1));
        if (idx > (long long)nv - 1) idx = (long long)nv - 1;
        const unsigned long long key = sorted_desc[idx];
        tau[k] = cmake<float>(f32_from_ordered((unsigned int)(key >> 32)), f32_from_ordered((unsigned int)(key & 0xffffffffu)));
    }
}

// ---- percentile operators (functions/POCS.py:43-58): tau_k = np.percentile(|X_k|, q_k) per slice and iteration ----
__global__ void k_abs_keys(const Cx<float>* __restrict__ X, unsigned int* __restrict__ keys, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const Cx<float> v = X[i];
        keys[i] = __float_as_uint(sqrtf(v.x * v.x + v.y * v.y));      // non-negative floats order like their bit patterns
    }
}
// numpy's default 'linear' percentile: virtual index q/100 * (n-1), _lerp between the two neighbours
__global__ void k_pick_percentile(const unsigned int* __restrict__ sorted, long long n, Cx<float>* tau_sk) {
    const double q = (double)tau_sk->x;
    const double pos = q / 100.0 * (double)(n - 1);
    long long lo = (long long)floor(pos);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    const long long hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
    const double t = pos - (double)lo;
    const double a = (double)__uint_as_float(sorted[lo]), b = (double)__uint_as_float(sorted[hi]);
    const double diff = b - a;
    double r = a + diff * t;
    if (t >= 0.5) r = b - diff * (1.0 - t);
    *tau_sk = cmake<float>((float)r, 0.f);
}

// float statistics of a slice -> inputs of the host schedule (p3d_schedule.h)
ScheduleStats stats_f32(const SliceStats& st) {
    ScheduleStats s;
    s.z = cd(f32_from_ordered((unsigned int)(st.lexmax_key >> 32)), f32_from_ordered((unsigned int)(st.lexmax_key & 0xffffffffu)));
    float fmax, fmin; unsigned int ub = st.maxabs_bits, lb = st.minabs_bits;
    memcpy(&fmax, &ub, 4); memcpy(&fmin, &lb, 4);
    s.vmax = fmax; s.vmin = fmin; s.sumsq = st.sumsq;
    return s;
}

void ensure_mask(p3d_plan* P, const uint8_t* mask, int64_t bytes, int mem, cudaStream_t st, const uint8_t** dmask) {
    if (mem == P3D_MEM_DEVICE) { *dmask = mask; return; }
    if (bytes > P->d_mask_bytes) {
        if (P->d_mask) cudaFree(P->d_mask);
        P->d_mask = nullptr; P->d_mask_bytes = 0;
        P3D_CUDA(cudaMalloc(&P->d_mask, bytes));
        P->d_mask_bytes = bytes;
    }
    P3D_CUDA(cudaMemcpyAsync(P->d_mask, mask, bytes, cudaMemcpyHostToDevice, st));
    P3D_CUDA(cudaStreamSynchronize(st));
    *dmask = P->d_mask;
}

// packed mask words for the specialised row kernel (rebuilt every run: the mask may have changed)
const uint32_t* pack_mask(p3d_plan* P, const uint8_t* dmask, int64_t n_masks, cudaStream_t st) {
    if (!P->spec.pack_mask) return nullptr;
    const int64_t words = n_masks * (int64_t)P->n1 * P->spec.rows_T;
    if (words > P->d_mbits_words || !P->d_mbits) {
        if (P->d_mbits) cudaFree(P->d_mbits);
        P->d_mbits = nullptr; P->d_mbits_words = 0;
        P3D_CUDA(cudaMalloc(&P->d_mbits, sizeof(uint32_t) * words));
        P->d_mbits_words = words;
    }
    P->spec.pack_mask(dmask, P->d_mbits, (int)n_masks, P->n1, st);
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaStreamSynchronize(st));
    return P->d_mbits;
}

int64_t auto_capacity(p3d_plan* P, int64_t n_slices, int nbuf_per_lane, int lanes) {
    size_t free_b = 0, total_b = 0;
    P3D_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double per_slice = (double)P->n1 * P->n2 * sizeof(Cx<float>) * nbuf_per_lane;
    int64_t cap = (int64_t)((double)free_b * 0.85 / per_slice / lanes);
    cap = std::max<int64_t>(1, std::min<int64_t>(cap, n_slices));
    if (P->max_slices > 0) cap = std::min<int64_t>(cap, P->max_slices);
    return std::min<int64_t>(cap, 60000);
}

// tau[s][k] holds the scheduled percentile q_k on entry and the threshold np.percentile(|colFFT(W_s)|, q_k) on exit
void percentile_thresholds(p3d_plan* P, cudaStream_t st, const BandArgs<float>& B, int nb, int k) {
    const int64_t ne = (int64_t)P->n1 * P->n2;
    if (!P->pct_scr) {
        P->pct_slices = std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)(256e6 / (8.0 * (double)ne))));
        P3D_CUDA(cudaMalloc(&P->pct_scr, sizeof(Cx<float>) * ne * P->pct_slices));
        P3D_CUDA(cudaMalloc(&P->pct_keys, sizeof(unsigned int) * ne));
        P3D_CUDA(cudaMalloc(&P->pct_sorted, sizeof(unsigned int) * ne));
    }
    size_t need = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, (unsigned int*)nullptr, (unsigned int*)nullptr, (long long)ne, 0, 32, st);
    if (need > P->cub_temp_bytes) {
        P3D_CUDA(cudaStreamSynchronize(st));
        if (P->cub_temp) cudaFree(P->cub_temp);
        P->cub_temp = nullptr; P->cub_temp_bytes = 0;
        P3D_CUDA(cudaMalloc(&P->cub_temp, need)); P->cub_temp_bytes = need;
    }
    prof_begin(P, st, 6);
    for (int64_t s0 = 0; s0 < nb; s0 += P->pct_slices) {
        const int cnt = (int)std::min<int64_t>(P->pct_slices, nb - s0);
        P3D_CUDA(cudaMemcpyAsync(P->pct_scr, B.W + s0 * ne, sizeof(Cx<float>) * ne * cnt, cudaMemcpyDeviceToDevice, st));
        generic_fft_cols(generic_cfg(P), P->ax1.dev(), P->pct_scr, cnt, st);
        for (int i = 0; i < cnt; ++i) {
            k_abs_keys<<<(unsigned)std::min<long long>((ne + 255) / 256, 148 * 8), 256, 0, st>>>(P->pct_scr + (int64_t)i * ne, P->pct_keys, ne);
            size_t tb = P->cub_temp_bytes;
            cub::DeviceRadixSort::SortKeys(P->cub_temp, tb, P->pct_keys, P->pct_sorted, (long long)ne, 0, 32, st);
            k_pick_percentile<<<1, 1, 0, st>>>(P->pct_sorted, ne, const_cast<Cx<float>*>(B.tau) + (s0 + i) * B.niter + k);
        }
    }
    prof_end(P, st);
    P3D_CUDA(cudaGetLastError());
}

struct RunCtx {
    p3d_plan* P; const p3d_pocs_params* pr;
    const Cx<float>* x; int x_mem; const uint8_t* dmask; const uint32_t* dmbits; int64_t spm;
    Cx<float>* out; int out_mem;
    int32_t* niter_out; double* cost_out; double* costs_out;
    double* tau_out;            // schedule-only mode
    bool schedule_only;
    const float* filt;          // kx-ky filter mode (device pointer) or null
};

// collect results of the chunk pending on a lane
void collect_lane(RunCtx& R, Lane& L) {
    if (!L.pending) return;
    P3D_CUDA(cudaStreamSynchronize(L.stream));
    const int niter = R.pr->niter;
    for (int64_t i = 0; i < L.p_count; ++i) {
        const int64_t s = L.p_first + i;
        const int st = L.h_stop[i];
        const int nit = st < 0 ? 0 : (st > 0 ? st : niter);
        if (R.niter_out) R.niter_out[s] = nit;
        const double* S = L.h_S + i * (niter + 1);
        double last = 0.0;
        for (int k = 0; k < niter; ++k) {
            double c = std::numeric_limits<double>::quiet_NaN();
            if (k < nit) { const double d = S[k + 1] - S[k]; c = (d * d) / (S[k + 1] * S[k + 1]); last = c; }
            if (R.costs_out) R.costs_out[s * niter + k] = c;
        }
        if (R.cost_out) R.cost_out[s] = last;
    }
    L.pending = false;
}

void process_chunk(RunCtx& R, Lane& L, int64_t first, int64_t count) {
    p3d_plan* P = R.P;
    const p3d_pocs_params& pr = *R.pr;
    const int niter = pr.niter;
    const int64_t ne = (int64_t)P->n1 * P->n2;
    cudaStream_t st = L.stream;

    const Cx<float>* D;
    if (R.x_mem == P3D_MEM_HOST) {
        P3D_CUDA(cudaMemcpyAsync(L.D, R.x + first * ne, sizeof(Cx<float>) * ne * count, cudaMemcpyHostToDevice, st));
        D = L.D;
    } else {
        D = R.x + first * ne;
    }
    Cx<float>* OUT = (R.out_mem == P3D_MEM_HOST || R.schedule_only) ? L.OUT : R.out + first * ne;

    P3D_CUDA(cudaMemsetAsync(L.S, 0, sizeof(double) * count * (niter + 1), st));
    P3D_CUDA(cudaMemsetAsync(L.stop, 0, sizeof(int) * count, st));
    P3D_CUDA(cudaMemsetAsync(L.stats, 0, sizeof(SliceStats) * count, st));
    // minabs starts at +inf bits
    {
        std::vector<SliceStats> init((size_t)count);
        memset(init.data(), 0, sizeof(SliceStats) * count);
        for (auto& s : init) s.minabs_bits = 0x7f800000u;
        memcpy(L.h_stats, init.data(), sizeof(SliceStats) * count);
        P3D_CUDA(cudaMemcpyAsync(L.stats, L.h_stats, sizeof(SliceStats) * count, cudaMemcpyHostToDevice, st));
    }

    PocsGeom& G = P->geom;
    G.slices_per_mask = (int)std::min<int64_t>(R.spm, 0x7fffffff);

    BandArgs<float> A;
    memset(&A, 0, sizeof(A));
    A.mask = R.dmask; A.mbits = R.dmbits; A.niter = niter; A.eps = pr.eps; A.alpha = (float)pr.alpha;
    A.inv_n = (float)(1.0 / ((double)P->n1 * (double)P->n2));
    A.filt = R.filt;

    if (R.filt) {
        // kx-ky filter mode: out = ifft2(filt * fft2(x)) with the three fused passes of one POCS iteration
        // (row FFT | column FFT * filt, column IFFT | row IFFT / (N1 N2)); alpha = 0 and an all-zero mask
        // turn the re-insertion into the plain scaling.
        const int64_t band_max = 32768;
        for (int64_t b0 = 0; b0 < count; b0 += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, count - b0);
            BandArgs<float> B = A;
            B.W = L.W + b0 * ne; B.D = D + b0 * ne; B.OUT = OUT + b0 * ne; B.first_slice = 0;
            B.tau = L.tau + b0 * niter; B.S = L.S + b0 * (niter + 1); B.stop = L.stop + b0; B.stats = L.stats + b0;
            B.adaptive = 0; B.accum = 0; B.store_x0 = 0;
            launch_rows_init(P, st, B, nb);
            B.k = 0; B.last = 1; B.write_out = 1;
            launch_cols_iter(P, st, B, nb, P3D_OP_FILTER);
            launch_rows_iter(P, st, B, nb);
        }
        if (R.out_mem == P3D_MEM_HOST)
            P3D_CUDA(cudaMemcpyAsync(R.out + first * ne, OUT, sizeof(Cx<float>) * ne * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_S, L.S, sizeof(double) * count * (niter + 1), cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_stop, L.stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
        L.pending = true; L.p_first = first; L.p_count = count;
        return;
    }
    const bool data_driven = pr.thresh_model == P3D_MODEL_DATA_DRIVEN;
    const bool adaptive = pr.version == P3D_VERSION_ADAPTIVE;
    A.exact_tie = (pr.thresh_model == P3D_MODEL_INVERSE_PROPORTIONAL || data_driven || pr.decay_factors) ? 1 : 0;
    const int64_t band_max = 32768;

    auto band_args = [&](int64_t b0) {
        BandArgs<float> B = A;
        B.W = L.W + b0 * ne; B.D = D + b0 * ne; B.OUT = OUT + b0 * ne;
        B.first_slice = first + b0;
        B.tau = L.tau + b0 * niter; B.S = L.S + b0 * (niter + 1); B.stop = L.stop + b0; B.stats = L.stats + b0;
        return B;
    };

    // ---- setup: row FFT of d, statistics of X0 -----------------------------------------------
    for (int64_t b0 = 0; b0 < count; b0 += band_max) {
        const int nb = (int)std::min<int64_t>(band_max, count - b0);
        BandArgs<float> B = band_args(b0);
        B.adaptive = 0; B.accum = 1; B.store_x0 = data_driven ? 1 : 0;
        launch_rows_init(P, st, B, nb);
        launch_cols_stats(P, st, B, nb);
    }
    P3D_CUDA(cudaMemcpyAsync(L.h_stats, L.stats, sizeof(SliceStats) * count, cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaStreamSynchronize(st));

    // ---- schedule ---------------------------------------------------------------------------------
    std::vector<cd> tau;
    std::vector<char> real_flags((size_t)count, 0);
    if (!data_driven) {
        for (int64_t i = 0; i < count; ++i) {
            bool is_real = false;
            host_schedule(pr, stats_f32(L.h_stats[i]), ne, tau, is_real);
            if (pr.sqrt_decay) apply_sqrt_decay(tau, is_real);
            for (int k = 0; k < niter; ++k) {
                L.h_tau[i * niter + k] = cmake<float>((float)tau[k].real(), (float)tau[k].imag());
                if (R.tau_out) { R.tau_out[((first + i) * niter + k) * 2] = tau[k].real(); R.tau_out[((first + i) * niter + k) * 2 + 1] = tau[k].imag(); }
            }
            L.h_stop[i] = L.h_stats[i].nnz == 0 ? -1 : 0;
        }
    } else {
        // order statistics of the candidates inside (tau_min, tau_max), per slice, on the device
        size_t need = 0;
        cub::DeviceRadixSort::SortKeysDescending(nullptr, need, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long)ne, 0, 64, st);
        if (need > P->cub_temp_bytes) {
            if (P->cub_temp) cudaFree(P->cub_temp);
            P->cub_temp = nullptr; P->cub_temp_bytes = 0;
            P3D_CUDA(cudaMalloc(&P->cub_temp, need)); P->cub_temp_bytes = need;
        }
        for (int64_t i = 0; i < count; ++i) {
            L.h_stop[i] = L.h_stats[i].nnz == 0 ? -1 : 0;
            if (L.h_stop[i]) continue;
            cd tmin, tmax;
            schedule_bounds(pr, stats_f32(L.h_stats[i]), ne, tmin, tmax);
            const unsigned long long lo = lex_key((float)tmin.real(), (float)tmin.imag());
            const unsigned long long hi = lex_key((float)tmax.real(), (float)tmax.imag());
            unsigned long long* keys = reinterpret_cast<unsigned long long*>(OUT + i * ne);
            unsigned long long* sorted = reinterpret_cast<unsigned long long*>(L.W + i * ne);
            prof_begin(P, st, 6);
            k_make_keys<<<std::min<long long>((ne + 255) / 256, 148 * 8), 256, 0, st>>>(keys, ne, lo, hi, L.stats + i);
            size_t tb = P->cub_temp_bytes;
            cub::DeviceRadixSort::SortKeysDescending(P->cub_temp, tb, keys, sorted, (long long)ne, 0, 64, st);
            k_pick_tau<<<1, 128, 0, st>>>(sorted, L.stats + i, L.tau + i * niter, niter);
            prof_end(P, st);
        }
        P3D_CUDA(cudaGetLastError());
        P3D_CUDA(cudaMemcpyAsync(L.h_tau, L.tau, sizeof(Cx<float>) * count * niter, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_stats, L.stats, sizeof(SliceStats) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < count; ++i) {
            if (L.h_stop[i]) continue;
            P3D_REQUIRE(L.h_stats[i].n_cand > 0, P3D_ERR_NUMERIC,
                        "data-driven schedule: no coefficient between tau_min and tau_max in slice %lld",
                        (long long)(first + i));
            tau.resize(niter);
            for (int k = 0; k < niter; ++k) tau[k] = cd(L.h_tau[i * niter + k].x, L.h_tau[i * niter + k].y);
            if (pr.sqrt_decay) apply_sqrt_decay(tau, false);
            for (int k = 0; k < niter; ++k) {
                L.h_tau[i * niter + k] = cmake<float>((float)tau[k].real(), (float)tau[k].imag());
                if (R.tau_out) { R.tau_out[((first + i) * niter + k) * 2] = tau[k].real(); R.tau_out[((first + i) * niter + k) * 2 + 1] = tau[k].imag(); }
            }
        }
    }
    if (R.schedule_only) { L.pending = false; return; }

    P3D_CUDA(cudaMemcpyAsync(L.tau, L.h_tau, sizeof(Cx<float>) * count * niter, cudaMemcpyHostToDevice, st));
    P3D_CUDA(cudaMemcpyAsync(L.stop, L.h_stop, sizeof(int) * count, cudaMemcpyHostToDevice, st));

    // all-zero slices return the input unchanged (functions/POCS.py:515-521)
    for (int64_t i = 0; i < count; ++i)
        if (L.h_stop[i] < 0)
            P3D_CUDA(cudaMemcpyAsync(OUT + i * ne, D + i * ne, sizeof(Cx<float>) * ne, cudaMemcpyDeviceToDevice, st));

    // W was used as sort scratch (data-driven) or must hold FFT of the adaptive input: redo the row pass
    if (data_driven || adaptive) {
        for (int64_t b0 = 0; b0 < count; b0 += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, count - b0);
            BandArgs<float> B = band_args(b0);
            B.adaptive = adaptive ? 1 : 0; B.accum = 0;
            launch_rows_init(P, st, B, nb);
        }
    }

    // ---- iterations, band by band -----------------------------------------------------------------
    int64_t band = P->band_slices > 0 ? P->band_slices : count;
    band = std::min<int64_t>(band, band_max);
    for (int64_t b0 = 0; b0 < count; b0 += band) {
        const int nb = (int)std::min<int64_t>(band, count - b0);
        BandArgs<float> B = band_args(b0);
        B.adaptive = adaptive ? 1 : 0;
        for (int k = 0; k < niter; ++k) {
            B.k = k; B.last = (k == niter - 1) ? 1 : 0;
            // x_k only has to reach OUT when iteration k can be the last one executed
            B.write_out = (B.last || (pr.eps > 0.0 && k >= 3)) ? 1 : 0;
            if (pr.thresh_percentile) percentile_thresholds(P, st, B, nb, k);
            launch_cols_iter(P, st, B, nb, pr.thresh_op);
            launch_rows_iter(P, st, B, nb);
        }
    }

    // ---- results ------------------------------------------------------------------------------------
    if (R.out_mem == P3D_MEM_HOST)
        P3D_CUDA(cudaMemcpyAsync(R.out + first * ne, OUT, sizeof(Cx<float>) * ne * count, cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaMemcpyAsync(L.h_S, L.S, sizeof(double) * count * (niter + 1), cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaMemcpyAsync(L.h_stop, L.stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
    L.pending = true; L.p_first = first; L.p_count = count;
}

int run_impl(p3d_plan* P, const p3d_pocs_params* pr, const void* x, int x_mem, const uint8_t* mask,
             int64_t spm, void* out, int out_mem, int64_t n_slices, int32_t* niter_out, double* cost_out,
             double* costs_out, double* tau_out, bool schedule_only, const float* filt = nullptr) {
    P3D_REQUIRE(P && pr && x, P3D_ERR_BAD_ARG, "null plan / params / x");
    P3D_REQUIRE(n_slices >= 0, P3D_ERR_BAD_ARG, "n_slices must be >= 0");
    P3D_REQUIRE(pr->niter >= 1, P3D_ERR_BAD_ARG, "niter must be >= 1 (got %d)", pr->niter);
    P3D_REQUIRE(pr->thresh_op >= 0 && pr->thresh_op <= 2, P3D_ERR_NOT_IMPLEMENTED, "unsupported thresh_op %d", pr->thresh_op);
    P3D_REQUIRE(pr->thresh_model >= 0 && pr->thresh_model <= 3, P3D_ERR_NOT_IMPLEMENTED, "unsupported thresh_model %d", pr->thresh_model);
    P3D_REQUIRE(pr->version >= 0 && pr->version <= 2, P3D_ERR_BAD_ARG, "unsupported version %d", pr->version);
    if (!schedule_only) P3D_REQUIRE(mask && out, P3D_ERR_BAD_ARG, "null mask / out");
    if (pr->thresh_percentile) {
        P3D_REQUIRE(pr->decay_factors && pr->p_max >= 0.0 && pr->p_max <= 100.0 && pr->p_min >= 0.0 && pr->p_min <= 100.0,
                    P3D_ERR_BAD_ARG, "Percentiles must be in the range [0, 100]");
        P3D_REQUIRE(pr->thresh_model == P3D_MODEL_LINEAR || pr->thresh_model == P3D_MODEL_EXPONENTIAL, P3D_ERR_NOT_IMPLEMENTED,
                    "percentile operators need a linear or exponential schedule of percentiles");
        P3D_REQUIRE(P->precision != 64, P3D_ERR_NOT_IMPLEMENTED, "percentile operators run in the fp32 path only");
    }
    if (n_slices == 0) return P3D_OK;
    if (spm <= 0) spm = n_slices;
    DeviceGuard guard(P->device);

    if (P->precision == 64 && !filt) {
        if (!P->f64) { P->f64 = f64_create(P->device, P->n1, P->n2, &P->ax1, &P->ax2, P->smem_optin); f64_install_spec(P->f64, P->spec_variant64); }
        f64_set_force_generic(P->f64, P->force_generic ? 1 : 0);
        if (P->lanes.empty()) P->lanes.resize(1);
        if (!P->lanes[0].stream) P3D_CUDA(cudaStreamCreateWithFlags(&P->lanes[0].stream, cudaStreamNonBlocking));
        const uint8_t* dmask = nullptr;
        if (!schedule_only) {
            const int64_t n_masks = (n_slices + spm - 1) / spm;
            ensure_mask(P, mask, n_masks * (int64_t)P->n1 * P->n2, x_mem, P->lanes[0].stream, &dmask);
        }
        return f64_run(P->f64, pr, (const Cx<float>*)x, x_mem, dmask, spm, (Cx<float>*)out, out_mem, n_slices, niter_out,
                       cost_out, costs_out, tau_out, schedule_only, P->max_slices);
    }
    const bool host_in = x_mem == P3D_MEM_HOST, host_out = (out_mem == P3D_MEM_HOST) || schedule_only;
    // host data: four lanes (streams + buffer sets) rotate, so that the D2H of chunk i and the H2D of chunk
    // i+4 hide behind the iterations of chunks i+1..i+3 even when the PCIe path is slow (8 ranks sharing
    // one host: 8-11 GB/s per direction measured; e2e at 8 GPUs 444k -> 506k slice-it/s going from 2 to 4
    // lanes, 76.6k -> 80.9k on one GPU); device-resident data need one lane.
    // (the percentile operators share one scratch area and sort inside the iteration loop: one lane)
    const int lanes = pr->thresh_percentile ? 1 : (P->n_lanes > 0 ? P->n_lanes : ((host_in || host_out) ? 4 : 1));
    if ((int)P->lanes.size() < lanes) P->lanes.resize(lanes);
    const int nbuf = 1 + (host_in ? 1 : 0) + (host_out ? 1 : 0);
    for (auto& L : P->lanes) L.pending = false;
    // chunk size: everything at once on one lane; with two lanes at least two chunks per lane
    // (copy/compute overlap) but never tiny chunks
    int64_t want = n_slices;
    if (lanes > 1) {
        // ~4 chunks per lane keep the un-overlapped first H2D / last D2H short; chunks stay large
        // enough (>= ~48 MB of slices) to fill the GPU
        const int64_t ne_ = (int64_t)P->n1 * P->n2;
        const int64_t min_chunk = std::max<int64_t>(1, std::min<int64_t>(n_slices, (int64_t)(48e6 / (8.0 * (double)ne_)) + 1));
        // full-size chunks: ~2 per lane; the first and last chunks of a call are shorter (ramp below), so the
        // un-overlapped first H2D / last D2H stay short while the bulk of the launches is large enough to keep
        // the tail of a launch (its last, partially filled wave of CTAs) small
        want = std::max<int64_t>(min_chunk, (n_slices + 2 * lanes - 1) / (2 * lanes));
    }
    if (P->max_slices > 0) want = std::min<int64_t>(want, P->max_slices);
    int64_t have = 0;
    for (int i = 0; i < lanes; ++i) have = (i == 0) ? P->lanes[i].cap : std::min(have, P->lanes[i].cap);
    int64_t cap = want;
    if (have < want) {
        for (auto& L : P->lanes) { cudaStream_t st = L.stream; L.stream = nullptr; free_lane(L); L.stream = st; }
        cap = auto_capacity(P, want, nbuf, lanes);
    }
    for (int i = 0; i < lanes; ++i) ensure_lane(P, P->lanes[i], cap, pr->niter, host_in, host_out);

    RunCtx R;
    R.P = P; R.pr = pr; R.x = (const Cx<float>*)x; R.x_mem = x_mem; R.spm = spm;
    R.out = (Cx<float>*)out; R.out_mem = out_mem; R.niter_out = niter_out; R.cost_out = cost_out;
    R.costs_out = costs_out; R.tau_out = tau_out; R.schedule_only = schedule_only;
    R.dmask = nullptr; R.dmbits = nullptr; R.filt = filt;
    if (!schedule_only) {
        const int64_t n_masks = (n_slices + spm - 1) / spm;
        ensure_mask(P, mask, n_masks * (int64_t)P->n1 * P->n2, x_mem, P->lanes[0].stream, &R.dmask);
        R.dmbits = pack_mask(P, R.dmask, n_masks, P->lanes[0].stream);
    }

    // chunk sizes: cap/4, cap/2, cap, ..., cap, cap/2, cap/4 (host data on several lanes), else cap
    std::vector<int64_t> sizes;
    if (lanes > 1 && cap >= 8 && n_slices >= 3 * cap) {
        const int64_t ramp[2] = {std::max<int64_t>(1, cap / 4), std::max<int64_t>(1, cap / 2)};
        int64_t left = n_slices - 2 * (ramp[0] + ramp[1]);
        sizes.push_back(ramp[0]); sizes.push_back(ramp[1]);
        while (left > 0) { const int64_t c = std::min<int64_t>(cap, left); sizes.push_back(c); left -= c; }
        sizes.push_back(ramp[1]); sizes.push_back(ramp[0]);
    } else {
        for (int64_t first = 0; first < n_slices; first += cap) sizes.push_back(std::min<int64_t>(cap, n_slices - first));
    }
    int li = 0;
    int64_t first = 0;
    for (const int64_t count : sizes) {
        Lane& L = P->lanes[li];
        collect_lane(R, L);
        process_chunk(R, L, first, count);
        first += count;
        li = (li + 1) % lanes;
    }
    for (int i = 0; i < lanes; ++i) collect_lane(R, P->lanes[i]);
    for (int i = 0; i < lanes; ++i) P3D_CUDA(cudaStreamSynchronize(P->lanes[i].stream));
    prof_collect(P);
    return P3D_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
#define P3D_TRY try {
#define P3D_CATCH } catch (const P3dFail& f) { return f.code; } catch (const std::exception& e) { set_error("%s", e.what()); return P3D_ERR_CUDA; }

extern "C" {

int p3d_abi_version(void) { return P3D_ABI_VERSION; }
const char* p3d_last_error(void) { return get_error(); }
int p3d_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

int p3d_plan_create(p3d_plan** plan, int device, int n_iline, int n_xline, int64_t max_slices, int band_slices) {
    P3D_TRY
    P3D_REQUIRE(plan, P3D_ERR_BAD_ARG, "plan pointer is null");
    P3D_REQUIRE(n_iline >= 1 && n_xline >= 1, P3D_ERR_BAD_ARG, "slice shape must be positive (got %d x %d)", n_iline, n_xline);
    int ndev = 0;
    P3D_CUDA(cudaGetDeviceCount(&ndev));
    P3D_REQUIRE(device >= 0 && device < ndev, P3D_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard(device);
    p3d_plan* P = new p3d_plan();
    try {
        P->device = device; P->n1 = n_iline; P->n2 = n_xline; P->max_slices = max_slices; P->band_slices = band_slices;
        cudaDeviceProp prop;
        P3D_CUDA(cudaGetDeviceProperties(&prop, device));
        P->sm_count = prop.multiProcessorCount;
        P->smem_optin = prop.sharedMemPerBlockOptin;
        P->ax1.build(n_iline);
        P->ax2.build(n_xline);
        choose_geometry(P);
        { cudaError_t ce = generic_configure(generic_cfg(P)); P3D_CUDA(ce); }
        install_spec(P, 0);
    } catch (...) { p3d_plan_destroy(P); throw; }
    *plan = P;
    return P3D_OK;
    P3D_CATCH
}

int p3d_plan_destroy(p3d_plan* P) {
    if (!P) return P3D_OK;
    DeviceGuard guard(P->device);
    for (auto& L : P->lanes) { free_lane(L); if (L.stream) cudaStreamDestroy(L.stream); }
    P->ax1.release(); P->ax2.release();
    if (P->d_mask) cudaFree(P->d_mask);
    if (P->d_mbits) cudaFree(P->d_mbits);
    if (P->d_zero_mask) cudaFree(P->d_zero_mask);
    if (P->d_filt) cudaFree(P->d_filt);
    if (P->f64) f64_destroy(P->f64);
    if (P->spec_tw_cols) cudaFree(P->spec_tw_cols);
    if (P->spec_tw_rows) cudaFree(P->spec_tw_rows);
    if (P->cub_temp) cudaFree(P->cub_temp);
    if (P->pct_scr) cudaFree(P->pct_scr);
    if (P->pct_keys) cudaFree(P->pct_keys);
    if (P->pct_sorted) cudaFree(P->pct_sorted);
    for (auto& e : P->ev) if (e) cudaEventDestroy(e);
    delete P;
    return P3D_OK;
}

int p3d_pocs_run(p3d_plan* plan, const p3d_pocs_params* params, const void* x, int x_mem, const uint8_t* mask,
                 int64_t slices_per_mask, void* out, int out_mem, int64_t n_slices, int32_t* niter_out,
                 double* cost_out, double* costs_out) {
    P3D_TRY
    return run_impl(plan, params, x, x_mem, mask, slices_per_mask, out, out_mem, n_slices, niter_out, cost_out,
                    costs_out, nullptr, false);
    P3D_CATCH
}

int p3d_pocs_schedule(p3d_plan* plan, const p3d_pocs_params* params, const void* x, int x_mem, int64_t n_slices,
                      double* tau_out) {
    P3D_TRY
    P3D_REQUIRE(tau_out, P3D_ERR_BAD_ARG, "tau_out is null");
    return run_impl(plan, params, x, x_mem, nullptr, n_slices, nullptr, P3D_MEM_HOST, n_slices, nullptr, nullptr,
                    nullptr, tau_out, true);
    P3D_CATCH
}

int p3d_kxky_filter_run(p3d_plan* P, const void* x, int x_mem, const float* filt, int filt_mem, void* out, int out_mem,
                        int64_t n_slices) {
    P3D_TRY
    P3D_REQUIRE(P && x && filt && out, P3D_ERR_BAD_ARG, "null argument");
    if (n_slices == 0) return P3D_OK;
    const int64_t ne = (int64_t)P->n1 * P->n2;
    const float* dfilt = filt;
    {
        DeviceGuard guard(P->device);
        if (!P->d_zero_mask) { P3D_CUDA(cudaMalloc(&P->d_zero_mask, ne)); P3D_CUDA(cudaMemset(P->d_zero_mask, 0, ne)); }
        if (filt_mem == P3D_MEM_HOST) {
            if (!P->d_filt) P3D_CUDA(cudaMalloc(&P->d_filt, sizeof(float) * ne));
            P3D_CUDA(cudaMemcpy(P->d_filt, filt, sizeof(float) * ne, cudaMemcpyHostToDevice));
            dfilt = P->d_filt;
        }
    }
    p3d_pocs_params pr;
    memset(&pr, 0, sizeof(pr));
    pr.niter = 1; pr.thresh_op = P3D_OP_HARD; pr.thresh_model = P3D_MODEL_EXPONENTIAL; pr.q = 1.0; pr.alpha = 0.0; pr.p_max = 0.99; pr.p_min = 1e-5;
    // the zero mask lives on the device: tell run_impl so by passing it with a device-resident x, or stage it
    // through ensure_mask's host path (x_mem host => the mask pointer must be host memory)
    std::vector<uint8_t> hzero;
    const uint8_t* mask = P->d_zero_mask;
    if (x_mem == P3D_MEM_HOST) { hzero.assign((size_t)ne, 0); mask = hzero.data(); }
    return run_impl(P, &pr, x, x_mem, mask, n_slices, out, out_mem, n_slices, nullptr, nullptr, nullptr, nullptr, false, dfilt);
    P3D_CATCH
}

int p3d_fft2(p3d_plan* P, const void* x, int x_mem, void* out, int out_mem, int64_t n_slices, int inverse) {
    P3D_TRY
    P3D_REQUIRE(P && x && out, P3D_ERR_BAD_ARG, "null argument");
    if (n_slices == 0) return P3D_OK;
    DeviceGuard guard(P->device);
    const int64_t ne = (int64_t)P->n1 * P->n2;
    const size_t bytes = sizeof(Cx<float>) * ne * n_slices;
    Cx<float>* din = nullptr; Cx<float>* dout = nullptr;
    cudaStream_t st = nullptr;
    if (x_mem == P3D_MEM_HOST) { P3D_CUDA(cudaMalloc(&din, bytes)); P3D_CUDA(cudaMemcpy(din, x, bytes, cudaMemcpyHostToDevice)); }
    else din = (Cx<float>*)x;
    if (out_mem == P3D_MEM_HOST) P3D_CUDA(cudaMalloc(&dout, bytes)); else dout = (Cx<float>*)out;
    for (int64_t b0 = 0; b0 < n_slices; b0 += 32768) {
        const int nb = (int)std::min<int64_t>(32768, n_slices - b0);
        prof_begin(P, st, 7);
        generic_fft2(generic_cfg(P), P->ax1.dev(), P->ax2.dev(), din + b0 * ne, dout + b0 * ne, nb, inverse, st);
        prof_end(P, st);
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaDeviceSynchronize());
    if (out_mem == P3D_MEM_HOST) { P3D_CUDA(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost)); cudaFree(dout); }
    if (x_mem == P3D_MEM_HOST) cudaFree(din);
    prof_collect(P);
    return P3D_OK;
    P3D_CATCH
}

int p3d_plan_set_profiling(p3d_plan* P, int enabled) { if (!P) return P3D_ERR_BAD_ARG; P->profiling = enabled != 0; return P3D_OK; }

int p3d_plan_get_profile(p3d_plan* P, double* ms, int64_t* n, int reset) {
    if (!P) return P3D_ERR_BAD_ARG;
    for (int i = 0; i < P3D_PROFILE_KINDS; ++i) { if (ms) ms[i] = P->prof_ms[i]; if (n) n[i] = P->prof_n[i]; }
    if (reset) for (int i = 0; i < P3D_PROFILE_KINDS; ++i) { P->prof_ms[i] = 0; P->prof_n[i] = 0; }
    return P3D_OK;
}

int p3d_plan_event_record(p3d_plan* P, int slot) {
    P3D_TRY
    P3D_REQUIRE(P && slot >= 0 && slot < 8, P3D_ERR_BAD_ARG, "bad event slot");
    DeviceGuard guard(P->device);
    if (P->lanes.empty()) P->lanes.resize(1);
    if (!P->lanes[0].stream) P3D_CUDA(cudaStreamCreateWithFlags(&P->lanes[0].stream, cudaStreamNonBlocking));
    if (!P->ev[slot]) P3D_CUDA(cudaEventCreate(&P->ev[slot]));
    P3D_CUDA(cudaEventRecord(P->ev[slot], P->lanes[0].stream));
    return P3D_OK;
    P3D_CATCH
}

int p3d_plan_event_elapsed_ms(p3d_plan* P, int a, int b, double* ms) {
    P3D_TRY
    P3D_REQUIRE(P && ms && a >= 0 && a < 8 && b >= 0 && b < 8 && P->ev[a] && P->ev[b], P3D_ERR_BAD_ARG, "bad event slots");
    DeviceGuard guard(P->device);
    P3D_CUDA(cudaEventSynchronize(P->ev[b]));
    float f = 0.f;
    P3D_CUDA(cudaEventElapsedTime(&f, P->ev[a], P->ev[b]));
    *ms = f;
    return P3D_OK;
    P3D_CATCH
}

int p3d_plan_describe(p3d_plan* P, char* buf, int64_t buflen) {
    if (!P || !buf || buflen <= 0) return P3D_ERR_BAD_ARG;
    std::string s = "iline axis: " + P->ax1.describe() + "; xline axis: " + P->ax2.describe();
    s += "; generic tiles: C=" + std::to_string(P->geom.C) + " cols x " + std::to_string(P->col_threads) + " thr, smem " +
         std::to_string(P->col_smem) + " B; RB=" + std::to_string(P->geom.RB) + " rows x " + std::to_string(P->row_threads) +
         " thr, smem " + std::to_string(P->row_smem) + " B";
    s += std::string("; cols_iter=") + ((P->spec.cols_iter && !P->force_generic) ? P->spec.cols_name : "generic");
    s += std::string("; rows_iter=") + ((P->spec.rows_iter && !P->force_generic) ? P->spec.rows_name : "generic");
    s += "; precision=" + std::to_string(P->precision);
    if (P->precision == 64) {
        const SpecKernels64 k64 = select_spec_kernels64(P->n1, P->n2, P->spec_variant64);
        s += std::string("; cols_iter64=") + ((k64.cols_iter && !P->force_generic) ? k64.cols_name : "generic64");
        s += std::string("; rows_iter64=") + ((k64.rows_iter && !P->force_generic) ? k64.rows_name : "generic64");
    }
    s += "; band_slices=" + std::to_string(P->band_slices) + "; sms=" + std::to_string(P->sm_count);
    snprintf(buf, (size_t)buflen, "%s", s.c_str());
    return P3D_OK;
}

int p3d_plan_set_option(p3d_plan* P, const char* key, int64_t value) {
    if (!P || !key) return P3D_ERR_BAD_ARG;
    if (!strcmp(key, "band_slices")) P->band_slices = (int)value;
    else if (!strcmp(key, "force_generic")) P->force_generic = value != 0;
    else if (!strcmp(key, "lanes")) P->n_lanes = (int)value;
    else if (!strcmp(key, "precision")) {
        if (value != 32 && value != 64) { set_error("precision must be 32 or 64"); return P3D_ERR_BAD_ARG; }
        P->precision = (int)value;
    }
    else if (!strcmp(key, "spec_variant")) {
        try { DeviceGuard g(P->device); install_spec(P, (int)value); } catch (const P3dFail& f) { return f.code; }
    }
    else if (!strcmp(key, "spec_variant64")) {
        P->spec_variant64 = (int)value;
        if (P->f64) { try { DeviceGuard g(P->device); f64_install_spec(P->f64, (int)value); } catch (const P3dFail& f) { return f.code; } }
    }
    else if (!strcmp(key, "max_slices")) { P->max_slices = value; for (auto& L : P->lanes) { cudaStream_t st = L.stream; L.stream = nullptr; free_lane(L); L.stream = st; } }
    else { set_error("unknown option %s", key); return P3D_ERR_BAD_ARG; }
    return P3D_OK;
}

int p3d_host_alloc(void** ptr, int64_t bytes) {
    P3D_TRY
    P3D_REQUIRE(ptr && bytes >= 0, P3D_ERR_BAD_ARG, "bad argument");
    P3D_CUDA(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1)));
    return P3D_OK;
    P3D_CATCH
}
int p3d_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); return P3D_OK; }
int p3d_device_alloc(int device, void** ptr, int64_t bytes) {
    P3D_TRY
    DeviceGuard g(device);
    P3D_CUDA(cudaMalloc(ptr, (size_t)std::max<int64_t>(bytes, 1)));
    return P3D_OK;
    P3D_CATCH
}
int p3d_device_free(int device, void* ptr) { DeviceGuard g(device); if (ptr) cudaFree(ptr); return P3D_OK; }
int p3d_memcpy(int device, void* dst, const void* src, int64_t bytes, int kind) {
    P3D_TRY
    DeviceGuard g(device);
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    P3D_CUDA(cudaMemcpy(dst, src, (size_t)bytes, k));
    return P3D_OK;
    P3D_CATCH
}
int p3d_device_synchronize(int device) {
    P3D_TRY
    DeviceGuard g(device);
    P3D_CUDA(cudaDeviceSynchronize());
    return P3D_OK;
    P3D_CATCH
}

}  // extern "C"
