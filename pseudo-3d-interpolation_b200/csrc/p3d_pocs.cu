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
#include <string>
#include <thread>

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
    // escalating-precision mode: complex128 state of the slices that left the fp32 path, per-slice switch flags
    bool has_esc = false;
    Cx<double>* W64 = nullptr; Cx<double>* tau64 = nullptr; Cx<double>* h_tau64 = nullptr;
    int* esc = nullptr; int* h_esc = nullptr;
    float* guard = nullptr; float* h_guard = nullptr;
    int* list = nullptr; int* h_list = nullptr;
    double2* cand = nullptr; int cand_stride = 0;
    unsigned* arena = nullptr; int* acnt = nullptr; int* astart = nullptr; int* h_astart = nullptr; double2* yval = nullptr;
    int* wflag = nullptr; int* kfail = nullptr; int* kend = nullptr; int* h_wflag = nullptr; int* h_kfail = nullptr; int* h_kend = nullptr;
    int arena_cap = 0;
    int64_t n_escalated = 0, n_esc_iters = 0, n_verify_failed = 0;
    unsigned long long* dd_keys = nullptr; unsigned int* dd_vals = nullptr; int64_t dd_ne = 0;   // exact data-driven schedule scratch
    // CUB scratch of the data-driven schedule (per lane: lanes sort concurrently on their own streams)
    void* cub_temp = nullptr; size_t cub_temp_bytes = 0;
    std::vector<EventPair> events;      // profiling events of this lane's launches
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
    int precision = 0;               // 0: escalating (fp32 state, complex128 once a coefficient enters the guard band of the
                                     // threshold), 32: fp32 only, 64: float64 state mode (p3d_pocs_f64.cu)
    double guard_factor = 1024.0;    // guard half-width in units of eps32 * rms|X| (escalating mode)
    double watch_guard_factor = 4096.0;   // the same where a hit is only recorded and verified (watch list): a wider band costs
                                          // a few more replayed entries, not an earlier switch (fp32 deviation of coefficients
                                          // near the threshold measured on config 2: up to 260 units late in the pilot)
    int seg_iters = 4;               // escalating mode: iterations between two compactions of the active-slice list
    int arena_cap = 16384;           // escalating mode: support-record entries per slice (all pilot iterations together)
    int64_t pilot_min_elems = 0;     // escalating mode: slices smaller than this run in complex128 from the first iteration (0: every
                                     // size has a pilot; with the one-launch replay it pays even for 200 x 200 slices: +3.5 %)
    int fused_replay_max = 1024;     // largest per-iteration support (rounded up to a power of two) replayed by the one-launch kernel
    int debug_fail_iter = -1;        // testing: the replay reports a verification failure at this iteration for every slice it replays
    int use_tma = 1;                 // column tiles fetched with cp.async.bulk.tensor where the tile shape allows it
    int watch_mode = -1;             // escalating mode: guard-band hits are verified by the float64 replay instead of freezing the slice
                                     // (-1 = where it pays: slices of 400 k points and more; small slices freeze on the first hit -
                                     // their replay launches cost more than the complex128 iterations they would save)
    int support_cap = 0;             // escalating mode: largest support replayed per iteration (0 = 2.2 sqrt(n1 n2): where one more
                                     // replayed iteration, |S|^2 gathers at the measured 2.4 ps each, costs what the complex128
                                     // iteration it replaces does, 11.7 us per 10^6 points more than an fp32 one)
    Cx<double>* mhat = nullptr; int64_t mhat_masks = 0;          // fft2 of the mask planes in complex128 (exact restart)
    Cx<float>* mask_c64 = nullptr; SliceStats* mh_stats = nullptr; double2* mh_cand = nullptr; int mh_cand_stride = 0;
    int64_t n_escalated = 0, n_esc_iters = 0;   // statistics of the last run (escalating mode)
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

void prof_begin(p3d_plan* P, std::vector<EventPair>& ev, cudaStream_t st, int kind) {
    if (!P->profiling) return;
    EventPair e; e.kind = kind;
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
    cudaEventRecord(e.a, st);
    ev.push_back(e);
}
void prof_end(p3d_plan* P, std::vector<EventPair>& ev, cudaStream_t st) {
    if (!P->profiling) return;
    cudaEventRecord(ev.back().b, st);
}
void prof_collect_vec(p3d_plan* P, std::vector<EventPair>& ev) {
    for (auto& e : ev) {
        float ms = 0.f;
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            P->prof_ms[e.kind] += ms; P->prof_n[e.kind] += 1;
        }
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    ev.clear();
}
void prof_collect(p3d_plan* P) {
    prof_collect_vec(P, P->events);
    for (auto& L : P->lanes) prof_collect_vec(P, L.events);
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
    if (L.W64) cudaFree(L.W64);
    if (L.tau64) cudaFree(L.tau64);
    if (L.esc) cudaFree(L.esc);
    if (L.guard) cudaFree(L.guard);
    if (L.list) cudaFree(L.list);
    if (L.cand) cudaFree(L.cand);
    if (L.arena) cudaFree(L.arena);
    if (L.acnt) cudaFree(L.acnt);
    if (L.astart) cudaFree(L.astart);
    if (L.yval) cudaFree(L.yval);
    if (L.h_astart) cudaFreeHost(L.h_astart);
    for (int* p : {L.wflag, L.kfail, L.kend}) if (p) cudaFree(p);
    for (int* p : {L.h_wflag, L.h_kfail, L.h_kend}) if (p) cudaFreeHost(p);
    L.wflag = L.kfail = L.kend = L.h_wflag = L.h_kfail = L.h_kend = nullptr;
    L.arena = nullptr; L.acnt = nullptr; L.astart = nullptr; L.yval = nullptr; L.h_astart = nullptr; L.arena_cap = 0;
    if (L.h_tau64) cudaFreeHost(L.h_tau64);
    if (L.h_esc) cudaFreeHost(L.h_esc);
    if (L.h_guard) cudaFreeHost(L.h_guard);
    if (L.h_list) cudaFreeHost(L.h_list);
    if (L.cub_temp) cudaFree(L.cub_temp);
    if (L.dd_keys) cudaFree(L.dd_keys);
    if (L.dd_vals) cudaFree(L.dd_vals);
    L.dd_keys = nullptr; L.dd_vals = nullptr; L.dd_ne = 0;
    L.cub_temp = nullptr; L.cub_temp_bytes = 0; L.cand_stride = 0;
    L.W64 = L.tau64 = L.h_tau64 = nullptr; L.esc = L.h_esc = L.list = L.h_list = nullptr; L.guard = L.h_guard = nullptr; L.cand = nullptr;
    L.has_esc = false;
    L.W = L.D = L.OUT = L.tau = nullptr; L.S = nullptr; L.stop = nullptr; L.stats = nullptr;
    L.h_tau = nullptr; L.h_S = nullptr; L.h_stop = nullptr; L.h_stats = nullptr;
    L.cap = 0; L.niter_cap = 0; L.own_d = L.own_out = false;
}

void ensure_lane(p3d_plan* P, Lane& L, int64_t cap, int niter, bool need_d, bool need_out, bool need_esc, int cand_stride) {
    const int64_t ne = (int64_t)P->n1 * P->n2;
    if (!L.stream) P3D_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    const bool regrow = cap > L.cap || niter > L.niter_cap || (need_d && !L.own_d) || (need_out && !L.own_out) ||
                        (need_esc && (!L.has_esc || cand_stride > L.cand_stride || P->arena_cap != L.arena_cap));
    if (!regrow) return;
    cudaStream_t st = L.stream;
    cap = std::max(cap, L.cap); niter = std::max(niter, L.niter_cap);
    need_d = need_d || L.own_d; need_out = need_out || L.own_out; need_esc = need_esc || L.has_esc;
    cand_stride = std::max(cand_stride, L.cand_stride);
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
    P3D_CUDA(cudaMalloc(&L.list, sizeof(int) * cap));
    P3D_CUDA(cudaMallocHost(&L.h_list, sizeof(int) * cap));
    if (need_esc) {
        P3D_CUDA(cudaMalloc(&L.W64, sizeof(Cx<double>) * ne * cap));
        P3D_CUDA(cudaMalloc(&L.tau64, sizeof(Cx<double>) * cap * niter));
        P3D_CUDA(cudaMalloc(&L.esc, sizeof(int) * cap));
        P3D_CUDA(cudaMalloc(&L.guard, sizeof(float) * cap));
        P3D_CUDA(cudaMalloc(&L.cand, sizeof(double2) * cap * std::max(1, cand_stride)));
        P3D_CUDA(cudaMallocHost(&L.h_tau64, sizeof(Cx<double>) * cap * niter));
        P3D_CUDA(cudaMallocHost(&L.h_esc, sizeof(int) * cap));
        P3D_CUDA(cudaMallocHost(&L.h_guard, sizeof(float) * cap));
        L.arena_cap = P->arena_cap;
        P3D_CUDA(cudaMalloc(&L.arena, sizeof(unsigned) * cap * L.arena_cap));
        P3D_CUDA(cudaMalloc(&L.yval, sizeof(double2) * cap * L.arena_cap));
        P3D_CUDA(cudaMalloc(&L.acnt, sizeof(int) * cap));
        P3D_CUDA(cudaMalloc(&L.astart, sizeof(int) * cap * (niter + 1)));
        P3D_CUDA(cudaMallocHost(&L.h_astart, sizeof(int) * cap * (niter + 1)));
        P3D_CUDA(cudaMalloc(&L.wflag, sizeof(int) * cap));
        P3D_CUDA(cudaMalloc(&L.kfail, sizeof(int) * cap));
        P3D_CUDA(cudaMalloc(&L.kend, sizeof(int) * cap));
        P3D_CUDA(cudaMallocHost(&L.h_wflag, sizeof(int) * cap));
        P3D_CUDA(cudaMallocHost(&L.h_kfail, sizeof(int) * cap));
        P3D_CUDA(cudaMallocHost(&L.h_kend, sizeof(int) * cap));
        L.has_esc = true; L.cand_stride = std::max(1, cand_stride);
    }
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
void launch_rows_init(p3d_plan* P, Lane& L, const BandArgs<float>& A, int nslices) {
    cudaStream_t st = L.stream;
    prof_begin(P, L.events, st, 0);
    if (P->spec.rows_init && !P->force_generic) P->spec.rows_init(P->geom, P->spec_tw_rows, A, nslices, st);
    else generic_rows_init(generic_cfg(P), P->ax2.dev(), A, nslices, st);
    prof_end(P, L.events, st);
    P3D_CUDA(cudaGetLastError());
}
void launch_cols_stats(p3d_plan* P, Lane& L, const BandArgs<float>& A, int nslices) {
    cudaStream_t st = L.stream;
    prof_begin(P, L.events, st, 1);
    if (P->spec.cols_stats && !P->force_generic) P->spec.cols_stats(P->geom, P->spec_tw_cols, A, nslices, st);
    else generic_cols_stats(generic_cfg(P), P->ax1.dev(), A, nslices, st);
    prof_end(P, L.events, st);
    P3D_CUDA(cudaGetLastError());
}
void launch_cols_iter(p3d_plan* P, Lane& L, const BandArgs<float>& A, int nslices, int op) {
    cudaStream_t st = L.stream;
    prof_begin(P, L.events, st, 2);
    if (P->spec.cols_iter && !P->force_generic) P->spec.cols_iter(P->geom, P->spec_tw_cols, A, nslices, op, st);
    else generic_cols_iter(generic_cfg(P), P->ax1.dev(), A, nslices, op, st);
    prof_end(P, L.events, st);
    P3D_CUDA(cudaGetLastError());
}
void launch_rows_iter(p3d_plan* P, Lane& L, const BandArgs<float>& A, int nslices) {
    cudaStream_t st = L.stream;
    prof_begin(P, L.events, st, 3);
    if (P->spec.rows_iter && !P->force_generic) P->spec.rows_iter(P->geom, P->spec_tw_rows, A, nslices, st);
    else generic_rows_iter(generic_cfg(P), P->ax2.dev(), A, nslices, st);
    prof_end(P, L.events, st);
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

// ---- data-driven schedule from the complex128 X0 (escalating mode): exact order statistics in numpy's ordering ----
// key = ordered real part of the candidates inside (tau_min, tau_max) (lexicographic bounds), 0 for everything else
__global__ void k_dd_keys64(const Cx<double>* __restrict__ X0, long long n, double lo_re, double lo_im, double hi_re, double hi_im,
                            unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals, SliceStats* st) {
    unsigned long long cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const Cx<double> v = X0[i];
        const bool above_lo = (lo_re < v.x) || (lo_re == v.x && lo_im < v.y);
        const bool below_hi = (v.x < hi_re) || (v.x == hi_re && v.y < hi_im);
        const bool in = above_lo && below_hi;
        keys[i] = in ? f64_ordered(v.x) : 0ull;
        vals[i] = (unsigned int)i;
        cnt += in ? 1ull : 0ull;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&st->n_cand, cnt);
}
// tau_k = V[ceil(k (Nv - 1) / (niter - 1))] of the descending lexicographic order V; equal real parts (Hermitian spectra
// of real slices) are ordered by their imaginary parts inside the run
__global__ void k_pick_tau64(const unsigned long long* __restrict__ skeys, const unsigned int* __restrict__ svals, const Cx<double>* __restrict__ X0,
                             const SliceStats* st, Cx<double>* tau64, int niter) {
    const long long nv = (long long)st->n_cand;
    for (int k = threadIdx.x; k < niter; k += blockDim.x) {
        if (nv == 0) { tau64[k] = cmake<double>(nan(""), nan("")); continue; }
        long long idx = 0;
        if (k > 0) idx = (long long)ceil((double)((long long)k * (nv - 1)) / (double)(niter - 1));
        if (idx > nv - 1) idx = nv - 1;
        const unsigned long long key = skeys[idx];
        long long a = idx, b = idx + 1;
        while (a > 0 && skeys[a - 1] == key) --a;
        while (b < nv && skeys[b] == key) ++b;
        Cx<double> pick = X0[svals[idx]];
        if (b - a > 1) {
            // rank (idx - a) by descending imaginary part inside the run of equal real parts
            const long long want = idx - a;
            for (long long c = a; c < b; ++c) {
                const double im = X0[svals[c]].y;
                long long larger = 0, equal_before = 0;
                for (long long e = a; e < b; ++e) {
                    const double ime = X0[svals[e]].y;
                    if (ime > im) ++larger;
                    else if (ime == im && e < c) ++equal_before;
                }
                if (larger + equal_before == want) { pick = X0[svals[c]]; break; }
            }
        }
        tau64[k] = pick;
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
    prof_begin(P, P->events, st, 6);
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
    prof_end(P, P->events, st);
    P3D_CUDA(cudaGetLastError());
}

struct RunCtx {
    p3d_plan* P; const p3d_pocs_params* pr;
    const Cx<float>* x; int x_mem; const uint8_t* dmask; const uint32_t* dmbits; int64_t spm;
    const uint32_t* dmbits64 = nullptr;     // packed mask of the complex128 row kernel (escalating mode)
    bool escalate = false;
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
            launch_rows_init(P, L, B, nb);
            B.k = 0; B.last = 1; B.write_out = 1;
            launch_cols_iter(P, L, B, nb, P3D_OP_FILTER);
            launch_rows_iter(P, L, B, nb);
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
        if (P->use_tma) { B.tma_base = L.W; B.tma_slices = L.cap; B.tma_slice0 = (int)b0; }
        B.tau = L.tau + b0 * niter; B.S = L.S + b0 * (niter + 1); B.stop = L.stop + b0; B.stats = L.stats + b0;
        return B;
    };

    // ---- setup: row FFT of d, statistics of X0 -----------------------------------------------
    for (int64_t b0 = 0; b0 < count; b0 += band_max) {
        const int nb = (int)std::min<int64_t>(band_max, count - b0);
        BandArgs<float> B = band_args(b0);
        B.adaptive = 0; B.accum = 1; B.store_x0 = data_driven ? 1 : 0;
        launch_rows_init(P, L, B, nb);
        launch_cols_stats(P, L, B, nb);
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
        if (need > L.cub_temp_bytes) {
            if (L.cub_temp) cudaFree(L.cub_temp);
            L.cub_temp = nullptr; L.cub_temp_bytes = 0;
            P3D_CUDA(cudaMalloc(&L.cub_temp, need)); L.cub_temp_bytes = need;
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
            prof_begin(P, L.events, st, 6);
            k_make_keys<<<std::min<long long>((ne + 255) / 256, 148 * 8), 256, 0, st>>>(keys, ne, lo, hi, L.stats + i);
            size_t tb = L.cub_temp_bytes;
            cub::DeviceRadixSort::SortKeysDescending(L.cub_temp, tb, keys, sorted, (long long)ne, 0, 64, st);
            k_pick_tau<<<1, 128, 0, st>>>(sorted, L.stats + i, L.tau + i * niter, niter);
            prof_end(P, L.events, st);
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
            launch_rows_init(P, L, B, nb);
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
            launch_cols_iter(P, L, B, nb, pr.thresh_op);
            launch_rows_iter(P, L, B, nb);
        }
    }

    // ---- results ------------------------------------------------------------------------------------
    if (R.out_mem == P3D_MEM_HOST)
        P3D_CUDA(cudaMemcpyAsync(R.out + first * ne, OUT, sizeof(Cx<float>) * ne * count, cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaMemcpyAsync(L.h_S, L.S, sizeof(double) * count * (niter + 1), cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaMemcpyAsync(L.h_stop, L.stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
    L.pending = true; L.p_first = first; L.p_count = count;
}

// =====================================================================================================
// Escalating-precision engine ("precision" = 0, the default).
//
// The reference iterates in float64.  An fp32 iterate follows that trajectory to ~1e-7 as long as every
// threshold decision is the same; once tau_k has sunk into the dense part of the spectrum a coefficient sits within
// fp32 rounding of the threshold sooner or later, one decision differs and the late iterations amplify the difference
// (1e-3 on configs 1 and 2).  Continuing in complex128 from the fp32 iterate is not enough either: the fp32 rounding the
// iterate has collected decays by only ~10 % per iteration, so it is still there when the thresholds reach the dense
// spectrum (measured: 20 of 1025 slices of config 2 off by more than 1e-4 even with a guard band of 4096 eps).
//
// So the fp32 iterations are only a PILOT that finds out which coefficients survive each threshold while the spectrum
// is sparse, and the float64 trajectory is rebuilt exactly from that:
//   * the column kernel tests every coefficient against a guard band | |X| - c_k | < G * eps32 * rms|X| around the
//     modulus c_k at which the operator jumps, and appends the index of every survivor to the slice's support record;
//   * the first guard hit (or a full record) in iteration k_e freezes the slice: iterations 0 .. k_e - 1 were decided
//     unambiguously, iteration k_e is left to complex128;
//   * with the supports S_i known the float64 iterates obey a linear recursion in the SPARSE domain,
//         X_i[j] = alpha X0[j] + y_{i-1}[j] - (alpha / N) sum_{s in S_{i-1}} mhat[j - s] y_{i-1}[s],   y_i = T_i(X_i) on S_i,
//     (mhat = fft2(mask), X0 = fft2(d), both in complex128): |S_i| |S_{i-1}| complex multiply-adds per iteration
//     instead of two 2-D FFTs (k_replay);
//   * y_{k_e - 1} is scattered into a zeroed spectrum, one complex128 inverse column pass + row pass rebuild
//     x_{k_e - 1} and its row FFT exactly, and the complex128 iteration kernels run iterations k_e .. niter - 1.
// The schedule needs the lexicographic maximum of X0 to double accuracy as well (tau = p * z), so the statistics pass
// runs in complex128 and leaves X0 in the complex128 work array for the replay.
// Phase 1 (fp32) runs in segments of seg_iters iterations over a compacted list of the slices still in fp32; phase 2
// sorts the frozen slices by k_e and launches iteration k over the prefix already switched.
// =====================================================================================================
__global__ void k_lexmax_finish(const double2* __restrict__ cand, int stride, int ntiles, SliceStats* stats) {
    __shared__ double red[64];
    const int s = blockIdx.x;
    double re = -INFINITY, im = -INFINITY;
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) {
        const double2 c = cand[(long long)s * stride + i];
        if (c.x > re || (c.x == re && c.y > im)) { re = c.x; im = c.y; }
    }
    block_lexmax(re, im, red);
    if (threadIdx.x == 0) { stats[s].re64_key = f64_ordered(re); stats[s].im64_key = f64_ordered(im); }
}

__global__ void k_mask_to_c64(const uint8_t* __restrict__ m, Cx<float>* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = cmake<float>(m[i] ? 1.f : 0.f, 0.f);
}

// ascending sort of every (slice, iteration) segment of the support record: the replay sums in a fixed order
__global__ void k_sort_segments(const int* __restrict__ list, const int* __restrict__ kend, unsigned* arena, const int* __restrict__ astart,
                                int acap, int niter) {
    extern __shared__ unsigned seg_sh[];
    const int s = list[blockIdx.y], i = blockIdx.x;
    if (i >= kend[s]) return;
    const int a0 = astart[(long long)s * (niter + 1) + i], n = astart[(long long)s * (niter + 1) + i + 1] - a0;
    if (n <= 1) return;
    int m = 1; while (m < n) m <<= 1;
    unsigned* seg = arena + (long long)s * acap + a0;
    for (int t = threadIdx.x; t < m; t += blockDim.x) seg_sh[t] = t < n ? seg[t] : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < m; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const unsigned a = seg_sh[t], b = seg_sh[p];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { seg_sh[t] = b; seg_sh[p] = a; }
                }
            }
            __syncthreads();
        }
    for (int t = threadIdx.x; t < n; t += blockDim.x) seg[t] = seg_sh[t];
}

// one iteration of the sparse-domain float64 recursion (see the header of this section)
template <int OP>
__global__ void __launch_bounds__(128)
k_replay(const int i, const int* __restrict__ list, const int* __restrict__ kend, int* __restrict__ kfail, const unsigned* __restrict__ arena,
         const int* __restrict__ astart, double2* __restrict__ yval, const int acap, const int niter,
         const Cx<double>* __restrict__ X0, const Cx<double>* __restrict__ mhat, const Cx<double>* __restrict__ tau,
         const long long first_slice, const int spm, const int n1, const int n2, const double alpha, const double inv_n, const int debug_fail) {
    __shared__ unsigned sp[128];
    __shared__ double2 sy[128];
    const int s = list[blockIdx.y];
    if (i >= kend[s]) return;
    const int* as = astart + (long long)s * (niter + 1);
    const int a0 = as[i], n = as[i + 1] - a0;
    if ((int)blockIdx.x * 128 >= n) return;
    const int p0 = i > 0 ? as[i - 1] : 0, np = i > 0 ? a0 - p0 : 0;
    const int t = blockIdx.x * 128 + threadIdx.x;
    const bool valid = t < n;
    const unsigned* ar = arena + (long long)s * acap;
    double2* yv = yval + (long long)s * acap;
    const unsigned ent = valid ? ar[a0 + t] : 0u;
    const bool watched = (ent >> 31) != 0u;           // killed by the pilot inside the guard band: verify, contributes nothing
    const unsigned pj = ent & 0x7fffffffu;
    const int jr = (int)(pj >> 16), jc = (int)(pj & 0xffffu);
    const long long ne = (long long)n1 * n2;
    const Cx<double> x0 = valid ? X0[(long long)s * ne + (long long)jr * n2 + jc] : cmake<double>(0.0, 0.0);
    const Cx<double>* __restrict__ mh = mhat + ((first_slice + s) / spm) * ne;
    double accx = 0.0, accy = 0.0, selfx = 0.0, selfy = 0.0;
    for (int q0 = 0; q0 < np; q0 += 128) {
        const int nq = min(128, np - q0);
        __syncthreads();
        if ((int)threadIdx.x < nq) { sp[threadIdx.x] = ar[p0 + q0 + threadIdx.x] & 0x7fffffffu; sy[threadIdx.x] = yv[p0 + q0 + threadIdx.x]; }
        __syncthreads();
        if (valid) {
            // eight gathers in flight per thread (the mhat plane lives in L2: latency, not bandwidth, bounds this loop)
            int q = 0;
            for (; q + 8 <= nq; q += 8) {
                double2 m[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned ps = sp[q + u];
                    int dr = jr - (int)(ps >> 16); if (dr < 0) dr += n1;
                    int dc = jc - (int)(ps & 0xffffu); if (dc < 0) dc += n2;
                    m[u] = __ldg(reinterpret_cast<const double2*>(mh) + (long long)dr * n2 + dc);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double2 ys = sy[q + u];
                    if (sp[q + u] == pj) { selfx = ys.x; selfy = ys.y; }
                    accx += m[u].x * ys.x - m[u].y * ys.y;
                    accy += m[u].x * ys.y + m[u].y * ys.x;
                }
            }
            for (; q < nq; ++q) {
                const unsigned ps = sp[q];
                const double2 ys = sy[q];
                if (ps == pj) { selfx = ys.x; selfy = ys.y; }
                int dr = jr - (int)(ps >> 16); if (dr < 0) dr += n1;
                int dc = jc - (int)(ps & 0xffffu); if (dc < 0) dc += n2;
                const double2 m = __ldg(reinterpret_cast<const double2*>(mh) + (long long)dr * n2 + dc);
                accx += m.x * ys.x - m.y * ys.y;
                accy += m.x * ys.y + m.y * ys.x;
            }
        }
    }
    if (!valid) return;
    Cx<double> X = x0;
    if (i > 0) X = cmake<double>(alpha * x0.x + selfx - alpha * inv_n * accx, alpha * x0.y + selfy - alpha * inv_n * accy);
    const Cx<double> tk = tau[(long long)s * niter + i];
    const double a = tk.x, b = tk.y;
    Cx<double> y = apply_threshold<OP, double>(X, a, b, a * a - b * b, 2.0 * a * b);
    // the pilot's decision against the exact one: a kept coefficient must survive the float64 threshold, a watched one
    // must not; the first iteration with a mismatch is where complex128 has to take over
    const bool survives = (y.x != 0.0) || (y.y != 0.0);
    if (survives == watched || i == debug_fail) atomicMin(&kfail[s], i);
    if (watched) y = cmake<double>(0.0, 0.0);
    yv[a0 + t] = make_double2(y.x, y.y);
}


// The whole replay of a slice in ONE launch (small supports: launch-bound otherwise): one CTA per slice walks through the
// pilot's iterations, keeps the previous support and its values in shared memory, sorts each iteration's record
// (fixed summation order) and verifies the pilot's decisions like k_replay.  Dynamic shared memory: 2 x maxseg x 20 bytes.
template <int OP>
__global__ void __launch_bounds__(256)
k_replay_fused(const int* __restrict__ list, const int* __restrict__ kend, int* __restrict__ kfail, unsigned* __restrict__ arena,
               const int* __restrict__ astart, double2* __restrict__ yval, const int acap, const int niter, const int maxseg,
               const Cx<double>* __restrict__ X0, const Cx<double>* __restrict__ mhat, const Cx<double>* __restrict__ tau,
               const long long first_slice, const int spm, const int n1, const int n2, const double alpha, const double inv_n,
               const int debug_fail) {
    extern __shared__ __align__(16) unsigned char fused_sh[];
    double2* ybuf = reinterpret_cast<double2*>(fused_sh);                          // [2][maxseg]
    unsigned* ibuf = reinterpret_cast<unsigned*>(fused_sh + (size_t)2 * maxseg * sizeof(double2));   // [2][maxseg]
    const int s = list[blockIdx.x];
    const int ke = kend[s];
    const int* as = astart + (long long)s * (niter + 1);
    unsigned* ar = arena + (long long)s * acap;
    double2* yv = yval + (long long)s * acap;
    const long long ne = (long long)n1 * n2;
    const Cx<double>* __restrict__ x0s = X0 + (long long)s * ne;
    const double2* __restrict__ mh = reinterpret_cast<const double2*>(mhat + ((first_slice + s) / spm) * ne);
    int cur = 0, np = 0;
    for (int i = 0; i < ke; ++i) {
        const int a0 = as[i], n = as[i + 1] - a0;
        unsigned* ic = ibuf + (size_t)cur * maxseg;
        double2* yc = ybuf + (size_t)cur * maxseg;
        const unsigned* ip = ibuf + (size_t)(cur ^ 1) * maxseg;
        const double2* yp = ybuf + (size_t)(cur ^ 1) * maxseg;
        // this iteration's record, sorted ascending
        int m = 1; while (m < n) m <<= 1;
        for (int t = threadIdx.x; t < m; t += blockDim.x) ic[t] = t < n ? ar[a0 + t] : 0xffffffffu;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < m; t += blockDim.x) {
                    const int p = t ^ j;
                    if (p > t) {
                        const unsigned a = ic[t], b = ic[p];
                        if ((a > b) == ((t & k) == 0)) { ic[t] = b; ic[p] = a; }
                    }
                }
                __syncthreads();
            }
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const unsigned ent = ic[t];
            ar[a0 + t] = ent;                                   // the restart scatters from the sorted record
            const bool watched = (ent >> 31) != 0u;
            const unsigned pj = ent & 0x7fffffffu;
            const int jr = (int)(pj >> 16), jc = (int)(pj & 0xffffu);
            const Cx<double> x0 = x0s[(long long)jr * n2 + jc];
            double accx = 0.0, accy = 0.0, selfx = 0.0, selfy = 0.0;
            int q = 0;
            for (; q + 8 <= np; q += 8) {
                double2 mm[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned ps = ip[q + u] & 0x7fffffffu;
                    int dr = jr - (int)(ps >> 16); if (dr < 0) dr += n1;
                    int dc = jc - (int)(ps & 0xffffu); if (dc < 0) dc += n2;
                    mm[u] = __ldg(mh + (long long)dr * n2 + dc);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double2 ys = yp[q + u];
                    if ((ip[q + u] & 0x7fffffffu) == pj) { selfx = ys.x; selfy = ys.y; }
                    accx += mm[u].x * ys.x - mm[u].y * ys.y;
                    accy += mm[u].x * ys.y + mm[u].y * ys.x;
                }
            }
            for (; q < np; ++q) {
                const unsigned ps = ip[q] & 0x7fffffffu;
                const double2 ys = yp[q];
                if (ps == pj) { selfx = ys.x; selfy = ys.y; }
                int dr = jr - (int)(ps >> 16); if (dr < 0) dr += n1;
                int dc = jc - (int)(ps & 0xffffu); if (dc < 0) dc += n2;
                const double2 mv = __ldg(mh + (long long)dr * n2 + dc);
                accx += mv.x * ys.x - mv.y * ys.y;
                accy += mv.x * ys.y + mv.y * ys.x;
            }
            Cx<double> X = x0;
            if (i > 0) X = cmake<double>(alpha * x0.x + selfx - alpha * inv_n * accx, alpha * x0.y + selfy - alpha * inv_n * accy);
            const Cx<double> tk = tau[(long long)s * niter + i];
            const double a = tk.x, b = tk.y;
            Cx<double> y = apply_threshold<OP, double>(X, a, b, a * a - b * b, 2.0 * a * b);
            const bool survives = (y.x != 0.0) || (y.y != 0.0);
            if (survives == watched || i == debug_fail) atomicMin(&kfail[s], i);
            if (watched) y = cmake<double>(0.0, 0.0);
            yc[t] = make_double2(y.x, y.y);
            yv[a0 + t] = make_double2(y.x, y.y);
        }
        __syncthreads();
        np = n; cur ^= 1;
    }
}

// restart spectrum of a frozen slice: zero, then y_{k_e - 1} at its support
__global__ void k_zero_slices(const int* __restrict__ list, Cx<double>* W, long long ne) {
    const int s = list[blockIdx.y];
    double2* w = reinterpret_cast<double2*>(W + (long long)s * ne);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ne; i += (long long)gridDim.x * blockDim.x) w[i] = make_double2(0.0, 0.0);
}
// sums |x_k| that complex128 will (re)compute: from the restart iteration on (a slice reopened by the verification has
// fp32 sums there)
__global__ void k_zero_sums(const int* __restrict__ list, const int* __restrict__ esc, double* S, int niter, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int s = list[r];
    const int ke = esc[s] - 1;
    for (int q = ke > 1 ? ke : 1; q <= niter; ++q) S[(long long)s * (niter + 1) + q] = 0.0;
}
__global__ void k_scatter_restart(const int* __restrict__ list, const int* __restrict__ esc, const unsigned* __restrict__ arena,
                                  const int* __restrict__ astart, const double2* __restrict__ yval, int acap, int niter,
                                  Cx<double>* W, int n2, long long ne, double* S) {
    const int s = list[blockIdx.y];
    const int ke = esc[s] - 1;              // >= 1 for every slice of this list
    const int a0 = astart[(long long)s * (niter + 1) + ke - 1], n = astart[(long long)s * (niter + 1) + ke] - a0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const unsigned pj = arena[(long long)s * acap + a0 + t];
        if (pj >> 31) continue;                     // a watched kill: not part of the spectrum
        const double2 y = yval[(long long)s * acap + a0 + t];
        W[(long long)s * ne + (long long)(pj >> 16) * n2 + (pj & 0xffffu)] = cmake<double>(y.x, y.y);
    }
}

void process_chunk_esc(RunCtx& R, Lane& L, int64_t first, int64_t count) {
    p3d_plan* P = R.P;
    const p3d_pocs_params& pr = *R.pr;
    const int niter = pr.niter;
    const int64_t ne = (int64_t)P->n1 * P->n2;
    cudaStream_t st = L.stream;
    const F64Kernels K = f64_kernels(P->f64);
    const int64_t band_max = 32768;
    const int acap = L.arena_cap;

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
    P3D_CUDA(cudaMemsetAsync(L.esc, 0, sizeof(int) * count, st));
    P3D_CUDA(cudaMemsetAsync(L.acnt, 0, sizeof(int) * count, st));
    P3D_CUDA(cudaMemsetAsync(L.astart, 0, sizeof(int) * count * (niter + 1), st));
    P3D_CUDA(cudaMemsetAsync(L.wflag, 0, sizeof(int) * count, st));
    P3D_CUDA(cudaMemsetAsync(L.kfail, 0x7f, sizeof(int) * count, st));
    for (int64_t i = 0; i < count; ++i) {
        memset(&L.h_stats[i], 0, sizeof(SliceStats));
        L.h_stats[i].minabs_bits = 0x7f800000u;
        L.h_stats[i].minabs64_key = ~0ull;
        L.h_list[i] = (int)i;
    }
    P3D_CUDA(cudaMemcpyAsync(L.stats, L.h_stats, sizeof(SliceStats) * count, cudaMemcpyHostToDevice, st));
    P3D_CUDA(cudaMemcpyAsync(L.list, L.h_list, sizeof(int) * count, cudaMemcpyHostToDevice, st));

    P->geom.slices_per_mask = (int)std::min<int64_t>(R.spm, 0x7fffffff);
    PocsGeom G64 = K.cfg.geom; G64.slices_per_mask = P->geom.slices_per_mask;
    GenericCfg cfg64 = K.cfg; cfg64.geom = G64;

    const bool data_driven = pr.thresh_model == P3D_MODEL_DATA_DRIVEN;
    const bool adaptive = pr.version == P3D_VERSION_ADAPTIVE;

    BandArgs<float> A;
    memset(&A, 0, sizeof(A));
    A.mask = R.dmask; A.mbits = R.dmbits; A.niter = niter; A.eps = pr.eps; A.alpha = (float)pr.alpha;
    A.inv_n = (float)(1.0 / ((double)P->n1 * (double)P->n2));
    A.W = L.W; A.D = D; A.OUT = OUT; A.first_slice = first; A.tau = L.tau; A.S = L.S; A.stop = L.stop; A.stats = L.stats;
    A.exact_tie = (pr.thresh_model == P3D_MODEL_INVERSE_PROPORTIONAL || data_driven || pr.decay_factors) ? 1 : 0;
    A.esc = L.esc;
    if (P->use_tma) { A.tma_base = L.W; A.tma_slices = L.cap; A.tma_slice0 = 0; }

    BandArgs<double> A64;
    memset(&A64, 0, sizeof(A64));
    A64.mask = R.dmask; A64.mbits = R.dmbits64; A64.niter = niter; A64.eps = pr.eps; A64.alpha = pr.alpha;
    A64.inv_n = 1.0 / ((double)P->n1 * (double)P->n2);
    A64.W = L.W64; A64.D32 = D; A64.OUT32 = OUT; A64.first_slice = first; A64.tau = L.tau64; A64.S = L.S; A64.stop = L.stop;
    A64.stats = L.stats; A64.exact_tie = 1; A64.cand = L.cand; A64.cand_stride = L.cand_stride; A64.esc = L.esc;
    if (P->use_tma) { A64.tma_base = L.W64; A64.tma_slices = L.cap; A64.tma_slice0 = 0; }

    // launch helpers over a (compacted) list of slices of this chunk
    auto rows_init64 = [&](BandArgs<double> B, const int* list, int n) {
        prof_begin(P, L.events, st, 10);
        for (int64_t o = 0; o < n; o += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, n - o);
            B.list = list + o;
            if (K.spec_rows) K.spec.rows_init_io32(G64, K.tw_rows, B, nb, st);
            else generic64_rows_init(cfg64, K.a2, B, nb, st);
        }
        prof_end(P, L.events, st);
        P3D_CUDA(cudaGetLastError());
    };
    auto cols64 = [&](BandArgs<double> B, const int* list, int n, int op) {
        prof_begin(P, L.events, st, 8);
        for (int64_t o = 0; o < n; o += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, n - o);
            B.list = list + o;
            if (K.spec_cols) K.spec.cols_iter(G64, K.tw_cols, B, nb, op, st);
            else generic64_cols_iter(cfg64, K.a1, B, nb, op, st);
        }
        prof_end(P, L.events, st);
        P3D_CUDA(cudaGetLastError());
    };
    auto rows64 = [&](BandArgs<double> B, const int* list, int n) {
        prof_begin(P, L.events, st, 9);
        for (int64_t o = 0; o < n; o += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, n - o);
            B.list = list + o;
            if (K.spec_rows) K.spec.rows_iter_io32(G64, K.tw_rows, B, nb, st);
            else generic64_rows_iter(cfg64, K.a2, B, nb, st);
        }
        prof_end(P, L.events, st);
        P3D_CUDA(cudaGetLastError());
    };
    auto for_list32 = [&](BandArgs<float> B, const int* list, int n, auto&& fn) {
        for (int64_t o = 0; o < n; o += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, n - o);
            B.list = list + o;
            fn(B, nb);
        }
    };

    // ---- setup: row FFT of d (fp32 state), statistics of X0 in complex128 (X0 stays in W64 for the replay) -----------
    {
        BandArgs<float> B = A;
        B.adaptive = 0; B.accum = 1; B.store_x0 = 0;
        for_list32(B, L.list, (int)count, [&](const BandArgs<float>& b, int nb) { launch_rows_init(P, L, b, nb); });
        BandArgs<double> B64 = A64;
        B64.adaptive = 0; B64.accum = 0; B64.store_x0_inplace = 1;
        rows_init64(B64, L.list, (int)count);
        prof_begin(P, L.events, st, 10);
        for (int64_t o = 0; o < count; o += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, count - o);
            B64.list = L.list + o;
            if (K.spec_cols) K.spec.cols_stats(G64, K.tw_cols, B64, nb, st);
            else generic64_cols_stats(cfg64, K.a1, B64, nb, st);
        }
        k_lexmax_finish<<<(unsigned)count, 128, 0, st>>>(L.cand, L.cand_stride, K.cand_stride, L.stats);
        prof_end(P, L.events, st);
        P3D_CUDA(cudaGetLastError());
    }
    P3D_CUDA(cudaMemcpyAsync(L.h_stats, L.stats, sizeof(SliceStats) * count, cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaStreamSynchronize(st));

    // ---- schedule (host, double) ---------------------------------------------------------------------------------
    std::vector<cd> tau;
    const double eps32 = 5.9604644775390625e-08;      // 2^-24
    // adaptive POCS has no sparse recursion of this form: complex128 from the first iteration (an infinite guard band)
    // watch list (kernels that record the guard band; bit 31 of an entry must be free): a hit is verified by the replay
    // instead of freezing the slice
    const bool want_watch = P->watch_mode < 0 ? (ne >= 400000) : (P->watch_mode != 0);
    const bool watch = want_watch && P->spec.cols_iter && !P->force_generic && P->n1 < 32768 && !adaptive;
    const double gfac = (adaptive || ne < P->pilot_min_elems) ? std::numeric_limits<double>::infinity()
                                                               : (watch ? std::max(P->guard_factor, P->watch_guard_factor) : P->guard_factor);
    auto store_tau = [&](int64_t i) {
        for (int k = 0; k < niter; ++k) {
            L.h_tau[i * niter + k] = cmake<float>((float)tau[k].real(), (float)tau[k].imag());
            L.h_tau64[i * niter + k] = cmake<double>(tau[k].real(), tau[k].imag());
            if (R.tau_out) { R.tau_out[((first + i) * niter + k) * 2] = tau[k].real(); R.tau_out[((first + i) * niter + k) * 2 + 1] = tau[k].imag(); }
        }
    };
    auto guard_of = [&](const SliceStats& ss) {
        const double g = gfac * eps32 * std::sqrt(ss.sumsq / (double)std::max<unsigned long long>(1ull, ss.nnz));
        return (float)std::min(g, 3.0e38);
    };
    auto stats64 = [&](const SliceStats& ss) {
        ScheduleStats sc;
        sc.z = cd(f64_from_ordered(ss.re64_key), f64_from_ordered(ss.im64_key));
        sc.sumsq = ss.sumsq; sc.vmax = f64_from_ordered(ss.maxabs64_key); sc.vmin = f64_from_ordered(ss.minabs64_key);
        return sc;
    };
    if (!data_driven) {
        for (int64_t i = 0; i < count; ++i) {
            const SliceStats& ss = L.h_stats[i];
            L.h_stop[i] = ss.nnz == 0 ? -1 : 0;
            bool is_real = false;
            host_schedule(pr, stats64(ss), ne, tau, is_real);
            if (pr.sqrt_decay) apply_sqrt_decay(tau, is_real);
            store_tau(i);
            L.h_guard[i] = guard_of(ss);
        }
    } else {
        // thresholds = order statistics of X0 in numpy's complex ordering, taken from the complex128 X0 left in W64:
        // radix sort of the ordered real parts (payload = position), ties resolved on the imaginary parts
        if (L.dd_ne < ne) {
            if (L.dd_keys) cudaFree(L.dd_keys);
            if (L.dd_vals) cudaFree(L.dd_vals);
            L.dd_keys = nullptr; L.dd_vals = nullptr; L.dd_ne = 0;
            P3D_CUDA(cudaMalloc(&L.dd_keys, sizeof(unsigned long long) * 2 * ne));
            P3D_CUDA(cudaMalloc(&L.dd_vals, sizeof(unsigned int) * 2 * ne));
            L.dd_ne = ne;
        }
        size_t need = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, need, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                  (unsigned int*)nullptr, (unsigned int*)nullptr, (long long)ne, 0, 64, st);
        if (need > L.cub_temp_bytes) {
            if (L.cub_temp) cudaFree(L.cub_temp);
            L.cub_temp = nullptr; L.cub_temp_bytes = 0;
            P3D_CUDA(cudaMalloc(&L.cub_temp, need)); L.cub_temp_bytes = need;
        }
        for (int64_t i = 0; i < count; ++i) {
            L.h_stop[i] = L.h_stats[i].nnz == 0 ? -1 : 0;
            if (L.h_stop[i]) continue;
            cd tmin, tmax;
            schedule_bounds(pr, stats64(L.h_stats[i]), ne, tmin, tmax);
            prof_begin(P, L.events, st, 6);
            k_dd_keys64<<<std::min<long long>((ne + 255) / 256, 148 * 8), 256, 0, st>>>(L.W64 + i * ne, ne, tmin.real(), tmin.imag(), tmax.real(), tmax.imag(),
                                                                                          L.dd_keys, L.dd_vals, L.stats + i);
            size_t tb = L.cub_temp_bytes;
            cub::DeviceRadixSort::SortPairsDescending(L.cub_temp, tb, L.dd_keys, L.dd_keys + ne, L.dd_vals, L.dd_vals + ne, (long long)ne, 0, 64, st);
            k_pick_tau64<<<1, 128, 0, st>>>(L.dd_keys + ne, L.dd_vals + ne, L.W64 + i * ne, L.stats + i, L.tau64 + i * niter, niter);
            prof_end(P, L.events, st);
        }
        P3D_CUDA(cudaGetLastError());
        P3D_CUDA(cudaMemcpyAsync(L.h_tau64, L.tau64, sizeof(Cx<double>) * count * niter, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_stats, L.stats, sizeof(SliceStats) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < count; ++i) {
            tau.assign(niter, cd(0, 0));
            L.h_guard[i] = 0.f;
            if (!L.h_stop[i]) {
                P3D_REQUIRE(L.h_stats[i].n_cand > 0, P3D_ERR_NUMERIC,
                            "data-driven schedule: no coefficient between tau_min and tau_max in slice %lld", (long long)(first + i));
                for (int k = 0; k < niter; ++k) tau[k] = cd(L.h_tau64[i * niter + k].x, L.h_tau64[i * niter + k].y);
                if (pr.sqrt_decay) apply_sqrt_decay(tau, false);
                L.h_guard[i] = guard_of(L.h_stats[i]);
            }
            store_tau(i);
        }
    }
    if (R.schedule_only) { L.pending = false; return; }

    P3D_CUDA(cudaMemcpyAsync(L.tau, L.h_tau, sizeof(Cx<float>) * count * niter, cudaMemcpyHostToDevice, st));
    P3D_CUDA(cudaMemcpyAsync(L.tau64, L.h_tau64, sizeof(Cx<double>) * count * niter, cudaMemcpyHostToDevice, st));
    P3D_CUDA(cudaMemcpyAsync(L.stop, L.h_stop, sizeof(int) * count, cudaMemcpyHostToDevice, st));
    P3D_CUDA(cudaMemcpyAsync(L.guard, L.h_guard, sizeof(float) * count, cudaMemcpyHostToDevice, st));
    for (int64_t i = 0; i < count; ++i)
        if (L.h_stop[i] < 0)
            P3D_CUDA(cudaMemcpyAsync(OUT + i * ne, D + i * ne, sizeof(Cx<float>) * ne, cudaMemcpyDeviceToDevice, st));

    // ---- phase 1: fp32 pilot over the slices still in fp32 -----------------------------------------------------------
    const bool guard_on = gfac > 0.0;
    std::vector<int> active;
    active.reserve((size_t)count);
    for (int64_t i = 0; i < count; ++i) if (L.h_stop[i] == 0) active.push_back((int)i);
    auto upload_list = [&](const std::vector<int>& v) {
        // the previous launches were waited for (or never used the list): the pinned mirror is free
        memcpy(L.h_list, v.data(), sizeof(int) * v.size());
        if (!v.empty()) P3D_CUDA(cudaMemcpyAsync(L.list, L.h_list, sizeof(int) * v.size(), cudaMemcpyHostToDevice, st));
    };
    upload_list(active);
    const bool need_sync = guard_on || pr.eps > 0.0;
    // iterations between two compactions: at least seg_iters, and enough work (~2 ms) that the wait is amortised
    const double est_iter_ms = 12e-9 * (double)ne * (double)count;       // ~12 ns per 1000 elements and fp32 iteration
    const int seg = std::max(std::max(1, P->seg_iters), (int)std::min(16.0, std::ceil(2.0 / std::max(est_iter_ms, 1e-6))));
    int k = 0;
    while (k < niter && !active.empty()) {
        const int kend = need_sync ? std::min(niter, k + (k == 0 && guard_on ? 1 : seg)) : niter;
        for (; k < kend; ++k) {
            BandArgs<float> B = A;
            B.guard = guard_on ? L.guard : nullptr;
            if (guard_on) {
                B.arena = L.arena; B.acnt = L.acnt; B.astart = L.astart; B.arena_cap = acap;
                B.scap = P->support_cap > 0 ? P->support_cap : (int)(2.2 * std::sqrt((double)ne));
                B.watch = watch ? 1 : 0;
                B.wflag = L.wflag;
            }
            B.k = k; B.last = (k == niter - 1) ? 1 : 0;
            B.write_out = (B.last || (pr.eps > 0.0 && k >= 3)) ? 1 : 0;
            for_list32(B, L.list, (int)active.size(), [&](const BandArgs<float>& b, int nb) {
                launch_cols_iter(P, L, b, nb, pr.thresh_op);
                launch_rows_iter(P, L, b, nb);
            });
        }
        if (k < niter && need_sync) {
            P3D_CUDA(cudaMemcpyAsync(L.h_esc, L.esc, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
            P3D_CUDA(cudaMemcpyAsync(L.h_stop, L.stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
            P3D_CUDA(cudaStreamSynchronize(st));
            size_t w = 0;
            for (size_t r = 0; r < active.size(); ++r) {
                const int i = active[r];
                if (L.h_esc[i] == 0 && L.h_stop[i] == 0) active[w++] = i;
            }
            if (w != active.size()) { active.resize(w); upload_list(active); }
        }
    }

    // ---- phase 2: replay + verification, exact restart, complex128 iterations of the frozen slices -----------------------
    if (guard_on) {
        P3D_CUDA(cudaMemcpyAsync(L.h_esc, L.esc, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_stop, L.stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_wflag, L.wflag, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(L.h_astart, L.astart, sizeof(int) * count * (niter + 1), cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaStreamSynchronize(st));
        // replay: frozen slices up to their switch iteration; slices that finished (or stopped) in fp32 but recorded
        // watched decisions are replayed over everything they ran, for the verification alone
        std::vector<int> rl;
        int kmax = 0, seg_max = 0;
        for (int64_t i = 0; i < count; ++i) {
            int kend = 0;
            if (L.h_stop[i] < 0) kend = 0;
            else if (L.h_esc[i] > 0 && L.h_stop[i] == 0) kend = L.h_esc[i] - 1;
            else if (L.h_wflag[i]) kend = L.h_stop[i] > 0 ? L.h_stop[i] : niter;
            L.h_kend[i] = kend;
            if (kend > 0) { rl.push_back((int)i); kmax = std::max(kmax, kend); }
        }
        if (!rl.empty()) {
            prof_begin(P, L.events, st, 11);
            std::vector<int> smax((size_t)kmax, 0);
            for (int i : rl) {
                const int* as = L.h_astart + (int64_t)i * (niter + 1);
                for (int q = 0; q < L.h_kend[i]; ++q) { smax[q] = std::max(smax[q], as[q + 1] - as[q]); seg_max = std::max(seg_max, as[q + 1] - as[q]); }
            }
            upload_list(rl);
            P3D_CUDA(cudaMemcpyAsync(L.kend, L.h_kend, sizeof(int) * count, cudaMemcpyHostToDevice, st));
            const int nrl = (int)rl.size();
            int pow2 = 1; while (pow2 < seg_max) pow2 <<= 1;
            const int spm_i = (int)std::min<int64_t>(R.spm, 0x7fffffff);
            if (pow2 <= P->fused_replay_max) {
                // small supports: the whole replay of a slice in one launch (one CTA per slice)
                const size_t smem = (size_t)2 * pow2 * (sizeof(double2) + sizeof(unsigned));
                static bool fused_cfg = false;
                if (!fused_cfg) {
                    cudaFuncSetAttribute(k_replay_fused<P3D_OP_HARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 20);
                    cudaFuncSetAttribute(k_replay_fused<P3D_OP_SOFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 20);
                    cudaFuncSetAttribute(k_replay_fused<P3D_OP_GARROTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 20);
                    fused_cfg = true;
                }
#define P3D_REPLAY_F(OPV) k_replay_fused<OPV><<<(unsigned)nrl, 256, smem, st>>>(L.list, L.kend, L.kfail, L.arena, L.astart, L.yval, acap, niter, pow2, L.W64, P->mhat, \
                                                                               L.tau64, (long long)first, spm_i, P->n1, P->n2, pr.alpha, A64.inv_n, P->debug_fail_iter)
                if (pr.thresh_op == P3D_OP_HARD) P3D_REPLAY_F(P3D_OP_HARD);
                else if (pr.thresh_op == P3D_OP_SOFT) P3D_REPLAY_F(P3D_OP_SOFT);
                else P3D_REPLAY_F(P3D_OP_GARROTE);
#undef P3D_REPLAY_F
            } else {
            static bool sort_cfg = false;
            if (!sort_cfg) { cudaFuncSetAttribute(k_sort_segments, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4); sort_cfg = true; }
            for (int64_t o = 0; o < nrl; o += band_max) {
                const int nb = (int)std::min<int64_t>(band_max, nrl - o);
                k_sort_segments<<<dim3((unsigned)kmax, (unsigned)nb), 256, sizeof(unsigned) * pow2, st>>>(L.list + o, L.kend, L.arena, L.astart, acap, niter);
            }
            for (int i = 0; i < kmax; ++i) {
                if (smax[i] == 0) continue;
                const unsigned gx = (unsigned)((smax[i] + 127) / 128);
                for (int64_t o = 0; o < nrl; o += band_max) {
                    const int nb = (int)std::min<int64_t>(band_max, nrl - o);
                    const dim3 grid(gx, (unsigned)nb);
#define P3D_REPLAY(OPV) k_replay<OPV><<<grid, 128, 0, st>>>(i, L.list + o, L.kend, L.kfail, L.arena, L.astart, L.yval, acap, niter, L.W64, P->mhat, L.tau64, \
                                                            (long long)first, spm_i, P->n1, P->n2, pr.alpha, A64.inv_n, P->debug_fail_iter)
                    if (pr.thresh_op == P3D_OP_HARD) P3D_REPLAY(P3D_OP_HARD);
                    else if (pr.thresh_op == P3D_OP_SOFT) P3D_REPLAY(P3D_OP_SOFT);
                    else P3D_REPLAY(P3D_OP_GARROTE);
#undef P3D_REPLAY
                }
            }
            }
            prof_end(P, L.events, st);
            P3D_CUDA(cudaGetLastError());
            P3D_CUDA(cudaMemcpyAsync(L.h_kfail, L.kfail, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
            P3D_CUDA(cudaStreamSynchronize(st));
        }
        // first complex128 iteration of every slice: its freeze point, or earlier where the verification found the pilot's
        // decision to differ from the exact one (everything the pilot did from there on is discarded)
        std::vector<std::pair<int, int>> e64;       // (k_e, slice)
        bool esc_changed = false;
        for (int64_t i = 0; i < count; ++i) {
            if (L.h_stop[i] < 0) continue;
            int ke = (L.h_esc[i] > 0 && L.h_stop[i] == 0) ? L.h_esc[i] - 1 : -1;
            if (L.h_kend[i] > 0 && L.h_kfail[i] < L.h_kend[i]) {
                ke = L.h_kfail[i];
                L.h_esc[i] = ke + 1; L.h_stop[i] = 0; esc_changed = true;
                L.n_verify_failed += 1;
            }
            if (ke >= 0) e64.push_back(std::make_pair(ke, (int)i));
        }
        if (esc_changed) {
            // (a slice the pilot had stopped early on its cost is reopened: the complex128 kernels decide again)
            P3D_CUDA(cudaMemcpyAsync(L.esc, L.h_esc, sizeof(int) * count, cudaMemcpyHostToDevice, st));
            P3D_CUDA(cudaMemcpyAsync(L.stop, L.h_stop, sizeof(int) * count, cudaMemcpyHostToDevice, st));
        }
        std::sort(e64.begin(), e64.end());
        if (!e64.empty()) {
            const int n64 = (int)e64.size();
            std::vector<int> order((size_t)n64);
            int n0 = 0;                              // slices whose first iteration already belongs to complex128: restart state = d
            int smax_restart = 1;
            for (int r = 0; r < n64; ++r) {
                order[r] = e64[r].second;
                if (e64[r].first == 0) ++n0;
                else {
                    const int* as = L.h_astart + (int64_t)e64[r].second * (niter + 1);
                    smax_restart = std::max(smax_restart, as[e64[r].first] - as[e64[r].first - 1]);
                }
            }
            upload_list(order);
            k_zero_sums<<<(unsigned)((n64 + 127) / 128), 128, 0, st>>>(L.list, L.esc, L.S, niter, n64);
            const int n1r = n64 - n0;                // slices restarted from a replayed spectrum
            const int* list1 = L.list + n0;
            if (n0 > 0) {
                BandArgs<double> B64 = A64;
                B64.adaptive = adaptive ? 1 : 0; B64.accum = 0;
                rows_init64(B64, L.list, n0);
            }
            if (n1r > 0) {
                prof_begin(P, L.events, st, 11);
                for (int64_t o = 0; o < n1r; o += band_max) {
                    const int nb = (int)std::min<int64_t>(band_max, n1r - o);
                    k_zero_slices<<<dim3((unsigned)std::min<long long>((ne + 255) / 256, 592), (unsigned)nb), 256, 0, st>>>(list1 + o, L.W64, ne);
                    k_scatter_restart<<<dim3((unsigned)((smax_restart + 255) / 256), (unsigned)nb), 256, 0, st>>>(list1 + o, L.esc, L.arena, L.astart, L.yval, acap, niter,
                                                                                                            L.W64, P->n2, ne, L.S);
                }
                prof_end(P, L.events, st);
                P3D_CUDA(cudaGetLastError());
                // inverse column pass + row pass of iteration k_e - 1 in complex128: x_{k_e - 1} and its row FFT
                BandArgs<double> B64 = A64;
                B64.restart = 1; B64.k = 0; B64.last = 0; B64.write_out = 0;
                cols64(B64, list1, n1r, P3D_OP_RESTART);
                rows64(B64, list1, n1r);
            }
            BandArgs<double> B64 = A64;
            B64.adaptive = adaptive ? 1 : 0;
            int ptr = 0;
            int64_t its = 0;
            for (int kk = e64[0].first; kk < niter; ++kk) {
                while (ptr < n64 && e64[ptr].first <= kk) ++ptr;
                B64.k = kk; B64.last = (kk == niter - 1) ? 1 : 0;
                B64.write_out = (B64.last || (pr.eps > 0.0 && kk >= 3)) ? 1 : 0;
                cols64(B64, L.list, ptr, pr.thresh_op);
                rows64(B64, L.list, ptr);
                its += ptr;
            }
            L.n_escalated += n64; L.n_esc_iters += its;
        }
    }

    // ---- results ------------------------------------------------------------------------------------
    if (R.out_mem == P3D_MEM_HOST)
        P3D_CUDA(cudaMemcpyAsync(R.out + first * ne, OUT, sizeof(Cx<float>) * ne * count, cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaMemcpyAsync(L.h_S, L.S, sizeof(double) * count * (niter + 1), cudaMemcpyDeviceToHost, st));
    P3D_CUDA(cudaMemcpyAsync(L.h_stop, L.stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
    L.pending = true; L.p_first = first; L.p_count = count;
}

// fft2 of the mask planes in complex128 (the kernel of the sparse recursion), once per run
void compute_mhat(p3d_plan* P, const uint8_t* dmask, int64_t n_masks, cudaStream_t st) {
    const int64_t ne = (int64_t)P->n1 * P->n2;
    const F64Kernels K = f64_kernels(P->f64);
    if (n_masks > P->mhat_masks || K.cand_stride > P->mh_cand_stride) {
        for (void* p : {(void*)P->mhat, (void*)P->mask_c64, (void*)P->mh_stats, (void*)P->mh_cand}) if (p) cudaFree(p);
        P->mhat = nullptr; P->mask_c64 = nullptr; P->mh_stats = nullptr; P->mh_cand = nullptr; P->mhat_masks = 0;
        P3D_CUDA(cudaMalloc(&P->mhat, sizeof(Cx<double>) * ne * n_masks));
        P3D_CUDA(cudaMalloc(&P->mask_c64, sizeof(Cx<float>) * ne * n_masks));
        P3D_CUDA(cudaMalloc(&P->mh_stats, sizeof(SliceStats) * n_masks));
        P3D_CUDA(cudaMalloc(&P->mh_cand, sizeof(double2) * n_masks * K.cand_stride));
        P->mhat_masks = n_masks; P->mh_cand_stride = K.cand_stride;
    }
    k_mask_to_c64<<<(unsigned)std::min<long long>((ne * n_masks + 255) / 256, 148 * 16), 256, 0, st>>>(dmask, P->mask_c64, ne * n_masks);
    PocsGeom G64 = K.cfg.geom; G64.slices_per_mask = 1;
    GenericCfg cfg64 = K.cfg; cfg64.geom = G64;
    BandArgs<double> M;
    memset(&M, 0, sizeof(M));
    M.W = P->mhat; M.D32 = P->mask_c64; M.OUT32 = P->mask_c64; M.mask = dmask; M.niter = 1; M.stats = P->mh_stats;
    M.cand = P->mh_cand; M.cand_stride = P->mh_cand_stride; M.store_x0_inplace = 1; M.alpha = 1.0; M.inv_n = 1.0;
    for (int64_t o = 0; o < n_masks; o += 32768) {
        const int nb = (int)std::min<int64_t>(32768, n_masks - o);
        BandArgs<double> B = M;
        B.W += o * ne; B.D32 += o * ne; B.OUT32 += o * ne; B.stats += o; B.cand += o * M.cand_stride; B.first_slice = o;
        if (K.spec_rows) K.spec.rows_init_io32(G64, K.tw_rows, B, nb, st); else generic64_rows_init(cfg64, K.a2, B, nb, st);
        if (K.spec_cols) K.spec.cols_stats(G64, K.tw_cols, B, nb, st); else generic64_cols_stats(cfg64, K.a1, B, nb, st);
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaStreamSynchronize(st));
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
    }
    if (n_slices == 0) return P3D_OK;
    if (spm <= 0) spm = n_slices;
    if (!schedule_only && x_mem == P3D_MEM_DEVICE && out_mem == P3D_MEM_DEVICE) {
        // the observed slices are read in place in every iteration while results (and scratch) go to `out`
        const char* xa = (const char*)x; const char* oa = (const char*)out;
        const size_t nb = sizeof(Cx<float>) * (size_t)P->n1 * P->n2 * (size_t)n_slices;
        P3D_REQUIRE(xa + nb <= oa || oa + nb <= xa, P3D_ERR_BAD_ARG, "device buffers x and out must not overlap");
    }
    DeviceGuard guard(P->device);

    // complex128 throughout: on request, and for the percentile operators in the default mode (every iteration thresholds at a
    // data value - the percentile lies between two sorted moduli - so an fp32 iterate decides a near-tie differently
    // sooner or later; there is no sparse support to replay either)
    if ((P->precision == 64 || (P->precision == 0 && pr->thresh_percentile)) && !filt) {
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
    // escalating precision: every ordinary POCS run unless the plan was pinned to fp32 (percentile operators: complex128
    // above, or fp32 on request; the kx-ky filter mode has no threshold decision that fp32 could get wrong)
    const bool escalate = P->precision == 0 && !filt && !pr->thresh_percentile;
    // host data: four lanes (streams + buffer sets), each fed by its own host thread, so that the D2H of chunk i and the
    // H2D of chunk i+4 hide behind the iterations of chunks i+1..i+3 even when the PCIe path is slow (8 ranks sharing
    // one host: 8-11 GB/s per direction measured; e2e at 8 GPUs 444k -> 506k slice-it/s going from 2 to 4
    // lanes, 76.6k -> 80.9k on one GPU); device-resident data need one lane.
    // (the percentile operators share one scratch area and sort inside the iteration loop: one lane)
    // (round 2, escalating mode: a chunk has several host waits - statistics, pilot verdict, list compaction -, and with
    // 8 lanes of ~32 slices some lane always has kernels queued: e2e / device-resident 0.943 -> 0.961 on config 2,
    // profiles/r2_sweep_e2e_chunks_v2.txt)
    // Every lane has a feeder thread that spins in cudaStreamSynchronize: 8 lanes only while all ranks of this host
    // (torchrun's LOCAL_WORLD_SIZE) keep that to half of the hardware threads - 4 lanes, round 1's setting, beyond.
    int host_lanes = 8;
    {
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw ? std::max(1, atoi(lw)) : 1;
        unsigned hc = std::thread::hardware_concurrency();
        if (hc == 0) hc = 16;
        if (2 * 8 * ranks > (int)hc) host_lanes = 4;
    }
    const int lanes = pr->thresh_percentile ? 1 : (P->n_lanes > 0 ? P->n_lanes : ((host_in || host_out) ? host_lanes : 1));
    if ((int)P->lanes.size() < lanes) P->lanes.resize(lanes);
    const int nbuf = 1 + (host_in ? 1 : 0) + (host_out ? 1 : 0) + (escalate ? 2 : 0);
    for (auto& L : P->lanes) { L.pending = false; L.n_escalated = 0; L.n_esc_iters = 0; L.n_verify_failed = 0; }
    int cand_stride = 0;
    if (escalate) {
        if (!P->f64) { P->f64 = f64_create(P->device, P->n1, P->n2, &P->ax1, &P->ax2, P->smem_optin); f64_install_spec(P->f64, P->spec_variant64); }
        f64_set_force_generic(P->f64, P->force_generic ? 1 : 0);
        cand_stride = f64_kernels(P->f64).cand_stride;
    }
    // chunk size: everything at once on one lane; with several lanes at least two chunks per lane
    // (copy/compute overlap) but never tiny chunks
    int64_t want = n_slices;
    if (lanes > 1) {
        const int64_t ne_ = (int64_t)P->n1 * P->n2;
        const int64_t min_chunk = std::max<int64_t>(1, std::min<int64_t>(n_slices, (int64_t)(48e6 / (8.0 * (double)ne_)) + 1));
        // full-size chunks: ~4 per lane (2 with 4 lanes); the first and last chunks of a call are shorter (ramp below), so the
        // un-overlapped first H2D / last D2H stay short while the bulk of the launches is large enough to keep
        // the tail of a launch (its last, partially filled wave of CTAs) small
        const int per_lane = lanes >= 8 ? 4 : 2;
        want = std::max<int64_t>(min_chunk, (n_slices + per_lane * lanes - 1) / (per_lane * lanes));
    }
    if (P->max_slices > 0) want = std::min<int64_t>(want, P->max_slices);
    int64_t have = 0;
    for (int i = 0; i < lanes; ++i) have = (i == 0) ? P->lanes[i].cap : std::min(have, P->lanes[i].cap);
    int64_t cap = want;
    if (have < want) {
        for (auto& L : P->lanes) { cudaStream_t st = L.stream; L.stream = nullptr; free_lane(L); L.stream = st; }
        cap = auto_capacity(P, want, nbuf, lanes);
    }
    for (int i = 0; i < lanes; ++i) ensure_lane(P, P->lanes[i], cap, pr->niter, host_in, host_out, escalate, cand_stride);

    RunCtx R;
    R.P = P; R.pr = pr; R.x = (const Cx<float>*)x; R.x_mem = x_mem; R.spm = spm;
    R.out = (Cx<float>*)out; R.out_mem = out_mem; R.niter_out = niter_out; R.cost_out = cost_out;
    R.costs_out = costs_out; R.tau_out = tau_out; R.schedule_only = schedule_only;
    R.dmask = nullptr; R.dmbits = nullptr; R.filt = filt; R.escalate = escalate;
    if (!schedule_only) {
        const int64_t n_masks = (n_slices + spm - 1) / spm;
        ensure_mask(P, mask, n_masks * (int64_t)P->n1 * P->n2, x_mem, P->lanes[0].stream, &R.dmask);
        R.dmbits = pack_mask(P, R.dmask, n_masks, P->lanes[0].stream);
        if (escalate) {
            R.dmbits64 = f64_pack_mask(P->f64, R.dmask, n_masks, P->lanes[0].stream);
            compute_mhat(P, R.dmask, n_masks, P->lanes[0].stream);
        }
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
    std::vector<int64_t> firsts(sizes.size());
    { int64_t f = 0; for (size_t i = 0; i < sizes.size(); ++i) { firsts[i] = f; f += sizes[i]; } }

    // one feeder thread per lane: chunk c runs on lane c % lanes; a lane's waits (statistics read-back, list compaction)
    // never hold up the enqueueing of the other lanes
    std::vector<int> codes((size_t)lanes, P3D_OK);
    std::vector<std::string> msgs((size_t)lanes);
    auto feed = [&](int li) {
        try {
            cudaSetDevice(P->device);
            Lane& L = P->lanes[li];
            for (size_t c = (size_t)li; c < sizes.size(); c += (size_t)lanes) {
                collect_lane(R, L);
                if (escalate) process_chunk_esc(R, L, firsts[c], sizes[c]);
                else process_chunk(R, L, firsts[c], sizes[c]);
            }
            collect_lane(R, L);
            P3D_CUDA(cudaStreamSynchronize(L.stream));
        } catch (const P3dFail& f) { codes[li] = f.code; msgs[li] = get_error(); }
        catch (const std::exception& e) { codes[li] = P3D_ERR_CUDA; msgs[li] = e.what(); }
    };
    if (lanes == 1) feed(0);
    else {
        std::vector<std::thread> th;
        for (int li = 0; li < lanes; ++li) th.emplace_back(feed, li);
        for (auto& t : th) t.join();
    }
    for (int li = 0; li < lanes; ++li)
        if (codes[li] != P3D_OK) {
            for (int i = 0; i < lanes; ++i) { if (P->lanes[i].stream) cudaStreamSynchronize(P->lanes[i].stream); P->lanes[i].pending = false; }
            prof_collect(P);
            set_error("%s", msgs[li].c_str());
            throw P3dFail{codes[li]};
        }
    P->n_escalated = 0; P->n_esc_iters = 0;
    for (int li = 0; li < lanes; ++li) { P->n_escalated += P->lanes[li].n_escalated; P->n_esc_iters += P->lanes[li].n_esc_iters; }
    prof_collect(P);
    return P3D_OK;
}

}  // namespace

// data-driven schedule from a complex128 X0 on the device (shared with the float64 state mode, p3d_pocs_f64.cu)
namespace p3d {
size_t dd_schedule64_temp_bytes(long long ne) {
    size_t need = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, need, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                              (unsigned int*)nullptr, (unsigned int*)nullptr, ne, 0, 64, (cudaStream_t)0);
    return need;
}
void dd_schedule64_device(const Cx<double>* X0, long long ne, double lo_re, double lo_im, double hi_re, double hi_im, SliceStats* stats,
                          Cx<double>* tau64, int niter, unsigned long long* keys, unsigned int* vals, void* temp, size_t temp_bytes,
                          cudaStream_t st) {
    k_dd_keys64<<<std::min<long long>((ne + 255) / 256, 148 * 8), 256, 0, st>>>(X0, ne, lo_re, lo_im, hi_re, hi_im, keys, vals, stats);
    cub::DeviceRadixSort::SortPairsDescending(temp, temp_bytes, keys, keys + ne, vals, vals + ne, ne, 0, 64, st);
    k_pick_tau64<<<1, 128, 0, st>>>(keys + ne, vals + ne, X0, stats, tau64, niter);
}

// complex128 percentile operators: tau_sk holds the scheduled percentile q_k on entry and np.percentile(|X|, q_k) on exit
// (numpy's 'linear' method: virtual index q / 100 (n - 1), _lerp between the two neighbours); keys: 2 * ne entries
__global__ void k_abs_keys64(const Cx<double>* __restrict__ X, unsigned long long* __restrict__ keys, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const Cx<double> v = X[i];
        keys[i] = (unsigned long long)__double_as_longlong(sqrt(v.x * v.x + v.y * v.y));      // non-negative doubles order like their bit patterns
    }
}
__global__ void k_pick_percentile64(const unsigned long long* __restrict__ sorted, long long n, Cx<double>* tau_sk) {
    const double q = tau_sk->x;
    const double pos = q / 100.0 * (double)(n - 1);
    long long lo = (long long)floor(pos);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    const long long hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
    const double t = pos - (double)lo;
    const double a = __longlong_as_double((long long)sorted[lo]), b = __longlong_as_double((long long)sorted[hi]);
    const double diff = b - a;
    double r = a + diff * t;
    if (t >= 0.5) r = b - diff * (1.0 - t);
    *tau_sk = cmake<double>(r, 0.0);
}
size_t percentile64_temp_bytes(long long ne) {
    size_t need = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, (unsigned long long*)nullptr, (unsigned long long*)nullptr, ne, 0, 64, (cudaStream_t)0);
    return need;
}
void percentile64_device(const Cx<double>* X, long long ne, Cx<double>* tau_sk, unsigned long long* keys, void* temp, size_t temp_bytes,
                         cudaStream_t st) {
    k_abs_keys64<<<(unsigned)std::min<long long>((ne + 255) / 256, 148 * 8), 256, 0, st>>>(X, keys, ne);
    cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys, keys + ne, ne, 0, 64, st);
    k_pick_percentile64<<<1, 1, 0, st>>>(keys + ne, ne, tau_sk);
}
}  // namespace p3d

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
    if (P->mhat) cudaFree(P->mhat);
    if (P->mask_c64) cudaFree(P->mask_c64);
    if (P->mh_stats) cudaFree(P->mh_stats);
    if (P->mh_cand) cudaFree(P->mh_cand);
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
        prof_begin(P, P->events, st, 7);
        generic_fft2(generic_cfg(P), P->ax1.dev(), P->ax2.dev(), din + b0 * ne, dout + b0 * ne, nb, inverse, st);
        prof_end(P, P->events, st);
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

int p3d_plan_get_escalation(p3d_plan* P, int64_t* n_slices, int64_t* n_slice_iterations) {
    if (!P) return P3D_ERR_BAD_ARG;
    if (n_slices) *n_slices = P->n_escalated;
    if (n_slice_iterations) *n_slice_iterations = P->n_esc_iters;
    return P3D_OK;
}

int p3d_plan_describe(p3d_plan* P, char* buf, int64_t buflen) {
    if (!P || !buf || buflen <= 0) return P3D_ERR_BAD_ARG;
    std::string s = "iline axis: " + P->ax1.describe() + "; xline axis: " + P->ax2.describe();
    s += "; generic tiles: C=" + std::to_string(P->geom.C) + " cols x " + std::to_string(P->col_threads) + " thr, smem " +
         std::to_string(P->col_smem) + " B; RB=" + std::to_string(P->geom.RB) + " rows x " + std::to_string(P->row_threads) +
         " thr, smem " + std::to_string(P->row_smem) + " B";
    s += std::string("; cols_iter=") + ((P->spec.cols_iter && !P->force_generic) ? P->spec.cols_name : "generic");
    s += std::string("; rows_iter=") + ((P->spec.rows_iter && !P->force_generic) ? P->spec.rows_name : "generic");
    s += "; precision=" + (P->precision == 0 ? std::string("escalating(fp32->complex128, guard ") + std::to_string((long long)P->guard_factor) + ")" : std::to_string(P->precision));
    if (P->precision != 32) {
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
        if (value != 0 && value != 32 && value != 64) { set_error("precision must be 0 (escalating), 32 or 64"); return P3D_ERR_BAD_ARG; }
        P->precision = (int)value;
    }
    else if (!strcmp(key, "guard_factor")) { P->guard_factor = (double)value; if (value == 0) P->watch_guard_factor = 0.0; }
    else if (!strcmp(key, "watch_guard_factor")) P->watch_guard_factor = (double)value;
    else if (!strcmp(key, "seg_iters")) P->seg_iters = (int)std::max<int64_t>(1, value);
    else if (!strcmp(key, "watch_mode")) P->watch_mode = (int)value;
    else if (!strcmp(key, "use_tma")) P->use_tma = value != 0;
    else if (!strcmp(key, "pilot_min_elems")) P->pilot_min_elems = value;
    else if (!strcmp(key, "debug_fail_iter")) P->debug_fail_iter = (int)value;
    else if (!strcmp(key, "fused_replay_max")) P->fused_replay_max = (int)std::min<int64_t>(4096, std::max<int64_t>(0, value));
    else if (!strcmp(key, "support_cap")) P->support_cap = (int)std::max<int64_t>(0, value);
    else if (!strcmp(key, "arena_cap")) P->arena_cap = (int)std::min<int64_t>(32768, std::max<int64_t>(128, value));
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
