// Float64 state mode of the POCS path ("precision" = 64): complex128 iterate, spectrum and
// thresholds on the device, generic kernels instantiated for double.  Same ABI as the fp32 path
// (complex64 in, complex64 out = the float64 result rounded once at the end, like the reference's
// np.vectorize(otypes=[cube.dtype]) cast, cube_POCS_interpolation_3D.py:324).  One stream,
// chunked by device memory; built for exactness, not for speed.
#include "p3d_pocs_f64.h"
#include "p3d_pocs_launch.h"
#include "p3d_pocs_spec.cuh"
#include "p3d_schedule.h"

#include <algorithm>
#include <cstring>
#include <limits>

namespace p3d {

struct F64Runner {
    int device = 0, n1 = 0, n2 = 0;
    AxisPlan* ax1 = nullptr; AxisPlan* ax2 = nullptr;
    GenericCfg cfg{};
    cudaStream_t st = nullptr;
    int64_t cap = 0; int niter_cap = 0;
    Cx<double>* W = nullptr; Cx<double>* D = nullptr; Cx<double>* OUT = nullptr; Cx<double>* tau = nullptr;
    Cx<float>* io32 = nullptr;           // complex64 staging for host input / output
    double* S = nullptr; int* stop = nullptr; SliceStats* stats = nullptr;
    SpecKernels64 spec;                  // register-resident iteration kernels (1000 / 2000 / 256 / 200), else generic
    Cx<double>* tw_cols = nullptr; Cx<double>* tw_rows = nullptr;
    uint32_t* mbits = nullptr; int64_t mbits_words = 0;
    int force_generic = 0;
    unsigned long long* dd_keys = nullptr; unsigned int* dd_vals = nullptr; void* dd_temp = nullptr; size_t dd_temp_bytes = 0; long long dd_ne = 0;
    // percentile operators: spectra of a few slices, sort keys (2 ne) and sort scratch
    Cx<double>* pct_scr = nullptr; int64_t pct_slices = 0; unsigned long long* pct_keys = nullptr; void* pct_temp = nullptr; size_t pct_temp_bytes = 0;
    std::vector<Cx<double>> h_tau; std::vector<double> h_S; std::vector<int> h_stop; std::vector<SliceStats> h_stats;
};

static void f64_free_buffers(F64Runner* R) {
    void* ptrs[] = {R->W, R->D, R->OUT, R->tau, R->io32, R->S, R->stop, R->stats};
    for (void* p : ptrs) if (p) cudaFree(p);
    R->W = R->D = R->OUT = R->tau = nullptr; R->io32 = nullptr; R->S = nullptr; R->stop = nullptr; R->stats = nullptr;
    R->cap = 0; R->niter_cap = 0;
}

void f64_install_spec(F64Runner* R, int variant) {
    R->spec = select_spec_kernels64(R->n1, R->n2, variant);
    auto upload = [](const std::vector<int>& radices, Cx<double>** dst) {
        if (*dst) { cudaFree(*dst); *dst = nullptr; }
        if (radices.empty()) return;
        std::vector<Cx<double>> t = spec_twiddle_table64(radices);
        P3D_CUDA(cudaMalloc(dst, sizeof(Cx<double>) * t.size()));
        P3D_CUDA(cudaMemcpy(*dst, t.data(), sizeof(Cx<double>) * t.size(), cudaMemcpyHostToDevice));
    };
    if (R->spec.cols_table) {
        if (R->tw_cols) { cudaFree(R->tw_cols); R->tw_cols = nullptr; }
        std::vector<Cx<double>> t = R->spec.cols_table();
        P3D_CUDA(cudaMalloc(&R->tw_cols, sizeof(Cx<double>) * t.size()));
        P3D_CUDA(cudaMemcpy(R->tw_cols, t.data(), sizeof(Cx<double>) * t.size(), cudaMemcpyHostToDevice));
    } else upload(R->spec.cols_radices, &R->tw_cols);
    upload(R->spec.rows_radices, &R->tw_rows);
    R->mbits_words = 0;
}
void f64_set_force_generic(F64Runner* R, int on) { R->force_generic = on; }
const char* f64_cols_name(const F64Runner* R) { return (R->spec.cols_iter && !R->force_generic) ? R->spec.cols_name : "generic64"; }
const char* f64_rows_name(const F64Runner* R) { return (R->spec.rows_iter && !R->force_generic) ? R->spec.rows_name : "generic64"; }

F64Runner* f64_create(int device, int n1, int n2, AxisPlan* ax1, AxisPlan* ax2, size_t smem_optin) {
    F64Runner* R = new F64Runner();
    try {
        R->device = device; R->n1 = n1; R->n2 = n2; R->ax1 = ax1; R->ax2 = ax2;
        ax1->build64(); ax2->build64();
        const size_t budget = smem_optin - 2048, two = (228 * 1024) / 2 - 2048;
        const size_t es = sizeof(Cx<double>);
        const int L1 = ax1->L, L2 = ax2->L;
        int C = 8;
        while (C > 1 && (size_t)2 * L1 * C * es > two) C >>= 1;
        if (C < 2) { C = 2; while (C > 1 && (size_t)2 * L1 * C * es > budget) C >>= 1; }
        P3D_REQUIRE((size_t)2 * L1 * C * es <= budget, P3D_ERR_NOT_IMPLEMENTED, "iline axis %d too long for the float64 mode", n1);
        C = std::min(C, std::max(1, n2));
        const int pitch2 = L2 + 1;
        int RB = 8;
        while (RB > 1 && (size_t)2 * pitch2 * RB * es > two) RB >>= 1;
        P3D_REQUIRE((size_t)2 * pitch2 * RB * es <= budget, P3D_ERR_NOT_IMPLEMENTED, "xline axis %d too long for the float64 mode", n2);
        RB = std::min(RB, std::max(1, n1));
        PocsGeom& G = R->cfg.geom;
        G.n1 = n1; G.n2 = n2; G.C = C; G.RB = RB; G.pitch2 = pitch2; G.slices_per_mask = 1;
        R->cfg.col_smem = (size_t)2 * L1 * C * es; R->cfg.row_smem = (size_t)2 * pitch2 * RB * es;
        auto pick = [](long elems) { long t = ((elems / 8 + 31) / 32) * 32; return (int)std::min<long>(512, std::max<long>(128, t)); };
        R->cfg.col_threads = pick((long)L1 * C); R->cfg.row_threads = pick((long)L2 * RB);
        { cudaError_t e = generic64_configure(R->cfg); P3D_CUDA(e); }
        P3D_CUDA(cudaStreamCreateWithFlags(&R->st, cudaStreamNonBlocking));
        f64_install_spec(R, 0);
    } catch (...) { f64_destroy(R); throw; }
    return R;
}

void f64_destroy(F64Runner* R) {
    if (!R) return;
    f64_free_buffers(R);
    if (R->tw_cols) cudaFree(R->tw_cols);
    if (R->tw_rows) cudaFree(R->tw_rows);
    if (R->mbits) cudaFree(R->mbits);
    if (R->dd_keys) cudaFree(R->dd_keys);
    if (R->dd_vals) cudaFree(R->dd_vals);
    if (R->dd_temp) cudaFree(R->dd_temp);
    if (R->pct_scr) cudaFree(R->pct_scr);
    if (R->pct_keys) cudaFree(R->pct_keys);
    if (R->pct_temp) cudaFree(R->pct_temp);
    if (R->st) cudaStreamDestroy(R->st);
    delete R;
}

F64Kernels f64_kernels(F64Runner* R) {
    F64Kernels K;
    K.cfg = R->cfg; K.a1 = R->ax1->dev64(); K.a2 = R->ax2->dev64(); K.spec = R->spec;
    K.tw_cols = R->tw_cols; K.tw_rows = R->tw_rows;
    K.spec_cols = R->spec.cols_iter && R->spec.cols_stats && !R->force_generic;
    K.spec_rows = R->spec.rows_iter_io32 && R->spec.rows_init_io32 && !R->force_generic;
    const int C = K.spec_cols ? R->spec.cols_C : R->cfg.geom.C;
    K.cand_stride = (R->n2 + C - 1) / C;
    return K;
}

const uint32_t* f64_pack_mask(F64Runner* R, const uint8_t* dmask, int64_t n_masks, cudaStream_t st) {
    if (!R->spec.pack_mask || R->force_generic) return nullptr;
    const int64_t words = n_masks * (int64_t)R->n1 * R->spec.rows_T;
    if (words > R->mbits_words || !R->mbits) {
        if (R->mbits) cudaFree(R->mbits);
        R->mbits = nullptr; R->mbits_words = 0;
        P3D_CUDA(cudaMalloc(&R->mbits, sizeof(uint32_t) * words));
        R->mbits_words = words;
    }
    R->spec.pack_mask(dmask, R->mbits, (int)n_masks, R->n1, st);
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaStreamSynchronize(st));
    return R->mbits;
}

static void f64_ensure(F64Runner* R, int64_t n_slices, int niter, int64_t max_slices) {
    const int64_t ne = (int64_t)R->n1 * R->n2;
    int64_t want = n_slices;
    if (max_slices > 0) want = std::min(want, max_slices);
    if (R->cap >= want && R->niter_cap >= niter) return;
    f64_free_buffers(R);
    size_t free_b = 0, total_b = 0;
    P3D_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double per_slice = (double)ne * (3 * sizeof(Cx<double>) + sizeof(Cx<float>));
    int64_t cap = (int64_t)((double)free_b * 0.8 / per_slice);
    cap = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(cap, want), 30000));
    P3D_CUDA(cudaMalloc(&R->W, sizeof(Cx<double>) * ne * cap));
    P3D_CUDA(cudaMalloc(&R->D, sizeof(Cx<double>) * ne * cap));
    P3D_CUDA(cudaMalloc(&R->OUT, sizeof(Cx<double>) * ne * cap));
    P3D_CUDA(cudaMalloc(&R->io32, sizeof(Cx<float>) * ne * cap));
    P3D_CUDA(cudaMalloc(&R->tau, sizeof(Cx<double>) * cap * niter));
    P3D_CUDA(cudaMalloc(&R->S, sizeof(double) * cap * (niter + 1)));
    P3D_CUDA(cudaMalloc(&R->stop, sizeof(int) * cap));
    P3D_CUDA(cudaMalloc(&R->stats, sizeof(SliceStats) * cap));
    R->cap = cap; R->niter_cap = niter;
    R->h_tau.resize((size_t)cap * niter); R->h_S.resize((size_t)cap * (niter + 1)); R->h_stop.resize(cap); R->h_stats.resize(cap);
}

// numpy ordering of complex numbers
// percentile operators (threshold_operator.py:98-123 via functions/POCS.py:595-599): before iteration k the threshold of every
// slice becomes np.percentile(|colFFT(W_s)|, q_k) - column transform of the band into a scratch area (the statistics
// kernel with store_x0, W untouched), 64-bit radix sort of the moduli, linear interpolation between the two neighbours
static void f64_percentile_thresholds(F64Runner* R, const AxisDev<double>& a1, const BandArgs<double>& B, int nb, int k, cudaStream_t st) {
    const int64_t ne = (int64_t)R->n1 * R->n2;
    if (!R->pct_scr) {
        R->pct_slices = std::max<int64_t>(1, std::min<int64_t>(32, (int64_t)(256e6 / (16.0 * (double)ne))));
        P3D_CUDA(cudaMalloc(&R->pct_scr, sizeof(Cx<double>) * ne * R->pct_slices));
        P3D_CUDA(cudaMalloc(&R->pct_keys, sizeof(unsigned long long) * 2 * ne));
        R->pct_temp_bytes = percentile64_temp_bytes(ne);
        P3D_CUDA(cudaMalloc(&R->pct_temp, R->pct_temp_bytes));
    }
    for (int64_t s0 = 0; s0 < nb; s0 += R->pct_slices) {
        const int cnt = (int)std::min<int64_t>(R->pct_slices, nb - s0);
        BandArgs<double> Q = B;
        Q.W = B.W + s0 * ne; Q.OUT = R->pct_scr; Q.stats = B.stats + s0; Q.stop = B.stop + s0; Q.S = B.S + s0 * (B.niter + 1);
        Q.first_slice = B.first_slice + s0; Q.tau = B.tau + s0 * B.niter;
        Q.adaptive = 0; Q.accum = 1; Q.store_x0 = 1;          // the statistics they also accumulate are not used any more
        generic64_cols_stats(R->cfg, a1, Q, cnt, st);
        for (int i = 0; i < cnt; ++i)
            percentile64_device(R->pct_scr + (int64_t)i * ne, ne, const_cast<Cx<double>*>(B.tau) + (s0 + i) * B.niter + k, R->pct_keys,
                                R->pct_temp, R->pct_temp_bytes, st);
    }
    P3D_CUDA(cudaGetLastError());
}

static inline bool lex_less(const cd& a, const cd& b) { return a.real() < b.real() || (a.real() == b.real() && a.imag() < b.imag()); }

int f64_run(F64Runner* R, const p3d_pocs_params* prp, const Cx<float>* x, int x_mem, const uint8_t* dmask, int64_t spm,
            Cx<float>* out, int out_mem, int64_t n_slices, int32_t* niter_out, double* cost_out, double* costs_out,
            double* tau_out, bool schedule_only, int64_t max_slices) {
    const p3d_pocs_params& pr = *prp;
    const int niter = pr.niter;
    const int64_t ne = (int64_t)R->n1 * R->n2;
    f64_ensure(R, n_slices, niter, max_slices);
    cudaStream_t st = R->st;
    const bool data_driven = pr.thresh_model == P3D_MODEL_DATA_DRIVEN;
    const bool adaptive = pr.version == P3D_VERSION_ADAPTIVE;
    const AxisDev<double> a1 = R->ax1->dev64(), a2 = R->ax2->dev64();
    R->cfg.geom.slices_per_mask = (int)std::min<int64_t>(spm, 0x7fffffff);
    const bool spec_cols = R->spec.cols_iter && !R->force_generic, spec_rows = R->spec.rows_iter && !R->force_generic;
    if (spec_rows && !schedule_only) {
        // packed mask words of the register-resident row kernel (rebuilt every run: the mask may have changed)
        const int64_t n_masks = (n_slices + spm - 1) / spm;
        const int64_t words = n_masks * (int64_t)R->n1 * R->spec.rows_T;
        if (words > R->mbits_words || !R->mbits) {
            if (R->mbits) cudaFree(R->mbits);
            R->mbits = nullptr; R->mbits_words = 0;
            P3D_CUDA(cudaMalloc(&R->mbits, sizeof(uint32_t) * words));
            R->mbits_words = words;
        }
        R->spec.pack_mask(dmask, R->mbits, (int)n_masks, R->n1, st);
        P3D_CUDA(cudaGetLastError());
    }

    for (int64_t first = 0; first < n_slices; first += R->cap) {
        const int64_t count = std::min<int64_t>(R->cap, n_slices - first);
        // ---- input -> complex128
        if (x_mem == P3D_MEM_HOST) {
            P3D_CUDA(cudaMemcpyAsync(R->io32, x + first * ne, sizeof(Cx<float>) * ne * count, cudaMemcpyHostToDevice, st));
            convert_c64_to_c128(R->io32, R->D, ne * count, st);
        } else {
            convert_c64_to_c128(x + first * ne, R->D, ne * count, st);
        }
        P3D_CUDA(cudaMemsetAsync(R->S, 0, sizeof(double) * count * (niter + 1), st));
        P3D_CUDA(cudaMemsetAsync(R->stop, 0, sizeof(int) * count, st));
        for (int64_t i = 0; i < count; ++i) {
            memset(&R->h_stats[i], 0, sizeof(SliceStats));
            R->h_stats[i].minabs_bits = 0x7f800000u;
            R->h_stats[i].minabs64_key = ~0ull;
        }
        P3D_CUDA(cudaMemcpyAsync(R->stats, R->h_stats.data(), sizeof(SliceStats) * count, cudaMemcpyHostToDevice, st));
        P3D_CUDA(cudaStreamSynchronize(st));      // h_stats is pageable and reused below

        BandArgs<double> A;
        memset(&A, 0, sizeof(A));
        A.mask = dmask; A.mbits = spec_rows ? R->mbits : nullptr; A.exact_tie = 1; A.niter = niter; A.eps = pr.eps; A.alpha = pr.alpha;
        A.inv_n = 1.0 / ((double)R->n1 * (double)R->n2);
        A.W = R->W; A.D = R->D; A.OUT = R->OUT; A.first_slice = first; A.tau = R->tau; A.S = R->S; A.stop = R->stop; A.stats = R->stats;
        const int64_t band_max = 32768;

        for (int64_t b0 = 0; b0 < count; b0 += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, count - b0);
            BandArgs<double> B = A;
            B.W += b0 * ne; B.D += b0 * ne; B.OUT += b0 * ne; B.first_slice += b0; B.S += b0 * (niter + 1); B.stop += b0; B.stats += b0;
            B.adaptive = 0; B.accum = 1; B.store_x0 = 1;
            generic64_rows_init(R->cfg, a2, B, nb, st);
            generic64_cols_stats(R->cfg, a1, B, nb, st);
            generic64_lexmax_imag(B.OUT, B.stats, ne, nb, st);
        }
        P3D_CUDA(cudaGetLastError());
        P3D_CUDA(cudaMemcpyAsync(R->h_stats.data(), R->stats, sizeof(SliceStats) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaStreamSynchronize(st));

        // ---- schedule (host, double)
        std::vector<cd> tau;
        std::vector<cd> x0;
        for (int64_t i = 0; i < count; ++i) {
            const SliceStats& ss = R->h_stats[i];
            R->h_stop[i] = ss.nnz == 0 ? -1 : 0;
            ScheduleStats s;
            s.z = cd(f64_from_ordered(ss.re64_key), f64_from_ordered(ss.im64_key));
            s.sumsq = ss.sumsq; s.vmax = f64_from_ordered(ss.maxabs64_key); s.vmin = f64_from_ordered(ss.minabs64_key);
            bool is_real = false;
            if (!data_driven) {
                host_schedule(pr, s, ne, tau, is_real);
            } else {
                tau.assign(niter, cd(0, 0));
                if (!R->h_stop[i]) {
                    cd tmin, tmax;
                    schedule_bounds(pr, s, ne, tmin, tmax);
                    // order statistics on the device: radix sort of the ordered real parts, ties on the imaginary parts
                    if (R->dd_ne < ne) {
                        if (R->dd_keys) cudaFree(R->dd_keys);
                        if (R->dd_vals) cudaFree(R->dd_vals);
                        if (R->dd_temp) cudaFree(R->dd_temp);
                        R->dd_keys = nullptr; R->dd_vals = nullptr; R->dd_temp = nullptr; R->dd_ne = 0;
                        P3D_CUDA(cudaMalloc(&R->dd_keys, sizeof(unsigned long long) * 2 * ne));
                        P3D_CUDA(cudaMalloc(&R->dd_vals, sizeof(unsigned int) * 2 * ne));
                        R->dd_temp_bytes = dd_schedule64_temp_bytes(ne);
                        P3D_CUDA(cudaMalloc(&R->dd_temp, R->dd_temp_bytes));
                        R->dd_ne = ne;
                    }
                    dd_schedule64_device(R->OUT + i * ne, ne, tmin.real(), tmin.imag(), tmax.real(), tmax.imag(), R->stats + i,
                                         R->tau + i * niter, niter, R->dd_keys, R->dd_vals, R->dd_temp, R->dd_temp_bytes, st);
                    P3D_CUDA(cudaGetLastError());
                    std::vector<Cx<double>> ht((size_t)niter);
                    SliceStats hs;
                    P3D_CUDA(cudaMemcpyAsync(ht.data(), R->tau + i * niter, sizeof(Cx<double>) * niter, cudaMemcpyDeviceToHost, st));
                    P3D_CUDA(cudaMemcpyAsync(&hs, R->stats + i, sizeof(SliceStats), cudaMemcpyDeviceToHost, st));
                    P3D_CUDA(cudaStreamSynchronize(st));
                    P3D_REQUIRE(hs.n_cand > 0, P3D_ERR_NUMERIC, "data-driven schedule: no coefficient between tau_min and tau_max in slice %lld", (long long)(first + i));
                    for (int k = 0; k < niter; ++k) tau[k] = cd(ht[k].x, ht[k].y);
                }
            }
            if (pr.sqrt_decay) apply_sqrt_decay(tau, is_real);
            for (int k = 0; k < niter; ++k) {
                R->h_tau[i * niter + k] = cmake<double>(tau[k].real(), tau[k].imag());
                if (tau_out) { tau_out[((first + i) * niter + k) * 2] = tau[k].real(); tau_out[((first + i) * niter + k) * 2 + 1] = tau[k].imag(); }
            }
        }
        if (schedule_only) continue;
        P3D_CUDA(cudaMemcpyAsync(R->tau, R->h_tau.data(), sizeof(Cx<double>) * count * niter, cudaMemcpyHostToDevice, st));
        P3D_CUDA(cudaMemcpyAsync(R->stop, R->h_stop.data(), sizeof(int) * count, cudaMemcpyHostToDevice, st));
        P3D_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < count; ++i)
            if (R->h_stop[i] < 0)
                P3D_CUDA(cudaMemcpyAsync(R->OUT + i * ne, R->D + i * ne, sizeof(Cx<double>) * ne, cudaMemcpyDeviceToDevice, st));

        for (int64_t b0 = 0; b0 < count; b0 += band_max) {
            const int nb = (int)std::min<int64_t>(band_max, count - b0);
            BandArgs<double> B = A;
            B.W += b0 * ne; B.D += b0 * ne; B.OUT += b0 * ne; B.first_slice += b0; B.tau += b0 * niter; B.S += b0 * (niter + 1); B.stop += b0; B.stats += b0;
            B.adaptive = adaptive ? 1 : 0;
            if (adaptive) { B.accum = 0; B.store_x0 = 0; generic64_rows_init(R->cfg, a2, B, nb, st); }
            for (int k = 0; k < niter; ++k) {
                B.k = k; B.last = (k == niter - 1) ? 1 : 0;
                B.write_out = (B.last || (pr.eps > 0.0 && k >= 3)) ? 1 : 0;
                if (pr.thresh_percentile) f64_percentile_thresholds(R, a1, B, nb, k, st);
                if (spec_cols) R->spec.cols_iter(R->cfg.geom, R->tw_cols, B, nb, pr.thresh_op, st);
                else generic64_cols_iter(R->cfg, a1, B, nb, pr.thresh_op, st);
                if (spec_rows) R->spec.rows_iter(R->cfg.geom, R->tw_rows, B, nb, st);
                else generic64_rows_iter(R->cfg, a2, B, nb, st);
            }
        }
        P3D_CUDA(cudaGetLastError());

        // ---- results: one rounding to complex64
        if (out_mem == P3D_MEM_HOST) {
            convert_c128_to_c64(R->OUT, R->io32, ne * count, st);
            P3D_CUDA(cudaMemcpyAsync(out + first * ne, R->io32, sizeof(Cx<float>) * ne * count, cudaMemcpyDeviceToHost, st));
        } else {
            convert_c128_to_c64(R->OUT, out + first * ne, ne * count, st);
        }
        P3D_CUDA(cudaMemcpyAsync(R->h_S.data(), R->S, sizeof(double) * count * (niter + 1), cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaMemcpyAsync(R->h_stop.data(), R->stop, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
        P3D_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < count; ++i) {
            const int64_t s = first + i;
            const int sp = R->h_stop[i];
            const int nit = sp < 0 ? 0 : (sp > 0 ? sp : niter);
            if (niter_out) niter_out[s] = nit;
            const double* S = R->h_S.data() + i * (niter + 1);
            double last = 0.0;
            for (int k = 0; k < niter; ++k) {
                double c = std::numeric_limits<double>::quiet_NaN();
                if (k < nit) { const double d = S[k + 1] - S[k]; c = (d * d) / (S[k + 1] * S[k + 1]); last = c; }
                if (costs_out) costs_out[s * niter + k] = c;
            }
            if (cost_out) cost_out[s] = last;
        }
    }
    return P3D_OK;
}

}  // namespace p3d
