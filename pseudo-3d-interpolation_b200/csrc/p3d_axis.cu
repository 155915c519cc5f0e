// Axis plans (factorisation, twiddle / chirp tables) and small host utilities.
#include "p3d_host.h"

#include <cmath>
#include <cstring>

namespace p3d {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

bool is_smooth(int n) {
    if (n < 1) return false;
    const int primes[] = {2, 3, 5, 7, 11, 13};
    for (int p : primes) while (n % p == 0) n /= p;
    return n == 1;
}

// Greedy choice of pass radices: few passes first (16, 10, 8 ...), primes last.
std::vector<int> factorize_radices(int n) {
    std::vector<int> r;
    const int pref[] = {16, 10, 8, 5, 4, 3, 2, 7, 11, 13};
    int m = n;
    while (m > 1) {
        bool found = false;
        for (int p : pref) {
            if (m % p == 0) {
                // avoid a trailing radix-2 when 16 could be split as 8*... (e.g. 32 = 8*4, not 16*2)
                if (p == 16 && (m / 16) == 2) continue;
                r.push_back(p); m /= p; found = true; break;
            }
        }
        if (!found) return {};
    }
    return r;          // n == 1: no pass at all
}

int bluestein_length(int n) {
    const long need = 2L * n - 1;
    long best = -1; double best_cost = 1e300;
    // candidates 2^a 3^b 5^c in [need, 4*need]
    for (long a = 1; a <= 8 * need; a *= 2)
        for (long b = a; b <= 8 * need; b *= 3)
            for (long c = b; c <= 8 * need; c *= 5) {
                if (c < need || c > 4 * need) continue;
                auto rad = factorize_radices((int)c);
                double cost = (double)c * (double)rad.size();
                if (cost < best_cost) { best_cost = cost; best = c; }
            }
    return (int)best;
}

static void fft_rec(std::complex<double>* a, int n, int stride, std::complex<double>* out, double sign) {
    // out[k] = sum_j a[j*stride] exp(sign 2 pi i jk/n)
    if (n == 1) { out[0] = a[0]; return; }
    int p = 0;
    for (int q = 2; q * q <= n; ++q) if (n % q == 0) { p = q; break; }
    if (p == 0) {           // prime: direct DFT
        for (int k = 0; k < n; ++k) {
            std::complex<double> s = 0;
            for (int j = 0; j < n; ++j) {
                double ang = sign * 2.0 * M_PI * (double)(((long)j * k) % n) / n;
                s += a[(long)j * stride] * std::complex<double>(cos(ang), sin(ang));
            }
            out[k] = s;
        }
        return;
    }
    int m = n / p;
    std::vector<std::complex<double>> sub((size_t)n);
    for (int r = 0; r < p; ++r) fft_rec(a + (long)r * stride, m, stride * p, sub.data() + (long)r * m, sign);
    for (int k = 0; k < n; ++k) {
        std::complex<double> s = 0;
        for (int r = 0; r < p; ++r) {
            double ang = sign * 2.0 * M_PI * (double)(((long)r * k) % n) / n;
            s += sub[(long)r * m + (k % m)] * std::complex<double>(cos(ang), sin(ang));
        }
        out[k] = s;
    }
}

void host_fft(std::vector<std::complex<double>>& a, bool inverse) {
    std::vector<std::complex<double>> out(a.size());
    fft_rec(a.data(), (int)a.size(), 1, out.data(), inverse ? +1.0 : -1.0);
    a.swap(out);
}

void AxisPlan::build(int n_) {
    release();
    n = n_;
    P3D_REQUIRE(n >= 1, P3D_ERR_BAD_ARG, "axis length must be >= 1 (got %d)", n);
    bluestein = !is_smooth(n);
    L = bluestein ? bluestein_length(n) : n;
    radix = factorize_radices(L);
    P3D_REQUIRE((L == 1 || !radix.empty()) && (int)radix.size() <= P3D_MAX_PASSES, P3D_ERR_NOT_IMPLEMENTED,
                "cannot factorise transform length %d", L);
    std::vector<Cx<float>> tw((size_t)L);
    for (int t = 0; t < L; ++t) {
        double ang = -2.0 * M_PI * (double)t / (double)L;
        tw[t] = cmake<float>((float)cos(ang), (float)sin(ang));
    }
    P3D_CUDA(cudaMalloc(&d_tw, sizeof(Cx<float>) * L));
    P3D_CUDA(cudaMemcpy(d_tw, tw.data(), sizeof(Cx<float>) * L, cudaMemcpyHostToDevice));
    if (bluestein) {
        std::vector<std::complex<double>> w((size_t)n);
        for (int j = 0; j < n; ++j) {
            long j2 = ((long)j * j) % (2L * n);
            double ang = -M_PI * (double)j2 / (double)n;
            w[j] = std::complex<double>(cos(ang), sin(ang));
        }
        std::vector<std::complex<double>> b((size_t)L, 0.0);
        for (int j = 0; j < n; ++j) {
            b[j] = std::conj(w[j]);
            if (j > 0) b[L - j] = std::conj(w[j]);
        }
        host_fft(b, false);
        std::vector<Cx<float>> ch((size_t)n), bf((size_t)L);
        for (int j = 0; j < n; ++j) ch[j] = cmake<float>((float)w[j].real(), (float)w[j].imag());
        for (int j = 0; j < L; ++j) bf[j] = cmake<float>((float)(b[j].real() / L), (float)(b[j].imag() / L));
        P3D_CUDA(cudaMalloc(&d_chirp, sizeof(Cx<float>) * n));
        P3D_CUDA(cudaMemcpy(d_chirp, ch.data(), sizeof(Cx<float>) * n, cudaMemcpyHostToDevice));
        P3D_CUDA(cudaMalloc(&d_bfilt, sizeof(Cx<float>) * L));
        P3D_CUDA(cudaMemcpy(d_bfilt, bf.data(), sizeof(Cx<float>) * L, cudaMemcpyHostToDevice));
    }
}

void AxisPlan::build64() {
    if (d_tw64) return;
    std::vector<Cx<double>> tw((size_t)L);
    for (int t = 0; t < L; ++t) {
        const double ang = -2.0 * M_PI * (double)t / (double)L;
        tw[t] = cmake<double>(cos(ang), sin(ang));
    }
    P3D_CUDA(cudaMalloc(&d_tw64, sizeof(Cx<double>) * L));
    P3D_CUDA(cudaMemcpy(d_tw64, tw.data(), sizeof(Cx<double>) * L, cudaMemcpyHostToDevice));
    if (bluestein) {
        std::vector<std::complex<double>> w((size_t)n);
        for (int j = 0; j < n; ++j) {
            const long j2 = ((long)j * j) % (2L * n);
            const double ang = -M_PI * (double)j2 / (double)n;
            w[j] = std::complex<double>(cos(ang), sin(ang));
        }
        std::vector<std::complex<double>> b((size_t)L, 0.0);
        for (int j = 0; j < n; ++j) { b[j] = std::conj(w[j]); if (j > 0) b[L - j] = std::conj(w[j]); }
        host_fft(b, false);
        std::vector<Cx<double>> ch((size_t)n), bf((size_t)L);
        for (int j = 0; j < n; ++j) ch[j] = cmake<double>(w[j].real(), w[j].imag());
        for (int j = 0; j < L; ++j) bf[j] = cmake<double>(b[j].real() / L, b[j].imag() / L);
        P3D_CUDA(cudaMalloc(&d_chirp64, sizeof(Cx<double>) * n));
        P3D_CUDA(cudaMemcpy(d_chirp64, ch.data(), sizeof(Cx<double>) * n, cudaMemcpyHostToDevice));
        P3D_CUDA(cudaMalloc(&d_bfilt64, sizeof(Cx<double>) * L));
        P3D_CUDA(cudaMemcpy(d_bfilt64, bf.data(), sizeof(Cx<double>) * L, cudaMemcpyHostToDevice));
    }
}

void AxisPlan::release() {
    if (d_tw) cudaFree(d_tw);
    if (d_chirp) cudaFree(d_chirp);
    if (d_bfilt) cudaFree(d_bfilt);
    if (d_tw64) cudaFree(d_tw64);
    if (d_chirp64) cudaFree(d_chirp64);
    if (d_bfilt64) cudaFree(d_bfilt64);
    d_tw = d_chirp = d_bfilt = nullptr;
    d_tw64 = d_chirp64 = d_bfilt64 = nullptr;
    radix.clear();
}

AxisDev<double> AxisPlan::dev64() const {
    AxisDev<double> a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.L = L; a.npass = (int)radix.size();
    for (int i = 0; i < a.npass; ++i) a.radix[i] = radix[i];
    a.bluestein = bluestein ? 1 : 0;
    a.tw = d_tw64; a.chirp = d_chirp64; a.bfilt = d_bfilt64;
    return a;
}

AxisDev<float> AxisPlan::dev() const {
    AxisDev<float> a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.L = L; a.npass = (int)radix.size();
    for (int i = 0; i < a.npass; ++i) a.radix[i] = radix[i];
    a.bluestein = bluestein ? 1 : 0;
    a.tw = d_tw; a.chirp = d_chirp; a.bfilt = d_bfilt;
    return a;
}

std::string AxisPlan::describe() const {
    std::string s = "n=" + std::to_string(n);
    if (bluestein) s += " bluestein L=" + std::to_string(L);
    s += " radices=";
    for (size_t i = 0; i < radix.size(); ++i) s += (i ? "x" : "") + std::to_string(radix[i]);
    return s;
}

}  // namespace p3d
