// Host-side helpers shared by the translation units of libp3d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/p3d_b200.h"
#include "p3d_fft_generic.cuh"

namespace p3d {

void set_error(const char* fmt, ...);
const char* get_error();

struct P3dFail { int code; };

#define P3D_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            p3d::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__,       \
                           __LINE__, cudaGetErrorString(e_));                                  \
            throw p3d::P3dFail{e_ == cudaErrorMemoryAllocation ? P3D_ERR_OOM : P3D_ERR_CUDA};  \
        }                                                                                      \
    } while (0)

#define P3D_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                       \
        if (!(cond)) { p3d::set_error(__VA_ARGS__); throw p3d::P3dFail{code}; }                \
    } while (0)

// ------------------------------------------------------------------------------------------
// Axis plan: factorisation + device tables for one transform length
// ------------------------------------------------------------------------------------------
struct AxisPlan {
    int n = 0, L = 0;
    bool bluestein = false;
    std::vector<int> radix;
    Cx<float>* d_tw = nullptr;
    Cx<float>* d_chirp = nullptr;
    Cx<float>* d_bfilt = nullptr;
    Cx<double>* d_tw64 = nullptr;       // float64 tables, built on demand (build64)
    Cx<double>* d_chirp64 = nullptr;
    Cx<double>* d_bfilt64 = nullptr;

    void build(int n);          // allocates device tables on the current device
    void build64();             // additionally the float64 tables (precision = 64 mode)
    void release();
    AxisDev<float> dev() const;
    AxisDev<double> dev64() const;
    std::string describe() const;
};

bool is_smooth(int n);                       // all prime factors in {2,3,5,7,11,13}
std::vector<int> factorize_radices(int n);   // radices for a smooth n
int  bluestein_length(int n);                // smooth L >= 2n-1 with the cheapest passes
void host_fft(std::vector<std::complex<double>>& a, bool inverse);   // any length, O(n^2) fallback for table setup

}  // namespace p3d
