// Float64 state mode runner (p3d_pocs_f64.cu).
#pragma once
#include "p3d_host.h"
#include "p3d_pocs_launch.h"
#include "p3d_pocs_spec.cuh"

namespace p3d {

struct F64Runner;
F64Runner* f64_create(int device, int n1, int n2, AxisPlan* ax1, AxisPlan* ax2, size_t smem_optin);
void f64_destroy(F64Runner* R);
void f64_install_spec(F64Runner* R, int variant);
void f64_set_force_generic(F64Runner* R, int on);
const char* f64_cols_name(const F64Runner* R);
const char* f64_rows_name(const F64Runner* R);
// x / out: complex64 in host or device memory (x_mem / out_mem); dmask: DEVICE pointer
int f64_run(F64Runner* R, const p3d_pocs_params* pr, const Cx<float>* x, int x_mem, const uint8_t* dmask, int64_t spm,
            Cx<float>* out, int out_mem, int64_t n_slices, int32_t* niter_out, double* cost_out, double* costs_out,
            double* tau_out, bool schedule_only, int64_t max_slices);


// data-driven schedule from a complex128 X0 on the device (p3d_pocs.cu): tau64[k] = V[ceil(k (Nv - 1) / (niter - 1))] of the
// candidates strictly between (lo) and (hi) in numpy's complex ordering, sorted descending; keys / vals: 2 * ne entries each
size_t dd_schedule64_temp_bytes(long long ne);
void dd_schedule64_device(const Cx<double>* X0, long long ne, double lo_re, double lo_im, double hi_re, double hi_im, SliceStats* stats,
                          Cx<double>* tau64, int niter, unsigned long long* keys, unsigned int* vals, void* temp, size_t temp_bytes,
                          cudaStream_t st);

// percentile operators on complex128 state (p3d_pocs.cu): *tau_sk = np.percentile(|X|, tau_sk->x); keys: 2 * ne entries
size_t percentile64_temp_bytes(long long ne);
void percentile64_device(const Cx<double>* X, long long ne, Cx<double>* tau_sk, unsigned long long* keys, void* temp, size_t temp_bytes,
                         cudaStream_t st);

// ---- complex128 kernels for the escalating-precision engine (p3d_pocs.cu): the runner owns tile geometry, tables,
// register plans and the packed mask of its row kernel; the engine owns the state buffers
struct F64Kernels {
    GenericCfg cfg;
    AxisDev<double> a1, a2;
    SpecKernels64 spec;
    const Cx<double>* tw_cols; const Cx<double>* tw_rows;
    bool spec_cols, spec_rows;
    int cand_stride;             // per-CTA lexicographic maxima per slice written by the statistics kernel
};
F64Kernels f64_kernels(F64Runner* R);
const uint32_t* f64_pack_mask(F64Runner* R, const uint8_t* dmask, int64_t n_masks, cudaStream_t st);

}  // namespace p3d
