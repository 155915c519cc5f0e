/* p3d_b200.h -- C ABI of the B200-native FFT-POCS hot path.
 *
 * Drop-in boundary for the one data-parallel path of fwrnke/pseudo-3D-interpolation:
 *   - per-slice POCS with the FFT transform  (pseudo_3D_interpolation/functions/POCS.py:371-656,
 *     called once per slice at pseudo_3D_interpolation/cube_POCS_interpolation_3D.py:314-340)
 *   - threshold schedule                      (functions/POCS.py:169-368)
 *   - threshold operators                     (functions/threshold_operator.py:20-123)
 *   - time-axis forward / inverse transforms  (cube_apply_FFT.py:240-254, cube_apply_IFFT.py:83-94)
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative p3d_status on failure (no exception crosses the ABI); p3d_last_error() gives
 * the thread-local message.  The caller owns every buffer it passes.  A plan is bound to
 * one CUDA device and is not re-entrant: use one host thread (or process) per GPU.
 * Complex values are interleaved (re, im) float32 pairs (numpy complex64).
 */
#ifndef P3D_B200_H
#define P3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3D_ABI_VERSION 3

typedef enum {
    P3D_OK = 0,
    P3D_ERR_BAD_ARG = -1,       /* ValueError in the Python mirror                      */
    P3D_ERR_NOT_IMPLEMENTED = -2, /* NotImplementedError in the Python mirror           */
    P3D_ERR_CUDA = -3,          /* CUDA runtime failure                                  */
    P3D_ERR_OOM = -4,           /* device or pinned-host allocation failed               */
    P3D_ERR_NUMERIC = -5        /* e.g. data-driven schedule with an empty candidate set */
} p3d_status;

/* thresh_op of POCS_algorithm (functions/POCS.py:91-102) */
typedef enum { P3D_OP_HARD = 0, P3D_OP_SOFT = 1, P3D_OP_GARROTE = 2 } p3d_thresh_op;
/* thresh_model of get_threshold_decay (functions/POCS.py:251,348-362) */
typedef enum {
    P3D_MODEL_LINEAR = 0, P3D_MODEL_EXPONENTIAL = 1, P3D_MODEL_DATA_DRIVEN = 2,
    P3D_MODEL_INVERSE_PROPORTIONAL = 3
} p3d_thresh_model;
/* version of POCS_algorithm (functions/POCS.py:564-575); FAST is numerically REGULAR (x_old aliases x_inv) */
typedef enum { P3D_VERSION_REGULAR = 0, P3D_VERSION_FAST = 1, P3D_VERSION_ADAPTIVE = 2 } p3d_pocs_version;
/* where a data pointer lives */
typedef enum { P3D_MEM_HOST = 0, P3D_MEM_DEVICE = 1 } p3d_mem;

/* Keyword arguments of POCS_algorithm (functions/POCS.py:371-391). */
typedef struct {
    int32_t niter;            /* maximum number of iterations                                   */
    int32_t thresh_op;        /* p3d_thresh_op                                                   */
    int32_t thresh_model;     /* p3d_thresh_model                                                */
    int32_t version;          /* p3d_pocs_version                                                */
    double  q;                /* exponent of 'exponential-q' / 'inverse_proportional-q' (1.0)    */
    double  eps;              /* convergence threshold on the cost (stop when iter > 2)          */
    double  alpha;            /* re-insertion weight                                             */
    double  p_max;
    double  p_min;            /* ignored when p_min_adaptive != 0                                */
    int32_t p_min_adaptive;   /* p_min == 'adaptive': tau_min = 0.01 * rms(|X0|)                 */
    int32_t sqrt_decay;       /* use sqrt(tau_k) (principal complex root)                        */
    int32_t decay_factors;    /* decay_kind == 'factors': tau_max = p_max, tau_min = p_min       */
    int32_t absmax_threshold; /* 0 = reference-exact complex lexicographic max (default);
                                 1 = real max|X0| ("max-amplitude" of the docs), opt-in extra   */
    int32_t thresh_percentile;/* '<op>-percentile' operators (functions/POCS.py:43-58): the scheduled value
                                 tau_k in [0, 100] is a percentile and the threshold of iteration k is
                                 np.percentile(|X_k|, tau_k) of the slice's current spectrum; requires
                                 decay_factors != 0 and a linear / exponential schedule (as the reference,
                                 where anything else fails inside np.percentile)                    */
} p3d_pocs_params;

typedef struct p3d_plan p3d_plan;

int         p3d_abi_version(void);
const char* p3d_last_error(void);
int         p3d_device_count(void);

/* Plan for slices of n_iline x n_xline (row-major, xline contiguous) on CUDA device `device`.
 * max_slices: upper bound of slices processed per internal chunk (0 = choose from free memory).
 * band_slices: slices per launch group ("band"); 0 = auto (band scheduler). */
int p3d_plan_create(p3d_plan** plan, int device, int n_iline, int n_xline, int64_t max_slices,
                    int band_slices);
int p3d_plan_destroy(p3d_plan* plan);

/* Interpolate n_slices independent slices: the loop of cube_POCS_interpolation_3D.py:314-340
 * with POCS_algorithm(x2d, mask2d, None, transform_kind='FFT', **params) per slice.
 *   x, out : (n_slices, n_iline, n_xline) complex64, host or device memory (x_mem/out_mem)
 *   mask   : (n_masks, n_iline, n_xline) uint8 in {0,1}, same memory kind as x; slice s uses
 *            mask[s / slices_per_mask] (n_masks = ceil(n_slices / slices_per_mask));
 *            pass slices_per_mask = n_slices for the single shared mask of one cube
 *   niter_out[n_slices]          : iterations executed per slice (host, may be NULL)
 *   cost_out[n_slices]           : last cost per slice (host, may be NULL)
 *   costs_out[n_slices * niter]  : cost history, NaN-padded after the last iteration (host, may be NULL)
 * Host buffers may be pageable; pinned buffers (p3d_host_alloc) make the copies asynchronous.
 * Device buffers: x is read in place during every iteration and out is written while x is still needed, so
 * [x, x + n) and [out, out + n) must not overlap (P3D_ERR_BAD_ARG otherwise); x is never modified. */
int p3d_pocs_run(p3d_plan* plan, const p3d_pocs_params* params,
                 const void* x, int x_mem, const uint8_t* mask, int64_t slices_per_mask,
                 void* out, int out_mem, int64_t n_slices,
                 int32_t* niter_out, double* cost_out, double* costs_out);

/* Threshold schedule only (get_threshold_decay, functions/POCS.py:169-368) for n_slices
 * slices: tau_out[n_slices * niter * 2] doubles (re, im).  x as in p3d_pocs_run. */
int p3d_pocs_schedule(p3d_plan* plan, const p3d_pocs_params* params, const void* x, int x_mem,
                      int64_t n_slices, double* tau_out);

/* Forward / inverse 2-D FFT of slices (numpy.fft.fft2 / ifft2 semantics, the `transform` /
 * `itransform` callables of cube_POCS_interpolation_3D.py:255-257); diagnostic + test entry. */
int p3d_fft2(p3d_plan* plan, const void* x, int x_mem, void* out, int out_mem, int64_t n_slices,
             int inverse);

/* kx-ky domain filter of a stack of slices: out = ifft2(filt * fft2(x)) per slice, the arithmetic of
 * remove_acquisition_footprint (cube_postprocessing_3D.py:254) and spatial_antialiasing (:342); the
 * caller takes the real part.  Runs the three fused passes of one POCS iteration (fp32).
 *   x, out : (n_slices, n_iline, n_xline) complex64, host or device
 *   filt   : (n_iline, n_xline) float32 in FFT order (the reference's np.fft.ifftshift(ffilter)) */
int p3d_kxky_filter_run(p3d_plan* plan, const void* x, int x_mem, const float* filt, int filt_mem,
                        void* out, int out_mem, int64_t n_slices);

/* Time-axis forward transform of step 12 (cube_apply_FFT.py:240-254):
 *   x   : (nt, n_traces) float32, time-major            (n_traces = n_iline * n_xline)
 *   out : (nf, n_traces) complex64, nf = nfft/2+1 (compute_real) or nfft, fftfreq/rfftfreq order
 *   F[k] = dt * exp(-2 pi i f_k t0) * sum_n x[n] exp(-2 pi i k n / nfft) * window[k]
 * nfft >= nt zero-pads (``--upsampling-factor``); window may be NULL (nf doubles otherwise). */
int p3d_time_fft(int device, const float* x, int x_mem, void* out, int out_mem,
                 int64_t nt, int64_t nfft, int64_t n_traces, double dt, double t0,
                 int compute_real, const double* window);
/* Time-axis inverse transform of step 14 (cube_apply_IFFT.py:83-94):
 *   x   : (nf, n_traces) complex64; rows ascending in frequency (fftshift order) when
 *         ascending != 0 and compute_real == 0, else fftfreq / rfftfreq order
 *   out : (nt_out, n_traces) float32 = first nt_out samples of (1/dt) * Re IDFT(F * exp(+2 pi i f t0)) */
int p3d_time_ifft(int device, const void* x, int x_mem, float* out, int out_mem,
                  int64_t nfft, int64_t nt_out, int64_t n_traces, double dt, double t0,
                  int compute_real, int ascending);

/* Amplitude envelope along the time axis: env = |x + i Hilbert(x)|, the arithmetic of functions/signal.py:672-690
 * (scipy.signal.hilbert + abs) used by step 11 (cube_preprocessing_3D.py:341-353) to make the `env` variable.
 *   x, out : (nt, n_traces) float32, time-major; any nt (register plans for 512 / 1000 / 1024 / 2000 / 2048 / 2500 /
 *            3000 / 4000 / 4096 / 5000 samples, generic kernels otherwise) */
int p3d_time_envelope(int device, const float* x, int x_mem, float* out, int out_mem, int64_t nt, int64_t n_traces);

/* Device time (CUDA events) spent in the kernels of this thread's last p3d_time_fft / p3d_time_ifft
 * call (0 when the call used the generic direct kernels). */
int p3d_time_last_kernel_ms(double* ms);
/* Which kernels served this thread's last p3d_time_* call: "tma" (one pass: TMA-staged tiles, register FFT; record
 * lengths 512 / 1000 / 1024 / 2000 / 2048 / 2500 / 4000 / 4096 with a trace count divisible by 4 and 16-byte aligned
 * device buffers; 2048 / 4096 / 2000 / 4000 as two half-length transforms and a radix-2 step), "pipeline"
 * (transpose / FFT / transpose through L2), "direct" or "generic" (any length).  Environment: P3D_TIME_PATH =
 * tma | pipeline | direct selects the first candidate (read at every call). */
const char* p3d_time_last_path(void);

/* Pinned host memory for asynchronous copies. */
int p3d_host_alloc(void** ptr, int64_t bytes);
int p3d_host_free(void* ptr);
/* Device memory helpers for callers without a CUDA runtime of their own (tests, bench). */
int p3d_device_alloc(int device, void** ptr, int64_t bytes);
int p3d_device_free(int device, void* ptr);
int p3d_memcpy(int device, void* dst, const void* src, int64_t bytes, int kind /*0 h2d,1 d2h,2 d2d*/);
int p3d_device_synchronize(int device);

/* Per-kernel device-time accounting (CUDA events on the launch stream).
 * kinds: 0 rows_init, 1 cols_stats, 2 cols_iter, 3 rows_iter, 4 time_fft, 5 time_ifft, 6 sort, 7 fft2,
 *        8 cols_iter64, 9 rows_iter64, 10 init64, 11 replay (complex128 side of the escalating-precision mode) */
#define P3D_PROFILE_KINDS 12
int p3d_plan_set_profiling(p3d_plan* plan, int enabled);
int p3d_plan_get_profile(p3d_plan* plan, double* ms_per_kind, int64_t* launches_per_kind, int reset);
/* Device-side timing of whole runs: record a CUDA event in slot (0..7) on the plan's first
 * stream (every p3d_pocs_run returns with all of its streams drained), and read the elapsed
 * time between two recorded slots (synchronises on the later one). */
int p3d_plan_event_record(p3d_plan* plan, int slot);
int p3d_plan_event_elapsed_ms(p3d_plan* plan, int slot_a, int slot_b, double* ms);
/* Human-readable description of the chosen kernels / tiles / bands (for DESIGN.md and bench). */
int p3d_plan_describe(p3d_plan* plan, char* buf, int64_t buflen);
/* Options: key in {"band_slices","force_generic","lanes","max_slices","spec_variant","spec_variant64",
 *   "precision"    0 (default) = escalating: an fp32 pilot decides the thresholds while the spectrum is sparse and records the
 *                  surviving coefficients, the float64 trajectory is rebuilt exactly from that record (sparse-domain
 *                  replay) and the remaining iterations run on complex128 state (meets the 1e-4 of the float64 reference -
 *                  in fact reproduces the float64 mode - at about 60 % of the fp32 rate); 32 = fp32 only (fastest;
 *                  threshold decisions may differ from float64 once tau_k reaches the dense part of the spectrum: 1e-3);
 *                  64 = complex128 throughout,
 *   "guard_factor" half-width of the guard band in units of 2^-24 * rms|X| (default 1024; 0 disables the switch),
 *   "watch_mode"   -1 (default) / 0 / 1: guard-band hits are recorded and verified exactly by the float64 replay instead
 *                  of freezing the slice (-1: for slices of 400 k points and more),
 *   "arena_cap"    support-record entries per slice (default 16384), "support_cap" largest support replayed per
 *                  iteration (0 = 2.2 sqrt(n_iline n_xline)), "pilot_min_elems" slices smaller than this skip the fp32
 *                  pilot (default 0), "fused_replay_max" largest support replayed by the one-launch kernel (1024), "seg_iters" iterations between two compactions of the fp32 slice list (4),
 *   "lanes"        copy / compute lanes of the host-buffer path, each with its own stream, buffers and feeder thread
 *                  (0 = default: 8, or 4 when LOCAL_WORLD_SIZE ranks x 8 spinning threads would exceed half of the host's
 *                  hardware threads; device-resident data use one lane), "max_slices" largest chunk of slices per lane,
 *   "use_tma"      1 (default) / 0: column tiles fetched with cp.async.bulk.tensor where the tile shape allows it,
 *   "debug_fail_iter" testing only: the replay reports a failed verification at this iteration (-1 = off)} */
int p3d_plan_set_option(p3d_plan* plan, const char* key, int64_t value);
/* Escalating mode, last p3d_pocs_run of this plan: slices that switched to complex128 and the slice-iterations they
 * ran there (either pointer may be NULL). */
int p3d_plan_get_escalation(p3d_plan* plan, int64_t* n_slices, int64_t* n_slice_iterations);

#ifdef __cplusplus
}
#endif
#endif /* P3D_B200_H */
