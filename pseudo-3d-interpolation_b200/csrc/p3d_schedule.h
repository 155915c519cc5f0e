// Host-side threshold schedule in double / complex<double>: restates get_threshold_decay
// (functions/POCS.py:286-362 of the reference) for the FFT transform from per-slice statistics.
#pragma once
#include <cmath>
#include <complex>
#include <vector>

#include "../../include/p3d_b200.h"

namespace p3d {

typedef std::complex<double> cd;

struct ScheduleStats {
    cd z;                 // lexicographic maximum of X0 (largest real part, ties: largest imaginary part)
    double sumsq;         // sum |X0|^2
    double vmax, vmin;    // max / min |X0|
};

// linear / exponential[-q] / inverse-proportional[-q]; data-driven takes its tau from order statistics
inline void host_schedule(const p3d_pocs_params& pr, const ScheduleStats& st, long long size, std::vector<cd>& tau, bool& is_real) {
    const int niter = pr.niter;
    tau.assign(niter, cd(0, 0));
    is_real = false;
    const double dn = (double)(niter - 1);
    if (pr.thresh_model == P3D_MODEL_INVERSE_PROPORTIONAL) {
        is_real = true;
        const double q = pr.q;
        const double nq = std::pow((double)niter, q);
        const double a = (nq * (st.vmax - st.vmin)) / (nq - 1), b = (nq * st.vmin - st.vmax) / (nq - 1);
        for (int k = 0; k < niter; ++k) tau[k] = cd(a / std::pow((double)(k + 1), q) + b, 0.0);
        return;
    }
    cd z = pr.absmax_threshold ? cd(st.vmax, 0.0) : st.z;
    cd tmax, tmin;
    bool real_sched = false;
    if (pr.decay_factors) { tmax = cd(pr.p_max, 0); tmin = cd(pr.p_min, 0); real_sched = true; }
    else {
        tmax = pr.p_max * z;
        tmin = pr.p_min_adaptive ? cd(0.01 * std::sqrt(st.sumsq / (double)size), 0.0) : pr.p_min * z;
        real_sched = pr.absmax_threshold != 0;
    }
    is_real = real_sched;
    for (int k = 0; k < niter; ++k) {
        const double mu = (double)k / dn;        // niter == 1 -> 0/0 = nan, as in the reference
        if (pr.thresh_model == P3D_MODEL_LINEAR) {
            tau[k] = tmax - (tmax - tmin) * mu;
        } else if (pr.thresh_model == P3D_MODEL_EXPONENTIAL) {
            if (real_sched) {
                const double c = std::log(tmin.real() / tmax.real());
                tau[k] = cd(tmax.real() * std::exp(c * std::pow(mu, pr.q)), 0.0);
            } else {
                const cd c = std::log(tmin / tmax);
                tau[k] = tmax * std::exp(c * std::pow(mu, pr.q));
            }
        }
    }
}

inline void schedule_bounds(const p3d_pocs_params& pr, const ScheduleStats& st, long long size, cd& tmin, cd& tmax) {
    cd z = pr.absmax_threshold ? cd(st.vmax, 0.0) : st.z;
    tmax = pr.decay_factors ? cd(pr.p_max, 0) : pr.p_max * z;
    tmin = pr.decay_factors ? cd(pr.p_min, 0)
                            : (pr.p_min_adaptive ? cd(0.01 * std::sqrt(st.sumsq / (double)size), 0.0) : pr.p_min * z);
}

inline void apply_sqrt_decay(std::vector<cd>& tau, bool is_real) {
    for (auto& t : tau) {
        if (is_real) t = cd(std::sqrt(t.real()), 0.0);   // nan for negative values, like numpy
        else t = std::sqrt(t);
    }
}

}  // namespace p3d
