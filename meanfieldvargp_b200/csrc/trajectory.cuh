// Lorenz 96 sample path on the device: what
// Lorenz96.make_trajectory does (src/dynamical_systems/lorenz_96.py:103-185) --
//   burn-in : 125 D explicit Euler steps of size 1e-3 from x = theta (+0.01 at D/2) (:128-139)
//   path    : x[t] = x[t-1] + f(x[t-1]) dt + ek[:, t],  ek = sqrt(sigma dt) N(0,1)   (:153-178)
//   drift   : f_i = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + theta (circular)            (_l96, :9-30)
// Lorenz 96 is chaotic (the burn-in alone spans ~100 Lyapunov times), so "the same
// trajectory as the reference" means the same BITS: every operation is a separately
// rounded IEEE multiply / add in numpy's order (no FMA contraction), and the noise is an
// input (the reference draws it all up front, :159-162).
//
// Time is sequential, so one CTA owns the whole state in shared memory (double buffered,
// one barrier per step); threads are strided over the dimensions.  Latency bound by
// construction: ~0.2 us per step.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mfvgp {

struct TrajArgs {
    const double *ek;      // [dim_t][D] noise increments, time major (row 0 unused)
    double *x_out;         // [dim_t][D] sample path
    double *scratch;       // [2][D] state buffers when D does not fit shared memory
    int32_t *status;       // bit 0: NaN after burn-in (:142-145), bit 1: NaN on the path (:170-174)
    double theta, dt, delta_t;
    int D, dim_t, nburn, use_smem;
};

// one Euler step of _l96: (fwd_1x - bwd_2x) * bwd_1x - x + u, then x + f * h   (each op rounded)
__device__ __forceinline__ double l96_euler(const double *x, int i, int D, double u, double h) {
    const int ip = i + 1 == D ? 0 : i + 1, im1 = i == 0 ? D - 1 : i - 1,
              im2 = i >= 2 ? i - 2 : i - 2 + D;
    const double xi = x[i];
    const double f = __dadd_rn(__dsub_rn(__dmul_rn(__dsub_rn(x[ip], x[im2]), x[im1]), xi), u);
    return __dadd_rn(xi, __dmul_rn(f, h));
}

__global__ void __launch_bounds__(1024)
l96_trajectory_kernel(const TrajArgs a) {
    extern __shared__ double traj_smem[];
    const int D = a.D, tid = threadIdx.x, nt = blockDim.x;
    double *xa = a.use_smem ? traj_smem : a.scratch;
    double *xb = xa + D;
    __shared__ int sh_nan;
    if (tid == 0) sh_nan = 0;
    for (int i = tid; i < D; i += nt) xa[i] = a.theta + (i == D / 2 ? 0.01 : 0.0);   // :125-131
    __syncthreads();
    for (int t = 0; t < a.nburn; ++t) {                                              // :136-139
        for (int i = tid; i < D; i += nt) xb[i] = l96_euler(xa, i, D, a.theta, a.delta_t);
        __syncthreads();
        double *tmp = xa; xa = xb; xb = tmp;
    }
    int bad = 0;
    for (int i = tid; i < D; i += nt) {
        const double v = xa[i];
        bad |= isnan(v);
        a.x_out[i] = v;                                                              // :151
    }
    if (bad) atomicOr(&sh_nan, 1);
    __syncthreads();
    for (int t = 1; t < a.dim_t; ++t) {                                              // :165-178
        const double *ek = a.ek + (int64_t)t * D;
        double *out = a.x_out + (int64_t)t * D;
        for (int i = tid; i < D; i += nt) {
            const double v = __dadd_rn(l96_euler(xa, i, D, a.theta, a.dt), ek[i]);
            xb[i] = v;
            out[i] = v;
            if (isnan(v)) bad |= 2;
        }
        __syncthreads();
        double *tmp = xa; xa = xb; xb = tmp;
    }
    if (bad & 2) atomicOr(&sh_nan, 2);
    __syncthreads();
    if (tid == 0) a.status[0] = sh_nan;
}

// ---- DW / OU / L63 sample paths ---------------------------------------------------------------
//   DoubleWell.make_trajectory         double_well.py:40-91
//       x[t] = x[t-1] + 4.0 * x[t-1] * (theta - x[t-1] ** 2) * dt + ek[t]          (:80-81)
//   OrnsteinUhlenbeck.make_trajectory  ornstein_uhlenbeck.py:38-79
//       x[t] = x[t-1] + theta * (mu - x[t-1]) * dt + ek[t]                         (:68-69)
//   Lorenz63.make_trajectory           lorenz_63.py:102-181, _l63 :9-35
//       burn-in of 5000 Euler steps (1e-3) from (1,1,1), then
//       x[t] = x[t-1] + _l63(x[t-1]) * dt + ek.T[t]                                (:131-172)
// One to three dimensions and a sequential recurrence: a path is one thread's work.  What the
// device adds is the ENSEMBLE: npaths independent paths (own start, own noise), one thread each,
// path-major buffers.  Same rule as above: every operation separately rounded in the
// reference's order (Lorenz 63 is chaotic too), `** 2` as one rounded product.
template <int SYS> struct SmallDrift;
template <> struct SmallDrift<0> {      // DW, theta[0] = theta
    static constexpr int DIM = 1;
    static __device__ __forceinline__ void f(const double *x, const double *th, double *d) {
        d[0] = __dmul_rn(__dmul_rn(4.0, x[0]), __dsub_rn(th[0], __dmul_rn(x[0], x[0])));
    }
};
template <> struct SmallDrift<1> {      // OU, theta = (theta, mu)
    static constexpr int DIM = 1;
    static __device__ __forceinline__ void f(const double *x, const double *th, double *d) {
        d[0] = __dmul_rn(th[0], __dsub_rn(th[1], x[0]));
    }
};
template <> struct SmallDrift<2> {      // L63, theta = (sigma, rho, beta)
    static constexpr int DIM = 3;
    static __device__ __forceinline__ void f(const double *x, const double *th, double *d) {
        d[0] = __dmul_rn(th[0], __dsub_rn(x[1], x[0]));
        d[1] = __dsub_rn(__dmul_rn(__dsub_rn(th[1], x[2]), x[0]), x[1]);
        d[2] = __dsub_rn(__dmul_rn(x[0], x[1]), __dmul_rn(th[2], x[2]));
    }
};

struct SmallTrajArgs {
    const double *x0;      // [npaths][DIM] state before the burn-in
    const double *ek;      // [npaths][dim_t][DIM] noise increments (row 0 unused)
    double *x_out;         // [npaths][dim_t][DIM]
    int32_t *status;       // [npaths] bit 0: NaN after the burn-in, bit 1: NaN drift on the path
    double theta[3];
    double dt, delta_t;
    int npaths, dim_t, nburn;
};

template <int SYS>
__global__ void __launch_bounds__(128)
small_trajectory_kernel(const SmallTrajArgs a) {
    constexpr int DIM = SmallDrift<SYS>::DIM;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npaths) return;
    double x[DIM], d[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) x[k] = a.x0[(int64_t)p * DIM + k];
    for (int t = 0; t < a.nburn; ++t) {                                   // lorenz_63.py:131-133
        SmallDrift<SYS>::f(x, a.theta, d);
#pragma unroll
        for (int k = 0; k < DIM; ++k) x[k] = __dadd_rn(x[k], __dmul_rn(d[k], a.delta_t));
    }
    int bad = 0;
    const double *ek = a.ek + (int64_t)p * a.dim_t * DIM;
    double *out = a.x_out + (int64_t)p * a.dim_t * DIM;
#pragma unroll
    for (int k = 0; k < DIM; ++k) { bad |= isnan(x[k]); out[k] = x[k]; }
    for (int t = 1; t < a.dim_t; ++t) {
        SmallDrift<SYS>::f(x, a.theta, d);
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            if (isnan(d[k])) bad |= 2;                                    // lorenz_63.py:164-168
            x[k] = __dadd_rn(__dadd_rn(x[k], __dmul_rn(d[k], a.dt)), ek[(int64_t)t * DIM + k]);
            out[(int64_t)t * DIM + k] = x[k];
        }
    }
    a.status[p] = bad;
}

}  // namespace mfvgp
