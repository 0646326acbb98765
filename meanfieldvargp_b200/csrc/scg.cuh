// Device-resident vector kernels for the scaled conjugate gradient loop
// (reference: SCG.minimize, src/numerical/scaled_cg.py:107-386).  Iterates,
// gradients and search direction never leave the device; every dot product is
// a fixed-order two-stage reduction so the accept/reject branches
// (scaled_cg.py:194, :233, :254) are reproducible run to run.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mfvgp {

constexpr int kRedBlock = 256;
constexpr int kRedSlots = 4;

// programmatic dependent launch (see l96_coop.cuh)
__device__ __forceinline__ void scg_pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Single-GPU runs fold the per-block partials in the LAST block of the producing kernel
// (atomic ticket, fixed order => deterministic) and publish the scalars the host control
// loop branches on straight into host-mapped memory: [0..3] this kernel's sums,
// [4..7] F, E0, Esde, Eobs of the last E_cost, [8] its status bits, then the epoch flag.
// The host spins on the flag: no memcpy, no stream synchronisation (scaled_cg.py's
// branches at :194, :233, :254 stay on the host, the vectors never leave the device).
struct FoldArgs {
    unsigned *ticket;          // nullptr: partials only, scg_fold_kernel follows (multi-GPU)
    double *scal;              // [8] device scalars: [0..3] sums, [4..7] E_cost outputs
    const int32_t *status;     // status word of the last E_cost
    double *host_rec;          // [16] host-mapped record
    volatile long long *host_flag;
    long long epoch;
    // Speculative chain of one successful iteration (mfvgp_scg): the scalars between the
    // kernels stay on the device.  mode 0: publish (above).  mode 1 (direction): mu, kappa,
    // sigma = 1e-3/sqrt(kappa) (scaled_cg.py:203-212) -> ctrl.  mode 2 (theta): theta, delta,
    // beta, alpha (:228-239) -> ctrl.  mode 3 (stats): publish sums + ctrl.  ctrl [16]:
    // 0 sigma, 1 alpha, 2 theta, 3 delta, 4 beta, 5 mu, 6 kappa, 7 flags (1: mu >= 0,
    // 2: kappa < eps), 8 status of the sigma-step evaluation.
    int mode;
    double *ctrl;
    double beta_in;
};

// Index map of the optimisation vector a rank works on.  Single GPU: the whole
// x.  Sharded: the rank's active columns of the mean block [D][Nm] and of the
// log-variance block [D][Ns]; the last active column may be a ghost column
// shared with the next rank -- it is updated here too (both ranks hold the same
// bits) but only its owner counts it in reductions.
struct VecMap {
    int64_t n;        // number of local entries
    int64_t num_mp;   // D*Nm
    int full;         // 1: identity map
    int D, Nm, Ns, cm0, cs0, wm, ws, own_wm, own_ws;
    __device__ __forceinline__ int64_t at(int64_t idx, bool &owned) const {
        if (full) { owned = true; return idx; }
        const int64_t nm = (int64_t)D * wm;
        if (idx < nm) {
            const int64_t j = idx / wm;
            const int c = (int)(idx - j * wm);
            owned = c < own_wm;
            return j * Nm + cm0 + c;
        }
        const int64_t r = idx - nm;
        const int64_t j = r / ws;
        const int c = (int)(r - j * ws);
        owned = c < own_ws;
        return num_mp + j * Ns + cs0 + c;
    }
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Block-level combine of kRedSlots values (slots 0..2 summed, slot 3 max'ed)
// and store to part[blockIdx.x][*].
__device__ __forceinline__ void fold_partials(const double *part, int nblk, double (&r)[kRedSlots],
                                              double (*sh)[kRedBlock / 32]) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblk; b += kRedBlock) {
        v[0] += __ldcg(part + (int64_t)b * kRedSlots + 0);
        v[1] += __ldcg(part + (int64_t)b * kRedSlots + 1);
        v[2] += __ldcg(part + (int64_t)b * kRedSlots + 2);
        v[3] = fmax(v[3], __ldcg(part + (int64_t)b * kRedSlots + 3));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = warp_sum_d(v[i]);
    v[3] = warp_max_d(v[3]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) sh[i][w] = v[i];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) r[i] = sh[i][0];
        for (int k = 1; k < kRedBlock / 32; ++k) {
            r[0] += sh[0][k]; r[1] += sh[1][k]; r[2] += sh[2][k];
            r[3] = fmax(r[3], sh[3][k]);
        }
    }
}

__device__ __forceinline__ void red_store(double (&v)[kRedSlots], double *part,
                                          const FoldArgs &fa) {
    __shared__ double sh[kRedSlots][kRedBlock / 32];
    __shared__ int sh_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = warp_sum_d(v[i]);
    v[3] = warp_max_d(v[3]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) sh[i][w] = v[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double r[kRedSlots];
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) r[i] = sh[i][0];
        for (int k = 1; k < kRedBlock / 32; ++k) {
            r[0] += sh[0][k]; r[1] += sh[1][k]; r[2] += sh[2][k];
            r[3] = fmax(r[3], sh[3][k]);
        }
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) part[(int64_t)blockIdx.x * kRedSlots + i] = r[i];
        sh_last = 0;
        if (fa.ticket != nullptr) {
            __threadfence();
            const unsigned k = atomicAdd(fa.ticket, 1u);
            sh_last = (k == gridDim.x - 1u);
            if (sh_last) fa.ticket[0] = 0u;
        }
    }
    if (fa.ticket == nullptr) return;
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    double r[kRedSlots];
    fold_partials(part, (int)gridDim.x, r, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) fa.scal[i] = r[i];
        double *c = fa.ctrl;
        if (fa.mode == 1) {                    // the host's arithmetic, operation for operation
            const double mu = r[0], kappa = r[1];
            c[5] = mu; c[6] = kappa;
            c[7] = (mu >= 0.0 ? 1.0 : 0.0) + (kappa < 2.220446049250313e-16 ? 2.0 : 0.0);
            c[0] = 1.0e-3 / sqrt(kappa);
            return;
        }
        if (fa.mode == 2) {
            const double sigma = c[0], mu = c[5], kappa = c[6];
            const double theta = r[0] / sigma;
            double beta = fa.beta_in;
            double delta = __dadd_rn(theta, __dmul_rn(beta, kappa));
            if (delta <= 0.0) {
                delta = __dmul_rn(beta, kappa);
                beta = __dsub_rn(beta, theta / kappa);
            }
            c[1] = -(mu / delta); c[2] = theta; c[3] = delta; c[4] = beta;
            c[8] = (double)__ldcg(fa.status);
            return;
        }
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) fa.host_rec[i] = r[i];
#pragma unroll
        for (int i = 4; i < 8; ++i) fa.host_rec[i] = __ldcg(fa.scal + i);
        fa.host_rec[8] = (double)__ldcg(fa.status);
        if (fa.mode == 3) {
            fa.host_rec[9] = c[5]; fa.host_rec[10] = c[6]; fa.host_rec[11] = c[8];
            fa.host_rec[12] = c[2]; fa.host_rec[13] = c[1]; fa.host_rec[14] = c[4];
            fa.host_rec[15] = c[7];
        }
        __threadfence_system();
        *fa.host_flag = fa.epoch;
    }
}

// d = gamma*d - g (gamma = 0 and d arbitrary gives the restart d = -g);
// partials: [0] mu = d.g, [1] kappa = d.d     (scaled_cg.py:192-200, 361-369)
__global__ void __launch_bounds__(kRedBlock)
scg_direction_kernel(double *d, const double *g, double gamma, int restart, const VecMap vm,
                     double *part, const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    scg_pdl_enter();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < vm.n;
         t += (int64_t)gridDim.x * blockDim.x) {
        bool own;
        const int64_t i = vm.at(t, own);
        const double gi = g[i];
        const double di = restart ? -gi : gamma * d[i] - gi;
        d[i] = di;
        if (own) {
            v[0] = fma(di, gi, v[0]);
            v[1] = fma(di, di, v[1]);
        }
    }
    red_store(v, part, fa);
}

// y = x + a*d                                   (scaled_cg.py:222, :242)
__global__ void __launch_bounds__(kRedBlock)
scg_axpy_kernel(double *y, const double *x, const double *d, double a, const double *a_dev,
                const VecMap vm) {
    scg_pdl_enter();
    if (a_dev != nullptr) a = __ldcg(a_dev);      // step length computed on the device
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < vm.n;
         t += (int64_t)gridDim.x * blockDim.x) {
        bool own;
        const int64_t i = vm.at(t, own);
        y[i] = x[i] + a * d[i];
    }
}

// partial [0] = d.(gp - g)                      (scaled_cg.py:228)
__global__ void __launch_bounds__(kRedBlock)
scg_theta_kernel(const double *d, const double *gp, const double *g, const VecMap vm,
                 double *part, const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    scg_pdl_enter();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < vm.n;
         t += (int64_t)gridDim.x * blockDim.x) {
        bool own;
        const int64_t i = vm.at(t, own);
        if (own) v[0] = fma(d[i], gp[i] - g[i], v[0]);
    }
    red_store(v, part, fa);
}

// After the trial evaluation at x + alpha d:
// [0] sum |g_now| (:281), [1] g_now.g_now (:333), [2] g_now.(g_new - g_now)
// (Polak-Ribiere numerator once g_now becomes grad_new and g_new grad_old,
// :367), [3] max |d| (:306).
__global__ void __launch_bounds__(kRedBlock)
scg_stats_kernel(const double *g_now, const double *g_new, const double *d, const VecMap vm,
                 double *part, const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    scg_pdl_enter();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < vm.n;
         t += (int64_t)gridDim.x * blockDim.x) {
        bool own;
        const int64_t i = vm.at(t, own);
        if (!own) continue;
        const double a = g_now[i];
        v[0] += fabs(a);
        v[1] = fma(a, a, v[1]);
        v[2] = fma(a, g_new[i] - a, v[2]);
        v[3] = fmax(v[3], fabs(d[i]));
    }
    red_store(v, part, fa);
}

// Second stage: one block folds the per-block partials in a fixed order.
__global__ void __launch_bounds__(kRedBlock)
scg_fold_kernel(const double *part, int nblk, double *out) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        v[0] += part[(int64_t)b * kRedSlots + 0];
        v[1] += part[(int64_t)b * kRedSlots + 1];
        v[2] += part[(int64_t)b * kRedSlots + 2];
        v[3] = fmax(v[3], part[(int64_t)b * kRedSlots + 3]);
    }
    __shared__ double sh[kRedSlots][kRedBlock / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = warp_sum_d(v[i]);
    v[3] = warp_max_d(v[3]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) sh[i][w] = v[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double r[kRedSlots];
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) r[i] = sh[i][0];
        for (int k = 1; k < kRedBlock / 32; ++k) {
            r[0] += sh[0][k]; r[1] += sh[1][k]; r[2] += sh[2][k];
            r[3] = fmax(r[3], sh[3][k]);
        }
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) out[i] = r[i];
    }
}

}  // namespace mfvgp
