// Posterior reconstruction: the marginal mean m(t) and variance s(t) on a fine time
// grid from the optimised support points -- the step that follows find_minimum in every
// notebook of the reference (examples/example_500D.ipynb cell 29, :537-575; the local
// polynomials are src/numerical/symbolics.py:102-149, LagrangePolynomial :9-100).
//
// Tx = [t0, obs_times..., tf] has M+1 intervals; interval n interpolates the mean points
// 3n..3n+3 (cubic Lagrange on ti + k dt/3) and the variance points 2n..2n+2 (quadratic on
// ti + k dt/2), sampled at `num` points of linspace(ti, tj, num, endpoint=False); the last
// sample is tf itself, evaluated with the last interval's polynomial.  Outputs in the
// notebook's final layout: tt [P], mt [D][P], st [D][P], P = (M+1) num + 1.
//
// One thread per (dimension, sample): 7 broadcast loads, ~40 FLOP, 16 B written ->
// bound by the HBM write stream.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mfvgp {

struct ReconArgs {
    const double *Tx;        // [M+2]
    const double *mean;      // [D][ldm]
    const double *vars;      // [D][lds]  variances
    int64_t ldm, lds;
    double *tt, *mt, *st;    // [P], [D][P], [D][P]
    int64_t P;
    int D, M, num;
};

// exp() of the log-variance block once per support point (not once per sample)
__global__ void __launch_bounds__(256)
exp_rows_kernel(const double *src, int64_t lds, int ncols, int D, double *dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * ncols) return;
    const int64_t j = i / ncols, c = i - j * ncols;
    dst[i] = exp(src[j * lds + c]);
}

constexpr int kReconRows = 8;      // state dimensions per staged group
constexpr int kReconGroups = 8;    // groups a CTA sweeps: 64 rows per CTA
constexpr int kReconMaxNI = 16;    // intervals a CTA's 256 samples may span on the staged path

__device__ __forceinline__ void recon_cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// A thread owns one sample: it evaluates the seven Lagrange basis values once and sweeps
// kReconGroups x kReconRows state dimensions -- 7 FMAs and 16 B written per (dimension, sample):
// the kernel sits on the HBM write stream.  The support points of the next group of rows arrive
// in shared memory (cp.async, double buffered) while the current group is written, so a CTA
// lives for 64 rows instead of 8 (ncu on the 8-row version: no pipe above 21 %, DRAM at 40 %:
// short-lived CTAs spent their time in the load -> barrier -> store -> drain sequence).
__global__ void __launch_bounds__(256, 4)
reconstruct_kernel(const ReconArgs a, int groups /* <= kReconGroups: row groups per CTA */) {
    __shared__ double sh_m[2][kReconRows][3 * kReconMaxNI + 1];
    __shared__ double sh_s[2][kReconRows][2 * kReconMaxNI + 1];
    const int64_t p0 = (int64_t)blockIdx.x * blockDim.x, p = p0 + threadIdx.x;
    const int jbase = blockIdx.y * (kReconRows * groups);
    // P < 2^32 is checked by the host entry: 32-bit division
    const unsigned num = (unsigned)a.num;
    const int64_t p_last = min(p0 + (int64_t)blockDim.x, a.P) - 1;
    const int n_lo = min(a.M, (int)((unsigned)p0 / num)), n_hi = min(a.M, (int)((unsigned)p_last / num));
    const int NI = n_hi - n_lo + 1;
    const bool staged = NI <= kReconMaxNI;          // CTA-uniform
    const int wm = 3 * NI + 1, ws = 2 * NI + 1;
    const int ngroups = min(groups, (a.D - jbase + kReconRows - 1) / kReconRows);

    auto stage = [&](int g, int buf) {               // support points of group g -> buffer buf
        const int j0 = jbase + g * kReconRows, rows = min(kReconRows, a.D - j0);
        for (int i = threadIdx.x; i < rows * wm; i += blockDim.x) {
            const int r = i / wm, c = i - r * wm;
            recon_cp_async8(&sh_m[buf][r][c], a.mean + (int64_t)(j0 + r) * a.ldm + 3 * (int64_t)n_lo + c);
        }
        for (int i = threadIdx.x; i < rows * ws; i += blockDim.x) {
            const int r = i / ws, c = i - r * ws;
            recon_cp_async8(&sh_s[buf][r][c], a.vars + (int64_t)(j0 + r) * a.lds + 2 * (int64_t)n_lo + c);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (staged) stage(0, 0);

    const bool valid = p < a.P;
    const bool last = p == a.P - 1;
    const unsigned pu = (unsigned)(valid ? p : 0), nu = pu / num;
    const int n = last ? a.M : min(a.M, (int)nu);
    const int k = last ? 0 : (int)(pu - nu * num);
    const double ti = a.Tx[n], tj = a.Tx[n + 1];
    // np.linspace(ti, tj, num, endpoint=False)[k] = ti + k * ((tj - ti) / num); tf for the last
    const double t = last ? tj : ti + k * ((tj - ti) / a.num);
    const double delta = fabs(tj - ti);
    // Lagrange bases of symbolics.py:64-90 on the knots ti + k h (h = dt/3) and ti + k c
    // (c = dt/2), written in the knot-spacing units u = (t - ti)/h, w = (t - ti)/c: the knot
    // differences become the constants 1, 2, 3.
    const double d = (t - ti) / delta;                  // in [0, 1]
    const double u = 3.0 * d, w = d + d;
    const double u1 = u - 1.0, u2 = u - 2.0, u3 = u - 3.0, w1 = w - 1.0, w2 = w - 2.0;
    const double p01 = u * u1, p23 = u2 * u3;
    const double B0 = -(1.0 / 6.0) * (u1 * p23), B1 = 0.5 * (u * p23), B2 = -0.5 * (p01 * u3),
                 B3 = (1.0 / 6.0) * (p01 * u2);
    const double C0 = 0.5 * (w1 * w2), C1 = -(w * w2), C2 = 0.5 * (w * w1);
    if (blockIdx.y == 0 && valid) a.tt[p] = t;
    const int cm = 3 * (n - n_lo), cs = 2 * (n - n_lo);

    for (int g = 0; g < ngroups; ++g) {
        const int buf = g & 1;
        const int j0 = jbase + g * kReconRows, rows = min(kReconRows, a.D - j0);
        if (staged) {
            if (g + 1 < ngroups) {
                stage(g + 1, buf ^ 1);               // in flight while group g is written
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
        }
        if (valid) {
#pragma unroll 4
            for (int r = 0; r < rows; ++r) {
                const int j = j0 + r;
                const double *mp = staged ? &sh_m[buf][r][cm] : a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
                const double *sp = staged ? &sh_s[buf][r][cs] : a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
                const double mt = fma(mp[0], B0, fma(mp[1], B1, fma(mp[2], B2, mp[3] * B3)));
                const double st = fma(sp[0], C0, fma(sp[1], C1, sp[2] * C2));
                __stcs(a.mt + (int64_t)j * a.P + p, mt);          // written once, never re-read here
                __stcs(a.st + (int64_t)j * a.P + p, st);
            }
        }
        if (staged) __syncthreads();                 // buffer buf is refilled two groups on
    }
}

// Initial state x0 for find_minimum from an initial sample path on the device (SURVEY.md 8(f)
// rank 2; examples/example_500D.ipynb cells 17 and 19, :402-428): the means are the path at the
// mean support indices, the variances var_base + var_spread * u at the variance support
// indices, stored as logs; x = [means [D][Nm] | log-variances [D][Ns]] (free_energy.py:678-684).
struct InitStateArgs {
    const double *path;       // [D][T] initial path x0(t), row = state dimension
    const int64_t *mp_idk;    // [Nm] time indices of the mean support points
    const double *u;          // [D][Ns] uniform draws in [0, 1)
    double *x;                // [D*Nm + D*Ns]
    double var_base, var_spread;
    int64_t T;
    int D, Nm, Ns;
};

__global__ void __launch_bounds__(256)
initial_state_kernel(const InitStateArgs a) {
    const int64_t nm = (int64_t)a.D * a.Nm, n = nm + (int64_t)a.D * a.Ns;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < nm) {
            const int64_t j = i / a.Nm, c = i - j * a.Nm;
            a.x[i] = a.path[j * a.T + a.mp_idk[c]];
        } else {
            a.x[i] = log(fma(a.var_spread, a.u[i - nm], a.var_base));
        }
    }
}

}  // namespace mfvgp
