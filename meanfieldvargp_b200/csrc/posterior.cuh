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
    const double *vars;      // [D][lds]  variances, or log-variances if log_vars
    int64_t ldm, lds;
    double *tt, *mt, *st;    // [P], [D][P], [D][P]
    int64_t P;
    int D, M, num, log_vars;
};

__global__ void __launch_bounds__(256)
reconstruct_kernel(const ReconArgs a) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (p >= a.P) return;
    const bool last = p == a.P - 1;
    const int n = last ? a.M : (int)(p / a.num);
    const int k = last ? 0 : (int)(p - (int64_t)n * a.num);
    const double ti = a.Tx[n], tj = a.Tx[n + 1];
    // np.linspace(ti, tj, num, endpoint=False)[k] = ti + k * ((tj - ti) / num); tf for the last
    const double t = last ? tj : ti + k * ((tj - ti) / a.num);
    const double delta = fabs(tj - ti), h = delta / 3.0, c = delta / 2.0;
    const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
    const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
    const double m0 = mp[0], m1 = mp[1], m2 = mp[2], m3 = mp[3];
    double s0 = sp[0], s1 = sp[1], s2 = sp[2];
    if (a.log_vars) { s0 = exp(s0); s1 = exp(s1); s2 = exp(s2); }
    // Lagrange bases on the absolute knots, as symbolics.py:64-90 builds them
    const double h0 = ti, h1 = ti + h, h2 = ti + 2.0 * h, h3 = ti + 3.0 * h;
    const double d0 = t - h0, d1 = t - h1, d2 = t - h2, d3 = t - h3;
    const double mt = m0 * (d1 / (h0 - h1)) * (d2 / (h0 - h2)) * (d3 / (h0 - h3))
                    + m1 * (d0 / (h1 - h0)) * (d2 / (h1 - h2)) * (d3 / (h1 - h3))
                    + m2 * (d0 / (h2 - h0)) * (d1 / (h2 - h1)) * (d3 / (h2 - h3))
                    + m3 * (d0 / (h3 - h0)) * (d1 / (h3 - h1)) * (d2 / (h3 - h2));
    const double c0 = ti, c1 = ti + c, c2 = ti + 2.0 * c;
    const double e0 = t - c0, e1 = t - c1, e2 = t - c2;
    const double st = s0 * (e1 / (c0 - c1)) * (e2 / (c0 - c2))
                    + s1 * (e0 / (c1 - c0)) * (e2 / (c1 - c2))
                    + s2 * (e0 / (c2 - c0)) * (e1 / (c2 - c1));
    a.mt[(int64_t)j * a.P + p] = mt;
    a.st[(int64_t)j * a.P + p] = st;
    if (j == 0) a.tt[p] = t;
}

}  // namespace mfvgp
