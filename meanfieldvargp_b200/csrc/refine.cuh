// Adaptive Gauss-Kronrod refinement on the device.
//
// The reference integrates with scipy.integrate.quad_vec(limit=100,
// epsabs=1e-6, epsrel=1e-6) (free_energy.py:405-417).  quad_vec bisects at
// least once; when its error estimate after that bisection is below tol/8 the
// answer is the fixed 42-node rule of the main kernels.  Otherwise it keeps
// bisecting the panel with the largest error.  guard_reduce_kernel flags the
// (interval, integrand) pairs for which convergence after the first bisection
// cannot be guaranteed; this kernel recomputes exactly those, following
// scipy/integrate/_quad_vec.py step by step:
//   _quadrature_gk   : s_k, s_g, |f| and |f - s_k/2| sums, the error model
//                      err = dabs*min(1,(200 err/dabs)^1.5), round-off floor
//   quad_vec         : max-error heap, batch selection (parallel_count=128,
//                      "err_sum > global_error - tol/8"), termination tests
//   _subdivide_interval : midpoint split, dint = s1 + s2 - old_int
// Only the energy integrand and the dS integrand can refine; the dM
// integrand is a polynomial of degree <= 18 for all four drifts, so both the
// Kronrod and the embedded Gauss rule integrate it exactly.
//
// Layout: one CTA per flagged work item, threads strided over the rows i
// (energy index) of the integrand; row i owns NJ*3 dS elements
// (NJ = 4 neighbour slots for L96, D for the dense systems).
#pragma once

#include "kernels.cuh"

namespace mfvgp {

constexpr int kRefThreads = 128;
constexpr int kHeapCap = 128;
constexpr int kLimit = 100;              // quad_vec(limit=100)
constexpr int kMaxNE = 12;               // elements per row (L96: 4 slots x 3)

__constant__ double c_gk_x[kNodesPerPanel];
__constant__ double c_gk_v[kNodesPerPanel];
__constant__ double c_gk_w[10];

struct RefineArgs {
    const double *mean, *vars;
    int64_t ldm, lds;
    const double *obs_times, *sigma, *theta;
    const int32_t *flags;      // [L] bit0: energy, bit1: dS
    const int32_t *nflagged;   // [1] number of flagged intervals (nullptr: scan the flags)
    double *esde_n;            // [L]   overwritten for refined energy integrals
    double *ds_out;            // [L][D][3] overwritten for refined dS integrals (or nullptr)
    double *ws;                // [gridDim.x][2][D*kMaxNE] scratch: elementwise integral, and
                               // (delta mode) the fixed-rule value s1+s2 of the first bisection
    int delta_mode;            // 1: ds_out receives (refined - fixed rule), for the walk path (gradient already assembled)
    int32_t *info;             // [L][2] number of bisections done (0 = not refined)
    int32_t *status;           // OR-ed MFVGP_ST_REFINE_LIMIT
    int D, n_begin, n_end, log_vars, want_grad;
    int group;                 // CTAs per item in refine_kernel (GroupCtx)
    double *gpart;             // [groups][2][group][8]
    unsigned *gbar;            // [groups] arrival counters, zero at kernel start
};

__device__ __forceinline__ void basis3(double tau, double *B, double *dB) {
    const double a = tau, b = tau - 1.0 / 3.0, c = tau - 2.0 / 3.0, d = tau - 1.0;
    B[0] = -4.5 * b * c * d;   dB[0] = -4.5 * (c * d + b * d + b * c);
    B[1] = 13.5 * a * c * d;   dB[1] = 13.5 * (c * d + a * d + a * c);
    B[2] = -13.5 * a * b * d;  dB[2] = -13.5 * (b * d + a * d + a * b);
    B[3] = 4.5 * a * b * c;    dB[3] = 4.5 * (b * c + a * c + a * b);
}
__device__ __forceinline__ void basis2(double tau, double *B, double *dB) {
    const double a = tau, b = tau - 0.5, c = tau - 1.0;
    B[0] = 2.0 * b * c;   dB[0] = 2.0 * (b + c);
    B[1] = -4.0 * a * c;  dB[1] = -4.0 * (a + c);
    B[2] = 2.0 * a * b;   dB[2] = 2.0 * (a + b);
}

// Row evaluator: energy E_i and the dS elements of row i at local time tau.
template <int SYS> struct RowEval;

template <> struct RowEval<3> {                     // Lorenz 96
    static constexpr int NJ = 4;
    double P[4][4], S[4][3], Sig, theta, inv_dt;
    __device__ void load(const RefineArgs &a, int n, int i, double dt) {
        const int D = a.D;
        const int idx[4] = {i, (i + 1) % D, (i + D - 1) % D, (i + D - 2) % D};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double *mp = a.mean + (int64_t)idx[s] * a.ldm + 3 * (int64_t)n;
            const double *sp = a.vars + (int64_t)idx[s] * a.lds + 2 * (int64_t)n;
#pragma unroll
            for (int k = 0; k < 4; ++k) P[s][k] = mp[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) S[s][k] = a.log_vars ? exp(sp[k]) : sp[k];
        }
        Sig = a.sigma[i];
        theta = a.theta[0];
        inv_dt = 1.0 / dt;
    }
    // which output dimension a slot of row i contributes to
    __device__ static int target(int D, int i, int slot) {
        return slot == 0 ? i : (slot == 1 ? (i + 1) % D : (slot == 2 ? (i + D - 1) % D
                                                                     : (i + D - 2) % D));
    }
    __device__ void eval(double tau, double &E, double *el /*[12]*/) const {
        double B3[4], dB3[4], B2[3], dB2[3];
        basis3(tau, B3, dB3);
        basis2(tau, B2, dB2);
        double m[4], s[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            m[q] = P[q][0] * B3[0] + P[q][1] * B3[1] + P[q][2] * B3[2] + P[q][3] * B3[3];
            s[q] = S[q][0] * B2[0] + S[q][1] * B2[1] + S[q][2] * B2[2];
        }
        const double dm = (P[0][0] * dB3[0] + P[0][1] * dB3[1] + P[0][2] * dB3[2]
                           + P[0][3] * dB3[3]) * inv_dt;
        const double ds = (S[0][0] * dB2[0] + S[0][1] * dB2[1] + S[0][2] * dB2[2]) * inv_dt;
        const double u = m[1] - m[3], su = s[1] + s[3], mb = m[2], sb = s[2];
        const double R = fma(u, mb, theta - m[0]) - dm;
        const Rational r(s[0], ds, Sig);
        E = R * R + s[0] + sb * u * u + mb * mb * su + su * sb + r.rat + r.q;
        const double es[4] = {1.0 + r.rat_s, mb * mb + sb, u * u + su, mb * mb + sb};
        const double eds = (r.rat_ds + 1.0) * inv_dt;
#pragma unroll
        for (int slot = 0; slot < 4; ++slot)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                el[slot * 3 + k] = es[slot] * B2[k] + (slot == 0 ? eds * dB2[k] : 0.0);
    }
};

template <int SYS> struct RowEvalDense {            // DW, OU, L63
    using Dr = Drift<SYS>;
    static constexpr int D = Dr::D;
    static constexpr int NJ = D;
    double P[D][4], S[D][3], Sig[D], th[Dr::NTHETA], inv_dt;
    int row;
    __device__ void load(const RefineArgs &a, int n, int i, double dt) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
            const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
#pragma unroll
            for (int k = 0; k < 4; ++k) P[j][k] = mp[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) S[j][k] = a.log_vars ? exp(sp[k]) : sp[k];
            Sig[j] = a.sigma[j];
        }
#pragma unroll
        for (int k = 0; k < Dr::NTHETA; ++k) th[k] = a.theta[k];
        inv_dt = 1.0 / dt;
        row = i;
    }
    __device__ static int target(int, int, int slot) { return slot; }
    __device__ void eval(double tau, double &E, double *el) const {
        double B3[4], dB3[4], B2[3], dB2[3];
        basis3(tau, B3, dB3);
        basis2(tau, B2, dB2);
        double m[D], dm[D], s[D], ds[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            m[j] = P[j][0] * B3[0] + P[j][1] * B3[1] + P[j][2] * B3[2] + P[j][3] * B3[3];
            dm[j] = (P[j][0] * dB3[0] + P[j][1] * dB3[1] + P[j][2] * dB3[2]
                     + P[j][3] * dB3[3]) * inv_dt;
            s[j] = S[j][0] * B2[0] + S[j][1] * B2[1] + S[j][2] * B2[2];
            ds[j] = (S[j][0] * dB2[0] + S[j][1] * dB2[1] + S[j][2] * dB2[2]) * inv_dt;
        }
        double Ev[D], Em[D][D], Edm[D], Es[D][D], Eds[D];
        Dr::nodal(m, dm, s, ds, Sig, th, Ev, Em, Edm, Es, Eds);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (i != row) continue;
            E = Ev[i];
#pragma unroll
            for (int j = 0; j < D; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    el[j * 3 + k] = Es[i][j] * B2[k]
                                    + (i == j ? Eds[i] * inv_dt * dB2[k] : 0.0);
        }
    }
};
template <> struct RowEval<0> : RowEvalDense<0> {};
template <> struct RowEval<1> : RowEvalDense<1> {};
template <> struct RowEval<2> : RowEvalDense<2> {};

// A flagged (interval, integrand) item may be shared by a GROUP of G co-resident CTAs
// (refine_kernel at large D: one CTA would walk 8192 rows alone for milliseconds).  The rows
// are dealt round-robin over the group's threads; every reduction of the scalar control loop
// goes through global memory -- each CTA's partial, a group barrier, then every CTA folds the
// G partials in rank order -- so all CTAs see identical scalars and run quad_vec's heap logic
// redundantly, in lock step, with no leader.  G = 1 (tiny problems) costs nothing.
struct GroupCtx {
    int G, rank;               // CTAs in the group, this CTA's rank
    double *part;              // [2][G][8] partial sums (double buffered by the barrier count)
    unsigned *bar;             // monotone arrival counter of the group (0 at kernel start)
    unsigned seq;              // barriers this CTA has passed
    int32_t *status;           // OR-ed 32 if a barrier wait gives up
};

__device__ __forceinline__ void group_barrier(GroupCtx &gc) {
    __syncthreads();
    if (gc.G > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(gc.bar, 1u);
            const unsigned target = (gc.seq + 1u) * (unsigned)gc.G;
            long spin = 0;
            while (*((volatile unsigned *)gc.bar) < target) {
                __nanosleep(40);
                if (++spin > (1L << 25)) { atomicOr(gc.status, 32); break; }   // never expected
            }
            __threadfence();
        }
        gc.seq++;
        __syncthreads();
    }
}

// block_sum over the group: result valid in thread 0 of EVERY CTA of the group
template <int NV, int NT>
__device__ __forceinline__ void group_sum(double (&v)[NV], double *sh, GroupCtx &gc) {
    block_sum<NV, NT>(v, sh);
    if (gc.G == 1) return;
    double *slot = gc.part + (int64_t)(gc.seq & 1u) * gc.G * 8;
    if (threadIdx.x == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) slot[gc.rank * 8 + i] = v[i];
    group_barrier(gc);
    if (threadIdx.x < 32) {            // warp 0: lanes over the ranks, then a butterfly (fixed tree)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = 0.0;
            for (int r = threadIdx.x; r < gc.G; r += 32) t += __ldcg(slot + r * 8 + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            v[i] = t;
        }
    }
}

// _quadrature_gk on [pa, pb] (local coordinates) for one row, NE elements, by ONE THREAD
// (large D: a thread per row keeps all 32 lanes on arithmetic; the nodes are walked twice,
// the second time for the |f - s_k/2| sum).
// integ[e] = h*s_k ; acc[0..2] += squared contributions to the three norms.
template <int SYS, int NE, bool ENERGY>
__device__ __forceinline__ void gk21_row_thread(const RowEval<SYS> &ev, double pa, double pb,
                                         double dt, double *integ, double *acc) {
    const double c = 0.5 * (pa + pb), hl = 0.5 * (pb - pa);
    const double h = hl * dt;
    double sk[NE], sg[NE], sabs[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) sk[e] = sg[e] = sabs[e] = 0.0;
    for (int r = 0; r < kNodesPerPanel; ++r) {
        double E, el[kMaxNE];
        ev.eval(c + hl * c_gk_x[r], E, el);
        const double v = c_gk_v[r];
        const double w = (r & 1) ? c_gk_w[(r - 1) >> 1] : 0.0;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const double f = ENERGY ? E : el[e];
            sk[e] += v * f;
            sabs[e] += v * fabs(f);
            sg[e] += w * f;
        }
    }
    double sdabs[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) sdabs[e] = 0.0;
    for (int r = 0; r < kNodesPerPanel; ++r) {
        double E, el[kMaxNE];
        ev.eval(c + hl * c_gk_x[r], E, el);
        const double v = c_gk_v[r];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const double f = ENERGY ? E : el[e];
            sdabs[e] += v * fabs(f - 0.5 * sk[e]);
        }
    }
    const double eps50 = 50.0 * 2.220446049250313e-16;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        integ[e] = h * sk[e];
        const double d1 = (sk[e] - sg[e]) * h, d2 = sdabs[e] * h, d3 = eps50 * h * sabs[e];
        acc[0] += d1 * d1;
        acc[1] += d2 * d2;
        acc[2] += d3 * d3;
    }
}

// _quadrature_gk on [pa, pb] (local coordinates) for one row, NE elements, by ONE WARP
// (small D, where rows are scarce and latency decides): lane r
// evaluates Kronrod node r (lanes 21..31 idle), the node sums are butterflies, and the node values
// stay in registers for the second sum (|f - s_k/2|), so every node is evaluated once.
// integ[e] = h*s_k (all lanes); acc[0..2] += squared contributions to the three norms (lane 0).
__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int SYS, int NE, bool ENERGY>
__device__ __forceinline__ void gk21_row_warp(const RowEval<SYS> &ev, double pa, double pb,
                                              double dt, double *integ, double *acc) {
    const int r = threadIdx.x & 31;
    const bool node = r < kNodesPerPanel;
    const double c = 0.5 * (pa + pb), hl = 0.5 * (pb - pa);
    const double h = hl * dt;
    double f[NE];
    {
        double E, el[kMaxNE];
        ev.eval(c + hl * c_gk_x[node ? r : 0], E, el);
#pragma unroll
        for (int e = 0; e < NE; ++e) f[e] = ENERGY ? E : el[e];
    }
    const double v = node ? c_gk_v[r] : 0.0;
    const double w = (node && (r & 1)) ? c_gk_w[(r - 1) >> 1] : 0.0;
    const double eps50 = 50.0 * 2.220446049250313e-16;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const double fe = node ? f[e] : 0.0;
        const double sk = warp_allsum(v * fe);
        const double sg = warp_allsum(w * fe);
        const double sabs = warp_allsum(v * fabs(fe));
        const double sdabs = warp_allsum(v * fabs(fe - 0.5 * sk));
        integ[e] = h * sk;
        const double d1 = (sk - sg) * h, d2 = sdabs * h, d3 = eps50 * h * sabs;
        if (r == 0) {
            acc[0] += d1 * d1;
            acc[1] += d2 * d2;
            acc[2] += d3 * d3;
        }
    }
}

template <int SYS, int NE, bool ENERGY, bool WARP>
__device__ __forceinline__ void gk21_row(const RowEval<SYS> &ev, double pa, double pb,
                                         double dt, double *integ, double *acc) {
    if (WARP) gk21_row_warp<SYS, NE, ENERGY>(ev, pa, pb, dt, integ, acc);
    else gk21_row_thread<SYS, NE, ENERGY>(ev, pa, pb, dt, integ, acc);
}

// scipy's error model from the three norms of one panel
__device__ __forceinline__ void gk_error(double n_kg2, double n_dabs2, double n_rnd2,
                                         double &err, double &rnd) {
    err = sqrt(n_kg2);
    const double dabs = sqrt(n_dabs2);
    if (dabs != 0.0 && err != 0.0) err = dabs * fmin(1.0, pow(200.0 * err / dabs, 1.5));
    rnd = sqrt(n_rnd2);
    if (rnd > 2.2250738585072014e-308) err = fmax(err, rnd);
}

// WARP: one warp per row (nodes across the lanes) instead of one thread per row
template <int SYS, bool ENERGY, bool WARP, int NT = kRefThreads>
__device__ void refine_item(const RefineArgs &a, int n, double *ws, double *ws_fixed,
                            GroupCtx &gc) {
    using RE = RowEval<SYS>;
    constexpr int NE = ENERGY ? 1 : RE::NJ * 3;
    __shared__ double h_err[kHeapCap], h_a[kHeapCap], h_b[kHeapCap];
    __shared__ double sh_red[8 * (NT / 32)];
    __shared__ int h_cnt, batch_n, stop_flag;
    __shared__ double b_err[kHeapCap], b_a[kHeapCap], b_b[kHeapCap];

    const int D = a.D;
    const int lane = threadIdx.x & 31;
    // rows of this thread (or of its warp) and who records their results
    const int wid0 = WARP ? gc.rank * (NT / 32) + (threadIdx.x >> 5) : gc.rank * NT + threadIdx.x;
    const int wstride = WARP ? gc.G * (NT / 32) : gc.G * NT;
    const bool writer = WARP ? lane == 0 : true;
    const double dt = fabs(a.obs_times[n + 1] - a.obs_times[n]);
    double gerr = 0.0, rnd = 0.0, normG = 0.0;   // meaningful in thread 0
    int nbis = 0;

    // ---- initial panel [0,1] ------------------------------------------------
    {
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = wid0; i < D; i += wstride) {          // one warp per row
            RE ev;
            ev.load(a, n, i, dt);
            double integ[NE];
            gk21_row<SYS, NE, ENERGY, WARP>(ev, 0.0, 1.0, dt, integ, acc);
            if (writer)
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    ws[(int64_t)i * NE + e] = integ[e];
                    acc[3] += integ[e] * integ[e];
                }
        }
        group_sum<7, NT>(acc, sh_red, gc);
        if (threadIdx.x == 0) {
            double e0, r0;
            gk_error(acc[0], acc[1], acc[2], e0, r0);
            gerr = e0; rnd = r0; normG = sqrt(acc[3]);
            h_err[0] = e0; h_a[0] = 0.0; h_b[0] = 1.0;
            h_cnt = 1;
            stop_flag = 0;
        }
        __syncthreads();
    }

    // ---- quad_vec main loop ---------------------------------------------------
    while (true) {
        if (threadIdx.x == 0) {
            batch_n = 0;
            if (h_cnt > 0 && h_cnt < kLimit && !stop_flag) {
                const double tol = fmax(1.0e-6, 1.0e-6 * normG);
                double err_sum = 0.0;
                for (int j = 0; j < 128; ++j) {
                    if (h_cnt == 0) break;
                    if (j > 0 && err_sum > gerr - tol / 8.0) break;
                    // heappop of (-err, a, b): largest err, ties by smaller a, then b
                    int best = 0;
                    for (int k = 1; k < h_cnt; ++k) {
                        if (h_err[k] > h_err[best] ||
                            (h_err[k] == h_err[best] &&
                             (h_a[k] < h_a[best] ||
                              (h_a[k] == h_a[best] && h_b[k] < h_b[best]))))
                            best = k;
                    }
                    b_err[batch_n] = h_err[best]; b_a[batch_n] = h_a[best];
                    b_b[batch_n] = h_b[best];
                    batch_n++;
                    err_sum += h_err[best];
                    h_cnt--;
                    h_err[best] = h_err[h_cnt]; h_a[best] = h_a[h_cnt]; h_b[best] = h_b[h_cnt];
                }
            }
        }
        __syncthreads();
        const int nb = batch_n;
        if (nb == 0) break;
        for (int t = 0; t < nb; ++t) {
            const double pa = b_a[t], pb = b_b[t], pc = 0.5 * (pa + pb);
            double acc[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int i = wid0; i < D; i += wstride) {      // one warp per row
                RE ev;
                ev.load(a, n, i, dt);
                double i_old[NE], i_l[NE], i_r[NE], dummy[3] = {0, 0, 0};
                gk21_row<SYS, NE, ENERGY, WARP>(ev, pa, pc, dt, i_l, acc);
                gk21_row<SYS, NE, ENERGY, WARP>(ev, pc, pb, dt, i_r, acc + 3);
                gk21_row<SYS, NE, ENERGY, WARP>(ev, pa, pb, dt, i_old, dummy);   // interval_cache value
                if (writer)
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        const double dint = i_l[e] + i_r[e] - i_old[e];
                        ws[(int64_t)i * NE + e] += dint;
                        // the root panel's two halves are the fixed 42-node rule
                        if (!ENERGY && pa == 0.0 && pb == 1.0)
                            ws_fixed[(int64_t)i * NE + e] = i_l[e] + i_r[e];
                    }
            }
            __syncthreads();       // ws of this CTA's rows: written by lanes 0, read thread-strided below
            group_sum<7, NT>(acc, sh_red, gc);
            if (threadIdx.x == 0) {
                double e1, r1, e2, r2;
                gk_error(acc[0], acc[1], acc[2], e1, r1);
                gk_error(acc[3], acc[4], acc[5], e2, r2);
                const double derr = e1 + e2 - b_err[t];
                gerr += derr;
                rnd += r1 + r2;
                h_err[h_cnt] = e1; h_a[h_cnt] = pa; h_b[h_cnt] = pc; h_cnt++;
                h_err[h_cnt] = e2; h_a[h_cnt] = pc; h_b[h_cnt] = pb; h_cnt++;
                nbis++;
            }
            __syncthreads();
        }
        // |global_integral| after the batch
        {
            double acc[1] = {0.0};
            for (int i = wid0; i < D; i += wstride) {      // own rows only
                if (WARP) {                                 // lanes over the elements
                    if (lane < NE) {
                        const double g = ws[(int64_t)i * NE + lane];
                        acc[0] += g * g;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        const double g = ws[(int64_t)i * NE + e];
                        acc[0] += g * g;
                    }
                }
            }
            group_sum<1, NT>(acc, sh_red, gc);
            if (threadIdx.x == 0) {
                normG = sqrt(acc[0]);
                if (h_cnt >= 2) {
                    const double tol = fmax(1.0e-6, 1.0e-6 * normG);
                    if (gerr < tol / 8.0) stop_flag = 1;            // CONVERGED
                    else if (gerr < rnd) stop_flag = 2;             // ROUNDING_ERROR
                }
                if (!(isfinite(gerr) && isfinite(rnd))) stop_flag = 3;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && gc.rank == 0) {
        a.info[2 * n + (ENERGY ? 0 : 1)] = nbis;
        if (stop_flag == 0 && h_cnt >= kLimit) atomicOr(a.status, 16);
    }

    // ---- write the refined integral back ---------------------------------------
    if (ENERGY) {
        double acc[1] = {0.0};
        for (int i = wid0; i < D; i += wstride)
            if (writer) acc[0] += ws[i] / a.sigma[i];
        group_sum<1, NT>(acc, sh_red, gc);
        if (threadIdx.x == 0 && gc.rank == 0) a.esde_n[n] = 0.5 * acc[0];
    } else if (a.ds_out != nullptr) {
        group_barrier(gc);                 // the neighbours' rows may belong to other CTAs
        for (int j = gc.rank * NT + threadIdx.x; j < D; j += gc.G * NT) {
            double o[3] = {0.0, 0.0, 0.0};
            if (SYS == 3) {
                const int src[4] = {j, (j + D - 1) % D, (j + 1) % D, (j + 2) % D};
#pragma unroll
                for (int slot = 0; slot < 4; ++slot) {
                    const int i = src[slot];      // row whose `slot` neighbour is j
                    const double is = 1.0 / a.sigma[i];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int64_t q = (int64_t)i * NE + slot * 3 + k;
                        o[k] += is * (a.delta_mode ? __ldcg(ws + q) - __ldcg(ws_fixed + q) : __ldcg(ws + q));   // other CTAs' rows: through L2
                    }
                }
            } else {
                for (int i = 0; i < D; ++i) {
                    const double is = 1.0 / a.sigma[i];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int64_t q = (int64_t)i * NE + j * 3 + k;
                        o[k] += is * (a.delta_mode ? __ldcg(ws + q) - __ldcg(ws_fixed + q) : __ldcg(ws + q));   // other CTAs' rows: through L2
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) a.ds_out[((int64_t)n * D + j) * 3 + k] = 0.5 * o[k];
        }
    }
    __syncthreads();
}

// persistent CTAs in groups of a.group (co-resident: the grid never exceeds one wave);
// work item w = (interval, integrand), one group per item at a time
template <int SYS, bool WARP>
__global__ void __launch_bounds__(kRefThreads)
refine_kernel(const RefineArgs a) {
    const int nl = a.n_end - a.n_begin;
    pdl_wait();                      // flags, partials and x belong to the predecessor until here
    pdl_launch_dependents();
    if (a.nflagged != nullptr && a.nflagged[0] == 0) {        // the common case: nothing to do
        for (int w = blockIdx.x * kRefThreads + threadIdx.x; w < 2 * nl; w += gridDim.x * kRefThreads)
            a.info[2 * a.n_begin + w] = 0;
        return;
    }
    const int G = a.group, grp = blockIdx.x / G, ngrp = gridDim.x / G;
    GroupCtx gc;
    gc.G = G; gc.rank = blockIdx.x % G;
    gc.part = a.gpart + (int64_t)grp * 2 * G * 8;
    gc.bar = a.gbar + grp;
    gc.seq = 0; gc.status = a.status;
    double *ws = a.ws + (int64_t)grp * 2 * a.D * kMaxNE;
    double *ws_fixed = ws + (int64_t)a.D * kMaxNE;
    for (int w = grp; w < 2 * nl; w += ngrp) {
        const int n = a.n_begin + (w >> 1);
        const int which = w & 1;
        const int f = a.flags[n];
        if (threadIdx.x == 0 && gc.rank == 0) a.info[2 * n + which] = 0;
        if (which == 0 && (f & 1)) refine_item<SYS, true, WARP>(a, n, ws, ws_fixed, gc);
        if (which == 1 && (f & 2) && a.want_grad) refine_item<SYS, false, WARP>(a, n, ws, ws_fixed, gc);
    }
}

// Fused E_cost path: the gradient has already been assembled from the fixed rule;
// add (refined - fixed) dS integrals of the flagged intervals, times exp(x)
// (free_energy.py:743).  CTA n owns columns 2n and 2n+1, so every column has one
// writer and a fixed order of its (at most two) addends.
struct ApplyRefineArgs {
    const int32_t *flags;      // [L]
    const int32_t *nflagged;   // [1] or nullptr
    const double *delta;       // [L][D][3]  (refined - fixed), already * 1/2 / Sigma
    const double *x;           // optimisation vector (log-variances after D*Nm)
    double *gs;                // gradient, variance block [D][Ns]
    int D, Nm, Ns, n_begin, n_end;
};

__global__ void __launch_bounds__(256)
apply_refine_kernel(const ApplyRefineArgs a) {
    pdl_wait();                      // refine_kernel's deltas (no-ops after a plain launch)
    pdl_launch_dependents();
    if (a.nflagged != nullptr && a.nflagged[0] == 0) return;
    const int n = a.n_begin + blockIdx.x;             // n_end itself: last shared column only
    const bool f_own = n < a.n_end && (a.flags[n] & 2);
    const bool f_prev = n > a.n_begin && (a.flags[n - 1] & 2);
    if (!f_own && !f_prev) return;
    for (int j = threadIdx.x; j < a.D; j += blockDim.x) {
        const int64_t base = (int64_t)j * a.Ns + 2 * n;
        const double *xs = a.x + (int64_t)a.D * a.Nm;
        double d0 = 0.0;
        if (f_prev) d0 += a.delta[((int64_t)(n - 1) * a.D + j) * 3 + 2];
        if (f_own) d0 += a.delta[((int64_t)n * a.D + j) * 3 + 0];
        a.gs[base] += d0 * exp(xs[base]);
        if (f_own) a.gs[base + 1] += a.delta[((int64_t)n * a.D + j) * 3 + 1] * exp(xs[base + 1]);
    }
}

}  // namespace mfvgp
