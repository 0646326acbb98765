// CUDA kernels (sm_100a, FP64) for the MeanFieldVarGP free-energy path.
//
// Reference functions replaced (files under the reference repository):
//   FreeEnergy.single_interval  src/var_bayes/free_energy.py:338-459
//   FreeEnergy.E_sde            src/var_bayes/free_energy.py:461-565
//   FreeEnergy.E_kl0 / E_obs    src/var_bayes/free_energy.py:281-336, 567-649
//   FreeEnergy.E_cost           src/var_bayes/free_energy.py:651-751
//   Lorenz96._construct_functions src/dynamical_systems/lorenz_96.py:187-381
//   StochasticProcess.energy/grad_mean/grad_variance
//                               src/dynamical_systems/stochastic_process.py:299-448
//
// Data layout in HBM is the reference's: mean points [D][Nm], variance points
// [D][Ns] (row = state dimension, time contiguous), per-interval gradients
// dEsde_dm [L][D][4], dEsde_ds [L][D][3].
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "drift.cuh"
#include "tables.h"

namespace mfvgp {

// [42][16] table of the fixed 42-node rule (tables.h)
__constant__ double c_tab[kNodes * kTabStride];

// split rule (tables.h): rational part on the 42 GK nodes, polynomial part on 7 GL nodes
__constant__ double c_rat[kNodes * kTabStride];
__constant__ double c_gl[kGL * kTabStride];

__constant__ double c_gkx_main[kNodesPerPanel];     // Kronrod abscissae (descending)
__constant__ double c_vxB[kNodes], c_vxD[kNodes], c_m1poly[4];   // tables.h build_moment_table

constexpr int kPartStride = 12;  // doubles per (interval, dim-tile) partial
constexpr int kStatusBits = 6;   // MFVGP_ST_* bits carried through the multi-GPU record
constexpr int kRecordLen = 4 + kStatusBits;   // F, E0, Esde, Eobs + one 0/1 slot per status bit
// partial slots (raw sums; scaled by h = dt/4 in guard_reduce_kernel)
//  0: sum_i sum_r v_r E_i / Sigma_i         (weighted energy)
//  1: sum_i (sum_r v_r E_i)^2               (|I_E|^2 for quad_vec's tol)
//  2,3: sum_i (sum_r c_r E_i)^2, panels 1,2 (Kronrod - Gauss, quad_vec's err)
//  4,5: sum_i sum_k (sum_r c_r dS_own_ik)^2, panels 1,2
//  6: sum_i sum_k (sum_r v_r dS_own_ik)^2   (own-slot part: lower bound of |I_dS|^2)
//  7,8: sum_i (sum_r v_r x_r dS_own_i0)^2, panels 1,2 (lower bound of scipy's dabs^2)

struct EsdeArgs {
    const double *mean;     // [D][ldm]
    const double *vars;     // [D][lds]   variances, or log-variances if log_vars
    int64_t ldm, lds;
    const double *obs_times;  // [M]
    const double *sigma;      // [D]
    const double *theta;      // [ntheta]
    double *dm_out;           // [L][D][4] or nullptr
    double *ds_out;           // [L][D][3] or nullptr
    double *part;             // [(n_end-n_begin)*ntiles][kPartStride]
    int D, n_begin, n_end, ntiles, log_vars;
};

// Programmatic dependent launch (sm_90+): a kernel launched with the
// programmatic-stream-serialization attribute may be scheduled while its predecessor in
// the stream is still running; pdl_wait() blocks until that predecessor has completed and
// its writes are visible (no-op for a normal launch), pdl_launch_dependents() lets the
// successor's CTAs be scheduled early.  Hides ~2 us of launch latency per kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum of NV values per thread; result valid in thread 0.
template <int NV, int NT>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *sh /*[NV*NT/32]*/) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (NW > 1) {
        if (lane == 0)
#pragma unroll
            for (int i = 0; i < NV; ++i) sh[i * NW + w] = v[i];
        __syncthreads();
        if (threadIdx.x == 0)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                double s = sh[i * NW];
                for (int k = 1; k < NW; ++k) s += sh[i * NW + k];
                v[i] = s;
            }
    }
}

// ---------------------------------------------------------------------------
// Lorenz 96: one thread per (state dimension, interval) cell, 42-node loop.
// A CTA covers TE consecutive "extended" dimensions of one interval: TE-6
// output dimensions plus a wrap-around halo of 3 on each side, because the
// energy of dimension i couples (i, i+1, i-1, i-2) (lorenz_96.py:232-235) and
// the gradient of dimension j gathers from energies j, j-1, j+1, j+2.  The
// neighbour values travel through shared memory; the gather form needs no
// atomics, so results are deterministic.
// ---------------------------------------------------------------------------
template <int TE>
__global__ void __launch_bounds__(TE)
esde_l96_kernel(const EsdeArgs a) {
    constexpr int TD = TE - 6;
    __shared__ double sh_m[TE], sh_s[TE];
    __shared__ double sh_A[TE], sh_B[TE], sh_C[TE], sh_Dv[TE];
    __shared__ double sh_red[7 * (TE / 32)];

    const int e = threadIdx.x;
    const int tile = blockIdx.x % a.ntiles;
    const int n = a.n_begin + blockIdx.x / a.ntiles;
    const int d0 = tile * TD;
    const int D = a.D;
    int j = (d0 - 3 + e) % D;
    if (j < 0) j += D;
    const bool has_energy = (e >= 2) && (e <= TE - 2);
    const bool is_out = (e >= 3) && (e <= TE - 4) && (d0 + e - 3 < D);

    const double ti = a.obs_times[n], tj = a.obs_times[n + 1];
    const double dt = fabs(tj - ti);
    const double inv_dt = 1.0 / dt;

    const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
    const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
    const double P0 = mp[0], P1 = mp[1], P2 = mp[2], P3 = mp[3];
    double S0 = sp[0], S1 = sp[1], S2 = sp[2];
    if (a.log_vars) { S0 = exp(S0); S1 = exp(S1); S2 = exp(S2); }
    const double Sig = a.sigma[j];
    const double inv_sig = 1.0 / Sig;
    const double theta = a.theta[0];

    double accM[4] = {0, 0, 0, 0}, accS[3] = {0, 0, 0};
    double accE = 0.0, kgE = 0.0, kgS[3] = {0, 0, 0}, ownS1 = 0.0;
    double eE1 = 0.0, eS1 = 0.0;

#pragma unroll 1
    for (int q = 0; q < kNodes; ++q) {
        const double *T = c_tab + q * kTabStride;
        const double m = P0 * T[T_B3] + P1 * T[T_B3 + 1] + P2 * T[T_B3 + 2] + P3 * T[T_B3 + 3];
        const double dm = (P0 * T[T_DB3] + P1 * T[T_DB3 + 1] + P2 * T[T_DB3 + 2]
                           + P3 * T[T_DB3 + 3]) * inv_dt;
        const double s = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
        const double ds = (S0 * T[T_DB2] + S1 * T[T_DB2 + 1] + S2 * T[T_DB2 + 2]) * inv_dt;
        sh_m[e] = m;
        sh_s[e] = s;
        __syncthreads();

        double own_m = 0.0, own_s = 0.0, own_ds = 0.0;
        if (has_energy) {
            const double ma = sh_m[e + 1], mb = sh_m[e - 1], mc = sh_m[e - 2];
            const double sa = sh_s[e + 1], sb = sh_s[e - 1], sc = sh_s[e - 2];
            const double u = ma - mc, su = sa + sc;
            const double R = fma(u, mb, theta - m) - dm;       // E[f_i] - dm_i
            const Rational r(s, ds, Sig);
            const double E = R * R + s + sb * u * u + mb * mb * su + su * sb + r.rat + r.q;
            const double w = T[T_V] * inv_sig;
            const double wR2 = 2.0 * w * R;
            sh_A[e] = wR2 * mb + 2.0 * w * sb * u;             // w dE_i/dm_{i+1} (= -w dE_i/dm_{i-2})
            sh_B[e] = wR2 * u + 2.0 * w * mb * su;             // w dE_i/dm_{i-1}
            sh_C[e] = w * (mb * mb + sb);                      // w dE_i/ds_{i+1} = w dE_i/ds_{i-2}
            sh_Dv[e] = w * (u * u + su);                       // w dE_i/ds_{i-1}
            own_m = -wR2;                                      // w dE_i/dm_i = w dE_i/d(dm_i)
            const double es = 1.0 + r.rat_s, eds = r.rat_ds + 1.0;
            own_s = w * es;
            own_ds = w * eds;
            if (is_out) {
                accE = fma(w, E, accE);
                const double c = T[T_C];
                kgE = fma(c, E, kgE);
                const double ces = c * es, ceds = c * eds * inv_dt;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    kgS[k] += ces * T[T_B2 + k] + ceds * T[T_DB2 + k];
                ownS1 += T[T_V] * (es * T[T_B2 + 1] + eds * inv_dt * T[T_DB2 + 1]);
            }
        }
        __syncthreads();
        if (is_out) {
            const double Gm = own_m + sh_A[e - 1] + sh_B[e + 1] - sh_A[e + 2];
            const double Gs = own_s + sh_C[e - 1] + sh_Dv[e + 1] + sh_C[e + 2];
            const double Hm = own_m * inv_dt, Hs = own_ds * inv_dt;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                accM[k] += Gm * T[T_B3 + k] + Hm * T[T_DB3 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                accS[k] += Gs * T[T_B2 + k] + Hs * T[T_DB2 + k];
        }
        if (q == kNodesPerPanel - 1) {                         // end of panel 1
            eE1 = kgE * kgE;
            eS1 = kgS[0] * kgS[0] + kgS[1] * kgS[1] + kgS[2] * kgS[2];
            kgE = 0.0;
            kgS[0] = kgS[1] = kgS[2] = 0.0;
        }
    }

    // free_energy.py:446-458: the 1/2 factors; h = dt/4 is the panel half width
    const double scale = 0.5 * (0.25 * dt);
    if (is_out && a.dm_out != nullptr) {
        double *om = a.dm_out + ((int64_t)n * D + j) * 4;
        double *os = a.ds_out + ((int64_t)n * D + j) * 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) om[k] = scale * accM[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) os[k] = scale * accS[k];
    }
    const double KE = accE * Sig;                              // raw sum_r v_r E_i
    double red[7] = {accE, KE * KE, eE1, kgE * kgE, eS1,
                     kgS[0] * kgS[0] + kgS[1] * kgS[1] + kgS[2] * kgS[2], ownS1 * ownS1};
    if (!is_out)
#pragma unroll
        for (int i = 0; i < 7; ++i) red[i] = 0.0;
    block_sum<7, TE>(red, sh_red);
    if (threadIdx.x == 0) {
        double *p = a.part + (int64_t)blockIdx.x * kPartStride;
#pragma unroll
        for (int i = 0; i < 7; ++i) p[i] = red[i];
        p[7] = 0.0;
    }
}


// ---------------------------------------------------------------------------
// Lorenz 96, split rule (default).  Same cell/tile mapping and same outputs as
// esde_l96_kernel, about 3.5x fewer FP64 instructions:
//   E_i = [R^2 + sb u^2 + b^2 su + su sb]  coupled polynomial, degree <= 12
//       + [s_i + q_i]                      own polynomial, integrated in closed form
//       + q_i^2/(4 s_i)                    rational, own dimension only
// (A) the rational part and its d/dS run over scipy's 42 Kronrod nodes with no
//     neighbour traffic; they also carry quad_vec's Kronrod-minus-Gauss error
//     estimate (the polynomial parts contribute nothing to it);
// (B) the coupled polynomial part runs over 7 Gauss-Legendre nodes (exact to
//     degree 13) with the shared-memory neighbour exchange of the 42-node kernel.
// GK21 integrates those polynomials exactly as well, so both kernels return the
// reference's value up to rounding (tests/test_gpu_parity.py compares them).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp3(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);          // |e| <= 2^-20
    const double t = fma(e, e, e);             // e + e^2  -> error e^3 = 2^-60
    return fma(r, t, r);                       // measured max rel. error 2.2e-16 (tools/rcp_test.cu)
}

template <int TE>
__device__ __forceinline__ void esde_l96_split_body(const EsdeArgs &a) {
    constexpr int TD = TE - 6;
    __shared__ double sh_m[TE], sh_s[TE];
    __shared__ double sh_A[TE], sh_B[TE], sh_C[TE], sh_Dv[TE];
    __shared__ double sh_red[9 * (TE / 32)];

    const int e = threadIdx.x;
    const int tile = blockIdx.x % a.ntiles;
    const int n = a.n_begin + blockIdx.x / a.ntiles;
    const int d0 = tile * TD;
    const int D = a.D;
    int j = (d0 - 3 + e) % D;
    if (j < 0) j += D;
    const bool has_energy = (e >= 2) && (e <= TE - 2);
    const bool is_out = (e >= 3) && (e <= TE - 4) && (d0 + e - 3 < D);

    const double ti = a.obs_times[n], tj = a.obs_times[n + 1];
    const double dt = fabs(tj - ti);
    const double inv_dt = 1.0 / dt;

    const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
    const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
    const double P0 = mp[0], P1 = mp[1], P2 = mp[2], P3 = mp[3];
    double S0 = sp[0], S1 = sp[1], S2 = sp[2];
    if (a.log_vars) { S0 = exp(S0); S1 = exp(S1); S2 = exp(S2); }
    const double Sig = a.sigma[j];
    const double inv_sig = 1.0 / Sig;
    const double theta = a.theta[0];

    // ---- (A) rational part, own dimension, 42 Kronrod nodes -------------------
    // Per panel: Kronrod sums K (all 21 nodes) and embedded Gauss sums G (the 10
    // odd nodes); quad_vec's raw error estimate is |K - G| (its absolute accuracy
    // here, ~1e-16 |K|, is far below the ~1e-9 |K| level the guard decides at).
    // Fully unrolled so every table entry is a constant-bank operand of its FMA.
    double aE = 0.0, eE1 = 0.0, eE2 = 0.0, eS1 = 0.0, eS2 = 0.0, m1S1 = 0.0, m1S2 = 0.0;
    double aSv[3] = {0, 0, 0}, aSd[3] = {0, 0, 0};
    if (is_out) {
        // s(tau) = a0 + a1 tau + a2 tau^2 ; q(tau) = ds/dt - Sigma = b0 + b1 tau
        const double a0 = S0, a1 = -3.0 * S0 + 4.0 * S1 - S2, a2 = 2.0 * (S0 - 2.0 * S1 + S2);
        const double b0 = a1 * inv_dt - Sig, b1 = 2.0 * a2 * inv_dt;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            double kE = 0.0, gE = 0.0, mSv = 0.0, mSd = 0.0;
            double kSv[3] = {0, 0, 0}, kSd[3] = {0, 0, 0}, gSv[3] = {0, 0, 0}, gSd[3] = {0, 0, 0};
#pragma unroll
            for (int r = 0; r < kNodesPerPanel; ++r) {
                const double *T = c_rat + (p * kNodesPerPanel + r) * kTabStride;
                const double tau = T[R_TAU];
                const double s = fma(fma(a2, tau, a1), tau, a0);
                const double qq = fma(b1, tau, b0);
                const double qi = qq * fast_rcp3(s);   // q / s
                const double t = qq * qi;              // q^2 / s   (= 4 rat)
                const double qi2 = qi * qi;            // (q/s)^2   (= -4 d rat/d s)
                kE = fma(T[R_V], t, kE);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    kSv[k] = fma(T[R_VB + k], qi2, kSv[k]);
                    kSd[k] = fma(T[R_VD + k], qi, kSd[k]);
                }
                mSv = fma(c_vxB[p * kNodesPerPanel + r], qi2, mSv);
                mSd = fma(c_vxD[p * kNodesPerPanel + r], qi, mSd);
                if (r & 1) {
                    gE = fma(T[R_W], t, gE);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        gSv[k] = fma(T[R_WB + k], qi2, gSv[k]);
                        gSd[k] = fma(T[R_WD + k], qi, gSd[k]);
                    }
                }
            }
            const double dE = 0.25 * (kE - gE);
            double eS = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double kg = fma(0.5 * inv_dt, kSd[k] - gSd[k], -0.25 * (kSv[k] - gSv[k]));
                eS = fma(kg, kg, eS);
                aSv[k] += kSv[k];
                aSd[k] += kSd[k];
            }
            aE += kE;
            // first moment of the own-slot k=0 element of dS about the panel centre
            const double m1 = fma(0.5 * inv_dt, mSd, -0.25 * mSv)
                              + fma(c_m1poly[2 * p + 1], inv_dt, c_m1poly[2 * p]);
            if (p == 0) { eE1 = dE * dE; eS1 = eS; m1S1 = m1 * m1; }
            else { eE2 = dE * dE; eS2 = eS; m1S2 = m1 * m1; }
        }
    }

// ---- (B) coupled polynomial part, 7 Gauss-Legendre nodes -------------------
    double accM[4] = {0, 0, 0, 0}, accS[3] = {0, 0, 0}, accE = 0.0;
#pragma unroll
    for (int g = 0; g < kGL; ++g) {
        const double *T = c_gl + g * kTabStride;
        const double m = P0 * T[T_B3] + P1 * T[T_B3 + 1] + P2 * T[T_B3 + 2] + P3 * T[T_B3 + 3];
        const double dm = (P0 * T[T_DB3] + P1 * T[T_DB3 + 1] + P2 * T[T_DB3 + 2]
                           + P3 * T[T_DB3 + 3]) * inv_dt;
        const double s = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
        sh_m[e] = m;
        sh_s[e] = s;
        __syncthreads();
        double own_m = 0.0;
        if (has_energy) {
            const double ma = sh_m[e + 1], mb = sh_m[e - 1], mc = sh_m[e - 2];
            const double sa = sh_s[e + 1], sb = sh_s[e - 1], sc = sh_s[e - 2];
            const double u = ma - mc, su = sa + sc;
            const double R = fma(u, mb, theta - m) - dm;
            const double t1 = sb * u, t2 = mb * su;
            const double Ec = fma(R, R, fma(t1, u, fma(t2, mb, su * sb)));
            const double w = T[T_V] * inv_sig, w2 = w + w;
            const double wR2 = w2 * R;
            sh_A[e] = fma(wR2, mb, w2 * t1);
            sh_B[e] = fma(wR2, u, w2 * t2);
            sh_C[e] = w * fma(mb, mb, sb);
            sh_Dv[e] = w * fma(u, u, su);
            own_m = -wR2;
            accE = fma(w, Ec, accE);
        }
        __syncthreads();
        if (is_out) {
            const double Gm = own_m + sh_A[e - 1] + sh_B[e + 1] - sh_A[e + 2];
            const double Gs = sh_C[e - 1] + sh_Dv[e + 1] + sh_C[e + 2];
            const double Hm = own_m * inv_dt;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                accM[k] += Gm * T[T_B3 + k] + Hm * T[T_DB3 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) accS[k] = fma(Gs, T[T_B2 + k], accS[k]);
        }
    }

    // ---- epilogue: closed-form own polynomial, scaling, outputs ----------------
    // All "K-sums" below use the normalisation of the 42-node kernel
    // (integral = (dt/4) * Ksum) so guard_reduce_kernel serves both.
    const double four_inv_dt = 4.0 * inv_dt;
    // int (s_i + q_i) dt = dt (S0 + 4 S1 + S2)/6 + (S2 - S0) - Sigma dt
    const double own_poly = dt * (S0 + 4.0 * S1 + S2) * (1.0 / 6.0) + (S2 - S0) - Sig * dt;
    const double ksum_E = 0.25 * aE + 4.0 * accE * Sig + four_inv_dt * own_poly;   // raw, this i
    double kS[3];                         // rational d/dS K-sums of the own slot
#pragma unroll
    for (int k = 0; k < 3; ++k) kS[k] = fma(0.5 * inv_dt, aSd[k], -0.25 * aSv[k]);
    if (is_out && a.dm_out != nullptr) {
        double *om = a.dm_out + ((int64_t)n * D + j) * 4;
        double *os = a.ds_out + ((int64_t)n * D + j) * 3;
        const double sc_gl = 0.5 * dt;                 // 1/2 * interval length (GL weights sum to 1)
        const double sc_gk = 0.5 * 0.25 * dt * inv_sig;
        const double polyS[3] = {dt * (1.0 / 6.0) - 1.0, dt * (4.0 / 6.0), dt * (1.0 / 6.0) + 1.0};
#pragma unroll
        for (int k = 0; k < 4; ++k) om[k] = sc_gl * accM[k];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            os[k] = sc_gl * accS[k] + sc_gk * kS[k] + 0.5 * inv_sig * polyS[k];
    }
    double own2 = 0.0;                                    // own-slot K-sums, k = 0..2
    {
        const double polyK[3] = {four_inv_dt * (dt * (1.0 / 6.0) - 1.0), four_inv_dt * dt * (4.0 / 6.0),
                                 four_inv_dt * (dt * (1.0 / 6.0) + 1.0)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double o = kS[k] + polyK[k];
            own2 = fma(o, o, own2);
        }
    }
    double red[9] = {ksum_E * inv_sig, ksum_E * ksum_E, eE1, eE2, eS1, eS2, own2, m1S1, m1S2};
    if (!is_out)
#pragma unroll
        for (int i = 0; i < 9; ++i) red[i] = 0.0;
    block_sum<9, TE>(red, sh_red);
    if (threadIdx.x == 0) {
        double *p = a.part + (int64_t)blockIdx.x * kPartStride;
#pragma unroll
        for (int i = 0; i < 9; ++i) p[i] = red[i];
    }
}

template <int TE>
__global__ void __launch_bounds__(TE, 640 / TE)
esde_l96_split_kernel(const EsdeArgs a) { esde_l96_split_body<TE>(a); }

// Batched evaluation (mfvgp_ecost_batch): one of B independent states of the SAME problem per
// blockIdx.y / blockIdx.z; every per-evaluation buffer is rebased by these strides (doubles).
struct BatchStrides {
    int64_t x, dm, ds, part;          // state / gradient, per-interval workspaces, tile partials
    int64_t apart, esde_n, out;       // tail: partial rows, per-interval energies, [8] records
};

// one thread per cell: the throughput shape (the cooperative kernel's two lanes per cell and
// 250 registers are for latency); used when the batch alone fills the GPU
template <int TE>
__global__ void __launch_bounds__(TE, 640 / TE)
esde_l96_split_batch_kernel(const EsdeArgs a0, const BatchStrides bs) {
    EsdeArgs a = a0;
    const int64_t z = blockIdx.y;
    a.mean += z * bs.x; a.vars += z * bs.x;
    if (a.dm_out != nullptr) { a.dm_out += z * bs.dm; a.ds_out += z * bs.ds; }
    a.part += z * bs.part;
    esde_l96_split_body<TE>(a);
}


// ---------------------------------------------------------------------------
// DW / OU / L63 (D <= 3): one CTA per interval, one thread per quadrature
// node; each thread evaluates all dimensions at its node, the 42 node values
// are then summed in scipy's order (panel 1 then panel 2, r = 0..20).
// ---------------------------------------------------------------------------
// write_mask: bit 0 dm_out, bit 1 ds_out; slot = row of `part` to write (default: the CTA's)
template <int SYS>
__device__ __forceinline__ void esde_small_body(const EsdeArgs &a, int write_mask = 3,
                                                int slot = -1) {
    using Dr = Drift<SYS>;
    constexpr int D = Dr::D;
    constexpr int NV = 1 + 4 * D + 3 * D + D + 3 * D;   // see value layout below
    constexpr int V_E = 0, V_DM = 1, V_DS = V_DM + 4 * D, V_RE = V_DS + 3 * D,
                  V_RS = V_RE + D;
    __shared__ double vals[NV][kNodes];
    __shared__ double sums[NV][6];                       // K1, K2, KG1, KG2, M1 (panels 1, 2)

    const int n = a.n_begin + blockIdx.x;
    const int q = threadIdx.x;
    const double ti = a.obs_times[n], tj = a.obs_times[n + 1];
    const double dt = fabs(tj - ti);
    const double inv_dt = 1.0 / dt;

    if (q < kNodes) {
        const double *T = c_tab + q * kTabStride;
        double m[D], dm[D], s[D], ds[D], Sig[D], inv_sig[D], th[Dr::NTHETA];
#pragma unroll
        for (int i = 0; i < Dr::NTHETA; ++i) th[i] = a.theta[i];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
            const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
            double S0 = sp[0], S1 = sp[1], S2 = sp[2];
            if (a.log_vars) { S0 = exp(S0); S1 = exp(S1); S2 = exp(S2); }
            m[j] = mp[0] * T[T_B3] + mp[1] * T[T_B3 + 1] + mp[2] * T[T_B3 + 2]
                   + mp[3] * T[T_B3 + 3];
            dm[j] = (mp[0] * T[T_DB3] + mp[1] * T[T_DB3 + 1] + mp[2] * T[T_DB3 + 2]
                     + mp[3] * T[T_DB3 + 3]) * inv_dt;
            s[j] = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
            ds[j] = (S0 * T[T_DB2] + S1 * T[T_DB2 + 1] + S2 * T[T_DB2 + 2]) * inv_dt;
            Sig[j] = a.sigma[j];
            inv_sig[j] = 1.0 / Sig[j];
        }
        double E[D], Em[D][D], Edm[D], Es[D][D], Eds[D];
        Dr::nodal(m, dm, s, ds, Sig, th, E, Em, Edm, Es, Eds);
        double we = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) we += inv_sig[i] * E[i];
        vals[V_E][q] = we;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double gm = 0.0, gs = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                gm += inv_sig[i] * Em[i][j];
                gs += inv_sig[i] * Es[i][j];
            }
            const double hm = inv_sig[j] * Edm[j] * inv_dt;
            const double hs = inv_sig[j] * Eds[j] * inv_dt;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                vals[V_DM + 4 * j + k][q] = gm * T[T_B3 + k] + hm * T[T_DB3 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                vals[V_DS + 3 * j + k][q] = gs * T[T_B2 + k] + hs * T[T_DB2 + k];
                vals[V_RS + 3 * j + k][q] =
                    Es[j][j] * T[T_B2 + k] + Eds[j] * inv_dt * T[T_DB2 + k];
            }
            vals[V_RE + j][q] = E[j];
        }
    }
    __syncthreads();
    if (q < NV) {
        double K1 = 0.0, K2 = 0.0, G1 = 0.0, G2 = 0.0, M1 = 0.0, M2 = 0.0;
        for (int r = 0; r < kNodesPerPanel; ++r) {
            const double v = c_tab[r * kTabStride + T_V], c = c_tab[r * kTabStride + T_C];
            const double f1 = vals[q][r], f2 = vals[q][kNodesPerPanel + r];
            K1 = fma(v, f1, K1);
            K2 = fma(v, f2, K2);
            G1 = fma(c, f1, G1);
            G2 = fma(c, f2, G2);
            const double vx = v * c_gkx_main[r];          // first moment about the panel centre
            M1 = fma(vx, f1, M1);
            M2 = fma(vx, f2, M2);
        }
        sums[q][0] = K1; sums[q][1] = K2; sums[q][2] = G1; sums[q][3] = G2;
        sums[q][4] = M1; sums[q][5] = M2;
        const double scale = 0.5 * (0.25 * dt);
        if (a.dm_out != nullptr) {
            if (q >= V_DM && q < V_DS && (write_mask & 1))
                a.dm_out[(int64_t)n * D * 4 + (q - V_DM)] = scale * (K1 + K2);
            else if (q >= V_DS && q < V_RE && (write_mask & 2))
                a.ds_out[(int64_t)n * D * 3 + (q - V_DS)] = scale * (K1 + K2);
        }
    }
    __syncthreads();
    if (q == 0) {
        double p[kPartStride] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        p[0] = sums[V_E][0] + sums[V_E][1];
        for (int i = 0; i < D; ++i) {
            const double KE = sums[V_RE + i][0] + sums[V_RE + i][1];
            p[1] += KE * KE;
            p[2] += sums[V_RE + i][2] * sums[V_RE + i][2];
            p[3] += sums[V_RE + i][3] * sums[V_RE + i][3];
            for (int k = 0; k < 3; ++k) {
                const double *sv = sums[V_RS + 3 * i + k];
                p[4] += sv[2] * sv[2];
                p[5] += sv[3] * sv[3];
            }
            for (int k = 0; k < 3; ++k) {
                const double KS = sums[V_RS + 3 * i + k][0] + sums[V_RS + 3 * i + k][1];
                p[6] += KS * KS;
            }
            p[7] += sums[V_RS + 3 * i][4] * sums[V_RS + 3 * i][4];
            p[8] += sums[V_RS + 3 * i][5] * sums[V_RS + 3 * i][5];
        }
        double *out = a.part + (int64_t)(slot < 0 ? blockIdx.x : slot) * kPartStride;
        for (int i = 0; i < kPartStride; ++i) out[i] = p[i];
    }
}

template <int SYS>
__global__ void __launch_bounds__(64)
esde_small_kernel(const EsdeArgs a) { esde_small_body<SYS>(a); }

// ---------------------------------------------------------------------------
// Per-interval reduction of the dim-tile partials: the interval's Esde share
// and a conservative version of quad_vec's convergence test after its first
// bisection (scipy _quad_vec.py: err_p <= 200 |(K-G) h|, converged when
// sum_p err_p < max(epsabs, epsrel |I|)/8).  An interval whose bound fails is
// flagged for the adaptive kernels; otherwise the fixed rule IS quad_vec's
// answer.  One warp per interval.
// ---------------------------------------------------------------------------
struct GuardArgs {
    const double *part;        // [(n_end-n_begin)*ntiles][kPartStride]
    const double *obs_times;
    double *esde_n;            // [L] per-interval Esde (already * 1/2)
    int32_t *flags;            // [L] bit0 energy integrand, bit1 dS integrand
    int32_t *nflagged;         // [1] number of flagged intervals of this evaluation, or nullptr
                               //     (lets refine / apply_refine exit without scanning the flags)
    int n_begin, n_end, ntiles;
};

// one warp: interval n, lane = 0..31 (partials may come from other CTAs of the same launch:
// read them through L2)
// esde_out / flag_out (lane 0): what was written to esde_n[n] / flags[n], for callers that go on
// to sum them (saves reading back what this thread has just stored)
__device__ __forceinline__ void guard_interval(const GuardArgs &a, int n, int lane,
                                               double *esde_out = nullptr, int *flag_out = nullptr) {
    const int warp = n - a.n_begin;
    double p[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 3                     // all the loads of up to 96 tiles in flight at once
    for (int t = lane; t < a.ntiles; t += 32) {
        const double *src = a.part + ((int64_t)warp * a.ntiles + t) * kPartStride;
#pragma unroll
        for (int i = 0; i < 9; ++i) p[i] += __ldcg(src + i);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) p[i] = warp_sum(p[i]);
    if (lane == 0) {
        const double dt = fabs(a.obs_times[n + 1] - a.obs_times[n]);
        const double h = 0.25 * dt;
        a.esde_n[n] = 0.5 * h * p[0];
        const double tolE = fmax(1.0e-6, 1.0e-6 * h * sqrt(p[1]));
        const double tolS = fmax(1.0e-6, 1.0e-6 * h * sqrt(p[6]));
        // scipy: err_p = dabs_p min(1, (200 e_p/dabs_p)^1.5), e_p = |(K-G) h|.  It never
        // exceeds 200 e_p, and with a lower bound dl_p <= dabs_p it never exceeds
        // (200 e_p)^1.5 / sqrt(dl_p) once 200 e_p <= dl_p.  2 % margin on e_p.
        const double ubE = 204.0 * h * (sqrt(p[2]) + sqrt(p[3]));
        double ubS = 0.0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double e200 = 204.0 * h * sqrt(p[4 + k]), dl = h * sqrt(p[7 + k]);
            ubS += (e200 <= dl && dl > 0.0) ? e200 * sqrt(e200 / dl) : e200;
        }
        int32_t f = 0;
        if (!(ubE < 0.125 * tolE)) f |= 1;
        if (!(ubS < 0.125 * tolS)) f |= 2;
        a.flags[n] = f;
        if (esde_out != nullptr) { *esde_out = 0.5 * h * p[0]; *flag_out = f; }
        if (f && a.nflagged != nullptr) atomicAdd(a.nflagged, 1);
    }
}

__global__ void __launch_bounds__(256)
guard_reduce_kernel(const GuardArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n = a.n_begin + warp;
    if (n >= a.n_end) return;
    guard_interval(a, n, threadIdx.x & 31);
}

// ---------------------------------------------------------------------------
// E_cost assembly (free_energy.py:686-750): stitches the per-interval
// gradients (shared end points add), adds dE0 at column 0 and dEobs at the
// observation columns, applies the log-variance chain rule, and accumulates
// the E_kl0 / E_obs energy terms.  One CTA per state dimension (row of x).
// ---------------------------------------------------------------------------
struct AssembleArgs {
    const double *x;          // [D*Nm + D*Ns]
    const double *dm;         // [L][D][4]
    const double *ds;         // [L][D][3]
    const double *mu0, *tau0;
    const double *obs_values; // [M][d]
    const double *obs_noise;  // [d]
    const int32_t *obs_idx;   // [D] index into the observed dims or -1
    double *grad;             // same layout as x (may be nullptr)
    double *apart;            // [D][4]: sum log(tau0/s0), sum kl0 quad, sum obs, -
    int D, M, d, Nm, Ns;
    int n_begin, n_end, L;    // owned intervals; L = M-1
    int with_e0;              // this shard owns column 0
};

// row j of x / of the gradient; threads t0, t0+stride, ... share the columns
__device__ __forceinline__ void assemble_row(const AssembleArgs &a, int j, int t0, int stride,
                                             double (&acc)[3]) {
    const int jo = a.obs_idx[j];
    const double Ri = jo >= 0 ? 1.0 / a.obs_noise[jo] : 0.0;
    const double *xm = a.x + (int64_t)j * a.Nm;
    const double *xs = a.x + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
    double *gm = a.grad ? a.grad + (int64_t)j * a.Nm : nullptr;
    double *gs = a.grad ? a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns : nullptr;
    // Column ranges of this shard.  Owned columns [c0, own1) carry the E0 / obs
    // terms; the active range additionally holds the ghost column shared with
    // the next shard (its left edge), which only receives this shard's k=3 / k=2
    // contribution here -- the rest arrives through the edge exchange.
    const bool last = a.n_end == a.L;
    const int cm0 = 3 * a.n_begin, own_m1 = last ? a.Nm : 3 * a.n_end;
    const int cs0 = 2 * a.n_begin, own_s1 = last ? a.Ns : 2 * a.n_end;
    const int cm1 = last ? a.Nm : own_m1 + 1, cs1 = last ? a.Ns : own_s1 + 1;

    for (int c = cm0 + t0; c < cm1; c += stride) {
        const int n = c / 3, k = c - 3 * n;
        double g = 0.0;
        if (a.dm != nullptr) {
            if (n >= a.n_begin && n < a.n_end)
                g = a.dm[((int64_t)n * a.D + j) * 4 + k];
            if (k == 0 && n - 1 >= a.n_begin && n - 1 < a.n_end)
                g += a.dm[((int64_t)(n - 1) * a.D + j) * 4 + 3];
        }
        const double m = xm[c];
        if (c == 0 && a.with_e0) {
            const double z0 = m - a.mu0[j];
            g += z0 / a.tau0[j];
            acc[1] += z0 * z0 / a.tau0[j];
        }
        if (k == 0 && n >= 1 && n <= a.M && jo >= 0 && c < own_m1) {
            const double r = a.obs_values[(int64_t)(n - 1) * a.d + jo] - m;
            g -= Ri * r;
            acc[2] += Ri * r * r;
        }
        if (gm) gm[c] = g;
    }
    for (int c = cs0 + t0; c < cs1; c += stride) {
        const int n = c >> 1, k = c & 1;
        double g = 0.0;
        if (a.ds != nullptr) {
            if (n >= a.n_begin && n < a.n_end)
                g = a.ds[((int64_t)n * a.D + j) * 3 + k];
            if (k == 0 && n - 1 >= a.n_begin && n - 1 < a.n_end)
                g += a.ds[((int64_t)(n - 1) * a.D + j) * 3 + 2];
        }
        const double s = exp(xs[c]);
        if (c == 0 && a.with_e0) {
            const double t0_ = a.tau0[j];
            g += 0.5 * (1.0 / t0_ - 1.0 / s);
            acc[0] += log(t0_ / s);
            acc[1] += (s - t0_) / t0_;
        }
        if (k == 0 && n >= 1 && n <= a.M && jo >= 0 && c < own_s1) {
            g += 0.5 * Ri;
            acc[2] += Ri * s;
        }
        if (gs) gs[c] = g * s;                       // free_energy.py:743
    }
}

__global__ void __launch_bounds__(256)
assemble_kernel(const AssembleArgs a) {
    __shared__ double sh_red[3 * 8];
    const int j = blockIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    assemble_row(a, j, threadIdx.x, blockDim.x, acc);
    block_sum<3, 256>(acc, sh_red);
    if (threadIdx.x == 0) {
        double *p = a.apart + (int64_t)j * 4;
        p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2]; p[3] = 0.0;
    }
}

// ---------------------------------------------------------------------------
// Final scalar reduction (fixed order => deterministic), status bits.
// out[0..3] = F, E0, Esde, Eobs.
// ---------------------------------------------------------------------------
struct FinalArgs {
    const double *esde_n;     // [L]
    const int32_t *flags;     // [L]
    const double *apart;      // [D][4] or nullptr (E_sde only)
    double *out;              // [4]
    int32_t *status;          // [1]
    int32_t *refine_status;   // [1] bits set by refine_kernel, consumed here
    int32_t *nflagged;        // [1] GuardArgs::nflagged, reset here (or nullptr)
    unsigned *gbar;           // refine_kernel's group counters, zeroed here for the next evaluation
    int ngbar;
    int32_t *info;            // [L][2] bisection counts: zeroed here when nothing was flagged and
                              //        refine_kernel may not have been launched (or nullptr)
    int record9;              // out is the multi-GPU record: 4 energies + kStatusBits flag slots
    int n_begin, n_end, D;    // D = number of apart rows (state dims, or CTAs of the walk / tail kernel)
    int with_e0;
    double eobs_const;        // M (d log 2pi + safe_log(prod R)), 0 on non-owning shards
    // evaluation in two passes (mfvgp.cu ecost_impl, small problems on device pointers): the
    // first pass runs without the adaptive kernels; if the guard flagged an interval it raises
    // *late instead of cleaning up, refine_kernel then does its work and a second pass of the
    // tail kernel (which returns at once while *late is 0) assembles again and cleans up.
    int32_t *late = nullptr;        // first pass: set to 1 when an interval is flagged
    int32_t *late_clear = nullptr;  // second pass: reset to 0 at the end
};

// body for a CTA of NT threads (also the tail of assemble_tail_kernel, l96_coop.cuh);
// inputs written by other CTAs of the same launch are read through L2
template <int NT>
__device__ __forceinline__ void final_body(const FinalArgs &a) {
    __shared__ double sh_red[4 * (NT / 32)];
    __shared__ int sh_flag;
    if (threadIdx.x == 0) sh_flag = 0;
    // thread 0 needs two more words at the very end; they are final by now (refine_kernel ran
    // before this launch, the guard's count before this CTA saw the last ticket): fetch them
    // under the loads below instead of after the block sum
    int32_t rs_pre = 0, nf_pre = 0;
    if (threadIdx.x == 0) {
        rs_pre = __ldcg(a.refine_status);
        if (a.nflagged != nullptr) nf_pre = __ldcg(a.nflagged);
    }
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int anyflag = 0;
    for (int n = a.n_begin + threadIdx.x; n < a.n_end; n += NT) {
        acc[0] += __ldcg(a.esde_n + n);
        anyflag |= __ldcg(a.flags + n);
    }
    if (a.apart != nullptr)
        for (int j = threadIdx.x; j < a.D; j += NT) {
            acc[1] += __ldcg(a.apart + (int64_t)j * 4 + 0);
            acc[2] += __ldcg(a.apart + (int64_t)j * 4 + 1);
            acc[3] += __ldcg(a.apart + (int64_t)j * 4 + 2);
        }
    if (anyflag) atomicOr(&sh_flag, anyflag);
    if (a.info != nullptr && a.nflagged != nullptr && __ldcg(a.nflagged) == 0)
        for (int w = 2 * a.n_begin + threadIdx.x; w < 2 * a.n_end; w += NT) a.info[w] = 0;
    block_sum<4, NT>(acc, sh_red);
    __syncthreads();
    if (threadIdx.x == 0) {
        const double Esde = acc[0];
        double E0 = 0.0, Eobs = 0.0;
        if (a.apart != nullptr) {
            if (a.with_e0) {
                // safe_log(prod(tau0/s0)) (utilities.py:73-76): clamp in log space
                const double lo = -690.77552789821368, hi = 690.77552789821368;
                const double lg = fmin(fmax(acc[1], lo), hi);
                E0 = 0.5 * (lg + acc[2]);
            }
            Eobs = 0.5 * (acc[3] + a.eobs_const);
        }
        int32_t st = 0;
        if (!isfinite(Esde)) st |= 1;
        if (!isfinite(E0)) st |= 2;
        if (!isfinite(Eobs)) st |= 4;
        if (sh_flag) st |= 8;
        st |= rs_pre;
        const bool hand_over = a.late != nullptr && nf_pre != 0;   // a second pass follows
        if (hand_over) a.late[0] = 1;
        if (!hand_over) {
            a.refine_status[0] = 0;
            if (a.nflagged != nullptr) {
                if (nf_pre != 0)
                    for (int i = 0; i < a.ngbar; ++i) a.gbar[i] = 0u;
                a.nflagged[0] = 0;
            }
        }
        if (a.late_clear != nullptr) a.late_clear[0] = 0;
        // record[0..3] = F, E0, Esde, Eobs; record[4..9] = the six status bits as 0/1 so
        // that shards can be combined by summation (combine_record_kernel)
        a.out[0] = E0 + Esde + Eobs;
        a.out[1] = E0;
        a.out[2] = Esde;
        a.out[3] = Eobs;
        if (a.record9) {
#pragma unroll
            for (int b = 0; b < kStatusBits; ++b) a.out[4 + b] = (st >> b) & 1 ? 1.0 : 0.0;
        }
        a.status[0] = st;
    }
}

__global__ void __launch_bounds__(256)
final_kernel(const FinalArgs a) { final_body<256>(a); }

// many partial rows (walk path: one per CTA of the main kernel): 1024 threads
__global__ void __launch_bounds__(1024)
final_kernel_wide(const FinalArgs a) { final_body<1024>(a); }

// ---------------------------------------------------------------------------
// Multi-GPU (time intervals sharded, one rank per GPU).  Neighbouring shards
// share one column of support points (free_energy.py:504-505, 723-724): each
// computes its part of that column's gradient, the parts are exchanged over
// NVLink (ncclSend/ncclRecv) and added on both sides, so both hold the same
// bits.  buf layout: [0] mean left edge, [1] var left edge, [2] mean ghost,
// [3] var ghost, each [D].
// ---------------------------------------------------------------------------
struct EdgeArgs {
    double *grad;
    double *buf;              // [4][D]
    int D, Nm, Ns, n_begin, n_end, has_left, has_right;
};

__global__ void __launch_bounds__(256)
pack_edges_kernel(const EdgeArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.D) return;
    const double *gm = a.grad + (int64_t)j * a.Nm;
    const double *gs = a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
    if (a.has_left) {
        a.buf[j] = gm[3 * a.n_begin];
        a.buf[a.D + j] = gs[2 * a.n_begin];
    }
    if (a.has_right) {
        a.buf[2 * a.D + j] = gm[3 * a.n_end];
        a.buf[3 * a.D + j] = gs[2 * a.n_end];
    }
}

__global__ void __launch_bounds__(256)
add_edges_kernel(const EdgeArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.D) return;
    double *gm = a.grad + (int64_t)j * a.Nm;
    double *gs = a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
    if (a.has_left) {
        gm[3 * a.n_begin] += a.buf[j];
        gs[2 * a.n_begin] += a.buf[a.D + j];
    }
    if (a.has_right) {
        gm[3 * a.n_end] += a.buf[2 * a.D + j];
        gs[2 * a.n_end] += a.buf[3 * a.D + j];
    }
}

// One-collective exchange (E_cost with gradient): every rank all-gathers
// [record 16 | left edge 2D | right edge 2D]; stride = 16 + 4 D doubles.
struct XchgArgs {
    double *grad;
    const double *rec;        // [16] local record (final_kernel)
    double *send;             // [16 + 4 D]
    const double *gathered;   // [world][16 + 4 D]
    double *out;              // [4] F, E0, Esde, Eobs
    int32_t *status;
    int D, Nm, Ns, n_begin, n_end, rank, world;
};

__global__ void __launch_bounds__(256)
pack_all_kernel(const XchgArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < 16) a.send[threadIdx.x] = a.rec[threadIdx.x];
    if (j >= a.D) return;
    const double *gm = a.grad + (int64_t)j * a.Nm;
    const double *gs = a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
    double *e = a.send + 16;
    e[j] = gm[3 * a.n_begin];
    e[a.D + j] = gs[2 * a.n_begin];
    e[2 * a.D + j] = gm[3 * a.n_end];
    e[3 * a.D + j] = gs[2 * a.n_end];
}

// folds the records in rank order and adds the neighbours' parts of the shared columns:
// the left neighbour's RIGHT edge onto column 3 n_begin, the right neighbour's LEFT edge onto
// column 3 n_end (two addends each, so both owners end up with identical bits)
__global__ void __launch_bounds__(256)
combine_all_kernel(const XchgArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = 16 + 4 * (int64_t)a.D;
    if (blockIdx.x == 0 && threadIdx.x < kRecordLen) {
        const int i = threadIdx.x;
        double v = a.gathered[i];
        for (int r = 1; r < a.world; ++r) v += a.gathered[r * stride + i];
        if (i < 4) a.out[i] = v;
        // status bits travel as 0/1 in slots 4..9 (final_kernel record9)
        __shared__ double sh_bits[kRecordLen];
        sh_bits[i] = v;
        __syncwarp();
        if (i == 0) {
            int32_t st = 0;
            for (int b = 0; b < kStatusBits; ++b)
                if (sh_bits[4 + b] > 0.0) st |= 1 << b;
            a.status[0] = st;
        }
    }
    if (j >= a.D) return;
    double *gm = a.grad + (int64_t)j * a.Nm;
    double *gs = a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
    if (a.rank > 0) {
        const double *e = a.gathered + (a.rank - 1) * stride + 16;
        gm[3 * a.n_begin] += e[2 * a.D + j];
        gs[2 * a.n_begin] += e[3 * a.D + j];
    }
    if (a.rank < a.world - 1) {
        const double *e = a.gathered + (a.rank + 1) * stride + 16;
        gm[3 * a.n_end] += e[j];
        gs[2 * a.n_end] += e[a.D + j];
    }
}

// gathered [world][n]: slots < nsum are summed in rank order, the rest max'ed
__global__ void combine_record_kernel(const double *gathered, int world, int n, int nsum,
                                      double *out, int32_t *status /*or nullptr*/) {
    const int i = threadIdx.x;
    if (i < n) {
        double v = gathered[i];
        for (int r = 1; r < world; ++r)
            v = i < nsum ? v + gathered[r * n + i] : fmax(v, gathered[r * n + i]);
        out[i] = v;
    }
    if (status != nullptr) {
        __syncthreads();
        if (i == 0) {
            int32_t st = 0;
            for (int b = 0; b < kStatusBits; ++b)
                if (out[4 + b] > 0.0) st |= 1 << b;
            status[0] = st;
        }
    }
}

// ---------------------------------------------------------------------------
// Stand-alone E_kl0 (free_energy.py:281-336) and E_obs (:567-649), for callers
// that use those two methods directly; E_cost fuses them into assemble_kernel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ekl0_kernel(const double *m0, const double *s0, const double *mu0, const double *tau0, int D,
            double *out /*[1]*/, double *dm0, double *ds0, int32_t *status) {
    __shared__ double sh_red[2 * 8];
    double acc[2] = {0.0, 0.0};
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const double z0 = m0[j] - mu0[j], t0 = tau0[j], s = s0[j];
        acc[0] += log(t0 / s);
        acc[1] += (z0 * z0 + s - t0) / t0;
        if (dm0 != nullptr) {
            dm0[j] = z0 / t0;
            ds0[j] = 0.5 * (1.0 / t0 - 1.0 / s);
        }
    }
    block_sum<2, 256>(acc, sh_red);
    if (threadIdx.x == 0) {
        const double lg = fmin(fmax(acc[0], -690.77552789821368), 690.77552789821368);
        const double E0 = lg + acc[1];
        out[0] = 0.5 * E0;
        status[0] = isfinite(E0) ? 0 : 2;
    }
}

// one CTA per observed dimension; mean/vars are the [d][M] slices the reference
// passes (mean_points[ix_ikm], vars_points[ix_iks]); row partials -> rowsum[d]
__global__ void __launch_bounds__(256)
eobs_rows_kernel(const double *mean, const double *vars, const double *obs_values,
                 const double *obs_noise, int d, int M, double *rowsum, double *dm, double *ds) {
    __shared__ double sh_red[8];
    const int jo = blockIdx.x;
    const double Ri = 1.0 / obs_noise[jo];
    double acc[1] = {0.0};
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const double r = obs_values[(int64_t)k * d + jo] - mean[(int64_t)jo * M + k];
        acc[0] += Ri * r * r + Ri * vars[(int64_t)jo * M + k];
        if (dm != nullptr) {
            dm[(int64_t)jo * M + k] = -Ri * r;
            ds[(int64_t)jo * M + k] = 0.5 * Ri;
        }
    }
    block_sum<1, 256>(acc, sh_red);
    if (threadIdx.x == 0) rowsum[jo] = acc[0];
}

__global__ void __launch_bounds__(256)
eobs_final_kernel(const double *rowsum, int d, double eobs_const, double *out, int32_t *status) {
    __shared__ double sh_red[8];
    double acc[1] = {0.0};
    for (int j = threadIdx.x; j < d; j += blockDim.x) acc[0] += rowsum[j];
    block_sum<1, 256>(acc, sh_red);
    if (threadIdx.x == 0) {
        const double Eobs = acc[0] + eobs_const;
        out[0] = 0.5 * Eobs;
        status[0] = isfinite(Eobs) ? 0 : 4;
    }
}

// test hook (MFVGP_INJECT_STATUS): ORs bits into the status word the next final_kernel consumes
__global__ void or_status_kernel(int32_t *status, int32_t bits) { atomicOr(status, bits); }

// ---------------------------------------------------------------------------
// FP64 FMA throughput micro-benchmark: the roofline denominator for the Esde
// kernels (MEASURED_PEAKS.json has no FP64 entry).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace mfvgp
