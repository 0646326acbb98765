// Gradients of the free energy with respect to the drift parameters theta and the diffusion
// (system noise) coefficients Sigma -- SURVEY.md 8(f) rank 4.
//
// The reference leaves the hooks for this and implements nothing: the setters
// FreeEnergy.sde_theta / sde_sigma ("should be used only if we optimize the drift (theta) /
// diffusion (Sigma) parameters", src/var_bayes/free_energy.py:241-279) and the note at the end
// of E_cost (:745-748).  E_kl0 (:281-336) and E_obs (:567-649) do not depend on theta or
// Sigma, so dF/dtheta = dEsde/dtheta and dF/dSigma = dEsde/dSigma, with (free_energy.py:423-458)
//
//   Esde = sum_n 1/2 sum_i (1/Sigma_i) int_n E_i dt,
//   E_i  = (E[f_i] - dm_i)^2 + Var[f_i] + q_i^2/(4 s_i) - q_i E[df_i/dx_i],   q_i = ds_i - Sigma_i
//   (drift.cuh; energy_functions/*.sym), hence
//
//   dEsde/dtheta_k = 1/2 sum_n sum_i (1/Sigma_i) int_n dE_i/dtheta_k dt
//   dEsde/dSigma_i = 1/2 sum_n [ -(1/Sigma_i^2) int_n E_i dt + (1/Sigma_i) int_n dE_i/dSigma_i dt ],
//   dE_i/dSigma_i  = -q_i/(2 s_i) + E[df_i/dx_i]  = -dE_i/d(ds_i).
//
// Quadrature: the fixed rule that is quad_vec's answer whenever it converges after its first
// bisection (tables.h): the rational terms on scipy's 2 x 21 Kronrod nodes, polynomial terms
// by any exact rule (7-point Gauss-Legendre for Lorenz 96, the 42 nodes for DW / OU / L63).
// Where that rule is NOT converged -- quad_vec would bisect further there, and no reference
// integrand exists whose bisection order could be mirrored -- the rational terms are integrated
// by a composite GK21 rule on kFinePanels panels instead (exact to rounding for these
// integrands) and MFVGP_ST_REFINED is raised: the result then agrees with a quad_vec of the
// same integrand to quad_vec's own accuracy, not bit for bit.
// These are called once per outer (hyper-parameter) step, not per SCG iteration, so they are
// separate kernels: the E_cost kernels carry none of this.
#pragma once

#include "kernels.cuh"
#include "refine.cuh"

namespace mfvgp {

constexpr int kFinePanels = 16;
// the fixed rule counts as converged while scipy's error heuristic (200 |K - G|)^1.5 stays
// below ~1e-9 of the integral of |f|:  200 |K - G| <= 1e-6 sum v |f|
constexpr double kParamGuard = 1.0e-6 / 200.0;

struct EparamArgs {
    const double *x;           // [D*Nm + D*Ns]: means, then LOG-variances (free_energy.py:678-684)
    const double *obs_times, *sigma, *theta;
    double *part_th;           // [(n_end-n_begin) * rows_th][kMaxTheta] partial sums of dEsde/dtheta
    double *part_sig;          // [(n_end-n_begin)][D]: dEsde/dSigma_i of interval n
    int32_t *status;           // MFVGP_ST_REFINED is OR-ed in when a finer rule had to be used
    int D, Nm, Ns, n_begin, n_end, ntiles;
};

constexpr int kMaxTheta = 4;

// ---- d(sum_i E_i / Sigma_i)/dtheta_k at one node, DW / OU / L63 ------------------------------
template <int SYS>
__device__ __forceinline__ void drift_dtheta(const double *m, const double *dm, const double *s,
                                             const double *ds, const double *Sig,
                                             const double *th, const double *inv_sig,
                                             double *out /*[NTHETA]*/) {
    if (SYS == 0) {                       // DW: f = 4 x (theta - x^2), double_well.py:82-83
        const double t = th[0], m1 = m[0], s1 = s[0], m2 = m1 * m1;
        const double q = ds[0] - Sig[0];
        // E = E[f^2] - 2 E[f] dm + dm^2 + q^2/4s - q E[f'];  d/dtheta of each moment
        const double dEf2 = 16.0 * (2.0 * t * (m2 + s1)
                                    - 2.0 * (m2 * m2 + 6.0 * m2 * s1 + 3.0 * s1 * s1));
        out[0] = inv_sig[0] * (dEf2 - 8.0 * m1 * dm[0] - 4.0 * q);
    } else if (SYS == 1) {                // OU: f = theta (mu - x), ornstein_uhlenbeck.py:77
        const double t = th[0], mu = th[1];
        const double q = ds[0] - Sig[0];
        const double R = t * (mu - m[0]) - dm[0];
        out[0] = inv_sig[0] * (2.0 * R * (mu - m[0]) + 2.0 * t * s[0] + q);
        out[1] = inv_sig[0] * (2.0 * R * t);
    } else {                              // L63: theta = (sigma, rho, beta), lorenz_63.py:30-32
        const double a = th[0], rho = th[1], b = th[2];
        const double rz = rho - m[2];
        const double R0 = a * (m[1] - m[0]) - dm[0];
        const double R1 = m[0] * rz - m[1] - dm[1];
        const double R2 = m[0] * m[1] - b * m[2] - dm[2];
        out[0] = inv_sig[0] * (2.0 * R0 * (m[1] - m[0]) + 2.0 * a * (s[0] + s[1])
                               + (ds[0] - Sig[0]));
        out[1] = inv_sig[1] * (2.0 * R1 * m[0] + 2.0 * rz * s[0]);
        out[2] = inv_sig[2] * (-2.0 * R2 * m[2] + 2.0 * b * s[2] + (ds[2] - Sig[2]));
    }
}

// cubic / quadratic Lagrange bases on knots k/3, k/2 at local coordinate tau, d/dtau beside
__device__ __forceinline__ void lagrange_bases(double t, double *B3, double *D3, double *B2,
                                               double *D2) {
    const double a = t, b = t - 1.0 / 3.0, c = t - 2.0 / 3.0, d = t - 1.0;
    B3[0] = -4.5 * b * c * d;  D3[0] = -4.5 * (c * d + b * d + b * c);
    B3[1] = 13.5 * a * c * d;  D3[1] = 13.5 * (c * d + a * d + a * c);
    B3[2] = -13.5 * a * b * d; D3[2] = -13.5 * (b * d + a * d + a * b);
    B3[3] = 4.5 * a * b * c;   D3[3] = 4.5 * (b * c + a * c + a * b);
    const double e = t - 0.5;
    B2[0] = 2.0 * e * d;   D2[0] = 2.0 * (d + e);
    B2[1] = -4.0 * a * d;  D2[1] = -4.0 * (d + a);
    B2[2] = 2.0 * a * e;   D2[2] = 2.0 * (e + a);
}

// DW / OU / L63: one CTA per interval.  A composite GK21 rule on P panels is evaluated 42
// nodes (two panels) at a time, one thread per node, node values summed in scipy's order by
// one thread per quantity: P = 2 is the fixed rule; the embedded Gauss-10 sums of that pass
// decide whether the P = kFinePanels rule replaces it.
template <int SYS>
__global__ void __launch_bounds__(64)
eparam_small_kernel(const EparamArgs a) {
    using Dr = Drift<SYS>;
    constexpr int D = Dr::D, NTH = Dr::NTHETA;
    constexpr int NV = NTH + 2 * D;                     // dtheta_k | E_i | dE_i/dSigma_i
    __shared__ double vals[NV][kNodes];
    __shared__ double res[2][NV];                       // integrals over [0,1]: fixed, fine
    __shared__ int sh_fine;
    const int n = a.n_begin + blockIdx.x, q = threadIdx.x;
    const double dt = fabs(a.obs_times[n + 1] - a.obs_times[n]);
    const double inv_dt = 1.0 / dt;
    const double *lvars = a.x + (int64_t)a.D * a.Nm;
    double P[D][4], S[D][3], Sig[D], inv_sig[D], th[NTH];
#pragma unroll
    for (int i = 0; i < NTH; ++i) th[i] = a.theta[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double *mp = a.x + (int64_t)j * a.Nm + 3 * (int64_t)n;
        const double *sp = lvars + (int64_t)j * a.Ns + 2 * (int64_t)n;
#pragma unroll
        for (int k = 0; k < 4; ++k) P[j][k] = mp[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) S[j][k] = exp(sp[k]);
        Sig[j] = a.sigma[j];
        inv_sig[j] = 1.0 / Sig[j];
    }
    if (q == 0) sh_fine = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int np = pass == 0 ? 2 : kFinePanels;
        const double hw = 0.5 / np;                     // panel half width in tau
        double acc = 0.0, accG = 0.0, accA = 0.0;       // thread q < NV: Kronrod, Gauss, |f| sums
        for (int p0 = 0; p0 < np; p0 += 2) {
            __syncthreads();
            if (q < kNodes) {
                const int p = p0 + q / kNodesPerPanel, r = q % kNodesPerPanel;
                const double tau = fma(hw, c_gk_x[r], (2 * p + 1) * hw);
                double B3[4], D3[4], B2[3], D2[3];
                lagrange_bases(tau, B3, D3, B2, D2);
                double m[D], dm[D], sv[D], ds[D];
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    m[j] = P[j][0] * B3[0] + P[j][1] * B3[1] + P[j][2] * B3[2] + P[j][3] * B3[3];
                    dm[j] = (P[j][0] * D3[0] + P[j][1] * D3[1] + P[j][2] * D3[2] + P[j][3] * D3[3])
                            * inv_dt;
                    sv[j] = S[j][0] * B2[0] + S[j][1] * B2[1] + S[j][2] * B2[2];
                    ds[j] = (S[j][0] * D2[0] + S[j][1] * D2[1] + S[j][2] * D2[2]) * inv_dt;
                }
                double E[D], Em[D][D], Edm[D], Es[D][D], Eds[D], dth[NTH];
                Dr::nodal(m, dm, sv, ds, Sig, th, E, Em, Edm, Es, Eds);
                drift_dtheta<SYS>(m, dm, sv, ds, Sig, th, inv_sig, dth);
#pragma unroll
                for (int k = 0; k < NTH; ++k) vals[k][q] = dth[k];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    vals[NTH + i][q] = E[i];
                    vals[NTH + D + i][q] = -Eds[i];          // dE_i/dSigma_i
                }
            }
            __syncthreads();
            if (q < NV)
                for (int r = 0; r < kNodes; ++r) {
                    const int rr = r % kNodesPerPanel;
                    const double f = vals[q][r];
                    acc = fma(c_gk_v[rr], f, acc);
                    accA = fma(c_gk_v[rr], fabs(f), accA);
                    if (rr & 1) accG = fma(c_gk_w[rr >> 1], f, accG);
                }
        }
        if (q < NV) {
            res[pass][q] = hw * acc;
            if (pass == 0 && !(fabs(acc - accG) <= kParamGuard * accA)) atomicOr(&sh_fine, 1);
        }
        __syncthreads();
        if (!sh_fine) break;                            // the fixed rule is converged (CTA-uniform)
    }
    const int use = sh_fine ? 1 : 0;
    if (q == 0 && use) atomicOr(a.status, 8);           // MFVGP_ST_REFINED
    if (q < NTH)
        a.part_th[(int64_t)blockIdx.x * kMaxTheta + q] = 0.5 * dt * res[use][q];
    else if (q < NTH + D) {
        const int i = q - NTH;
        const double is = inv_sig[i];
        a.part_sig[(int64_t)blockIdx.x * a.D + i] =
            0.5 * is * dt * (res[use][NTH + D + i] - is * res[use][NTH + i]);
    }
}

// Lorenz 96: one thread per (dimension, interval) cell, dim tiles with the wrap-around halo of
// esde_l96_split_kernel (lorenz_96.py:232-235).  dE_i/dtheta = 2 R_i (polynomial, 7 GL nodes);
// int E_i = coupled polynomial (7 GL nodes) + closed-form own polynomial + rational part on
// the 42 Kronrod nodes; int dE_i/dSigma_i = -(1/2) int q_i/s_i - dt.
template <int TE>
__global__ void __launch_bounds__(TE)
eparam_l96_kernel(const EparamArgs a) {
    constexpr int TD = TE - 6;
    __shared__ double2 sh_ms[kGL][TE];
    __shared__ double sh_red[TE / 32];
    const int e = threadIdx.x;
    const int tile = blockIdx.x % a.ntiles;
    const int nl = blockIdx.x / a.ntiles, n = a.n_begin + nl;
    const int d0 = tile * TD, D = a.D;
    int j = (d0 - 3 + e) % D;
    if (j < 0) j += D;
    const bool is_out = (e >= 3) && (e <= TE - 4) && (d0 + e - 3 < D);
    const double dt = fabs(a.obs_times[n + 1] - a.obs_times[n]);
    const double inv_dt = 1.0 / dt;
    const double *mp = a.x + (int64_t)j * a.Nm + 3 * (int64_t)n;
    const double *sp = a.x + (int64_t)D * a.Nm + (int64_t)j * a.Ns + 2 * (int64_t)n;
    const double P0 = mp[0], P1 = mp[1], P2 = mp[2], P3 = mp[3];
    const double S0 = exp(sp[0]), S1 = exp(sp[1]), S2 = exp(sp[2]);
    const double Sig = a.sigma[j], inv_sig = 1.0 / Sig, theta = a.theta[0];

    double own[kGL];                                     // theta - m_i - dm_i at the GL nodes
#pragma unroll
    for (int g = 0; g < kGL; ++g) {
        const double *T = c_gl + g * kTabStride;
        const double m = P0 * T[T_B3] + P1 * T[T_B3 + 1] + P2 * T[T_B3 + 2] + P3 * T[T_B3 + 3];
        const double dm = (P0 * T[T_DB3] + P1 * T[T_DB3 + 1] + P2 * T[T_DB3 + 2]
                           + P3 * T[T_DB3 + 3]) * inv_dt;
        const double s = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
        sh_ms[g][e] = make_double2(m, s);
        own[g] = (theta - m) - dm;
    }
    __syncthreads();
    double dth = 0.0;
    if (is_out) {
        double accR = 0.0, accE = 0.0;
#pragma unroll
        for (int g = 0; g < kGL; ++g) {
            const double2 xa = sh_ms[g][e + 1], xb = sh_ms[g][e - 1], xc = sh_ms[g][e - 2];
            const double u = xa.x - xc.x, su = xa.y + xc.y, mb = xb.x, sb = xb.y;
            const double R = fma(u, mb, own[g]);                 // E[f_i] - dm_i
            const double Ec = fma(R, R, fma(sb * u, u, fma(mb * su, mb, su * sb)));
            const double w = c_gl[g * kTabStride + T_V];
            accR = fma(w, R, accR);
            accE = fma(w, Ec, accE);
        }
        // rational part: sum_42 v q^2/s and sum_42 v q/s (tau-polynomials of s and q)
        const double a0 = S0, a1 = -3.0 * S0 + 4.0 * S1 - S2, a2 = 2.0 * (S0 - 2.0 * S1 + S2);
        const double b0 = a1 * inv_dt - Sig, b1 = 2.0 * a2 * inv_dt;
        // RE = int_0^1 q^2/s dtau, RL = int_0^1 q/s dtau: fixed rule (2 panels), and the
        // composite rule on kFinePanels panels where its Kronrod - Gauss estimate is too large
        double RE = 0.0, RL = 0.0;
        bool fine = false;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const int np = pass == 0 ? 2 : kFinePanels;
            const double hw = 0.5 / np;
            double kE = 0.0, kL = 0.0, gE = 0.0, gL = 0.0, aL = 0.0;
#pragma unroll 1
            for (int p = 0; p < np; ++p)
#pragma unroll 1
                for (int r = 0; r < kNodesPerPanel; ++r) {
                    const double tau = fma(hw, c_gk_x[r], (2 * p + 1) * hw), v = c_gk_v[r];
                    const double s = fma(fma(a2, tau, a1), tau, a0), qq = fma(b1, tau, b0);
                    const double qi = qq / s;
                    kE = fma(v, qq * qi, kE);
                    kL = fma(v, qi, kL);
                    aL = fma(v, fabs(qi), aL);
                    if (r & 1) {
                        gE = fma(c_gk_w[r >> 1], qq * qi, gE);
                        gL = fma(c_gk_w[r >> 1], qi, gL);
                    }
                }
            RE = hw * kE;
            RL = hw * kL;
            if (pass == 0)
                fine = !(fabs(kE - gE) <= kParamGuard * fabs(kE) && fabs(kL - gL) <= kParamGuard * aL);
            if (!fine) break;
        }
        if (fine) atomicOr(a.status, 8);                 // MFVGP_ST_REFINED
        const double own_poly = dt * (S0 + 4.0 * S1 + S2) * (1.0 / 6.0) + (S2 - S0) - Sig * dt;
        const double I_E = fma(dt, accE, fma(0.25 * dt, RE, own_poly));      // int_n E_i dt
        const double I_dS = -fma(0.5 * dt, RL, dt);                          // int_n dE_i/dSigma_i dt
        a.part_sig[(int64_t)nl * D + j] = 0.5 * inv_sig * (I_dS - inv_sig * I_E);
        dth = inv_sig * dt * accR;                       // 1/2 (1/Sigma_i) int 2 R_i dt
    }
    // deterministic CTA sum of the theta partials
    dth = warp_sum(dth);
    if ((e & 31) == 0) sh_red[e >> 5] = dth;
    __syncthreads();
    if (e == 0) {
        double v = sh_red[0];
        for (int w = 1; w < TE / 32; ++w) v += sh_red[w];
        a.part_th[(int64_t)blockIdx.x * kMaxTheta] = v;
    }
}

// Fixed-order sums of the partials: CTAs 0 .. ceil(D/256)-1 take the Sigma entries (one thread
// per state dimension, intervals in order), the last CTA the theta entries.
__global__ void __launch_bounds__(256)
eparam_reduce_kernel(const double *part_th, int rows_th, int nth, const double *part_sig, int nl,
                     int D, double *out /*[nth + D]*/, int32_t *status) {
    __shared__ double sh[256];
    __shared__ int sh_bad;
    const int nb_sig = (D + 255) / 256;
    if (threadIdx.x == 0) sh_bad = 0;
    __syncthreads();
    int bad = 0;
    if ((int)blockIdx.x < nb_sig) {
        const int j = blockIdx.x * 256 + threadIdx.x;
        if (j < D) {
            double v = 0.0;
            for (int n = 0; n < nl; ++n) v += part_sig[(int64_t)n * D + j];
            out[nth + j] = v;
            bad = !isfinite(v);
        }
    } else {
        for (int k = 0; k < nth; ++k) {
            double v = 0.0;
            for (int r = threadIdx.x; r < rows_th; r += 256) v += part_th[(int64_t)r * kMaxTheta + k];
            sh[threadIdx.x] = v;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                out[k] = sh[0];
                bad |= !isfinite(sh[0]);
            }
            __syncthreads();
        }
    }
    if (bad) atomicOr(&sh_bad, 1);
    __syncthreads();
    if (threadIdx.x == 0 && sh_bad) atomicOr(status, 1);     // MFVGP_ST_ESDE_NONFINITE
}

// multi-GPU: gathered [world][n] summed in rank order (every rank ends with identical bits)
__global__ void __launch_bounds__(256)
eparam_combine_kernel(const double *gathered, int world, int n, double *out, int32_t *status) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        double v = gathered[i];
        for (int r = 1; r < world; ++r) v += gathered[(int64_t)r * n + i];
        out[i] = v;
        if (!isfinite(v)) atomicOr(status, 1);
    }
}

}  // namespace mfvgp
