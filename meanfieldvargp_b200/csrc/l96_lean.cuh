// Lorenz 96 E_sde of ONE interval per CTA row of dim tiles -- the throughput shape of the batched
// evaluation (mfvgp_ecost_batch, many states of the same problem per launch) -- with the
// instruction count of the walk kernel's formulation instead of esde_l96_split_body's:
//   (A) the rational part in panel-local coordinates with symmetric node pairs and moment sums
//       (rational_part, l96_walk.cuh), energy sums from the first-power sums;
//   (B) the coupled polynomial part with the rows' weights applied by the consumer (c_glw).
// Same per-interval outputs as esde_l96_split_body (kernels.cuh): dm_out / ds_out in the
// workspace layout [L][D][4] / [L][D][3] and the nine guard partials per (interval, tile), so the
// tail kernels serve both.  Replaces, like it, src/var_bayes/free_energy.py:461-565 for the
// drift of src/dynamical_systems/lorenz_96.py:187-381.  esde_l96_split_body stays as the
// independently written cross-check (MFVGP_PATH=split, MFVGP_BATCH_KERNEL=split).
#pragma once

#include "l96_walk.cuh"

namespace mfvgp {

template <int TE>
__device__ __forceinline__ void esde_l96_lean_body(const EsdeArgs &a) {
    constexpr int TD = TE - 6;
    __shared__ double2 sh_ms[TE], sh_AC[TE], sh_BD[TE];
    __shared__ double sh_isg[TE];
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

    const double dt = fabs(a.obs_times[n + 1] - a.obs_times[n]);
    const double inv_dt = 1.0 / dt;
    const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
    const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
    const double P0 = mp[0], P1 = mp[1], P2 = mp[2], P3 = mp[3];
    double S0 = sp[0], S1 = sp[1], S2 = sp[2];
    if (a.log_vars) { S0 = exp(S0); S1 = exp(S1); S2 = exp(S2); }
    const double Sig = a.sigma[j];
    const double inv_sig = 1.0 / Sig;
    const double theta = a.theta[0];
    sh_isg[e] = inv_sig;

    // ---- (A) rational part, own dimension ----------------------------------------------------
    RatOut ro;
    ro.aE = 0.0;
    ro.kS[0] = ro.kS[1] = ro.kS[2] = 0.0;
    ro.eE[0] = ro.eE[1] = ro.eS[0] = ro.eS[1] = ro.m1S[0] = ro.m1S[1] = 0.0;
    if (is_out) rational_part(S0, S1, S2, Sig, inv_dt, ro);
    __syncthreads();
    const double is_l = sh_isg[e >= 1 ? e - 1 : e], is_r = sh_isg[e + 1 < TE ? e + 1 : e],
                 is_rr = sh_isg[e + 2 < TE ? e + 2 : e];
    const double isdt = inv_sig * inv_dt;

    // ---- (B) coupled polynomial part, 7 Gauss-Legendre nodes ----------------------------------
    double accM[4] = {0, 0, 0, 0}, accS[3] = {0, 0, 0}, accE = 0.0;
#pragma unroll
    for (int g = 0; g < kGL; ++g) {
        const double *T = c_gl + g * kTabStride;
        const double *W = c_glw + g * kGlwStride;
        const double m = P0 * T[T_B3] + P1 * T[T_B3 + 1] + P2 * T[T_B3 + 2] + P3 * T[T_B3 + 3];
        const double dm = (P0 * T[T_DB3] + P1 * T[T_DB3 + 1] + P2 * T[T_DB3 + 2]
                           + P3 * T[T_DB3 + 3]) * inv_dt;
        const double s = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
        sh_ms[e] = make_double2(m, s);
        __syncthreads();
        double R = 0.0;
        if (has_energy) {
            const double2 xa = sh_ms[e + 1], xb = sh_ms[e - 1], xc = sh_ms[e - 2];
            const double mb = xb.x, sb = xb.y;
            const double u = xa.x - xc.x, su = xa.y + xc.y;
            R = fma(u, mb, (theta - m) - dm);                    // E[f_i] - dm_i
            const double t1 = sb * u, t2 = mb * su;
            const double Ec = fma(R, R, fma(t1, u, fma(t2, mb, su * sb)));
            sh_AC[e] = make_double2(fma(R, mb, t1), fma(mb, mb, sb));
            sh_BD[e] = make_double2(fma(R, u, t2), fma(u, u, su));
            accE = fma(T[T_V], Ec, accE);
        }
        __syncthreads();
        if (is_out) {
            const double2 l = sh_AC[e - 1], r = sh_BD[e + 1], rr = sh_AC[e + 2];
            const double Gm = fma(is_l, l.x, fma(is_r, r.x, fma(-is_rr, rr.x, -inv_sig * R)));
            const double Gs = fma(is_l, l.y, fma(is_r, r.y, is_rr * rr.y));
            const double Hm = -isdt * R;
#pragma unroll
            for (int k = 0; k < 4; ++k) accM[k] = fma(Gm, W[k], fma(Hm, W[4 + k], accM[k]));
#pragma unroll
            for (int k = 0; k < 3; ++k) accS[k] = fma(Gs, W[8 + k], accS[k]);
        }
    }

    // ---- per-interval results (the algebra of esde_l96_split_body's epilogue) -----------------
    const double four_inv_dt = 4.0 * inv_dt;
    const double own_poly = dt * (S0 + 4.0 * S1 + S2) * (1.0 / 6.0) + (S2 - S0) - Sig * dt;
    const double ksum_E = 0.25 * ro.aE + 4.0 * accE + four_inv_dt * own_poly;
    const double polyS[3] = {dt * (1.0 / 6.0) - 1.0, dt * (4.0 / 6.0), dt * (1.0 / 6.0) + 1.0};
    double own2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double o = fma(four_inv_dt, polyS[k], ro.kS[k]);
        own2 = fma(o, o, own2);
    }
    if (is_out && a.dm_out != nullptr) {
        double *om = a.dm_out + ((int64_t)n * D + j) * 4;
        double *os = a.ds_out + ((int64_t)n * D + j) * 3;
        const double sc_gl = 0.5 * dt, sc_gk = 0.125 * dt * inv_sig, hsig = 0.5 * inv_sig;
#pragma unroll
        for (int k = 0; k < 4; ++k) om[k] = sc_gl * accM[k];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            os[k] = fma(sc_gl, accS[k], fma(sc_gk, ro.kS[k], hsig * polyS[k]));
    }
    double red[9] = {ksum_E * inv_sig, ksum_E * ksum_E, ro.eE[0], ro.eE[1], ro.eS[0], ro.eS[1],
                     own2, ro.m1S[0], ro.m1S[1]};
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

// one of B independent states per blockIdx.y (BatchStrides, kernels.cuh)
template <int TE>
__global__ void __launch_bounds__(TE, 640 / TE)
esde_l96_lean_batch_kernel(const EsdeArgs a0, const BatchStrides bs) {
    EsdeArgs a = a0;
    const int64_t z = blockIdx.y;
    a.mean += z * bs.x; a.vars += z * bs.x;
    if (a.dm_out != nullptr) { a.dm_out += z * bs.dm; a.ds_out += z * bs.ds; }
    a.part += z * bs.part;
    esde_l96_lean_body<TE>(a);
}

}  // namespace mfvgp
