// Lorenz 96, small problems (the north-star size D=500, M=16 has 7 500 cells --
// 51 per SM): latency, not throughput, decides.  Here TWO lanes share one
// (dimension, interval) cell of FreeEnergy.single_interval
// (src/var_bayes/free_energy.py:338-459; drift lorenz_96.py:187-381):
//   (A) the rational part: lane p integrates Kronrod panel p (10 symmetric node pairs +
//       centre, moment form of l96_walk.cuh), one exchange of 7 values at the end;
//   (B) the coupled polynomial part: lane 0 takes Gauss-Legendre nodes 0-3, lane 1 nodes
//       4-6; all seven nodes travel through shared memory at once (two barriers per cell
//       instead of 14); lane 0 ends with dEsde_dm, lane 1 with dEsde_ds and the energy.
// Same [L][D][4] / [L][D][3] outputs and guard partials as esde_l96_split_kernel.
// The last CTA of an interval (atomic ticket) folds the tile partials and applies the guard
// (guard_interval, kernels.cuh): no guard launch.  Flagged intervals are counted and handed
// to refine_kernel, which follows in the stream and exits at once when there are none.
// (A first version with 8 lanes per cell was shuffle-bound: 19 us; see profiles/.)
#pragma once

#include "l96_walk.cuh"
#include "refine.cuh"

namespace mfvgp {

constexpr int kCoopTE = 32;                 // extended dims per CTA (26 outputs + halo)
constexpr int kCoopW = 2;                   // lanes per cell
constexpr int kCoopNT = kCoopTE * kCoopW;   // 64 threads

struct CoopArgs {
    EsdeArgs a;
    GuardArgs g;              // g.nflagged counts the flagged intervals for refine_kernel
    unsigned *tickets;        // [n_end - n_begin], zero between launches
    double theta;
    int guard_later;          // optimistic evaluations: the tail kernel folds the tiles and applies
                              // the guard in CTAs of its own (no ticket, no fold here)
};

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

__device__ __forceinline__ void esde_l96_coop_body(const CoopArgs &ca) {
    constexpr int TE = kCoopTE, TD = TE - 6;
    const EsdeArgs &a = ca.a;
    __shared__ double2 ms[kGL][TE], AC[kGL][TE], BD[kGL][TE];
    __shared__ double red[9][TE + 1];
    __shared__ int sh_last;

    const int tid = threadIdx.x, l = tid & 1, e = tid >> 1;
    const int tile = blockIdx.x, cta = blockIdx.y * gridDim.x + blockIdx.x;   // grid = (tiles, intervals)
    const int n = a.n_begin + blockIdx.y;
    const int d0 = tile * TD, D = a.D;
    int j = d0 - 3 + e;                  // circular index (lorenz_96.py:232-235), no division
    if (D >= TE) {
        if (j < 0) j += D;
        if (j >= D) j -= D;
    } else {
        j %= D;
        if (j < 0) j += D;
    }
    const bool has_energy = (e >= 2) && (e <= TE - 2);
    const bool is_out = (e >= 3) && (e <= TE - 4) && (d0 + e - 3 < D);

    // problem data never written by a kernel: load it (and warm the constant tables) while the
    // predecessor is still running
    const double ti = a.obs_times[n], tj = a.obs_times[n + 1];
    const double Sig = a.sigma[j];
    const double dt = fabs(tj - ti);
    const double inv_dt = fast_rcp(dt), inv_sig = fast_rcp(Sig), hid = 0.5 * inv_dt;
    double warm = c_pair[(tid * 8) % (kPairs * kPairStride)] + c_gl[(tid * 8) % (kGL * kTabStride)];
    pdl_wait();                     // x and the workspaces belong to the predecessor until here
    pdl_launch_dependents();        // only now: the tail kernel reads x before ITS wait
    const double *mp = a.mean + (int64_t)j * a.ldm + 3 * (int64_t)n;
    const double *sp = a.vars + (int64_t)j * a.lds + 2 * (int64_t)n;
    const double P0 = mp[0], P1 = mp[1], P2 = mp[2], P3 = mp[3];
    double S0 = sp[0], S1 = sp[1], S2 = sp[2];
    if (a.log_vars) { S0 = exp(S0); S1 = exp(S1); S2 = exp(S2); }
    if (warm == 1.2345e300) S0 = warm;          // never true: keeps the warm-up loads alive

    // ---- (A) rational part: lane l integrates Kronrod panel l ------------------------------
    // (every cell of the warp takes part: the shuffles need all 32 lanes; halo cells work
    // on their own valid, wrapped rows and their results are not used)
    double aE, kS[3], eE[2], eS[2], m1S[2];
    {
        const double c = 0.25 + 0.5 * l;
        const double a0 = S0, a1 = -3.0 * S0 + 4.0 * S1 - S2, a2 = 2.0 * (S0 - 2.0 * S1 + S2);
        const double b0 = fma(a1, inv_dt, -Sig), b1 = 2.0 * a2 * inv_dt;
        PanelSums s;
        panel_sums(fma(fma(a2, c, a1), c, a0), 0.25 * fma(2.0 * c, a2, a1), 0.0625 * a2,
                   fma(b1, c, b0), 0.25 * b1, s);
        // panel -> basis sums (l96_walk.cuh sv3/sd3, written for a run-time panel centre c)
        const double bc[3] = {fma(2.0 * c, c, fma(-3.0, c, 1.0)), 4.0 * c * (1.0 - c),
                              c * (2.0 * c - 1.0)};
        const double b1c[3] = {4.0 * c - 3.0, 4.0 - 8.0 * c, 4.0 * c - 1.0};
        const double bpp[3] = {4.0, -8.0, 4.0};
        double ks_p[3], es_p = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double sv = fma(bc[k], s.N0, fma(0.25 * b1c[k], s.N1, 0.03125 * bpp[k] * s.N2));
            const double sd = fma(b1c[k], s.L0, 0.25 * bpp[k] * s.L1);
            ks_p[k] = fma(hid, sd, -0.25 * sv);
            const double dv = fma(bc[k], s.N0 - s.GN0, fma(0.25 * b1c[k], s.N1 - s.GN1,
                                                          0.03125 * bpp[k] * (s.N2 - s.GN2)));
            const double dd = fma(b1c[k], s.L0 - s.GL0, 0.25 * bpp[k] * (s.L1 - s.GL1));
            const double kg = fma(hid, dd, -0.25 * dv);
            es_p = fma(kg, kg, es_p);
        }
        const double dE = 0.25 * (s.kE - s.gE);
        const double mv = fma(bc[0], s.N1, fma(0.25 * b1c[0], s.N2, 0.125 * s.N3));
        const double md = fma(b1c[0], s.L1, s.L2);
        const double m1 = fma(hid, md, -0.25 * mv)
                          + fma(c_m1poly[2 * l + 1], inv_dt, c_m1poly[2 * l]);
        const double eE_p = dE * dE, m1_p = m1 * m1;
        // the other panel's results; panel 0 first in every sum
        const double o_kE = shfl_xor_d(s.kE, 1), o_eE = shfl_xor_d(eE_p, 1),
                     o_eS = shfl_xor_d(es_p, 1), o_m1 = shfl_xor_d(m1_p, 1);
        aE = l ? o_kE + s.kE : s.kE + o_kE;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double o = shfl_xor_d(ks_p[k], 1);
            kS[k] = l ? o + ks_p[k] : ks_p[k] + o;
        }
        eE[0] = l ? o_eE : eE_p; eE[1] = l ? eE_p : o_eE;
        eS[0] = l ? o_eS : es_p; eS[1] = l ? es_p : o_eS;
        m1S[0] = l ? o_m1 : m1_p; m1S[1] = l ? m1_p : o_m1;
    }

    // ---- (B) coupled polynomial part: lane 0 takes GL nodes 0-3, lane 1 nodes 4-6 ---------------
    // all seven nodes travel through shared memory at once: two barriers per cell
    const int g_lo = 4 * l, g_n = l ? 3 : 4;
    double own[4];
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
        own[gg] = 0.0;
        if (gg < g_n) {
            const double *T = c_gl + (g_lo + gg) * kTabStride;
            const double m = P0 * T[T_B3] + P1 * T[T_B3 + 1] + P2 * T[T_B3 + 2] + P3 * T[T_B3 + 3];
            const double dm = (P0 * T[T_DB3] + P1 * T[T_DB3 + 1] + P2 * T[T_DB3 + 2]
                               + P3 * T[T_DB3 + 3]) * inv_dt;
            const double s = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
            ms[g_lo + gg][e] = make_double2(m, s);
            own[gg] = (ca.theta - m) - dm;
        }
    }
    __syncthreads();
    double accE = 0.0;
    if (has_energy) {
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
            if (gg < g_n) {
                const int g = g_lo + gg;
                const double *T = c_gl + g * kTabStride;
                const double2 xa = ms[g][e + 1], xb = ms[g][e - 1], xc = ms[g][e - 2];
                const double mb = xb.x, sb = xb.y;
                const double u = xa.x - xc.x, su = xa.y + xc.y;
                const double R = fma(u, mb, own[gg]);
                const double t1 = sb * u, t2 = mb * su;
                const double Ec = fma(R, R, fma(t1, u, fma(t2, mb, su * sb)));
                const double w = T[T_V] * inv_sig, w2 = w + w;
                const double wR2 = w2 * R;
                AC[g][e] = make_double2(fma(wR2, mb, w2 * t1), w * fma(mb, mb, sb));
                BD[g][e] = make_double2(fma(wR2, u, w2 * t2), w * fma(u, u, su));
                own[gg] = -wR2;
                accE = fma(w, Ec, accE);
            }
        }
    }
    __syncthreads();
    double accM[4] = {0.0, 0.0, 0.0, 0.0}, accS[3] = {0.0, 0.0, 0.0};
    if (is_out) {
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
            if (gg < g_n) {
                const int g = g_lo + gg;
                const double *T = c_gl + g * kTabStride;
                const double2 lft = AC[g][e - 1], r = BD[g][e + 1], rr = AC[g][e + 2];
                const double Gm = own[gg] + lft.x + r.x - rr.x;
                const double Gs = lft.y + r.y + rr.y;
                const double Hm = own[gg] * inv_dt;
#pragma unroll
                for (int k = 0; k < 4; ++k) accM[k] += Gm * T[T_B3 + k] + Hm * T[T_DB3 + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) accS[k] = fma(Gs, T[T_B2 + k], accS[k]);
            }
        }
    }
    // lane 0 ends with the dM sums, lane 1 with the dS sums and the energy (lane 0's part first)
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double mine = l ? (k < 3 ? accS[k] : accE) : accM[k];
        const double send = l ? accM[k] : (k < 3 ? accS[k] : accE);
        const double got = shfl_xor_d(send, 1);
        v[k] = l ? got + mine : mine + got;
    }

    // ---- outputs (same algebra as esde_l96_split_kernel) ------------------------------------------
    const double four_inv_dt = 4.0 * inv_dt;
    const double polyS[3] = {dt * (1.0 / 6.0) - 1.0, dt * (4.0 / 6.0), dt * (1.0 / 6.0) + 1.0};
    if (is_out && a.dm_out != nullptr) {
        const double sc_gl = 0.5 * dt;
        if (l == 0) {
            double *om = a.dm_out + ((int64_t)n * D + j) * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) om[k] = sc_gl * v[k];
        } else {
            double *os = a.ds_out + ((int64_t)n * D + j) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                os[k] = fma(sc_gl, v[k], fma(0.125 * dt * inv_sig, kS[k], 0.5 * inv_sig * polyS[k]));
        }
    }
    if (l == 1) {                                  // this cell's guard partials
        double r9[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (is_out) {
            const double own_poly = dt * (S0 + 4.0 * S1 + S2) * (1.0 / 6.0) + (S2 - S0) - Sig * dt;
            const double ksum_E = 0.25 * aE + 4.0 * v[3] * Sig + four_inv_dt * own_poly;
            double own2 = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double o = fma(four_inv_dt, polyS[k], kS[k]);
                own2 = fma(o, o, own2);
            }
            r9[0] = ksum_E * inv_sig; r9[1] = ksum_E * ksum_E;
            r9[2] = eE[0]; r9[3] = eE[1]; r9[4] = eS[0]; r9[5] = eS[1];
            r9[6] = own2; r9[7] = m1S[0]; r9[8] = m1S[1];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) red[i][e] = r9[i];
    }
    __syncthreads();
    if (tid < 9) {                                 // fixed-order fold over the TE cells
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;     // four chains, fixed order
#pragma unroll
        for (int c = 0; c < TE; c += 4) {
            x0 += red[tid][c]; x1 += red[tid][c + 1]; x2 += red[tid][c + 2]; x3 += red[tid][c + 3];
        }
        const double x = (x0 + x1) + (x2 + x3);
        a.part[(int64_t)cta * kPartStride + tid] = x;
        if (!ca.guard_later) __threadfence();      // the writers publish their partials
    }
    if (ca.guard_later) return;
    // ---- last CTA of this interval: fold the tiles, apply the guard ---------------------------------
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(ca.tickets + (n - a.n_begin), 1u);
        sh_last = (t == (unsigned)a.ntiles - 1u);
        if (sh_last) ca.tickets[n - a.n_begin] = 0u;
    }
    __syncthreads();
    if (sh_last && tid < 32) {
        // quad_vec keeps bisecting where the fixed rule is not provably its answer: the guard
        // flags the interval and counts it; refine_kernel (next in the stream, exits at once
        // when the count is zero) takes it from there
        __threadfence();
        guard_interval(ca.g, n, tid);
    }
}

__global__ void __launch_bounds__(kCoopNT)
esde_l96_coop_kernel(const CoopArgs ca) { esde_l96_coop_body(ca); }

// Batched evaluation (mfvgp_ecost_batch): blockIdx.z selects one of B independent states of the
// SAME problem -- B x 300 CTAs instead of 300 at the north-star size, which is what it takes to
// fill 148 SMs.  Always guard_later: the batched tail kernel folds the tiles.
__global__ void __launch_bounds__(kCoopNT)
esde_l96_coop_batch_kernel(const CoopArgs ca0, const BatchStrides bs) {
    CoopArgs ca = ca0;
    const int64_t z = blockIdx.z;
    ca.a.mean += z * bs.x; ca.a.vars += z * bs.x;
    if (ca.a.dm_out != nullptr) { ca.a.dm_out += z * bs.dm; ca.a.ds_out += z * bs.ds; }
    ca.a.part += z * bs.part;
    esde_l96_coop_body(ca);
}

// ---------------------------------------------------------------------------
// DW / OU / L63: esde_small_kernel with the guard folded in (one CTA per interval owns the
// whole interval, so no ticket is needed); flagged intervals go to refine_kernel.
// ---------------------------------------------------------------------------
template <int SYS>
__global__ void __launch_bounds__(64)
esde_small_inkernel(const CoopArgs ca) {
    const EsdeArgs &a = ca.a;
    const int n = a.n_begin + blockIdx.x, tid = threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();        // only now: the successors read x before THEIR wait
    esde_small_body<SYS>(a);
    __syncthreads();
    if (tid < 32) guard_interval(ca.g, n, tid);
}

// batched (mfvgp_ecost_batch): blockIdx.y selects the state; the guard runs in the kernel as
// above, on that state's partials, flags and flag counter
template <int SYS>
__global__ void __launch_bounds__(64)
esde_small_batch_kernel(const CoopArgs ca0, const BatchStrides bs, const int nl) {
    CoopArgs ca = ca0;
    const int64_t z = blockIdx.y;
    ca.a.mean += z * bs.x; ca.a.vars += z * bs.x;
    if (ca.a.dm_out != nullptr) { ca.a.dm_out += z * bs.dm; ca.a.ds_out += z * bs.ds; }
    ca.a.part += z * bs.part;
    ca.g.part += z * bs.part; ca.g.esde_n += z * bs.esde_n; ca.g.flags += z * (int64_t)nl;
    ca.g.nflagged += 2 * z;
    const EsdeArgs &a = ca.a;
    const int n = a.n_begin + blockIdx.x, tid = threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    esde_small_body<SYS>(a);
    __syncthreads();
    if (tid < 32) guard_interval(ca.g, n, tid);
}

// ---------------------------------------------------------------------------
// E_cost tail for small problems (single GPU): gradient assembly
// (free_energy.py:711-743), E_kl0 / E_obs terms (:309-310, :594-612, :733-738) and
// -- in the last CTA to finish (atomic ticket) -- the final scalar sums of
// final_kernel, in ONE launch.  One thread per entry of x; every load of a thread
// is independent of the others (row tables rinv_row / y_dense instead of the
// obs_idx -> obs_values chain), so the kernel is one memory round trip deep.
// ---------------------------------------------------------------------------
struct TailArgs {
    AssembleArgs a;
    FinalArgs f;
    const double *y_dense;     // [M][D] observation of state dimension j at time k (0 if unobserved)
    const double *rinv_row;    // [D] 1 / obs_noise of dimension j, 0 if unobserved
    unsigned *ticket;          // [1], zero between launches
    GuardArgs g;               // guard_ctas > 0: the last guard_ctas CTAs of the grid fold the
    int guard_ctas;            // main kernel's tile partials and apply the guard, a warp per
                               // interval, beside the CTAs that assemble the gradient
    const int32_t *run_if = nullptr;   // second pass of a two-pass evaluation: return at once
                                       // unless *run_if != 0 (FinalArgs::late)
};

constexpr int kTailThreads = 256;

__device__ __forceinline__ void assemble_tail_body(const TailArgs &t) {
    __shared__ double sh_red[3 * (kTailThreads / 32)];
    __shared__ int sh_last;
    const AssembleArgs &a = t.a;
    // small problems only: the element count fits 32 bits (mfvgp_create checks)
    const unsigned idx0 = blockIdx.x * kTailThreads + threadIdx.x;
    // one entry per thread (single evaluations: one memory round trip deep); batches whose main
    // kernel already rounds differently launch a quarter of the CTAs and stride over the entries
    const unsigned idx_step = (gridDim.x - (unsigned)t.guard_ctas) * kTailThreads;
    const unsigned nm = (unsigned)a.D * a.Nm, ntot = nm + (unsigned)a.D * a.Ns;
    const bool guard_cta = blockIdx.x >= gridDim.x - (unsigned)t.guard_ctas;
    double acc[3] = {0.0, 0.0, 0.0};
    pdl_launch_dependents();
    if (t.run_if != nullptr) {
        pdl_wait();
        if (__ldcg(t.run_if) == 0) return;       // nothing was flagged: the first pass is the result
    }
    if (idx0 >= ntot) pdl_wait();
    if (guard_cta) {
        const int nwarps = t.guard_ctas * (kTailThreads / 32);
        const int w = (int)(blockIdx.x - (gridDim.x - (unsigned)t.guard_ctas)) * (kTailThreads / 32)
                      + (int)(threadIdx.x >> 5);
        for (int n = t.g.n_begin + w; n < t.g.n_end; n += nwarps)
            guard_interval(t.g, n, (int)(threadIdx.x & 31));
    } else for (unsigned idx = idx0; idx < ntot; idx += idx_step) {
        const bool is_mean = idx < nm;
        const unsigned r = is_mean ? idx : idx - nm;
        const unsigned W = is_mean ? a.Nm : a.Ns;
        const int step = is_mean ? 3 : 2, K = step + 1;
        const int j = (int)(r / W), c = (int)(r - (unsigned)j * W);
        const int n = is_mean ? c / 3 : c >> 1, k = c - step * n;
        const double *ws = is_mean ? a.dm : a.ds;
        const bool obs_col = k == 0 && n >= 1 && n <= a.M;
        // inputs that do not depend on the preceding kernel first: their latency hides under
        // that kernel's tail (programmatic dependent launch); x itself was final before the
        // preceding kernel started
        const double xv = a.x[idx];
        double own = 0.0, prev = 0.0, y = 0.0, Ri = 0.0, mu = 0.0, t0 = 1.0;
        if (obs_col) {
            Ri = t.rinv_row[j];
            y = t.y_dense[(int64_t)(n - 1) * a.D + j];
        }
        if (c == 0) { mu = a.mu0[j]; t0 = a.tau0[j]; }
        const double s = is_mean ? 0.0 : exp(xv);
        pdl_wait();                                   // the per-interval gradients are ready
        if (ws != nullptr) {
            if (n < a.L) own = ws[((int64_t)n * a.D + j) * K + k];
            if (k == 0 && n >= 1 && n - 1 < a.L) prev = ws[((int64_t)(n - 1) * a.D + j) * K + step];
        }
        double g = own;
        g += prev;
        if (is_mean) {
            if (c == 0) {
                const double z0 = xv - mu;
                g += z0 / t0;
                acc[1] += z0 * z0 / t0;
            }
            if (obs_col && Ri != 0.0) {
                const double rr = y - xv;
                g -= Ri * rr;
                acc[2] += Ri * rr * rr;
            }
            if (a.grad) a.grad[idx] = g;
        } else {
            if (c == 0) {
                g += 0.5 * (1.0 / t0 - 1.0 / s);
                acc[0] += log(t0 / s);
                acc[1] += (s - t0) / t0;
            }
            if (obs_col && Ri != 0.0) {
                g += 0.5 * Ri;
                acc[2] += Ri * s;
            }
            if (a.grad) a.grad[idx] = g * s;             // free_energy.py:743
        }
    }
    block_sum<3, kTailThreads>(acc, sh_red);
    if (threadIdx.x == 0) {
        if (!guard_cta) {
            double *p = a.apart + (int64_t)blockIdx.x * 4;
            p[0] = acc[0]; p[1] = acc[1]; p[2] = acc[2]; p[3] = 0.0;
        }
        __threadfence();
        const unsigned k = atomicAdd(t.ticket, 1u);
        sh_last = (k == gridDim.x - 1u);
        if (sh_last) t.ticket[0] = 0u;
    }
    __syncthreads();
    if (sh_last) {
        __threadfence();
        final_body<kTailThreads>(t.f);
    }
}

__global__ void __launch_bounds__(kTailThreads)
assemble_tail_kernel(const TailArgs t) { assemble_tail_body(t); }

// batched: blockIdx.y selects the state; every per-evaluation buffer is rebased, the problem
// data (prior, observations, tables) is shared.  status / ticket / flag counters: one per entry.
__global__ void __launch_bounds__(kTailThreads)
assemble_tail_batch_kernel(const TailArgs t0, const BatchStrides bs, const int nl) {
    TailArgs t = t0;
    const int64_t z = blockIdx.y;
    t.a.x += z * bs.x;
    if (t.a.dm != nullptr) { t.a.dm += z * bs.dm; t.a.ds += z * bs.ds; }
    if (t.a.grad != nullptr) t.a.grad += z * bs.x;
    t.a.apart += z * bs.apart; t.f.apart += z * bs.apart;
    t.ticket += z;
    t.f.esde_n += z * bs.esde_n; t.g.esde_n += z * bs.esde_n;
    t.f.flags += z * (int64_t)nl; t.g.flags += z * (int64_t)nl;
    t.f.out += z * bs.out;
    t.f.status += z; t.f.refine_status += 2 * z; t.f.nflagged += 2 * z; t.g.nflagged += 2 * z;
    t.g.part += z * bs.part;
    assemble_tail_body(t);
}

}  // namespace mfvgp
