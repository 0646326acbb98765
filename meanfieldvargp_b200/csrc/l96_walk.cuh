// Lorenz 96 E_cost in one main kernel: "walk" layout.
//
// Replaces, for Lorenz 96, the reference's E_sde + gradient assembly of E_cost
// (src/var_bayes/free_energy.py:461-565 and :711-743; drift expressions
// src/dynamical_systems/lorenz_96.py:187-381).
//
// A CTA owns TE "extended" state dimensions (TE-6 output rows plus the
// wrap-around halo of 3 per side that the circular i-2/i-1/i+1 coupling needs,
// lorenz_96.py:232-235) and WALKS along a run of TN consecutive time intervals:
// thread e keeps row j's support points in registers, the end point of one
// interval is the start point of the next (free_energy.py:504-505), so
//   * every support point is loaded and exp()'d once,
//   * the loads of interval n+1 are issued before the arithmetic of interval n,
//   * the gradient is stitched in registers (:715-728: shared end points add),
//     gets dE0 (:733-734), dEobs (:737-738) and the log-variance chain rule
//     (:743) and is written straight into the gradient of x -- no per-interval
//     workspace, no assembly pass.
// Only the one column a run shares with the NEXT run needs another CTA: the next
// run publishes its part of that column (k=0 part + dEobs) through a mailbox after
// its first interval, and this run adds its own k=3 / k=2 part at its end and writes
// the column (one writer per element, fixed order: deterministic; CTAs are numbered
// by an atomic ticket so that a CTA only waits for CTAs that started before it).
//
// Quadrature = the split rule of esde_l96_split_kernel (kernels.cuh):
//   (A) rational part q^2/(4s) and its d/dS on scipy's 2 x 21 Kronrod nodes, plus
//       the embedded Gauss-10 sums (quad_vec's error estimate) -- here in
//       panel-local coordinates with symmetric node pairs and moment sums
//       (tables.h build_pair_table): 66 node constants instead of 560, ~20 % fewer
//       FP64 instructions;
//   (B) coupled polynomial part on 7 Gauss-Legendre nodes with the shared-memory
//       neighbour exchange.
#pragma once

#include "kernels.cuh"

namespace mfvgp {

__constant__ __align__(16) double c_pair[kPairs * kPairStride];
__constant__ double c_vcentre;
// part (B), weights applied by the consumer (MFVGP_WALK_DEFERW): per Gauss-Legendre node
// [0..3] 2 v B3_k, [4..7] 2 v dB3_k, [8..10] v B2_k, [11] v
constexpr int kGlwStride = 12;
__constant__ __align__(16) double c_glw[kGL * kGlwStride];

constexpr int kWalkMaxTN = 64;

// tuning knobs (tools/build_variant.sh builds variants with -D overrides)
// Gauss-Legendre nodes per pair of barriers in part (B).  Measured at D=8192, L=4096
// (tools/walk_ab.py): TE=128 -- 1: 3 237 us, 2: 3 141, 4: 3 178;  TE=64 -- 1: 3 155, 2: 3 066,
// 4: 3 029.  -DMFVGP_WALK_GB=n forces one value for every TE.
#ifdef MFVGP_WALK_GB
constexpr int walk_gb(int) { return MFVGP_WALK_GB; }
#else
constexpr int walk_gb(int te) { return te <= 64 ? 4 : 2; }
#endif
#ifndef MFVGP_PANEL_UNROLL
#define MFVGP_PANEL_UNROLL 1
#endif
#ifndef MFVGP_WALK_DEFERW
#define MFVGP_WALK_DEFERW 1
#endif
#ifndef MFVGP_WALK_THREADS_PER_SM
#define MFVGP_WALK_THREADS_PER_SM 512
#endif

constexpr int kPanelUnroll = MFVGP_PANEL_UNROLL;

// Sums of one Kronrod panel in panel-local coordinates x in [-1, 1].
struct PanelSums {
    double kE, N0, N1, N2, N3, L0, L1, L2;     // Kronrod:  sum v x^j {t, qi2, qi}
    double gE, GN0, GN1, GN2, GL0, GL1;        // Gauss-10: sum w x^j {t, qi2, qi}
};

// s(x) = al0 + al1 x + al2 x^2, q(x) = be0 + be1 x on the panel.
__device__ __forceinline__ void panel_sums(double al0, double al1, double al2, double be0,
                                           double be1, PanelSums &o) {
    double N0 = 0.0, N1 = 0.0, N2 = 0.0, N3 = 0.0, L0 = 0.0, L1 = 0.0, L2 = 0.0;
    double GN0 = 0.0, GN1 = 0.0, GN2 = 0.0, GL0 = 0.0, GL1 = 0.0;
    // two pairs per trip (even pair: Kronrod only; odd pair: also a Gauss-10 node); the loop
    // is kept rolled so the live ranges of the reciprocals stay short (fully unrolled --
    // -DMFVGP_PANEL_UNROLL=5, node constants as immediate constant-bank operands instead of
    // LDCU, 12 % of the warp instructions -- the walk kernel spills 400 bytes at 128 registers)
#pragma unroll kPanelUnroll
    for (int ii = 0; ii < kPairs / 2; ++ii) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // 8 constants of the pair as four 16-byte constant loads
            const double2 *T2 = reinterpret_cast<const double2 *>(c_pair) + (2 * ii + h) * 4;
            const double2 c01 = T2[0], c23 = T2[1], c45 = T2[2], c67 = T2[3];
            const double x = c01.x, v = c01.y, vx = c23.x, vx2 = c23.y, vx3 = c45.x;
            const double sp = fma(fma(al2, x, al1), x, al0);
            const double sm = fma(fma(al2, -x, al1), -x, al0);
            const double qp = fma(be1, x, be0), qm = fma(-be1, x, be0);
            const double ip = qp * fast_rcp3(sp), im = qm * fast_rcp3(sm);   // q / s
            const double p2 = ip * ip, m2 = im * im;                         // (q/s)^2
            const double ue = p2 + m2, uo = p2 - m2, le = ip + im, lo = ip - im;
            N0 = fma(v, ue, N0);
            N2 = fma(vx2, ue, N2);
            N1 = fma(vx, uo, N1);
            N3 = fma(vx3, uo, N3);
            L0 = fma(v, le, L0);
            L2 = fma(vx2, le, L2);
            L1 = fma(vx, lo, L1);
            if (h == 1) {
                const double w = c45.y, wx = c67.x, wx2 = c67.y;
                GN0 = fma(w, ue, GN0);
                GN2 = fma(wx2, ue, GN2);
                GN1 = fma(wx, uo, GN1);
                GL0 = fma(w, le, GL0);
                GL1 = fma(wx, lo, GL1);
            }
        }
    }
    {   // centre node x = 0 (Kronrod only)
        const double ic = be0 * fast_rcp3(al0), vc = c_vcentre;
        N0 = fma(vc, ic * ic, N0);
        L0 = fma(vc, ic, L0);
    }
    // q^2/s = q (q/s) with q = be0 + be1 x:  sum v q^2/s = be0 sum v (q/s) + be1 sum v x (q/s),
    // i.e. the energy sums follow from the first-power sums that the gradient needs anyway
    // (3.5 FP64 instructions per node pair less; the terms of a sum have one sign unless q
    // changes sign inside the panel, so nothing cancels beyond a small factor)
    const double kE = fma(be0, L0, be1 * L1), gE = fma(be0, GL0, be1 * GL1);
    o.kE = kE; o.N0 = N0; o.N1 = N1; o.N2 = N2; o.N3 = N3; o.L0 = L0; o.L1 = L1; o.L2 = L2;
    o.gE = gE; o.GN0 = GN0; o.GN1 = GN1; o.GN2 = GN2; o.GL0 = GL0; o.GL1 = GL1;
}

// Rational part of one cell: K-sums in the normalisation of the split kernel
// (integral = (dt/4) x sum) and the guard quantities.
struct RatOut {
    double aE;            // sum_42 v q^2/s
    double kS[3];         // own-slot K-sums of the rational d/dS part
    double eE[2], eS[2];  // per panel: (Kronrod - Gauss)^2 of E, of the 3 dS elements
    double m1S[2];        // per panel: squared first moment of the k = 0 dS element
};

// Quadratic Lagrange basis on knots 0, 1/2, 1 expanded about the panel centre c
// (tau = c + x/4):  sum v B2_k(tau) f = B2_k(c) M0 + B2_k'(c)/4 M1 + B2_k''/32 M2,
//                   sum v B2_k'(tau) f = B2_k'(c) M0 + B2_k''/4 M1.
// c = 1/4: B2 = (3/8, 3/4, -1/8), B2' = (-2, 2, 0);  c = 3/4: B2 = (-1/8, 3/4, 3/8),
// B2' = (0, -2, 2);  B2'' = (4, -8, 4).
template <int P>
__device__ __forceinline__ void sv3(double M0, double M1, double M2, double *o) {
    if (P == 0) {
        o[0] = fma(0.375, M0, fma(-0.5, M1, 0.125 * M2));
        o[1] = fma(0.75, M0, fma(0.5, M1, -0.25 * M2));
        o[2] = 0.125 * (M2 - M0);
    } else {
        o[0] = 0.125 * (M2 - M0);
        o[1] = fma(0.75, M0, fma(-0.5, M1, -0.25 * M2));
        o[2] = fma(0.375, M0, fma(0.5, M1, 0.125 * M2));
    }
}
template <int P>
__device__ __forceinline__ void sd3(double M0, double M1, double *o) {
    if (P == 0) {
        o[0] = fma(-2.0, M0, M1);
        o[1] = 2.0 * (M0 - M1);
        o[2] = M1;
    } else {
        o[0] = M1;
        o[1] = -2.0 * (M0 + M1);
        o[2] = fma(2.0, M0, M1);
    }
}

template <int P>
__device__ __forceinline__ void panel_post(const PanelSums &s, double hid /*0.5/dt*/,
                                           double inv_dt, RatOut &o) {
    double sv[3], sd[3];
    sv3<P>(s.N0, s.N1, s.N2, sv);
    sd3<P>(s.L0, s.L1, sd);
#pragma unroll
    for (int k = 0; k < 3; ++k) o.kS[k] += fma(hid, sd[k], -0.25 * sv[k]);
    o.aE += s.kE;
    const double dE = 0.25 * (s.kE - s.gE);
    o.eE[P] = dE * dE;
    sv3<P>(s.N0 - s.GN0, s.N1 - s.GN1, s.N2 - s.GN2, sv);
    sd3<P>(s.L0 - s.GL0, s.L1 - s.GL1, sd);
    double eS = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double kg = fma(hid, sd[k], -0.25 * sv[k]);
        eS = fma(kg, kg, eS);
    }
    o.eS[P] = eS;
    // first moment of the k = 0 element: sum v x [(1 - qi2/4) B2_0 + (1 + qi/2) B2_0'/dt]
    double mv, md;
    if (P == 0) {
        mv = fma(0.375, s.N1, fma(-0.5, s.N2, 0.125 * s.N3));
        md = fma(-2.0, s.L1, s.L2);
    } else {
        mv = 0.125 * (s.N3 - s.N1);
        md = s.L2;
    }
    const double m1 = fma(hid, md, -0.25 * mv) + fma(c_m1poly[2 * P + 1], inv_dt, c_m1poly[2 * P]);
    o.m1S[P] = m1 * m1;
}

__device__ __forceinline__ void rational_part(double S0, double S1, double S2, double Sig,
                                              double inv_dt, RatOut &o) {
    // s(tau) = a0 + a1 tau + a2 tau^2 ; q(tau) = ds/dt - Sigma = b0 + b1 tau
    const double a0 = S0, a1 = -3.0 * S0 + 4.0 * S1 - S2, a2 = 2.0 * (S0 - 2.0 * S1 + S2);
    const double b0 = fma(a1, inv_dt, -Sig), b1 = 2.0 * a2 * inv_dt;
    const double al2 = 0.0625 * a2, be1 = 0.25 * b1, hid = 0.5 * inv_dt;
    o.aE = 0.0;
    o.kS[0] = o.kS[1] = o.kS[2] = 0.0;
    PanelSums s;
    panel_sums(fma(fma(a2, 0.25, a1), 0.25, a0), 0.25 * fma(0.5, a2, a1), al2,
               fma(b1, 0.25, b0), be1, s);
    panel_post<0>(s, hid, inv_dt, o);
    panel_sums(fma(fma(a2, 0.75, a1), 0.75, a0), 0.25 * fma(1.5, a2, a1), al2,
               fma(b1, 0.75, b0), be1, s);
    panel_post<1>(s, hid, inv_dt, o);
}

struct WalkArgs {
    // inputs: x = [means [D][Nm] | log-variances [D][Ns]]
    const double *x;
    const double *obs_times, *sigma;
    const double *mu0, *tau0, *obs_values, *obs_noise;
    const int32_t *obs_idx;      // [D] index into the observed dims or -1
    double *gm, *gs;             // gradient blocks [D][Nm], [D][Ns] (nullptr: value only)
    double *head_m, *head_s;     // [nruns][D]: first column of each run (k=0 part + dEobs), mailbox
    int *head_flag;              // [nruns][ntiles]: == epoch once the run's heads are published
    unsigned *run_ticket;        // [1]: dynamic CTA numbering (0 between launches)
    int32_t *status;             // OR-ed MFVGP_ST_SYNC_TIMEOUT if a mailbox wait gives up
    int epoch;                   // launch counter of this handle
    double *part;                // [(n_end-n_begin)*ntiles][kPartStride] guard partials
    double *apart;               // [gridDim.x][4] by ticket: sum log(tau0/s0), kl0 quad, obs, -
    double theta;                // L96 forcing (lorenz_96.py:70)
    int D, M, d, Nm, Ns, L;
    int n_begin, n_end;          // intervals of this shard
    int ntiles, nruns;           // dim tiles, runs
    const int *run_start;        // [nruns + 1] first interval of every run, relative to n_begin.
                                 // Runs are taken from the LAST one backwards, and the schedule
                                 // (mfvgp_create) makes them shorter towards the START of the time
                                 // axis -- long runs first for few prologues, short runs at the
                                 // end of the kernel for a small ragged end
    int with_e0;
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// GB = Gauss-Legendre nodes exchanged per pair of barriers (7: one batch per interval)
template <int TE, int GB>
struct WalkSmem {
    double2 ms[GB][TE];           // (m, s) at the batch's nodes
    double2 AC[GB][TE];           // (w dE_i/dm_{i+1}, w dE_i/ds_{i+1})
    double2 BD[GB][TE];           // (w dE_i/dm_{i-1}, w dE_i/ds_{i-1})
    double pf[6][TE];             // next interval: P1, P2, P3, log S1, log S2, y
    double carry[2][TE];          // k=3 / k=2 part of the previous interval
    double e0[2][TE];             // E_kl0 energy terms of the rows (run 0 only)
    double red[9][TE / 32];
    double isg[TE];               // 1 / Sigma of the rows (the neighbours' weights, DEFERW)
    double dt[kWalkMaxTN], idt[kWalkMaxTN];
    // persistent CTAs: the next run's first interval, fetched during the last interval of the
    // current one (P0..P3, log S0..S2, Sigma of the row; observation index; interval times)
    double nx[8][TE];
    int nx_jo[TE];
    double nx_t[kWalkMaxTN + 1];
    unsigned cta_id, next_id;
    int wait_ok;
};

// PERSIST: the grid is one CTA per resident slot and a CTA keeps drawing runs until none is left.
// During the last interval of a run it draws its next ticket and fetches that run's first
// support points (cp.async into WalkSmem::nx), so the run-to-run hand-over -- CTA exit and
// launch, the ticket's round trip, un-prefetched loads: ~4.4 us of an SM slot per run, measured
// as 0.51 us of kernel time per run at D=8192 (tools/sched_sweep.py) -- shrinks to the part
// that cannot be hidden (three exp, the mailbox, the column stores).
template <int TE, int GB, bool PERSIST>
__global__ void __launch_bounds__(TE, MFVGP_WALK_THREADS_PER_SM / TE)
esde_l96_walk_kernel(const WalkArgs a) {
    constexpr int TD = TE - 6, NW = TE / 32;
    extern __shared__ __align__(16) unsigned char walk_smem_raw[];
    WalkSmem<TE, GB> &sm = *reinterpret_cast<WalkSmem<TE, GB> *>(walk_smem_raw);

    const int e = threadIdx.x;
    const unsigned total = (unsigned)a.nruns * (unsigned)a.ntiles;
    // CTAs number themselves in the order they start (atomic ticket) and take the runs from the
    // LAST one backwards: a run hands the k=0 part of its first column to the run before it
    // (mailbox below), so a CTA only ever waits for a run with a smaller ticket, i.e. one that a
    // CTA holds or has finished -- no deadlock, whatever order the hardware dispatches blocks in.
    // Reset of the ticket for the next launch: by whoever makes the last draw of this one
    // (PERSIST: every CTA makes exactly one draw that fails, so total + gridDim.x draws in all).
    if (e == 0) {
        const unsigned id = atomicAdd(a.run_ticket, 1u);
        if (id == (PERSIST ? total : 0u) + gridDim.x - 1u) *a.run_ticket = 0u;
        sm.cta_id = id;
    }
    // the kernels behind this one are launched as programmatic dependents and wait for this
    // grid to complete (griddepcontrol.wait): once every CTA has started they may take the SM
    // slots the ragged end frees, instead of being launched after the last CTA has exited
    pdl_launch_dependents();
    __syncthreads();
    unsigned cur_id = sm.cta_id;
    bool have_nx = false;                 // this run's first interval was prefetched
  for (;;) {
    if (PERSIST && cur_id >= total) break;
    const int lid = (int)cur_id;
    const int tile = lid % a.ntiles, run = a.nruns - 1 - lid / a.ntiles;
    const int rs0 = a.run_start[run];
    const int n0 = a.n_begin + rs0;
    const int tnv = a.run_start[run + 1] - rs0;                         // intervals in this run
    const int d0 = tile * TD, D = a.D;
    int j = (d0 - 3 + e) % D;
    if (j < 0) j += D;
    const bool has_energy = (e >= 2) && (e <= TE - 2);
    const bool is_out = (e >= 3) && (e <= TE - 4) && (d0 + e - 3 < D);
    const bool want_grad = a.gm != nullptr;

    for (int i = e; i < tnv; i += TE) {
        const double dt = have_nx ? fabs(sm.nx_t[i + 1] - sm.nx_t[i])
                                  : fabs(a.obs_times[n0 + i + 1] - a.obs_times[n0 + i]);
        sm.dt[i] = dt;
        sm.idt[i] = 1.0 / dt;
    }

    // row pointers at the run's first column; the gradient has the layout of x
    const double *mp = a.x + (int64_t)j * a.Nm + 3 * (int64_t)n0;
    const double *sp = a.x + (int64_t)D * a.Nm + (int64_t)j * a.Ns + 2 * (int64_t)n0;
    const int64_t g_off_m = a.gm - a.x, g_off_s = a.gs - (a.x + (int64_t)D * a.Nm);
    double P0, P1, P2, P3, S0, S1, S2, Sig_;
    int jo_;
    if (have_nx) {
        P0 = sm.nx[0][e]; P1 = sm.nx[1][e]; P2 = sm.nx[2][e]; P3 = sm.nx[3][e];
        S0 = sm.nx[4][e]; S1 = sm.nx[5][e]; S2 = sm.nx[6][e];
        Sig_ = sm.nx[7][e];
        jo_ = sm.nx_jo[e];
    } else {
        P0 = mp[0]; P1 = mp[1]; P2 = mp[2]; P3 = mp[3];
        S0 = sp[0]; S1 = sp[1]; S2 = sp[2];
        Sig_ = a.sigma[j];
        jo_ = a.obs_idx[j];
    }
    const double Sig = Sig_;
    const int jo = is_out ? jo_ : -1;
    double Ri = 0.0, y0 = 0.0;
    if (jo >= 0) {
        Ri = a.obs_noise[jo];
        if (n0 >= 1) y0 = a.obs_values[(int64_t)(n0 - 1) * a.d + jo];     // obs at column 3 n0
    }
    S0 = exp(S0); S1 = exp(S1); S2 = exp(S2);
    if (jo >= 0) Ri = 1.0 / Ri;
    const double inv_sig = 1.0 / Sig;
    sm.isg[e] = inv_sig;
    sm.carry[0][e] = 0.0;
    sm.carry[1][e] = 0.0;
    sm.e0[0][e] = 0.0;
    sm.e0[1][e] = 0.0;
    double acc_obs = 0.0;
    __syncthreads();
#if MFVGP_WALK_DEFERW
    // the exchanged derivative terms carry no weight: the row that adds them applies the
    // producing row's 1 / Sigma (a multiply-add where an add stood), the node weight sits in
    // the basis table c_glw -- 5 FP64 instructions per (row, node) less than weighting at the
    // producer
    const double is_l = sm.isg[e >= 1 ? e - 1 : e], is_r = sm.isg[e + 1 < TE ? e + 1 : e],
                 is_rr = sm.isg[e + 2 < TE ? e + 2 : e];
#endif

    for (int nl = 0; nl < tnv; ++nl) {
        const int n = n0 + nl;
        // ---- next interval's support points: global -> shared, asynchronously ---------
        const bool more = nl + 1 < tnv;
        if (more) {
            cp_async8(&sm.pf[0][e], mp + 3 * nl + 4);
            cp_async8(&sm.pf[1][e], mp + 3 * nl + 5);
            cp_async8(&sm.pf[2][e], mp + 3 * nl + 6);
            cp_async8(&sm.pf[3][e], sp + 2 * nl + 3);
            cp_async8(&sm.pf[4][e], sp + 2 * nl + 4);
            if (jo >= 0) cp_async8(&sm.pf[5][e], a.obs_values + (int64_t)n * a.d + jo);
        }
        if (PERSIST && !more && e == 0) {                // last interval of the run: the next ticket
            const unsigned nid = atomicAdd(a.run_ticket, 1u);
            if (nid == total + gridDim.x - 1u) *a.run_ticket = 0u;
            sm.next_id = nid;
        }
        const double dt = sm.dt[nl], inv_dt = sm.idt[nl];

        // ---- (A) rational part, own dimension ---------------------------------------------
        RatOut ro;
        ro.aE = 0.0;
        ro.kS[0] = ro.kS[1] = ro.kS[2] = 0.0;
        ro.eE[0] = ro.eE[1] = ro.eS[0] = ro.eS[1] = ro.m1S[0] = ro.m1S[1] = 0.0;
        if (is_out) rational_part(S0, S1, S2, Sig, inv_dt, ro);

        // ---- (B) coupled polynomial part, 7 Gauss-Legendre nodes ---------------------------
        double accM[4] = {0, 0, 0, 0}, accS[3] = {0, 0, 0}, accE = 0.0;
#if MFVGP_WALK_DEFERW
        const double isdt = inv_sig * inv_dt;
#endif
#pragma unroll
        for (int g0 = 0; g0 < kGL; g0 += GB) {
            double own[GB];          // phase 1-2: theta - m - dm;  phase 2-3: w dE_i/dm_i
#pragma unroll
            for (int gg = 0; gg < GB; ++gg) {
                if (g0 + gg >= kGL) break;
                const double *T = c_gl + (g0 + gg) * kTabStride;
                const double m = P0 * T[T_B3] + P1 * T[T_B3 + 1] + P2 * T[T_B3 + 2]
                                 + P3 * T[T_B3 + 3];
                const double dm = (P0 * T[T_DB3] + P1 * T[T_DB3 + 1] + P2 * T[T_DB3 + 2]
                                   + P3 * T[T_DB3 + 3]) * inv_dt;
                const double s = S0 * T[T_B2] + S1 * T[T_B2 + 1] + S2 * T[T_B2 + 2];
                sm.ms[gg][e] = make_double2(m, s);
                own[gg] = (a.theta - m) - dm;
            }
            __syncthreads();
            if (PERSIST && g0 == 0 && !more) {
                // the ticket drawn above is visible now: fetch the next run's first interval
                // (this row's support points, Sigma, observation index; the interval times)
                const unsigned nid = sm.next_id;
                if (nid < total) {
                    const int tile2 = (int)(nid % (unsigned)a.ntiles);
                    const int run2 = a.nruns - 1 - (int)(nid / (unsigned)a.ntiles);
                    const int r20 = a.run_start[run2], tn2 = a.run_start[run2 + 1] - r20;
                    int j2 = (tile2 * TD - 3 + e) % D;
                    if (j2 < 0) j2 += D;
                    const double *mp2 = a.x + (int64_t)j2 * a.Nm + 3 * (int64_t)(a.n_begin + r20);
                    const double *sp2 = a.x + (int64_t)D * a.Nm + (int64_t)j2 * a.Ns
                                        + 2 * (int64_t)(a.n_begin + r20);
                    cp_async8(&sm.nx[0][e], mp2);
                    cp_async8(&sm.nx[1][e], mp2 + 1);
                    cp_async8(&sm.nx[2][e], mp2 + 2);
                    cp_async8(&sm.nx[3][e], mp2 + 3);
                    cp_async8(&sm.nx[4][e], sp2);
                    cp_async8(&sm.nx[5][e], sp2 + 1);
                    cp_async8(&sm.nx[6][e], sp2 + 2);
                    cp_async8(&sm.nx[7][e], a.sigma + j2);
                    cp_async4(&sm.nx_jo[e], a.obs_idx + j2);
                    for (int i = e; i <= tn2; i += TE)
                        cp_async8(&sm.nx_t[i], a.obs_times + a.n_begin + r20 + i);
                }
            }
#pragma unroll
            for (int gg = 0; gg < GB; ++gg) {
                if (g0 + gg >= kGL) break;
                if (has_energy) {
                    const double *T = c_gl + (g0 + gg) * kTabStride;
                    const double2 xa = sm.ms[gg][e + 1], xb = sm.ms[gg][e - 1],
                                  xc = sm.ms[gg][e - 2];
                    const double mb = xb.x, sb = xb.y;
                    const double u = xa.x - xc.x, su = xa.y + xc.y;
                    const double R = fma(u, mb, own[gg]);          // E[f_i] - dm_i
                    const double t1 = sb * u, t2 = mb * su;
                    const double Ec = fma(R, R, fma(t1, u, fma(t2, mb, su * sb)));
#if MFVGP_WALK_DEFERW
                    sm.AC[gg][e] = make_double2(fma(R, mb, t1), fma(mb, mb, sb));
                    sm.BD[gg][e] = make_double2(fma(R, u, t2), fma(u, u, su));
                    own[gg] = R;
                    accE = fma(T[T_V], Ec, accE);
#else
                    const double w = T[T_V] * inv_sig, w2 = w + w;
                    const double wR2 = w2 * R;
                    sm.AC[gg][e] = make_double2(fma(wR2, mb, w2 * t1), w * fma(mb, mb, sb));
                    sm.BD[gg][e] = make_double2(fma(wR2, u, w2 * t2), w * fma(u, u, su));
                    own[gg] = -wR2;
                    accE = fma(w, Ec, accE);
#endif
                }
            }
            __syncthreads();
#pragma unroll
            for (int gg = 0; gg < GB; ++gg) {
                if (g0 + gg >= kGL) break;
                if (is_out) {
                    const double2 l = sm.AC[gg][e - 1], r = sm.BD[gg][e + 1], rr = sm.AC[gg][e + 2];
#if MFVGP_WALK_DEFERW
                    const double *W = c_glw + (g0 + gg) * kGlwStride;
                    const double Gm = fma(is_l, l.x, fma(is_r, r.x, fma(-is_rr, rr.x, -inv_sig * own[gg])));
                    const double Gs = fma(is_l, l.y, fma(is_r, r.y, is_rr * rr.y));
                    const double Hm = -isdt * own[gg];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        accM[k] = fma(Gm, W[k], fma(Hm, W[4 + k], accM[k]));
#pragma unroll
                    for (int k = 0; k < 3; ++k) accS[k] = fma(Gs, W[8 + k], accS[k]);
#else
                    const double *T = c_gl + (g0 + gg) * kTabStride;
                    const double Gm = own[gg] + l.x + r.x - rr.x;
                    const double Gs = l.y + r.y + rr.y;
                    const double Hm = own[gg] * inv_dt;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        accM[k] += Gm * T[T_B3 + k] + Hm * T[T_DB3 + k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) accS[k] = fma(Gs, T[T_B2 + k], accS[k]);
#endif
                }
            }
        }

        // ---- per-interval results (same algebra as esde_l96_split_kernel) -------------------
        const double four_inv_dt = 4.0 * inv_dt;
        const double own_poly = dt * (S0 + 4.0 * S1 + S2) * (1.0 / 6.0) + (S2 - S0) - Sig * dt;
#if MFVGP_WALK_DEFERW
        const double ksum_E = 0.25 * ro.aE + 4.0 * accE + four_inv_dt * own_poly;
#else
        const double ksum_E = 0.25 * ro.aE + 4.0 * accE * Sig + four_inv_dt * own_poly;
#endif
        const double polyS[3] = {dt * (1.0 / 6.0) - 1.0, dt * (4.0 / 6.0), dt * (1.0 / 6.0) + 1.0};
        double dsv[3], own2 = 0.0;
        {
            const double sc_gl = 0.5 * dt, sc_gk = 0.125 * dt * inv_sig, hsig = 0.5 * inv_sig;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dsv[k] = fma(sc_gl, accS[k], fma(sc_gk, ro.kS[k], hsig * polyS[k]));
                const double o = fma(four_inv_dt, polyS[k], ro.kS[k]);
                own2 = fma(o, o, own2);
            }
        }
        // guard partials of this (interval, tile).  The energy is summed in FP64; the eight
        // guard quantities (sums of squares feeding a test with a 2 % margin, kernels.cuh
        // guard_reduce_kernel) go through a transposing FP32 butterfly: 9 shuffles for all
        // eight instead of 5 x 2 each.  Underflow -> 0 (no flag, correct: far below the
        // threshold), overflow -> inf (flag: safe); |I_E|^2 is clamped so tol stays finite.
        {
            const int lane = e & 31, w = e >> 5;
            const double vE = warp_sum(is_out ? ksum_E * inv_sig : 0.0);
            float f[8] = {fminf((float)(ksum_E * ksum_E), 1.0e37f), (float)ro.eE[0],
                          (float)ro.eE[1], (float)ro.eS[0], (float)ro.eS[1], (float)own2,
                          (float)ro.m1S[0], (float)ro.m1S[1]};
            if (!is_out)
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = 0.0f;
#pragma unroll
            for (int half = 4, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
                const bool hi = lane & bit;
#pragma unroll
                for (int k = 0; k < half; ++k) {
                    const float send = hi ? f[k] : f[k + half], keep = hi ? f[k + half] : f[k];
                    f[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
            }
            f[0] += __shfl_xor_sync(0xffffffffu, f[0], 2);
            f[0] += __shfl_xor_sync(0xffffffffu, f[0], 1);
            if (lane == 0) sm.red[0][w] = vE;
            if ((lane & 3) == 0) sm.red[1 + (lane >> 2)][w] = (double)f[0];
        }
        __syncthreads();
        if (e < 9) {
            double v = sm.red[e][0];
#pragma unroll
            for (int w = 1; w < NW; ++w) v += sm.red[e][w];
            a.part[((int64_t)(n - a.n_begin) * a.ntiles + tile) * kPartStride + e] = v;
        }

        // ---- gradient columns 3n..3n+2 / 2n..2n+1 (free_energy.py:715-743) -------------------
        if (is_out) {
            const double sc_gl = 0.5 * dt;
            double gm0 = fma(sc_gl, accM[0], sm.carry[0][e]), gs0 = dsv[0] + sm.carry[1][e];
            if (n == 0 && a.with_e0) {                           // :733-734 and E_kl0 :309-310
                const double mu = a.mu0[j], t0 = a.tau0[j];
                const double z0 = P0 - mu, it0 = 1.0 / t0;
                gm0 = fma(z0, it0, gm0);
                gs0 += 0.5 * (it0 - 1.0 / S0);
                sm.e0[0][e] = log(t0 / S0);
                sm.e0[1][e] = (z0 * z0 + S0 - t0) * it0;
            }
            if (jo >= 0 && n >= 1) {                             // :737-738 and E_obs :594-612
                const double rr = y0 - P0;
                gm0 = fma(-Ri, rr, gm0);
                gs0 = fma(0.5, Ri, gs0);
                acc_obs = fma(Ri, fma(rr, rr, S0), acc_obs);
            }
            if (want_grad) {
                double *gmp = const_cast<double *>(mp) + g_off_m + 3 * nl;
                double *gsp = const_cast<double *>(sp) + g_off_s + 2 * nl;
                if (nl == 0 && run > 0) {
                    // this column also receives the k=3 / k=2 part of the previous run's last
                    // interval: that run adds it and writes the column (mailbox, see below)
                    a.head_m[(int64_t)run * D + j] = gm0;
                    a.head_s[(int64_t)run * D + j] = gs0;
                } else {
                    gmp[0] = gm0;
                    gsp[0] = gs0 * S0;                           // :743
                }
                gmp[1] = sc_gl * accM[1];
                gmp[2] = sc_gl * accM[2];
                gsp[1] = dsv[1] * S1;
            }
            sm.carry[0][e] = sc_gl * accM[3];
            sm.carry[1][e] = dsv[2];
        }
        if (nl == 0 && run > 0 && want_grad) {               // publish the heads of this run
            // the barrier orders every thread's head stores before thread 0's release store,
            // and a release is cumulative over that order (PTX memory model): no fence per
            // thread (it cost 0.1 us of kernel time per run: 462 -> 452 us at L=512, 86 runs)
            __syncthreads();
            if (e == 0) st_release_gpu(a.head_flag + run * a.ntiles + tile, a.epoch);
        }
        if (more) {
            cp_async_wait_all();
            P0 = P3; P1 = sm.pf[0][e]; P2 = sm.pf[1][e]; P3 = sm.pf[2][e];
            S0 = S2; S1 = exp(sm.pf[3][e]); S2 = exp(sm.pf[4][e]);
            if (jo >= 0) y0 = sm.pf[5][e];
        }
    }

    // ---- the column this run shares with the next one ------------------------------------
    const bool last_run = n0 + tnv == a.n_end;           // last run of this shard (CTA-uniform)
    if (!last_run && want_grad) {
        // the next run published the rest of that column after ITS first interval, long ago
        if (e == 0) {
            const int *flag = a.head_flag + (run + 1) * a.ntiles + tile;
            int ok = 0;
            for (long spin = 0; spin < (1L << 26); ++spin) {
                if (ld_acquire_gpu(flag) == a.epoch) { ok = 1; break; }
                __nanosleep(64);
            }
            if (!ok) atomicOr(a.status, 32);              // MFVGP_ST_SYNC_TIMEOUT: never expected
            sm.wait_ok = ok;
        }
        __syncthreads();
    }
    if (is_out) {
        const bool last_shard = a.n_end == a.L;
        const double carry_m = sm.carry[0][e], carry_s = sm.carry[1][e];
        if (!last_run) {
            if (want_grad) {
                const double hm = __ldcg(a.head_m + (int64_t)(run + 1) * D + j);
                const double hs = __ldcg(a.head_s + (int64_t)(run + 1) * D + j);
                double *gmp = const_cast<double *>(mp) + g_off_m + 3 * tnv;
                double *gsp = const_cast<double *>(sp) + g_off_s + 2 * tnv;
                gmp[0] = hm + carry_m;
                gsp[0] = (hs + carry_s) * S2;                    // :743, same support point
            }
        } else {
            // end column of the shard, 3 n_end / 2 n_end.  On the last shard observation
            // M-1 sits on it, observation M three (two) columns further on, and the
            // remaining columns are never touched by the reference (:711-712): zeros.
            double *gmp = const_cast<double *>(mp) + g_off_m + 3 * tnv;
            double *gsp = const_cast<double *>(sp) + g_off_s + 2 * tnv;
            double gme = carry_m, gse = carry_s;
            if (last_shard && jo >= 0) {
                const int64_t M = a.M;
                const double y1 = a.obs_values[(M - 2) * a.d + jo];
                const double y2 = a.obs_values[(M - 1) * a.d + jo];
                const double mT = a.x[(int64_t)j * a.Nm + 3 * M];
                const double sT = exp(a.x[(int64_t)D * a.Nm + (int64_t)j * a.Ns + 2 * M]);
                const double rr = y1 - P3, r2 = y2 - mT;
                gme = fma(-Ri, rr, gme);
                gse = fma(0.5, Ri, gse);
                acc_obs = fma(Ri, fma(rr, rr, S2), acc_obs);
                acc_obs = fma(Ri, fma(r2, r2, sT), acc_obs);
                if (want_grad) {
                    gmp[3] = -Ri * r2;
                    gsp[2] = 0.5 * Ri * sT;
                }
            } else if (last_shard && want_grad) {
                gmp[3] = 0.0;
                gsp[2] = 0.0;
            }
            if (want_grad) {
                gmp[0] = gme;
                gsp[0] = gse * S2;
                if (last_shard) {
                    gmp[1] = 0.0; gmp[2] = 0.0;
                    gmp[4] = 0.0; gmp[5] = 0.0; gmp[6] = 0.0;
                    gsp[1] = 0.0;
                    gsp[3] = 0.0; gsp[4] = 0.0;
                }
            }
        }
    }

    // ---- E_kl0 / E_obs energy partials of this CTA ------------------------------------------
    {
        const int lane = e & 31, w = e >> 5;
        const double acc[3] = {sm.e0[0][e], sm.e0[1][e], acc_obs};
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double v = warp_sum(acc[i]);
            if (lane == 0) sm.red[i][w] = v;
        }
        __syncthreads();
        if (e < 4) {
            double v = 0.0;
            if (e < 3) {
                v = sm.red[e][0];
#pragma unroll
                for (int w2 = 1; w2 < NW; ++w2) v += sm.red[e][w2];
            }
            // row = the ticket, i.e. the (run, tile) this CTA worked on -- not blockIdx.x, whose
            // pairing with (run, tile) depends on the order the CTAs happened to start in: the
            // fold that follows must add the same partials in the same order every time
            a.apart[(int64_t)lid * 4 + e] = v;
        }
    }
    if (!PERSIST) break;
    cur_id = sm.next_id;                  // drawn during the last interval (visible: barriers since)
    have_nx = cur_id < total;
    cp_async_wait_all();                  // the prefetched first interval of the next run
    __syncthreads();
  }
}

}  // namespace mfvgp
