// Device-resident vector kernels for the scaled conjugate gradient loop
// (reference: SCG.minimize, src/numerical/scaled_cg.py:107-386).  Iterates,
// gradients and search direction never leave the device; every dot product is
// a fixed-order two-stage reduction so the accept/reject branches
// (scaled_cg.py:194, :233, :254) are reproducible run to run.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "peer.cuh"

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
    // 2: kappa < eps), 8 status of the sigma-step evaluation, 9 f_old, 10 void flag,
    // 11 gamma and 12 beta for the chain that follows (decide, below).
    int mode;
    double *ctrl;
    double beta_in;
    // Pipelined chains: the stats kernel of a chain also takes the accept / continue decision
    // (scaled_cg.py:253-369) in the host's arithmetic -- Delta, the two termination tests, the
    // beta update, gamma for the next direction -- so the host can enqueue the NEXT iteration's
    // chain before it has read this one's record (spec = 1: gamma, beta, f_old come from ctrl).
    // If the decision is anything but "accepted, carry on", ctrl[10] voids that next chain: its
    // vector kernels return at once (x, d untouched) and its stats kernel publishes a void
    // record.  The host takes the same decision from the same numbers and checks the flag.
    int spec;                  // this chain was enqueued ahead of its predecessor's outcome
    int decide;                // mode 3: take the decision
    int refined_mask;          // status bits that void an optimistic chain (0 if not optimistic)
    double f_old_in, x_tol, f_tol;
    // sharded handles: the last block's fold goes through every rank before the scalars are
    // used (peer_allreduce4), so the chain logic above runs on identical numbers on every rank
    const PeerLayout *peer;    // nullptr on one GPU
    long long peer_epoch;
    int32_t *sync_status;      // MFVGP_ST_SYNC_TIMEOUT is OR-ed in here if a rank does not answer
};

__device__ __forceinline__ bool chain_is_void(const FoldArgs &fa) {
    return fa.spec && __ldcg(fa.ctrl + 10) != 0.0;
}

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

// Visits every local entry of the rank's part of x: f(index into x, owned).  Single GPU: a
// grid-stride loop over the contiguous vector.  Sharded: the rank's columns are one contiguous
// segment per row of the mean block and one per row of the log-variance block; a WARP takes
// chunks of kVmChunk consecutive columns of one row (coalesced, and no 64-bit division per
// element as an index -> (row, column) map would need; the per-chunk division is amortised
// over 256 elements).  The assignment of elements to threads is fixed, so the reductions
// built on it stay reproducible.
constexpr int kVmChunk = 256;
template <class F>
__device__ __forceinline__ void vm_for_each(const VecMap &vm, F f) {
    if (vm.full) {
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < vm.n;
             t += (int64_t)gridDim.x * blockDim.x)
            f(t, true);
        return;
    }
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int cm = (vm.wm + kVmChunk - 1) / kVmChunk, cs = (vm.ws + kVmChunk - 1) / kVmChunk;
    const int64_t tasks_m = (int64_t)vm.D * cm, tasks = tasks_m + (int64_t)vm.D * cs;
    for (int64_t task = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); task < tasks;
         task += (int64_t)gridDim.x * wpb) {
        const bool mean = task < tasks_m;
        const int64_t r = mean ? task : task - tasks_m;
        const int per = mean ? cm : cs;
        const int j = (int)(r / per), c0 = (int)(r - (int64_t)j * per) * kVmChunk;
        const int w = mean ? vm.wm : vm.ws, own_w = mean ? vm.own_wm : vm.own_ws;
        const int64_t base = mean ? (int64_t)j * vm.Nm + vm.cm0 : vm.num_mp + (int64_t)j * vm.Ns + vm.cs0;
        const int c1 = min(c0 + kVmChunk, w);
        for (int c = c0 + lane; c < c1; c += 32) f(base + c, c < own_w);
    }
}

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

// scaled_cg.py:228-239 from the folded d.(g+ - g): theta, delta, beta and the step alpha
struct ChainIn { double sigma, mu, kappa, beta; };      // what the step needs beside the dot product
__device__ __forceinline__ ChainIn load_chain_in(const FoldArgs &fa) {
    const double *c = fa.ctrl;
    ChainIn ci;
    ci.sigma = __ldcg(c + 0); ci.mu = __ldcg(c + 5); ci.kappa = __ldcg(c + 6);
    ci.beta = fa.spec ? __ldcg(c + 12) : fa.beta_in;   // 12: written by the previous chain's decision
    return ci;
}
__device__ __forceinline__ double chain_alpha_from(double dot, const ChainIn &ci, double &theta,
                                                   double &delta, double &beta) {
    theta = dot / ci.sigma;
    beta = ci.beta;
    delta = __dadd_rn(theta, __dmul_rn(beta, ci.kappa));
    if (delta <= 0.0) {
        delta = __dmul_rn(beta, ci.kappa);
        beta = __dsub_rn(beta, theta / ci.kappa);
    }
    return -(ci.mu / delta);
}
__device__ __forceinline__ double chain_alpha(double dot, const FoldArgs &fa, double &theta,
                                              double &delta, double &beta) {
    return chain_alpha_from(dot, load_chain_in(fa), theta, delta, beta);
}

// The accept / continue decision at the end of a chain (scaled_cg.py:253-369, the host's
// operations).  stats_decide_calc only reads: 1 if the trial point is accepted and the run
// carries on, with f_old, gamma and beta for the next chain in nxt[0..2].  stats_decide_commit
// leaves them in ctrl[9], [11], [12] and sets / clears the void flag ctrl[10].
struct DecideIn { double f_new, alpha, mu, flags, beta, f_old; int stt; };
__device__ __forceinline__ DecideIn load_decide_in(const FoldArgs &fa) {
    const double *c = fa.ctrl;
    DecideIn in;
    in.f_new = __ldcg(fa.scal + 4);
    in.stt = (int)__ldcg(fa.status) | (int)__ldcg(c + 8);
    in.alpha = __ldcg(c + 1); in.mu = __ldcg(c + 5); in.flags = __ldcg(c + 7);
    in.beta = __ldcg(c + 4);
    in.f_old = fa.spec ? __ldcg(c + 9) : fa.f_old_in;
    return in;
}
__device__ __forceinline__ double stats_decide_from(const double (&r)[kRedSlots],
                                                    const FoldArgs &fa, const DecideIn &in,
                                                    double (&nxt)[3]) {
    nxt[0] = nxt[1] = nxt[2] = 0.0;
    if (!fa.decide) return 0.0;
    const double f_new = in.f_new, f_old = in.f_old, alpha = in.alpha, mu = in.mu;
    double beta = in.beta;
    if (in.flags != 0.0 || (in.stt & 7) != 0 || (in.stt & fa.refined_mask) != 0) return 0.0;
    const double Delta = __ddiv_rn(__dmul_rn(2.0, __dsub_rn(f_new, f_old)),
                                   __dmul_rn(alpha, mu));                       // :253
    const bool stop = (__dmul_rn(fabs(alpha), r[3]) <= fa.x_tol &&
                       fabs(__dsub_rn(f_new, f_old)) <= fa.f_tol) ||             // :306
                      fabs(r[1]) <= 1.0e-8;                                       // :333
    if (!(Delta >= 0.0) || stop) return 0.0;
    if (Delta < 0.25) beta = fmin(__dmul_rn(4.0, beta), 1.0e100);                 // :350-356
    if (Delta > 0.75) beta = fmax(__dmul_rn(0.5, beta), 1.0e-15);
    nxt[0] = f_new;
    nxt[1] = fmax(__ddiv_rn(r[2], mu), 0.0);                                      // :367
    nxt[2] = beta;
    return 1.0;
}
__device__ __forceinline__ double stats_decide_calc(const double (&r)[kRedSlots],
                                                    const FoldArgs &fa, double (&nxt)[3]) {
    if (!fa.decide) { nxt[0] = nxt[1] = nxt[2] = 0.0; return 0.0; }
    return stats_decide_from(r, fa, load_decide_in(fa), nxt);
}
__device__ __forceinline__ void stats_decide_commit(double carry_on, const double (&nxt)[3],
                                                    const FoldArgs &fa) {
    if (!fa.decide) return;
    double *c = fa.ctrl;
    if (carry_on != 0.0) { c[9] = nxt[0]; c[11] = nxt[1]; c[12] = nxt[2]; }
    c[10] = carry_on != 0.0 ? 0.0 : 1.0;
}

// the host record (without the epoch flag that makes it visible)
constexpr int kRecLen = 18;
// one warp publishes a staged record: one coalesced store of the kRecLen values into the
// host-mapped slot, system fence, epoch flag -- beside the rest of the kernel, not in its way
__device__ __forceinline__ void publish_record(const double *staged, const FoldArgs &fa) {
    const int lane = threadIdx.x & 31;
    if (lane < kRecLen) fa.host_rec[lane] = staged[lane];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) *fa.host_flag = fa.epoch;
}
__device__ __forceinline__ void stats_record_to(double *rec, const double (&r)[kRedSlots],
                                                double carry_on, const FoldArgs &fa) {
    const double *c = fa.ctrl;
#pragma unroll
    for (int i = 0; i < kRedSlots; ++i) rec[i] = r[i];
#pragma unroll
    for (int i = 4; i < 8; ++i) rec[i] = __ldcg(fa.scal + i);
    rec[8] = (double)__ldcg(fa.status);
    if (fa.mode == 3) {
        rec[9] = c[5]; rec[10] = c[6]; rec[11] = c[8];
        rec[12] = c[2]; rec[13] = c[1]; rec[14] = c[4];
        rec[15] = c[7];
        rec[16] = 0.0;                         // not a void chain
        rec[17] = carry_on;
    }
}
__device__ __forceinline__ void stats_record(const double (&r)[kRedSlots], double carry_on,
                                             const FoldArgs &fa) {
    stats_record_to(fa.host_rec, r, carry_on, fa);
}

// What one thread does with the folded sums r (grid path: thread 0 of the last block; cluster
// path: thread 0 of CTA 0): store them, then -- by mode -- the chain scalars or the host record.
__device__ __forceinline__ void fold_finish(const double (&r)[kRedSlots], const FoldArgs &fa) {
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
        double theta, delta, beta;
        c[1] = chain_alpha(r[0], fa, theta, delta, beta);
        c[2] = theta; c[3] = delta; c[4] = beta;
        c[8] = (double)__ldcg(fa.status);
        return;
    }
    double nxt[3], carry_on = 0.0;
    if (fa.mode == 3) carry_on = stats_decide_calc(r, fa, nxt);
    stats_record(r, carry_on, fa);
    if (fa.mode == 3) stats_decide_commit(carry_on, nxt, fa);
    __threadfence_system();
    *fa.host_flag = fa.epoch;
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
        if (fa.peer != nullptr && !peer_allreduce4(*fa.peer, fa.peer_epoch, r))
            atomicOr(fa.sync_status, 32);
        fold_finish(r, fa);
    }
}

// d = gamma*d - g (gamma = 0 and d arbitrary gives the restart d = -g);
// partials: [0] mu = d.g, [1] kappa = d.d     (scaled_cg.py:192-200, 361-369)
__global__ void __launch_bounds__(kRedBlock)
scg_direction_kernel(double *d, const double *g, double gamma, int restart, const VecMap vm,
                     double *part, const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    scg_pdl_enter();
    if (fa.spec) {
        if (chain_is_void(fa)) return;
        gamma = __ldcg(fa.ctrl + 11);
    }
    vm_for_each(vm, [&](int64_t i, bool own) {
        const double gi = g[i];
        const double di = restart ? -gi : gamma * d[i] - gi;
        d[i] = di;
        if (own) {
            v[0] = fma(di, gi, v[0]);
            v[1] = fma(di, di, v[1]);
        }
    });
    red_store(v, part, fa);
}

// y = x + a*d                                   (scaled_cg.py:222, :242)
__global__ void __launch_bounds__(kRedBlock)
scg_axpy_kernel(double *y, const double *x, const double *d, double a, const double *a_dev,
                const double *void_flag, const VecMap vm) {
    scg_pdl_enter();
    if (void_flag != nullptr && __ldcg(void_flag) != 0.0) return;    // void chain: y untouched
    if (a_dev != nullptr) a = __ldcg(a_dev);      // step length computed on the device
    vm_for_each(vm, [&](int64_t i, bool) { y[i] = x[i] + a * d[i]; });
}

// partial [0] = d.(gp - g)                      (scaled_cg.py:228)
__global__ void __launch_bounds__(kRedBlock)
scg_theta_kernel(const double *d, const double *gp, const double *g, const VecMap vm,
                 double *part, const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    scg_pdl_enter();
    if (chain_is_void(fa)) return;
    vm_for_each(vm, [&](int64_t i, bool own) {
        if (own) v[0] = fma(d[i], gp[i] - g[i], v[0]);
    });
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
    if (chain_is_void(fa)) {                      // nothing was computed: say so and stop
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            fa.host_rec[16] = 1.0;
            fa.host_rec[17] = 0.0;
            __threadfence_system();
            *fa.host_flag = fa.epoch;
        }
        return;
    }
    vm_for_each(vm, [&](int64_t i, bool own) {
        if (!own) return;
        const double a = g_now[i];
        v[0] += fabs(a);
        v[1] = fma(a, a, v[1]);
        v[2] = fma(a, g_new[i] - a, v[2]);
        v[3] = fmax(v[3], fabs(d[i]));
    });
    red_store(v, part, fa);
}

// ---- cluster path: the whole vector in ONE thread-block cluster ------------------------------
// Up to kClMaxElems entries (every notebook size incl. the north-star L96-500: 43 500) fit the
// registers of one cluster of <= 8 CTAs x 1024 threads.  Then the grid-wide dependency that
// forces "reduce, then a second kernel for the step" disappears: the block sums cross the
// cluster through distributed shared memory (every CTA stores its four sums into every CTA's
// table, ONE cluster barrier, every CTA folds the table in rank order -- fixed order, identical
// bits in all CTAs, no global partials, no fence, no atomic ticket), and the same kernel goes on
// to take the step with d still in registers:
//   scg_cl_direction_kernel : d = gamma d - g, mu, kappa, sigma, then xt = x + sigma d
//   scg_cl_theta_kernel     : d.(g+ - g), theta / delta / beta / alpha, then xt = x + alpha d
//   scg_cl_stats_kernel     : trial statistics + the accept / continue decision + host record
// One SCG iteration is 3 vector launches + 2 evaluations instead of 5 + 2.
// Two shapes: 16 CTAs x 512 threads where the device schedules 16-CTA clusters (opt-in size;
// twice the SMs pulling the vectors out of L2, half the shuffles per SM), else 8 x 1024.
constexpr int kClPerThread = 8;
constexpr int kClMaxCtas = 16;
constexpr int64_t kClMaxElems = 8 * 1024 * kClPerThread;

__device__ __forceinline__ unsigned cl_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cl_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_store_remote(double *local, unsigned rank, double v) {
    unsigned la = (unsigned)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

// v (per-thread) -> r (cluster totals, in every thread of every CTA): slots 0..2 summed, 3 max'ed
// (round: a kernel that reduces twice uses a fresh table for the second round, because a fast
// CTA may deliver its second-round sums while a slow one still folds the first table)
// NS: the first NS slots are live sums, MAXS: slot 3 (a max) is live; dead slots cost nothing
template <int NT, int NS = 3, bool MAXS = true>
__device__ __forceinline__ void cl_reduce(double (&v)[kRedSlots], double (&r)[kRedSlots],
                                          int round = 0) {
    __shared__ double sh_w[kRedSlots][32];
    __shared__ double sh_tabs[2][kClMaxCtas][kRedSlots];
    double (*sh_tab)[kRedSlots] = sh_tabs[round];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = warp_sum_d(v[i]);
    if (MAXS) v[3] = warp_max_d(v[3]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) sh_w[i][w] = v[i];
        if (MAXS) sh_w[3][w] = v[3];
    }
    __syncthreads();
    const unsigned nc = cl_size(), me = cl_rank();
    if (w == 0) {
        double b[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < NS; ++i) b[i] = warp_sum_d(lane < NT / 32 ? sh_w[i][lane] : 0.0);
        if (MAXS) b[3] = warp_max_d(lane < NT / 32 ? sh_w[3][lane] : 0.0);
        // lane 0 holds this CTA's sums: lanes 0..nc-1 each deliver them to one CTA of the cluster
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i)
            if (i < NS || (MAXS && i == 3)) b[i] = __shfl_sync(0xffffffffu, b[i], 0);
        if ((unsigned)lane < nc)
#pragma unroll
            for (int i = 0; i < kRedSlots; ++i)
                if (i < NS || (MAXS && i == 3))
                    cl_store_remote(&sh_tab[me][i], (unsigned)lane, b[i]);
    }
    cl_sync();
#pragma unroll
    for (int i = 0; i < kRedSlots; ++i) r[i] = (i < NS || (MAXS && i == 3)) ? sh_tab[0][i] : 0.0;
    for (unsigned k = 1; k < nc; ++k) {
#pragma unroll
        for (int i = 0; i < NS; ++i) r[i] += sh_tab[k][i];
        if (MAXS) r[3] = fmax(r[3], sh_tab[k][3]);
    }
}

struct ClVec {
    int n;                     // entries (<= kClMaxElems), identity index map (single GPU)
    const double *x;           // step: xt = x + a d (nullptr: no step in this launch)
    double *xt;
};

template <int NT>
__global__ void __launch_bounds__(NT)
scg_cl_direction_kernel(double *d, const double *g, double gamma, int restart, const ClVec cv,
                        const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    double dreg[kClPerThread], xreg[kClPerThread];     // x rides along: no load after the reduction
    scg_pdl_enter();
    if (fa.spec) {
        if (chain_is_void(fa)) return;          // every CTA takes this exit: no barrier pending
        gamma = __ldcg(fa.ctrl + 11);
    }
    const int T = (int)cl_size() * NT, t0 = (int)cl_rank() * NT + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kClPerThread; ++k) {
        const int i = t0 + k * T;
        dreg[k] = xreg[k] = 0.0;
        if (i < cv.n) {
            if (cv.x != nullptr) xreg[k] = cv.x[i];
            const double gi = g[i];
            const double di = restart ? -gi : gamma * d[i] - gi;
            d[i] = di;
            dreg[k] = di;
            v[0] = fma(di, gi, v[0]);
            v[1] = fma(di, di, v[1]);
        }
    }
    double r[kRedSlots];
    cl_reduce<NT, 2, false>(v, r);
    if (cl_rank() == 0 && threadIdx.x == 0) fold_finish(r, fa);
    if (cv.x != nullptr) {                       // sigma step (scaled_cg.py:203-222)
        const double sigma = 1.0e-3 / sqrt(r[1]);
#pragma unroll
        for (int k = 0; k < kClPerThread; ++k) {
            const int i = t0 + k * T;
            if (i < cv.n) cv.xt[i] = xreg[k] + sigma * dreg[k];
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT)
scg_cl_theta_kernel(const double *d, const double *gp, const double *g, const ClVec cv,
                    const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    double dreg[kClPerThread], xreg[kClPerThread];
    scg_pdl_enter();
    if (chain_is_void(fa)) return;
    const bool lead = cl_rank() == 0 && threadIdx.x == 0;
    const bool step = cv.x != nullptr;           // chain (mode 2): the alpha step follows
    // the scalars of the direction kernel and the status of the evaluation just finished are
    // final: fetch them with the vectors, not after the reduction
    ChainIn ci = {1.0, 0.0, 1.0, 0.0};
    int32_t st_eval = 0;
    if (step) {
        ci = load_chain_in(fa);
        if (lead) st_eval = __ldcg(fa.status);
    }
    const int T = (int)cl_size() * NT, t0 = (int)cl_rank() * NT + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kClPerThread; ++k) {
        const int i = t0 + k * T;
        dreg[k] = xreg[k] = 0.0;
        if (i < cv.n) {
            if (step) xreg[k] = cv.x[i];
            dreg[k] = d[i];
            v[0] = fma(dreg[k], gp[i] - g[i], v[0]);
        }
    }
    double r[kRedSlots];
    cl_reduce<NT, 1, false>(v, r);
    if (!step) {
        if (lead) fold_finish(r, fa);
        return;
    }
    double theta, delta, beta;                   // scaled_cg.py:228-247
    const double alpha = chain_alpha_from(r[0], ci, theta, delta, beta);
    if (lead) {                                  // fold_finish, mode 2, with the values at hand
        double *c = fa.ctrl;
#pragma unroll
        for (int i = 0; i < kRedSlots; ++i) fa.scal[i] = r[i];
        c[1] = alpha; c[2] = theta; c[3] = delta; c[4] = beta;
        c[8] = (double)st_eval;
    }
#pragma unroll
    for (int k = 0; k < kClPerThread; ++k) {
        const int i = t0 + k * T;
        if (i < cv.n) cv.xt[i] = xreg[k] + alpha * dreg[k];
    }
}

template <int NT>
__global__ void __launch_bounds__(NT)
scg_cl_stats_kernel(const double *g_now, const double *g_new, const double *d, const ClVec cv,
                    const FoldArgs fa) {
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    scg_pdl_enter();
    if (chain_is_void(fa)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            fa.host_rec[16] = 1.0;
            fa.host_rec[17] = 0.0;
            __threadfence_system();
            *fa.host_flag = fa.epoch;
        }
        return;
    }
    const int T = (int)cl_size() * NT, t0 = (int)cl_rank() * NT + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kClPerThread; ++k) {
        const int i = t0 + k * T;
        if (i < cv.n) {
            const double a = g_now[i];
            v[0] += fabs(a);
            v[1] = fma(a, a, v[1]);
            v[2] = fma(a, g_new[i] - a, v[2]);
            v[3] = fmax(v[3], fabs(d[i]));
        }
    }
    double r[kRedSlots];
    cl_reduce<NT, 3, true>(v, r);
    if (cl_rank() == 0 && threadIdx.x == 0) fold_finish(r, fa);
}

// Pipelined runs: the statistics + decision that close iteration j and the direction + sigma
// step that open iteration j+1 in one launch (the gradient of the trial point is both the
// g_now of the statistics and the g of the new direction, so it is loaded once).  CTA 0 decides,
// the decision and gamma cross the cluster through distributed shared memory; if iteration j is
// not "accepted, carry on" the kernel stops after the statistics and the rest of chain j+1 is
// void (ctrl[10]).  fa3: the statistics' fold (mode 3, decide), fa1: the direction's (mode 1).
template <int NT>
__global__ void __launch_bounds__(NT)
scg_cl_stats_direction_kernel(const double *g_now, const double *g_new, double *d, int restart_next,
                              const ClVec cv, const FoldArgs fa3, const FoldArgs fa1) {
    __shared__ double sh_dec[2];
    __shared__ double sh_rec[kRecLen];      // the host record, staged (lead CTA)
    double v[kRedSlots] = {0.0, 0.0, 0.0, 0.0};
    double dreg[kClPerThread], greg[kClPerThread], xreg[kClPerThread];
#ifdef MFVGP_PHASE_TRACE
    unsigned long long tq[10];
    int nq = 0;
#define TQ() do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq[nq])); ++nq; } while (0)
    TQ();
#else
#define TQ() do { } while (0)
#endif
    scg_pdl_enter();
    TQ();
    const bool lead = cl_rank() == 0 && threadIdx.x == 0;
    if (chain_is_void(fa3)) {
        if (lead) {
            fa3.host_rec[16] = 1.0;
            fa3.host_rec[17] = 0.0;
            __threadfence_system();
            *fa3.host_flag = fa3.epoch;
        }
        return;
    }
    const int T = (int)cl_size() * NT, t0 = (int)cl_rank() * NT + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kClPerThread; ++k) {
        const int i = t0 + k * T;
        dreg[k] = greg[k] = xreg[k] = 0.0;
        if (i < cv.n) {
            xreg[k] = cv.x[i];
            const double a = g_now[i];
            greg[k] = a;
            dreg[k] = d[i];
            v[0] += fabs(a);
            v[1] = fma(a, a, v[1]);
            v[2] = fma(a, g_new[i] - a, v[2]);
            v[3] = fmax(v[3], fabs(dreg[k]));
        }
    }
    double r[kRedSlots];
    TQ();
    // thread 0 of every CTA fetches what the decision needs now: the latency hides under the
    // reduction instead of following it
    DecideIn din = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0};
    if (threadIdx.x == 0 && fa3.decide) din = load_decide_in(fa3);
    cl_reduce<NT, 3, true>(v, r, 0);
    TQ();
    // every CTA takes the decision from the same numbers (no broadcast across the cluster);
    // the lead commits it only after the second reduction's barrier, when all have read ctrl
    double nxt[3];
    if (threadIdx.x == 0) {
        const double carry = stats_decide_from(r, fa3, din, nxt);
        sh_dec[0] = carry;
        sh_dec[1] = nxt[1];
        if (lead) {
#pragma unroll
            for (int i = 0; i < kRedSlots; ++i) fa3.scal[i] = r[i];
            stats_record_to(sh_rec, r, carry, fa3);  // published by warp 1 after the last barrier
            if (carry == 0.0) stats_decide_commit(carry, nxt, fa3);
        }
    }
    TQ();
    __syncthreads();
    TQ();
    const bool carry_on = sh_dec[0] != 0.0;
    if (carry_on) {
        const double gamma = sh_dec[1];
        v[0] = v[1] = v[2] = v[3] = 0.0;
#pragma unroll
        for (int k = 0; k < kClPerThread; ++k) {
            const int i = t0 + k * T;
            if (i < cv.n) {
                const double gi = greg[k];
                const double di = restart_next ? -gi : gamma * dreg[k] - gi;
                d[i] = di;
                dreg[k] = di;
                v[0] = fma(di, gi, v[0]);
                v[1] = fma(di, di, v[1]);
            }
        }
        cl_reduce<NT, 2, false>(v, r, 1);
        TQ();
        if (cl_rank() == 0 && (threadIdx.x >> 5) == 1) publish_record(sh_rec, fa3);
        if (lead) {
            stats_decide_commit(1.0, nxt, fa3);
            fold_finish(r, fa1);
        }
        const double sigma = 1.0e-3 / sqrt(r[1]);
#pragma unroll
        for (int k = 0; k < kClPerThread; ++k) {
            const int i = t0 + k * T;
            if (i < cv.n) cv.xt[i] = xreg[k] + sigma * dreg[k];
        }
        TQ();
    } else if (cl_rank() == 0 && (threadIdx.x >> 5) == 1) {
        publish_record(sh_rec, fa3);
    }
    if (lead) {
        TQ();
#ifdef MFVGP_PHASE_TRACE
        if ((fa3.epoch & 15) == 0) {
            printf("phase trace (ns since entry):");
            for (int q = 1; q < nq; ++q) printf(" %llu", tq[q] - tq[0]);
            printf("\n");
        }
#endif
    }
#undef TQ
}

// Sharded runs: second stage and exchange in one -- folds this rank's per-block partials in
// fixed order and sends the four sums to every rank's mailbox (peer.cuh); peer_combine_kernel
// follows, folds the ranks' sums in rank order and publishes them to the host record.
__global__ void __launch_bounds__(kRedBlock)
scg_peer_fold_publish_kernel(const PeerArgs a) {
    __shared__ double sh[kRedSlots][kRedBlock / 32];
    scg_pdl_enter();
    double r[kRedSlots];
    fold_partials(a.fold_part, a.fold_nblk, r, sh);
    if (threadIdx.x == 0) {
        for (int p = 0; p < a.world; ++p) {
            double *dst = peer_rec(a.box[p], a, a.rank);
#pragma unroll
            for (int i = 0; i < kRedSlots; ++i) dst[i] = r[i];
        }
        // same thread: each release store orders the record stores before it
        for (int p = 0; p < a.world; ++p) st_release_sys(peer_flag(a.box[p], a, a.rank), a.epoch);
    }
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
