// What follows the Lorenz-96 walk kernel in one evaluation of E_cost (free_energy.py:651-751):
// the per-interval guard (quad_vec's convergence test, kernels.cuh guard_interval), the scalar
// sums F = E0 + Esde + Eobs (:701) with the safe_log clamp of E_kl0 (:309-310), the status
// word -- and, on a sharded handle, the exchange with the other ranks over peer memory
// (peer.cuh): this rank's record to every rank, its parts of the two shared gradient columns
// (free_energy.py:723-724) to the two neighbours.
//
//   walk_tail_kernel   guard CTAs (a warp per interval) also fold slices of the walk kernel's
//                      E_kl0 / E_obs partials; the last one to finish (ticket) folds the CTA
//                      partials in fixed order and -- when no interval was flagged, the normal
//                      case -- finishes the evaluation: scalars, status, record to the peers.
//                      Edge CTAs beside them send the shared columns to the neighbours, one
//                      flag per 256-row chunk (no ticket, one fence).
//   refine_kernel, apply_refine_kernel (refine.cuh): return at once unless an interval is flagged
//   walk_late_kernel   only after a flagged evaluation: the scalars once refine_kernel has
//                      replaced the flagged intervals' integrals, the record, and the shared
//                      columns again (apply_refine_kernel may have changed them); marked
//                      "late" in the record so that the neighbours wait for the second copy.
//   walk_combine_kernel (sharded only) waits for the records and the neighbours' chunks, folds
//                      the records in rank order, adds the neighbours' parts.
// All four are launched with programmatic dependent launch and release their successor at
// once, so the successors are resident and waiting when their turn comes.
#pragma once

#include "kernels.cuh"
#include "peer.cuh"

namespace mfvgp {

constexpr int kTailRecLate = 4 + kStatusBits;      // record slot: shared columns are sent twice

struct WalkTailArgs {
    GuardArgs g;
    const double *apart;      // [napart][4] E_kl0 / E_obs partials of the walk kernel's CTAs
    int napart;
    double *tpart;            // [guard CTAs][4] partial sums of this kernel
    unsigned *ticket;         // [1], zero between launches
    int32_t *late;            // [1] != 0 while a flagged evaluation waits for walk_late_kernel
    double *out;              // [4] F, E0, Esde, Eobs (world 1), else the local record [16]
    int32_t *status;          // [1]
    int32_t *refine_status;   // bits raised by the walk kernel / refine_kernel, consumed here
    unsigned *gbar;           // refine_kernel's group counters
    int ngbar;
    int with_e0;
    double eobs_const;
    int guard_ctas;
    // sharded handles
    PeerArgs p;               // p.world == 1: no exchange; p.grad == nullptr: records only
    int nchunk;
};

__device__ __forceinline__ long long *tail_chunk_flag(double *box, const PeerArgs &a, int side,
                                                     int chunk, int nchunk) {
    return peer_flag(box, a, a.world + 2 + side * nchunk + chunk);
}

// scalars of the evaluation from the four sums; record / output / status; clean-up for the
// next evaluation.  One thread.
__device__ __forceinline__ void tail_finish(const WalkTailArgs &a, const double *acc, int anyflag,
                                            int late, int32_t refine_status) {
    const double Esde = acc[0];
    double E0 = 0.0;
    if (a.with_e0) {
        // safe_log(prod(tau0/s0)) (utilities.py:73-76): clamp in log space
        const double lg = fmin(fmax(acc[1], -690.77552789821368), 690.77552789821368);
        E0 = 0.5 * (lg + acc[2]);
    }
    const double Eobs = 0.5 * (acc[3] + a.eobs_const);
    int32_t st = refine_status;
    if (!isfinite(Esde)) st |= 1;
    if (!isfinite(E0)) st |= 2;
    if (!isfinite(Eobs)) st |= 4;
    if (anyflag) st |= 8;
    a.refine_status[0] = 0;
    a.out[0] = E0 + Esde + Eobs;
    a.out[1] = E0;
    a.out[2] = Esde;
    a.out[3] = Eobs;
    a.status[0] = st;
    if (a.p.world > 1) {
#pragma unroll
        for (int b = 0; b < kStatusBits; ++b) a.out[4 + b] = (st >> b) & 1 ? 1.0 : 0.0;
        a.out[kTailRecLate] = late ? 1.0 : 0.0;
    }
}

// the record to every rank (a warp: lane i carries slot i), then the flags
__device__ __forceinline__ void tail_publish_record(const PeerArgs &p, const double *rec, int lane) {
    if (lane < kPeerRec) {
        const double v = rec[lane];
        for (int r = 0; r < p.world; ++r) peer_rec(p.box[r], p, p.rank)[lane] = v;
    }
    // __syncwarp orders the lanes' record stores before each lane's release store, and a
    // release is cumulative over that order: no separate system-scope fence per lane
    __syncwarp();
    if (lane < p.world) st_release_sys(peer_flag(p.box[lane], p, p.rank), p.epoch);
}

// rows [chunk*256, +256) of the two shared columns to the neighbours
__device__ __forceinline__ void tail_publish_edges(const PeerArgs &p, int chunk, int nchunk,
                                                   long long flag_value) {
    const int j = chunk * 256 + threadIdx.x;
    if (j < p.D) {
        const double *gm = p.grad + (int64_t)j * p.Nm;
        const double *gs = p.grad + (int64_t)p.D * p.Nm + (int64_t)j * p.Ns;
        if (p.rank > 0) {                       // my left-edge parts: slot [1] of the left neighbour
            double *e = peer_edge(p.box[p.rank - 1], p, 1);
            e[j] = __ldcg(gm + 3 * p.n_begin);
            e[p.D + j] = __ldcg(gs + 2 * p.n_begin);
        }
        if (p.rank < p.world - 1) {             // my right-edge parts: slot [0] of the right neighbour
            double *e = peer_edge(p.box[p.rank + 1], p, 0);
            e[j] = __ldcg(gm + 3 * p.n_end);
            e[p.D + j] = __ldcg(gs + 2 * p.n_end);
        }
    }
    // the barrier orders all 256 threads' stores before the two release stores below
    // (cumulative): one system-scope release per neighbour instead of a fence per thread
    __syncthreads();
    if (threadIdx.x == 0 && p.rank > 0)
        st_release_sys(tail_chunk_flag(p.box[p.rank - 1], p, 1, chunk, nchunk), flag_value);
    if (threadIdx.x == 32 && p.rank < p.world - 1)
        st_release_sys(tail_chunk_flag(p.box[p.rank + 1], p, 0, chunk, nchunk), flag_value);
}

__global__ void __launch_bounds__(256)
walk_tail_kernel(const WalkTailArgs a) {
    __shared__ double sh_red[4 * 8];
    __shared__ int sh_last, sh_flag;
    pdl_wait();                           // the walk kernel's partials and gradient
    pdl_launch_dependents();
    if ((int)blockIdx.x >= a.guard_ctas) {                  // edge CTAs (sharded, with gradient)
        tail_publish_edges(a.p, blockIdx.x - a.guard_ctas, a.nchunk, 2 * a.p.epoch);
        return;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) sh_flag = 0;
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // this CTA's slice of the E_kl0 / E_obs partial rows (issued first: independent of the guard)
    for (int r = blockIdx.x * 256 + threadIdx.x; r < a.napart; r += a.guard_ctas * 256) {
        acc[1] += __ldcg(a.apart + (int64_t)r * 4 + 0);
        acc[2] += __ldcg(a.apart + (int64_t)r * 4 + 1);
        acc[3] += __ldcg(a.apart + (int64_t)r * 4 + 2);
    }
    // guard: a warp per interval (writes esde_n, flags, counts the flagged ones)
    const int n = a.g.n_begin + blockIdx.x * 8 + w;
    if (n < a.g.n_end) {
        double esde = 0.0;
        int flag = 0;
        guard_interval(a.g, n, lane, &esde, &flag);
        if (lane == 0) {
            acc[0] = esde;
            if (flag) atomicOr(&sh_flag, flag);
        }
    }
    block_sum<4, 256>(acc, sh_red);
    if (threadIdx.x == 0) {
        double *t = a.tpart + (int64_t)blockIdx.x * 4;
        t[0] = acc[0]; t[1] = acc[1]; t[2] = acc[2];
        t[3] = acc[3];
        __threadfence();
        const unsigned k = atomicAdd(a.ticket, 1u);
        sh_last = k == (unsigned)a.guard_ctas - 1u;
        if (sh_last) *a.ticket = 0u;
    }
    __syncthreads();
    if (!sh_last) return;
    // last guard CTA: fold the CTA partials in fixed order
    __threadfence();
    int32_t rs_pre = 0, nf_pre = 0;       // thread 0 needs these at the very end: fetch them now
    if (threadIdx.x == 0) {
        rs_pre = __ldcg(a.refine_status);
        nf_pre = a.g.nflagged != nullptr ? __ldcg(a.g.nflagged) : 0;
    }
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = threadIdx.x; c < a.guard_ctas; c += 256) {
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[i] += __ldcg(a.tpart + (int64_t)c * 4 + i);
    }
    __syncthreads();
    block_sum<4, 256>(tot, sh_red);
    __shared__ int sh_pub;
    if (threadIdx.x == 0) {
        const int nflag = nf_pre;
        sh_pub = 0;
        if (nflag == 0) {
            tail_finish(a, tot, 0, 0, rs_pre);
            sh_pub = a.p.world > 1;
        } else {
            a.late[0] = 1;                // refine_kernel changes Esde: walk_late_kernel finishes
        }
    }
    __syncthreads();
    if (sh_pub && w == 0) tail_publish_record(a.p, a.out, lane);      // after the barrier above
}

// after refine_kernel / apply_refine_kernel of a flagged evaluation; grid = 1 + edge CTAs
struct WalkLateArgs {
    WalkTailArgs t;
    const double *esde_n;     // [L] (flagged intervals now hold the refined integrals)
    const int32_t *flags;
};

__global__ void __launch_bounds__(256)
walk_late_kernel(const WalkLateArgs a) {
    __shared__ double sh_red[4 * 8];
    __shared__ int sh_flag;
    pdl_wait();
    pdl_launch_dependents();
    if (__ldcg(a.t.late) == 0) return;            // nothing was flagged: walk_tail_kernel finished
    if (blockIdx.x > 0) {
        tail_publish_edges(a.t.p, blockIdx.x - 1, a.t.nchunk, 2 * a.t.p.epoch + 1);
        return;
    }
    if (threadIdx.x == 0) sh_flag = 0;
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int anyflag = 0;
    for (int n = a.t.g.n_begin + threadIdx.x; n < a.t.g.n_end; n += 256) {
        acc[0] += __ldcg(a.esde_n + n);
        anyflag |= __ldcg(a.flags + n);
    }
    for (int c = threadIdx.x; c < a.t.guard_ctas; c += 256) {
#pragma unroll
        for (int i = 1; i < 4; ++i) acc[i] += __ldcg(a.t.tpart + (int64_t)c * 4 + i);
    }
    if (anyflag) atomicOr(&sh_flag, anyflag);
    block_sum<4, 256>(acc, sh_red);
    __syncthreads();
    if (threadIdx.x == 0) {
        tail_finish(a.t, acc, sh_flag, 1, __ldcg(a.t.refine_status));
        for (int i = 0; i < a.t.ngbar; ++i) a.t.gbar[i] = 0u;
        a.t.g.nflagged[0] = 0;
        __threadfence();
    }
    __syncthreads();
    if (a.t.p.world > 1 && threadIdx.x < 32) tail_publish_record(a.t.p, a.t.out, threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) a.t.late[0] = 0;
}

__device__ __forceinline__ bool wait_flag(const long long *f, long long want) {
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (long spin = 0;; ++spin) {
        if (ld_acquire_sys(f) == want) return true;
        __nanosleep(20);
        if ((spin & 1023) == 1023) {                // a peer may lag by seconds, not by minutes
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 30000000000ull) return false;
        }
    }
}

// grid = nchunk (1 without gradient).  out / status_out receive the folded record.
__global__ void __launch_bounds__(256)
walk_combine_kernel(const PeerArgs a, int nchunk, double *out4) {
    __shared__ int sh_ok;
    double *mine = a.box[a.rank];
    if (threadIdx.x == 0) sh_ok = 1;
    pdl_wait();
    pdl_launch_dependents();
    __syncthreads();
    // records: CTA 0 needs all of them, every CTA its neighbours' (the "late" marker)
    if ((int)threadIdx.x < a.world) {
        const int r = threadIdx.x;
        const bool needed = blockIdx.x == 0 || r == a.rank - 1 || r == a.rank + 1;
        if (needed && !wait_flag(peer_flag(mine, a, r), a.epoch)) sh_ok = 0;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int nb = threadIdx.x == 0 ? a.rank - 1 : a.rank + 1;
        long long want = -1;
        if (a.grad != nullptr && nb >= 0 && nb < a.world)
            want = 2 * a.epoch + (peer_rec(mine, a, nb)[kTailRecLate] != 0.0 ? 1 : 0);
        if (want >= 0 &&
            !wait_flag(tail_chunk_flag(mine, a, threadIdx.x, blockIdx.x, nchunk), want))
            sh_ok = 0;
    }
    __syncthreads();
    if (!sh_ok && threadIdx.x == 0) atomicOr(a.status, 32);      // MFVGP_ST_SYNC_TIMEOUT
    if (blockIdx.x == 0 && threadIdx.x < 4 + kStatusBits) {
        __shared__ double sh_rec[kPeerRec];
        const int i = threadIdx.x;
        double v = peer_rec(mine, a, 0)[i];
        for (int r = 1; r < a.world; ++r) v += peer_rec(mine, a, r)[i];
        sh_rec[i] = v;
        if (i < 4) out4[i] = v;
        __syncwarp();
        if (i == 0) {
            int32_t st = sh_ok ? 0 : 32;
            for (int b = 0; b < kStatusBits; ++b)
                if (sh_rec[4 + b] > 0.0) st |= 1 << b;
            a.status_out[0] = st;
        }
    }
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (a.grad == nullptr || j >= a.D) return;
    double *gm = a.grad + (int64_t)j * a.Nm;
    double *gs = a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
    if (a.rank > 0) {
        const double *e = peer_edge(mine, a, 0);
        gm[3 * a.n_begin] += e[j];
        gs[2 * a.n_begin] += e[a.D + j];
    }
    if (a.rank < a.world - 1) {
        const double *e = peer_edge(mine, a, 1);
        gm[3 * a.n_end] += e[j];
        gs[2 * a.n_end] += e[a.D + j];
    }
}

}  // namespace mfvgp
