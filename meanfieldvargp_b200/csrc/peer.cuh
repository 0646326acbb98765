// Peer-memory exchange between the ranks of a sharded handle (one process per GPU), over
// NVLink / NVSwitch with plain stores into the neighbour's memory (CUDA IPC mappings).
//
// What is exchanged per evaluation (free_energy.py:504-505, 723-724): each rank's
// scalar record (F, E0, Esde, Eobs, status bits; or SCG dot products) goes to EVERY rank, its
// part of the gradient of the one column it shares with a neighbour goes to THAT neighbour only
// (2 x 8 x D bytes per boundary).  The NCCL all-gather this replaces moved every rank's record
// and both shared columns to every rank (world x (16 + 4 D) doubles per rank) and cost two
// launches plus the collective's latency; here one kernel publishes and one waits and folds.
//
// Mailbox of a rank (device memory, mapped by every peer), two slots (epoch & 1):
//   rec   [world][kPeerRec]   rec[r] written by rank r
//   edge  [2][2 D]            [0] by the LEFT neighbour (its right-edge mean / variance parts),
//                             [1] by the RIGHT neighbour (its left-edge parts)
//   flags [world + 2]         epoch of the last complete write of rec[r] / edge[side]
// A writer stores the data, fences at system scope and then releases the epoch into the
// flag; the reader acquires the flag and reads its own memory.  Exchanges are collective and
// stream-ordered on every rank, so a slot written for exchange k is not written again before
// every rank has folded exchange k (two slots are enough).  Waits are bounded:
// a rank that gives up raises MFVGP_ST_SYNC_TIMEOUT instead of hanging.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace mfvgp {

constexpr int kPeerRec = 16;          // doubles per record
constexpr int kPeerMaxWorld = 16;

struct PeerArgs {
    double *box[kPeerMaxWorld];       // every rank's mailbox as mapped here; box[rank] is local
    int64_t slot_doubles;             // doubles per slot
    long long epoch;
    unsigned *ticket;                 // [1], zero between launches
    int32_t *status;
    const double *rec;                // [n_rec] local record to publish
    double *out;                      // fold result [n_rec]
    int32_t *status_out;              // decoded status bits of a free-energy record, or nullptr
    double *grad;                     // gradient of x (edges), or nullptr: records only
    int n_rec, n_sum;                 // record length; slots < n_sum are summed, the rest max'ed
    int D, Nm, Ns, n_begin, n_end, rank, world;
    // SCG folds: the folded sums also go to the host-mapped record the control loop spins on
    // (scg.cuh FoldArgs: [0..3] sums, [4..7] F, E0, Esde, Eobs of the last E_cost, [8] its status)
    double *host_rec = nullptr;       // [32] host-mapped, or nullptr
    long long host_epoch = 0;
    const double *ecost_out = nullptr;
    const int32_t *ecost_status = nullptr;
    const double *fold_part = nullptr;   // scg_peer_fold_publish_kernel: [fold_nblk][4] partials
    int fold_nblk = 0;
};

__device__ __forceinline__ double *peer_rec(double *box, const PeerArgs &a, int r) {
    return box + (a.epoch & 1) * a.slot_doubles + (int64_t)r * kPeerRec;
}
__device__ __forceinline__ double *peer_edge(double *box, const PeerArgs &a, int side) {
    return box + (a.epoch & 1) * a.slot_doubles + (int64_t)a.world * kPeerRec
           + (int64_t)side * 2 * a.D;
}
__device__ __forceinline__ long long *peer_flag(double *box, const PeerArgs &a, int i) {
    return reinterpret_cast<long long *>(box + (a.epoch & 1) * a.slot_doubles
                                         + (int64_t)a.world * kPeerRec + 4 * (int64_t)a.D) + i;
}

// The mailboxes as a kernel that does its exchange itself (scg.cuh red_store) sees them: lives
// in device memory, written once by mfvgp_comm_init.  Same layout as PeerArgs describes.
struct PeerLayout {
    double *box[kPeerMaxWorld];
    int64_t slot_doubles;
    int rank, world, D;
};
__device__ __forceinline__ double *pl_rec(const PeerLayout &L, double *box, long long epoch, int r) {
    return box + (epoch & 1) * L.slot_doubles + (int64_t)r * kPeerRec;
}
__device__ __forceinline__ long long *pl_flag(const PeerLayout &L, double *box, long long epoch, int i) {
    return reinterpret_cast<long long *>(box + (epoch & 1) * L.slot_doubles
                                         + (int64_t)L.world * kPeerRec + 4 * (int64_t)L.D) + i;
}

__device__ __forceinline__ void st_release_sys(long long *p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long *p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One thread: sends four partial sums to every rank, waits for every rank's, folds them in rank
// order (slots 0-2 summed, slot 3 max'ed) -- identical bits on every rank.  false: a peer did
// not answer within 30 s.
__device__ __forceinline__ bool peer_allreduce4(const PeerLayout &L, long long epoch, double (&r)[4]) {
    for (int p = 0; p < L.world; ++p) {
        double *dst = pl_rec(L, L.box[p], epoch, L.rank);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = r[i];
    }
    for (int p = 0; p < L.world; ++p) st_release_sys(pl_flag(L, L.box[p], epoch, L.rank), epoch);
    double *mine = L.box[L.rank];
    bool ok = true;
    unsigned long long t0 = 0, t1 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int p = 0; p < L.world; ++p) {
        const long long *f = pl_flag(L, mine, epoch, p);
        for (long spin = 0; ld_acquire_sys(f) != epoch; ++spin) {
            __nanosleep(20);
            if ((spin & 1023) == 1023) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 30000000000ull) { ok = false; break; }
            }
        }
    }
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = pl_rec(L, mine, epoch, 0)[i];
    for (int p = 1; p < L.world; ++p) {
        const double *src = pl_rec(L, mine, epoch, p);
        acc[0] += src[0]; acc[1] += src[1]; acc[2] += src[2];
        acc[3] = fmax(acc[3], src[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = acc[i];
    return ok;
}

// grid = ceil(D / 256) (or 1 without edges).  The last CTA to finish releases the flags.
__global__ void __launch_bounds__(256)
peer_publish_kernel(const PeerArgs a) {
    __shared__ int sh_last;
    const int j = blockIdx.x * 256 + threadIdx.x;
    pdl_wait();                       // the record and the gradient belong to the predecessor
    pdl_launch_dependents();          // peer_combine_kernel may be scheduled: it waits for us
    if (blockIdx.x == 0 && (int)threadIdx.x < a.n_rec) {
        const double v = a.rec[threadIdx.x];
        for (int p = 0; p < a.world; ++p) peer_rec(a.box[p], a, a.rank)[threadIdx.x] = v;
    }
    if (a.grad != nullptr && j < a.D) {
        const double *gm = a.grad + (int64_t)j * a.Nm;
        const double *gs = a.grad + (int64_t)a.D * a.Nm + (int64_t)j * a.Ns;
        if (a.rank > 0) {                       // my left-edge parts: slot [1] of the left neighbour
            double *e = peer_edge(a.box[a.rank - 1], a, 1);
            e[j] = gm[3 * a.n_begin];
            e[a.D + j] = gs[2 * a.n_begin];
        }
        if (a.rank < a.world - 1) {             // my right-edge parts: slot [0] of the right neighbour
            double *e = peer_edge(a.box[a.rank + 1], a, 0);
            e[j] = gm[3 * a.n_end];
            e[a.D + j] = gs[2 * a.n_end];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(a.ticket, 1u);
        sh_last = t == gridDim.x - 1;
        if (sh_last) *a.ticket = 0u;
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < a.world)
        st_release_sys(peer_flag(a.box[threadIdx.x], a, a.rank), a.epoch);
    if (a.grad != nullptr) {
        if (threadIdx.x == 32 && a.rank > 0)
            st_release_sys(peer_flag(a.box[a.rank - 1], a, a.world + 1), a.epoch);
        if (threadIdx.x == 33 && a.rank < a.world - 1)
            st_release_sys(peer_flag(a.box[a.rank + 1], a, a.world + 0), a.epoch);
    }
}

// grid = ceil(D / 256) (or 1 without edges): waits for every rank's record (and the
// neighbours' edges), folds the records in rank order -- identical bits on every rank -- and
// adds the neighbours' parts of the two shared columns (two addends each: both owners of a
// shared column end up with identical bits).
__global__ void __launch_bounds__(256)
peer_combine_kernel(const PeerArgs a) {
    __shared__ int sh_ok;
    double *mine = a.box[a.rank];
    if (threadIdx.x == 0) sh_ok = 1;
    pdl_wait();                       // our own publish kernel has completed
    __syncthreads();
    const int nwait = a.world + (a.grad != nullptr ? 2 : 0);
    if ((int)threadIdx.x < nwait) {
        const int i = threadIdx.x;
        const bool needed = i < a.world || (i == a.world ? a.rank > 0 : a.rank < a.world - 1);
        if (needed) {
            const long long *f = peer_flag(mine, a, i);
            bool ok = false;
            unsigned long long t0 = 0, t1 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (long spin = 0;; ++spin) {
                if (ld_acquire_sys(f) == a.epoch) { ok = true; break; }
                __nanosleep(32);
                if ((spin & 1023) == 1023) {        // a peer may lag by seconds, not by minutes
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 30000000000ull) break;
                }
            }
            if (!ok) { sh_ok = 0; atomicOr(a.status, 32); }     // MFVGP_ST_SYNC_TIMEOUT
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && (int)threadIdx.x < a.n_rec) {
        __shared__ double sh_rec[kPeerRec];
        const int i = threadIdx.x;
        double v = peer_rec(mine, a, 0)[i];
        for (int r = 1; r < a.world; ++r) {
            const double w = peer_rec(mine, a, r)[i];
            v = i < a.n_sum ? v + w : fmax(v, w);
        }
        a.out[i] = v;
        sh_rec[i] = v;
        __syncwarp();
        if (i == 0 && a.host_rec != nullptr) {
            for (int k = 0; k < a.n_rec && k < 4; ++k) a.host_rec[k] = sh_rec[k];
            for (int k = 0; k < 4; ++k) a.host_rec[4 + k] = a.ecost_out[k];
            a.host_rec[8] = (double)(a.ecost_status[0] | (sh_ok ? 0 : 32));
            __threadfence_system();
            *reinterpret_cast<volatile long long *>(a.host_rec + 24) = a.host_epoch;
        }
        if (i == 0 && a.status_out != nullptr) {
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
