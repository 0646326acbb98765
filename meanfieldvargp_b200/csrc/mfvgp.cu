// C ABI (include/mfvgp.h) over the kernels in kernels.cuh / scg.cuh.
// Pure CUDA runtime: no torch types cross this boundary.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mfvgp.h"
#include "kernels.cuh"
#include "nccl_dyn.h"
#include "refine.cuh"
#include "l96_walk.cuh"
#include "l96_lean.cuh"
#include "l96_coop.cuh"
#include "posterior.cuh"
#include "trajectory.cuh"
#include "scg.cuh"
#include "eparam.cuh"
#include "peer.cuh"
#include "walk_tail.cuh"

using namespace mfvgp;

static thread_local std::string g_err;

#define CK(call)                                                              \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);       \
            return MFVGP_ERR_CUDA;                                            \
        }                                                                     \
    } while (0)

#define ARG(cond, msg)                                                        \
    do {                                                                      \
        if (!(cond)) {                                                        \
            g_err = msg;                                                      \
            return MFVGP_ERR_ARG;                                             \
        }                                                                     \
    } while (0)

struct mfvgp_handle {
    int system = 0, D = 0, M = 0, d = 0, device = 0;
    int L = 0, n_begin = 0, n_end = 0, Nm = 0, Ns = 0, ntheta = 0;
    int te = 64, ntiles = 1;
    int coop = 0;     // small L96 problems: esde_l96_coop_kernel + assemble_tail_kernel (l96_coop.cuh)
    unsigned *tickets = nullptr;
    int *head_flag = nullptr;     // walk kernel mailbox flags [fntn][fntiles]
    int walk_epoch = 0;
    // optimistic evaluation (no refine_kernel launch) pays only while flagged intervals are rare:
    // after a flagged evaluation the next 16 are run with the adaptive path in the chain
    int opt_cooldown = 0;
    bool guard_in_tail = true;   // MFVGP_GUARD_IN_TAIL=0: keep the guard in the main kernel
    double *y_dense = nullptr, *rinv_row = nullptr;   // row tables of assemble_tail_kernel
    int tail_grid = 1;
    double theta_host[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int rule = 1;     // 1: split rule (default), 0: 42-node coupled rule (MFVGP_RULE=gk42)
    // walk kernel (l96_walk.cuh; fused = 2): dim tile fte, ftn consecutive intervals per run,
    // grid = fntn runs x fntiles tiles.  fused = 0: split / coop kernels with a workspace
    int fte = 64, ftn = 1, fntiles = 1, fntn = 1, fused = 0;
    int *run_start = nullptr;    // walk kernel: [fntn + 1] first interval of every run (device)
    double *fedge_m = nullptr, *fedge_s = nullptr, *fapart = nullptr;
    bool have_sde = false, have_prior = false, have_obs = false;
    double eobs_const = 0.0;
    // problem data (device)
    double *sigma = nullptr, *theta = nullptr, *mu0 = nullptr, *tau0 = nullptr;
    double *obs_times = nullptr, *obs_values = nullptr, *obs_noise = nullptr;
    int32_t *obs_idx = nullptr;
    // workspaces (device)
    double *dm_ws = nullptr, *ds_ws = nullptr, *part = nullptr, *esde_n = nullptr;
    double *apart = nullptr, *out_ws = nullptr;
    int32_t *flags = nullptr, *status_ws = nullptr;
    // adaptive refinement (refine.cuh)
    double *refine_ws = nullptr;
    int32_t *refine_info = nullptr, *refine_status = nullptr;   // refine_status[1] = flagged count
    bool refine_warp = true;
    int refine_grid = 1, refine_group = 1;      // CTAs per flagged item (refine.cuh GroupCtx)
    double *refine_gpart = nullptr;
    unsigned *refine_gbar = nullptr;
    // batched evaluation (mfvgp_ecost_batch): per-entry workspaces for batch_cap entries
    int batch_cap = 0;
    double *b_dm = nullptr, *b_ds = nullptr, *b_part = nullptr, *b_apart = nullptr,
           *b_esde_n = nullptr;
    int32_t *b_flags = nullptr, *b_rstat = nullptr;
    unsigned *b_ticket = nullptr;
    // host-API staging
    double *x_dev = nullptr, *grad_dev = nullptr;
    double *pin = nullptr;   // pinned: out[4] + status
    // SCG workspace (device): two iterates, direction, 3 rotating gradients, g_plus
    double *scg_vec[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *scg_part = nullptr, *scg_scal = nullptr;
    double *scg_hostrec = nullptr, *scg_hostrec_dev = nullptr;   // host-mapped record + flag
    long long scg_epoch = 0;
    int scg_nblk = 0;
    // multi-GPU: one rank per GPU, intervals [n_begin, n_end) of this handle
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int one_collective = 1;       // MFVGP_XCHG=split: send/recv of the edges + all-gather of the record
    double *edge_send = nullptr, *edge_recv = nullptr;   // [4][D] each
    double *gather = nullptr;                            // [world][16]
    double *xchg_send = nullptr, *xchg_gather = nullptr; // one-collective exchange: [16 + 4D], [world][..]
    double *rec_ws = nullptr;                            // [16] local record before combine
    // peer-memory exchange (peer.cuh): this rank's mailbox and the peers' as mapped here
    double *mbox = nullptr;
    double *peer_box[kPeerMaxWorld] = {};
    bool peer_ok = false;
    int64_t peer_slot = 0;
    long long xchg_epoch = 0;
    unsigned *peer_ticket = nullptr;
    PeerLayout *peer_layout = nullptr;   // device copy for kernels that exchange by themselves
    int peer_nchunk = 1;
    // walk path tail (walk_tail.cuh)
    double *tail_part = nullptr;
    int32_t *tail_late = nullptr;
    bool walk_tail = true;        // MFVGP_WALK_TAIL=0: guard / final / exchange as separate launches
    bool walk_persist = false;    // MFVGP_WALK_PERSIST=1: persistent CTAs that prefetch their next run
                                  // (measured slower: 460 vs 445 us at L=512, 3 406 vs 3 335 at L=4096)
    bool two_pass = true;         // MFVGP_TWO_PASS=0: coop path on device pointers as main -> refine -> tail
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    double *eparam_ws = nullptr;  // mfvgp_eparam: [nl][D] + [nl * ntiles][kMaxTheta] + out + gather
    int inject_status = 0;        // MFVGP_INJECT_STATUS="bits[@rank]" (tests): OR-ed into every evaluation
    // optional per-launch timing of the dominant (Esde) kernel, for bench.py
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
};

// cudaSetDevice on every entry costs microseconds; query the current device first
static inline cudaError_t use_device(int device) {
    int cur = -1;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess || cur != device) e = cudaSetDevice(device);
    return e;
}

static int dev_alloc(double **p, size_t n) {
    CK(cudaMalloc((void **)p, (n ? n : 1) * sizeof(double)));
    return MFVGP_OK;
}

extern "C" int mfvgp_version(void) { return MFVGP_VERSION; }
extern "C" const char *mfvgp_last_error(void) { return g_err.c_str(); }

extern "C" int mfvgp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int64_t mfvgp_launch_count(const mfvgp_handle *h) { return h ? h->launches : 0; }

extern "C" int mfvgp_describe(const mfvgp_handle *h, char *buf, int buflen) {
    if (!h || !buf || buflen <= 0) return 0;
    static const char *sys[] = {"DW", "OU", "L63", "L96"};
    const char *path = "small";        // DW / OU / L63: esde_small_kernel + assemble + final
    const char *kern = "esde_small_kernel";
    int te = 0, tn = 1, nt = 1, nr = h->n_end - h->n_begin;
    if (h->system == MFVGP_SYS_L96) {
        if (h->rule == 0) { path = "gk42"; kern = "esde_l96_kernel"; te = h->te; nt = h->ntiles; }
        else if (h->fused == 2) { path = "walk"; kern = "esde_l96_walk_kernel"; te = h->fte; tn = h->ftn; nt = h->fntiles; nr = h->fntn; }
        else if (h->coop) { path = "coop"; kern = "esde_l96_coop_kernel"; te = h->te; nt = h->ntiles; }
        else { path = "split"; kern = "esde_l96_split_kernel"; te = h->te; nt = h->ntiles; }
    }
    const int n = snprintf(buf, (size_t)buflen,
                           "system=%s path=%s kernel=%s te=%d tn=%d ntiles=%d nruns=%d D=%d M=%d "
                           "intervals=%d..%d world=%d xchg=%s",
                           sys[h->system], path, kern, te, tn, nt, nr, h->D, h->M, h->n_begin,
                           h->n_end, h->world,
                           h->world == 1 ? "none" : (h->peer_ok ? "peer" : (h->one_collective ? "nccl" : "split")));
    return n < buflen ? n : buflen - 1;
}

extern "C" int mfvgp_create(mfvgp_handle **out, int system, int D, int M, int d,
                            int device, int n_begin, int n_end) {
    ARG(out != nullptr, "mfvgp_create: out is NULL");
    ARG(system >= MFVGP_SYS_DW && system <= MFVGP_SYS_L96, "mfvgp_create: unknown system");
    ARG(M >= 2, "mfvgp_create: need at least 2 observation times");
    ARG(d >= 1 && d <= D, "mfvgp_create: need 1 <= d <= D");
    if (system == MFVGP_SYS_DW || system == MFVGP_SYS_OU) ARG(D == 1, "DW/OU need D = 1");
    if (system == MFVGP_SYS_L63) ARG(D == 3, "L63 needs D = 3");
    // lorenz_96.py:60-63
    if (system == MFVGP_SYS_L96) ARG(D >= 4, "Lorenz96: Insufficient state vector dimensions");
    const int L = M - 1;
    if (n_end < 0) n_end = L;
    ARG(n_begin >= 0 && n_begin < n_end && n_end <= L, "mfvgp_create: bad interval range");
    if (mfvgp_device_count() <= device) {
        g_err = "mfvgp_create: no CUDA device (this library has no CPU fallback)";
        return MFVGP_ERR_CUDA;
    }
    CK(cudaSetDevice(device));
    // every error return below goes through `guard`, which releases the half-built handle
    mfvgp_handle *h = new mfvgp_handle();
    struct Guard {
        mfvgp_handle *h;
        ~Guard() { if (h) mfvgp_destroy(h); }
    } guard{h};
    h->system = system; h->D = D; h->M = M; h->d = d; h->device = device;
    h->L = L; h->n_begin = n_begin; h->n_end = n_end;
    h->Nm = 3 * M + 4; h->Ns = 2 * M + 3;
    const int64_t cells = (int64_t)D * (n_end - n_begin);
    if (system == MFVGP_SYS_L96) {
        h->te = D <= 26 ? 32 : (cells <= 148 * 256 ? 64 : 128);
        h->ntiles = (D + h->te - 7) / (h->te - 6);
    }
    const int nl = n_end - n_begin;
    // E_cost path for L96 (MFVGP_PATH=split|walk):
    //   split: esde_l96_coop_kernel (or esde_l96_split_kernel) + workspace + assembly (small)
    //   walk : esde_l96_walk_kernel, gradient written directly (default for large problems)
    const char *path_env = getenv("MFVGP_PATH");
    const char *rule_env = getenv("MFVGP_RULE");
    h->rule = (rule_env && std::string(rule_env) == "gk42") ? 0 : 1;
    // measured crossover (L96, one B200): up to ~40 k (dimension, interval) cells the
    // latency-oriented kernels of l96_coop.cuh win, above that the walk kernel
    int64_t small_cells = 148 * 256;
    if (const char *e = getenv("MFVGP_SMALL_CELLS")) small_cells = atoll(e);
    if (system == MFVGP_SYS_L96 && h->rule == 1) {
        std::string path = path_env ? path_env : "";
        if (path != "walk" && path != "split") path = cells > small_cells ? "walk" : "split";
        h->fused = path == "walk" ? 2 : 0;
    }
    if (system == MFVGP_SYS_L96 && h->rule == 1 && h->fused == 0 && cells <= small_cells) {
        const char *coop_env = getenv("MFVGP_COOP");
        h->coop = (coop_env && coop_env[0] == '0') ? 0 : 1;
        if (h->coop) {
            const char *git = getenv("MFVGP_GUARD_IN_TAIL");
            h->guard_in_tail = !(git && git[0] == '0');
            h->te = kCoopTE;
            h->ntiles = (D + kCoopTE - 7) / (kCoopTE - 6);
            h->tail_grid = (int)(((int64_t)D * (h->Nm + h->Ns) + kTailThreads - 1) / kTailThreads);
        }
    }
    if (system != MFVGP_SYS_L96) {
        // DW / OU / L63: guard inside esde_small_inkernel, refine_kernel, one-launch tail
        const char *coop_env = getenv("MFVGP_COOP");
        h->coop = (coop_env && coop_env[0] == '0') ? 0 : 1;
        if (h->coop)
            h->tail_grid = (int)(((int64_t)D * (h->Nm + h->Ns) + kTailThreads - 1) / kTailThreads);
    }
    if (system == MFVGP_SYS_L96) {
        if (h->fused == 2) {
            // dim tile TE in {32, 64, 128}: fewest threads spent on halo / padding, with a 7 %
            // handicap on 128 -- its CTAs of four warps wait longer at the barriers than twice as
            // many CTAs of two.  Measured at D=8192 (tools/walk_ab.py), where TE=64 runs 4.4 % more
            // threads (142 tiles of 58 rows against 68 of 122): kernel 3 066 vs 3 141 us at L=4096,
            // 412 vs 418 at L=512, evaluation 3 107 vs 3 172 and 435 vs 439 (twice the tile
            // partials for the tail kernel to read).  Before the instruction count of the kernel
            // fell by 11 % the two were a wash.  96 and 160 -- 3 / 5 warps over 4 schedulers --
            // are 8-20 % slower; 256 -- half the halo rows of 128 -- 3 600 us.
            int best = 128;
            double best_cost = 1.07 * (double)((D + 121) / 122) * 128;
            for (int te : {64, 32}) {
                const double cost = (double)((D + te - 7) / (te - 6)) * te;
                if (cost < best_cost) { best = te; best_cost = cost; }
            }
            h->fte = best;
            const int nt = (D + h->fte - 7) / (h->fte - 6);
            // intervals per run TN: every run pays a prologue (un-prefetched loads, three exp, a
            // mailbox hand-over) ~ 0.4 interval, so cost ~ 0.4/TN; fewer, longer CTAs lose
            // ~ half a wave at the end, cost ~ 0.5 TN R / (nl nt) with R resident CTAs.  The
            // minimum is at TN = sqrt(0.8 nl nt / R); never below 2.  (Measured, us: D=4000
            // L=250 167 at TN=1, 120 at 4; D=8192 L=512 463 at 7, 488 at 23; L=4096 flat 16..32.)
            const double resident = 148.0 * (512 / h->fte);
            const int tn = (int)lround(sqrt(0.8 * (double)nl * nt / resident));
            h->ftn = tn < 2 ? 2 : (tn > 32 ? 32 : tn);
            if (h->ftn > nl) h->ftn = nl;
            if (const char *ft = getenv("MFVGP_FTILE")) {   // experiments: "TE,TN"
                int te = 0, tn = 0;
                if (sscanf(ft, "%d,%d", &te, &tn) == 2) { h->fte = te; h->ftn = tn; }
            }
            ARG(h->fte == 32 || h->fte == 64 || h->fte == 128, "MFVGP_FTILE: walk TE must be 32|64|128");
            ARG(h->ftn >= 1 && h->ftn <= kWalkMaxTN, "MFVGP_FTILE: walk TN must be 1..64");
        }
        h->fntiles = (D + h->fte - 7) / (h->fte - 6);
        h->fntn = (nl + h->ftn - 1) / h->ftn;
        if (h->fused == 2) {
            // Run schedule.  Runs are taken from the last one backwards by CTAs in the order they
            // start, so the runs at the END of the time axis are processed first and those at its
            // START last.  Every run pays a prologue of ~0.65 interval (un-prefetched loads, three
            // exp, mailbox hand-over), which asks for few long runs; the kernel ends when the last
            // run does, which asks for short runs at the end.  Guided schedule: with c = resident
            // CTAs / dim tiles runs in flight, each run takes 1 / (f c) of the intervals still to
            // be handed out (f = 1.5), between tmin and tmax.  "MFVGP_WALK_SCHED=uniform" keeps
            // round 1's schedule (TN = sqrt(0.8 nl nt / R) everywhere, ~2 waves of TN/4 at the
            // start of the axis); "f,tmin,tmax" sets the guided one.
            const int resident = 148 * (512 / h->fte);
            std::vector<int> lens;                       // from the END of the time axis
            const char *se = getenv("MFVGP_WALK_SCHED");
            if (se && se[0] == 'u') {
                int tns = h->ftn / 4 < 2 ? 2 : h->ftn / 4;
                int ns = (2 * resident + h->fntiles - 1) / h->fntiles;
                if (tns >= h->ftn || tns < 1) ns = 0;
                while (ns > 0 && (int64_t)ns * tns > nl / 2) ns /= 2;
                int rest = nl - ns * tns;
                while (rest > 0) { const int t = rest < h->ftn ? rest : h->ftn; lens.push_back(t); rest -= t; }
                // the partial run of the uniform part sits next to the short ones
                if (!lens.empty() && lens.size() > 1) std::swap(lens.front(), lens.back());
                for (int i = 0; i < ns; ++i) lens.push_back(tns);
            } else {
                // measured at D=8192 (tools/sched_sweep.py): L=512 463 us (uniform) -> 443; L=4096
                // 3 358 -> 3 314; f = 1.25 is 0.3 % better still but f = 1 falls off a cliff
                // (first runs longer than the first wave can balance: 494 us at L=512)
                double f = 1.5;
                int tmin = 2, tmax = 64;
                if (se) sscanf(se, "%lf,%d,%d", &f, &tmin, &tmax);
                if (tmax > kWalkMaxTN) tmax = kWalkMaxTN;
                if (tmin < 1) tmin = 1;
                const double c = (double)resident / h->fntiles > 1.0 ? (double)resident / h->fntiles : 1.0;
                int rem = nl;
                while (rem > 0) {
                    int t = (int)lround(rem / (f * c));
                    t = t < tmin ? tmin : (t > tmax ? tmax : t);
                    if (t > rem) t = rem;
                    lens.push_back(t);
                    rem -= t;
                }
            }
            h->fntn = (int)lens.size();
            std::vector<int> start(lens.size() + 1, 0);
            for (size_t r = 0; r < lens.size(); ++r)          // run 0 = START of the axis = last of lens
                start[r + 1] = start[r] + lens[lens.size() - 1 - r];
            h->ftn = lens.front();
            CK(cudaMalloc((void **)&h->run_start, start.size() * sizeof(int)));
            CK(cudaMemcpy(h->run_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    double tab[kNodes * kTabStride];
    build_fixed_table(tab);
    CK(cudaMemcpyToSymbol(c_tab, tab, sizeof(tab)));
    build_rational_table(tab);
    CK(cudaMemcpyToSymbol(c_rat, tab, sizeof(tab)));
    {
        double gl[kGL * kTabStride];
        build_gl_table(gl);
        CK(cudaMemcpyToSymbol(c_gl, gl, sizeof(gl)));
        double glw[kGL * kGlwStride];
        for (int g = 0; g < kGL; ++g) {
            const double *T = gl + g * kTabStride;
            double *W = glw + g * kGlwStride;
            for (int k = 0; k < 4; ++k) {
                W[k] = 2.0 * T[T_V] * T[T_B3 + k];
                W[4 + k] = 2.0 * T[T_V] * T[T_DB3 + k];
            }
            for (int k = 0; k < 3; ++k) W[8 + k] = T[T_V] * T[T_B2 + k];
            W[11] = T[T_V];
        }
        CK(cudaMemcpyToSymbol(c_glw, glw, sizeof(glw)));
        double vxB[kNodes], vxD[kNodes], m1poly[4], gkx[kNodesPerPanel];
        build_moment_table(vxB, vxD, m1poly);
        for (int r = 0; r < kNodesPerPanel; ++r) gkx[r] = (double)gk21_x(r);
        CK(cudaMemcpyToSymbol(c_vxB, vxB, sizeof(vxB)));
        CK(cudaMemcpyToSymbol(c_vxD, vxD, sizeof(vxD)));
        CK(cudaMemcpyToSymbol(c_m1poly, m1poly, sizeof(m1poly)));
        CK(cudaMemcpyToSymbol(c_gkx_main, gkx, sizeof(gkx)));
        double pair[kPairs * kPairStride], vc;
        build_pair_table(pair, &vc);
        CK(cudaMemcpyToSymbol(c_pair, pair, sizeof(pair)));
        CK(cudaMemcpyToSymbol(c_vcentre, &vc, sizeof(vc)));
    }
    {
        double gx[kNodesPerPanel], gv[kNodesPerPanel], gw[10];
        for (int r = 0; r < kNodesPerPanel; ++r) { gx[r] = (double)gk21_x(r); gv[r] = (double)gk21_v(r); }
        for (int i = 0; i < 10; ++i) gw[i] = (double)gk21_w(i);
        CK(cudaMemcpyToSymbol(c_gk_x, gx, sizeof(gx)));
        CK(cudaMemcpyToSymbol(c_gk_v, gv, sizeof(gv)));
        CK(cudaMemcpyToSymbol(c_gk_w, gw, sizeof(gw)));
    }
    if (dev_alloc(&h->sigma, D) || dev_alloc(&h->theta, 8) || dev_alloc(&h->mu0, D) ||
        dev_alloc(&h->tau0, D) || dev_alloc(&h->obs_times, M) ||
        dev_alloc(&h->obs_values, (size_t)M * d) || dev_alloc(&h->obs_noise, d) ||
        dev_alloc(&h->dm_ws, (size_t)L * D * 4) || dev_alloc(&h->ds_ws, (size_t)L * D * 3) ||
        dev_alloc(&h->part, (size_t)nl * (h->ntiles > h->fntiles ? (h->ntiles > 2 ? h->ntiles : 2)
                                                                   : (h->fntiles > 2 ? h->fntiles : 2)) *
                                kPartStride) ||

        dev_alloc(&h->fedge_m, (size_t)h->fntn * D) || dev_alloc(&h->fedge_s, (size_t)h->fntn * D) ||
        dev_alloc(&h->fapart, (size_t)h->fntn * h->fntiles * 4) ||
        dev_alloc(&h->esde_n, L) || dev_alloc(&h->apart, (size_t)(D > h->tail_grid ? D : h->tail_grid) * 4) ||
        dev_alloc(&h->out_ws, 16) || dev_alloc(&h->rec_ws, 16))
        return MFVGP_ERR_CUDA;
    // adaptive path (refine.cuh): a flagged item is shared by a group of G co-resident CTAs.
    // D < 1024: rows are scarce and latency decides -- one WARP per state dimension (nodes
    // across the lanes), G ~ D/8 (two rows per warp).  D >= 1024: throughput decides -- one
    // THREAD per state dimension, groups of 2-16 CTAs.  (Measured at D=8192, forty flagged
    // intervals: 3.4 ms per evaluation with threads per row, 36 ms with warps per row; at
    // D=500, one flagged interval: 175 us with warps per row, 320 us with threads per row.)
    // At most two CTAs per SM fit (registers), so the grid never exceeds 296 CTAs.
    {
        h->refine_warp = D < 1024;
        if (const char *e = getenv("MFVGP_REFINE_WARP")) h->refine_warp = e[0] == '1';
        int gsz = h->refine_warp ? (D + 7) / 8
                                 : (D >= 8192 ? 16 : (D >= 4096 ? 8 : (D >= 2048 ? 4 : 2)));
        h->refine_group = gsz < 1 ? 1 : (gsz > 64 ? 64 : gsz);
        if (const char *e = getenv("MFVGP_REFINE_GROUP")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 64) h->refine_group = v;
        }
        int max_ctas = 296;
        if (const char *e = getenv("MFVGP_REFINE_CTAS")) max_ctas = atoi(e);
        int groups = max_ctas / h->refine_group;
        if (groups < 1) groups = 1;
        if (groups > 2 * nl) groups = 2 * nl;
        h->refine_grid = groups * h->refine_group;
        if (dev_alloc(&h->refine_ws, (size_t)groups * 2 * D * kMaxNE)) return MFVGP_ERR_CUDA;
        if (dev_alloc(&h->refine_gpart, (size_t)groups * 2 * h->refine_group * 8)) return MFVGP_ERR_CUDA;
        CK(cudaMalloc((void **)&h->refine_gbar, (size_t)groups * sizeof(unsigned)));
        CK(cudaMemset(h->refine_gbar, 0, (size_t)groups * sizeof(unsigned)));
    }
    if (h->coop && (dev_alloc(&h->y_dense, (size_t)M * D) || dev_alloc(&h->rinv_row, D)))
        return MFVGP_ERR_CUDA;
    CK(cudaMalloc((void **)&h->refine_info, 2 * L * sizeof(int32_t)));
    CK(cudaMalloc((void **)&h->refine_status, 2 * sizeof(int32_t)));
    CK(cudaMemset(h->refine_info, 0, 2 * L * sizeof(int32_t)));
    CK(cudaMemset(h->refine_status, 0, 2 * sizeof(int32_t)));
    CK(cudaMalloc((void **)&h->head_flag, (size_t)h->fntn * h->fntiles * sizeof(int)));
    CK(cudaMemset(h->head_flag, 0, (size_t)h->fntn * h->fntiles * sizeof(int)));
    // [0, L): coop kernel, one per interval; L: SCG folds; L+1: tail kernel; L+2: walk kernel
    CK(cudaMalloc((void **)&h->tickets, (L + 4) * sizeof(unsigned)));
    CK(cudaMemset(h->tickets, 0, (L + 4) * sizeof(unsigned)));
    CK(cudaMalloc((void **)&h->obs_idx, D * sizeof(int32_t)));
    CK(cudaMalloc((void **)&h->flags, L * sizeof(int32_t)));

    CK(cudaMalloc((void **)&h->status_ws, sizeof(int32_t)));
    if (dev_alloc(&h->tail_part, (size_t)((nl + 7) / 8) * 4)) return MFVGP_ERR_CUDA;
    CK(cudaMalloc((void **)&h->tail_late, sizeof(int32_t)));
    CK(cudaMemset(h->tail_late, 0, sizeof(int32_t)));
    {
        const char *e = getenv("MFVGP_WALK_TAIL");
        h->walk_tail = !(e && e[0] == '0');
        const char *e2 = getenv("MFVGP_TWO_PASS");
        h->two_pass = !(e2 && e2[0] == '0');
        const char *e3 = getenv("MFVGP_WALK_PERSIST");
        h->walk_persist = e3 && e3[0] == '1';
    }
    CK(cudaMemset(h->flags, 0, L * sizeof(int32_t)));
    CK(cudaMemset(h->esde_n, 0, L * sizeof(double)));
    CK(cudaMallocHost((void **)&h->pin, 16 * sizeof(double)));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // the memsets above ran on the legacy stream; evaluations run on non-blocking streams,
    // which are not ordered after it
    CK(cudaDeviceSynchronize());
    guard.h = nullptr;
    *out = h;
    return MFVGP_OK;
}

extern "C" int mfvgp_destroy(mfvgp_handle *h) {
    if (!h) return MFVGP_OK;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->sigma, h->theta, h->mu0, h->tau0, h->obs_times, h->obs_values,
                    h->obs_noise, h->obs_idx, h->dm_ws, h->ds_ws, h->part, h->esde_n,
                    h->apart, h->out_ws, h->flags, h->status_ws, h->x_dev, h->grad_dev,
                    h->scg_vec[0], h->scg_vec[1], h->scg_vec[2], h->scg_vec[3], h->scg_vec[4],
                    h->scg_vec[5], h->scg_vec[6], h->scg_part, h->scg_scal,
                    h->refine_ws, h->refine_info, h->refine_status, h->edge_send, h->edge_recv,
                    h->gather, h->xchg_send, h->xchg_gather, h->rec_ws, h->fedge_m, h->fedge_s, h->fapart, h->tickets,
                    h->y_dense, h->rinv_row, h->head_flag, h->refine_gpart,
                    h->refine_gbar, h->b_dm, h->b_ds, h->b_part, h->b_apart, h->b_esde_n,
                    h->b_flags, h->b_rstat, h->b_ticket, h->eparam_ws, h->tail_part, h->tail_late,
                    h->run_start};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int r = 0; r < h->world && r < kPeerMaxWorld; ++r)
        if (h->peer_ok && r != h->rank && h->peer_box[r]) cudaIpcCloseMemHandle(h->peer_box[r]);
    if (h->mbox) cudaFree(h->mbox);
    if (h->peer_layout) cudaFree(h->peer_layout);
    if (h->peer_ticket) cudaFree(h->peer_ticket);
    if (h->comm && nccl().lib) nccl().CommDestroy(h->comm);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->scg_hostrec) cudaFreeHost(h->scg_hostrec);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MFVGP_OK;
}

extern "C" int mfvgp_set_sde(mfvgp_handle *h, const double *sigma_h, const double *theta_h,
                             int ntheta) {
    ARG(h && sigma_h && theta_h, "mfvgp_set_sde: NULL argument");
    const int need = h->system == MFVGP_SYS_OU ? 2 : (h->system == MFVGP_SYS_L63 ? 3 : 1);
    ARG(ntheta == need, "mfvgp_set_sde: wrong number of drift parameters");
    for (int i = 0; i < h->D; ++i)   // stochastic_process.py:112-115
        ARG(sigma_h[i] > 0.0, "SDE noise vector is not positive");
    CK(use_device(h->device));
    CK(cudaDeviceSynchronize());     // evaluations in flight (any stream) still read the old values
    CK(cudaMemcpy(h->sigma, sigma_h, h->D * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->theta, theta_h, ntheta * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());     // pageable uploads may return before their DMA has landed
    h->ntheta = ntheta;
    for (int i = 0; i < ntheta && i < 8; ++i) h->theta_host[i] = theta_h[i];
    h->have_sde = true;
    return MFVGP_OK;
}

extern "C" int mfvgp_set_prior(mfvgp_handle *h, const double *mu0_h, const double *tau0_h) {
    ARG(h && mu0_h && tau0_h, "mfvgp_set_prior: NULL argument");
    CK(use_device(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h->mu0, mu0_h, h->D * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->tau0, tau0_h, h->D * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());
    h->have_prior = true;
    return MFVGP_OK;
}

extern "C" int mfvgp_set_obs(mfvgp_handle *h, const double *obs_times_h,
                             const double *obs_values_h, const double *obs_noise_h,
                             const uint8_t *obs_mask_h) {
    ARG(h && obs_times_h && obs_values_h && obs_noise_h, "mfvgp_set_obs: NULL argument");
    std::vector<int32_t> idx(h->D, -1);
    int cnt = 0;
    for (int j = 0; j < h->D; ++j)
        if (obs_mask_h == nullptr || obs_mask_h[j]) idx[j] = cnt++;
    ARG(cnt == h->d, "mfvgp_set_obs: observation mask does not select d dimensions");
    // free_energy.py:612 -- num_M (dim_d log2pi + safe_log(prod(obs_noise)))
    double prod = 1.0;
    for (int i = 0; i < h->d; ++i) prod *= obs_noise_h[i];
    const double clamped = fmax(fmin(prod, 1.0e300), 1.0e-300);
    h->eobs_const = h->M * (h->d * 1.8378770664093453 + log(clamped));
    CK(use_device(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h->obs_times, obs_times_h, h->M * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->obs_values, obs_values_h, (size_t)h->M * h->d * sizeof(double),
                  cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->obs_noise, obs_noise_h, h->d * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->obs_idx, idx.data(), h->D * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (h->y_dense) {
        std::vector<double> yd((size_t)h->M * h->D, 0.0), ri(h->D, 0.0);
        for (int j = 0; j < h->D; ++j) {
            if (idx[j] < 0) continue;
            ri[j] = 1.0 / obs_noise_h[idx[j]];
            for (int k = 0; k < h->M; ++k)
                yd[(size_t)k * h->D + j] = obs_values_h[(size_t)k * h->d + idx[j]];
        }
        CK(cudaMemcpy(h->y_dense, yd.data(), yd.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->rinv_row, ri.data(), ri.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    CK(cudaDeviceSynchronize());
    h->have_obs = true;
    return MFVGP_OK;
}

// ---- launch helpers ---------------------------------------------------------
// launch that may overlap the tail of its predecessor in the stream (the kernel calls
// pdl_wait() before touching global data); MFVGP_PDL=0 falls back to a plain launch
static bool pdl_enabled() {
    static const bool enabled = [] {
        const char *e = getenv("MFVGP_PDL");
        return !(e && e[0] == '0');
    }();
    return enabled;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, unsigned block, cudaStream_t st,
                              Args &&...args) {
    const bool enabled = pdl_enabled();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// one thread-block cluster of `nctas` CTAs (the whole grid), programmatic dependent launch
template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kern)(KArgs...), unsigned nctas, unsigned block,
                                  cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nctas); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static RefineArgs refine_args(mfvgp_handle *h, const double *mean, int64_t ldm, const double *vars,
                              int64_t lds, int log_vars, double *ds, int delta_mode) {
    RefineArgs r;
    r.mean = mean; r.vars = vars; r.ldm = ldm; r.lds = lds;
    r.obs_times = h->obs_times; r.sigma = h->sigma; r.theta = h->theta;
    r.flags = h->flags; r.esde_n = h->esde_n; r.ds_out = ds; r.ws = h->refine_ws;
    r.nflagged = h->refine_status + 1;
    r.group = h->refine_group; r.gpart = h->refine_gpart; r.gbar = h->refine_gbar;
    r.delta_mode = delta_mode;
    r.info = h->refine_info; r.status = h->refine_status;
    r.D = h->D; r.n_begin = h->n_begin; r.n_end = h->n_end; r.log_vars = log_vars;
    r.want_grad = ds != nullptr;
    return r;
}

// adaptive path for the flagged (interval, integrand) pairs; exits at once when none is flagged
static int launch_refine(mfvgp_handle *h, const double *mean, int64_t ldm, const double *vars,
                         int64_t lds, int log_vars, double *ds, int delta_mode, cudaStream_t st) {
    const RefineArgs r = refine_args(h, mean, ldm, vars, lds, log_vars, ds, delta_mode);
    const dim3 grid(h->refine_grid);
    switch (h->system) {
        case MFVGP_SYS_DW: CK(launch_pdl(refine_kernel<0, true>, grid, kRefThreads, st, r)); break;
        case MFVGP_SYS_OU: CK(launch_pdl(refine_kernel<1, true>, grid, kRefThreads, st, r)); break;
        case MFVGP_SYS_L63: CK(launch_pdl(refine_kernel<2, true>, grid, kRefThreads, st, r)); break;
        default:
            if (h->refine_warp) CK(launch_pdl(refine_kernel<3, true>, grid, kRefThreads, st, r));
            else CK(launch_pdl(refine_kernel<3, false>, grid, kRefThreads, st, r));
    }
    h->launches++;
    return MFVGP_OK;
}

static int launch_guard_refine(mfvgp_handle *h, const double *mean, int64_t ldm,
                               const double *vars, int64_t lds, int log_vars, double *ds,
                               int ntiles, int delta_mode, cudaStream_t st) {
    const int nl = h->n_end - h->n_begin;
    GuardArgs g;
    g.part = h->part; g.obs_times = h->obs_times; g.esde_n = h->esde_n; g.flags = h->flags;
    g.n_begin = h->n_begin; g.n_end = h->n_end; g.ntiles = ntiles;
    g.nflagged = h->refine_status + 1;
    guard_reduce_kernel<<<(nl * 32 + 255) / 256, 256, 0, st>>>(g);
    h->launches++;
    CK(cudaGetLastError());
    return launch_refine(h, mean, ldm, vars, lds, log_vars, ds, delta_mode, st);
}

static int launch_esde(mfvgp_handle *h, const double *mean, int64_t ldm, const double *vars,
                       int64_t lds, int log_vars, double *dm, double *ds, cudaStream_t st,
                       bool skip_refine = false, bool guard_later = false) {
    EsdeArgs a;
    a.mean = mean; a.vars = vars; a.ldm = ldm; a.lds = lds;
    a.obs_times = h->obs_times; a.sigma = h->sigma; a.theta = h->theta;
    a.dm_out = dm; a.ds_out = ds; a.part = h->part;
    a.D = h->D; a.n_begin = h->n_begin; a.n_end = h->n_end; a.ntiles = h->ntiles;
    a.log_vars = log_vars;
    const int nl = h->n_end - h->n_begin;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->timing) {
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        CK(cudaEventRecord(ev0, st));
    }
    CoopArgs c;                               // small-problem kernels: guard inside
    if (h->coop) {
        c.a = a;
        c.g.part = h->part; c.g.obs_times = h->obs_times; c.g.esde_n = h->esde_n;
        c.g.flags = h->flags; c.g.n_begin = h->n_begin; c.g.n_end = h->n_end;
        c.g.ntiles = h->ntiles;
        c.g.nflagged = h->refine_status + 1;
        c.tickets = h->tickets; c.theta = h->theta_host[0];
        c.guard_later = guard_later ? 1 : 0;
    }
    switch (h->system) {
        case MFVGP_SYS_DW:
            if (h->coop) CK(launch_pdl(esde_small_inkernel<0>, dim3(nl), 64, st, c));
            else esde_small_kernel<0><<<nl, 64, 0, st>>>(a);
            break;
        case MFVGP_SYS_OU:
            if (h->coop) CK(launch_pdl(esde_small_inkernel<1>, dim3(nl), 64, st, c));
            else esde_small_kernel<1><<<nl, 64, 0, st>>>(a);
            break;
        case MFVGP_SYS_L63:
            if (h->coop) CK(launch_pdl(esde_small_inkernel<2>, dim3(nl), 64, st, c));
            else esde_small_kernel<2><<<nl, 64, 0, st>>>(a);
            break;
        default: {
            const unsigned grid = (unsigned)nl * h->ntiles;
            if (h->coop) {
                CK(launch_pdl(esde_l96_coop_kernel, dim3(h->ntiles, nl), kCoopNT, st, c));
            } else if (h->rule == 0) {
                if (h->te == 32) esde_l96_kernel<32><<<grid, 32, 0, st>>>(a);
                else if (h->te == 64) esde_l96_kernel<64><<<grid, 64, 0, st>>>(a);
                else esde_l96_kernel<128><<<grid, 128, 0, st>>>(a);
            } else {
                if (h->te == 32) esde_l96_split_kernel<32><<<grid, 32, 0, st>>>(a);
                else if (h->te == 64) esde_l96_split_kernel<64><<<grid, 64, 0, st>>>(a);
                else esde_l96_split_kernel<128><<<grid, 128, 0, st>>>(a);
            }
        }
    }
    h->launches++;
    CK(cudaGetLastError());
    if (h->timing) {
        CK(cudaEventRecord(ev1, st));
        h->timing_events.emplace_back(ev0, ev1);
    }
    if (h->coop)                           // the guard ran inside the kernel
        return skip_refine ? MFVGP_OK : launch_refine(h, mean, ldm, vars, lds, log_vars, ds, 0, st);
    return launch_guard_refine(h, mean, ldm, vars, lds, log_vars, ds, h->ntiles, 0, st);
}

template <int TE, bool PERSIST>
static int launch_walk_kernel_p(const WalkArgs &w, unsigned grid, cudaStream_t st) {
    static uint64_t configured = 0;                 // one bit per device (attribute is per device)
    const size_t smem = sizeof(WalkSmem<TE, walk_gb(TE)>);
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (!(configured >> (dev & 63) & 1)) {
        CK(cudaFuncSetAttribute(esde_l96_walk_kernel<TE, walk_gb(TE), PERSIST>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= 1ull << (dev & 63);
    }
    if (PERSIST) {                                  // one CTA per resident slot, at most one per run
        const unsigned resident = 148u * (unsigned)(MFVGP_WALK_THREADS_PER_SM / TE);
        if (grid > resident) grid = resident;
    }
    esde_l96_walk_kernel<TE, walk_gb(TE), PERSIST><<<grid, TE, smem, st>>>(w);
    return MFVGP_OK;
}

template <int TE>
static int launch_walk_kernel(const WalkArgs &w, unsigned grid, cudaStream_t st, bool persist) {
    return persist ? launch_walk_kernel_p<TE, true>(w, grid, st)
                   : launch_walk_kernel_p<TE, false>(w, grid, st);
}

// L96 walk path: E_cost in one main kernel (l96_walk.cuh) + guard / adaptive path follow-ups
static int launch_walk(mfvgp_handle *h, const double *x, double *grad, cudaStream_t st) {
    const double *lvar = x + (int64_t)h->D * h->Nm;
    double *gs = grad ? grad + (int64_t)h->D * h->Nm : nullptr;
    const unsigned grid = (unsigned)h->fntn * h->fntiles;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->timing) {
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        CK(cudaEventRecord(ev0, st));
    }
    WalkArgs w;
    w.x = x; w.obs_times = h->obs_times; w.sigma = h->sigma; w.theta = h->theta_host[0];
    w.mu0 = h->mu0; w.tau0 = h->tau0; w.obs_values = h->obs_values;
    w.obs_noise = h->obs_noise; w.obs_idx = h->obs_idx;
    w.gm = grad; w.gs = gs; w.head_m = h->fedge_m; w.head_s = h->fedge_s;
    w.head_flag = h->head_flag; w.run_ticket = h->tickets + h->L + 2;
    w.status = h->refine_status; w.epoch = ++h->walk_epoch;
    w.part = h->part; w.apart = h->fapart;
    w.D = h->D; w.M = h->M; w.d = h->d; w.Nm = h->Nm; w.Ns = h->Ns; w.L = h->L;
    w.n_begin = h->n_begin; w.n_end = h->n_end;
    w.ntiles = h->fntiles; w.nruns = h->fntn; w.run_start = h->run_start;
    w.with_e0 = h->n_begin == 0;
    int rc;
    if (h->fte == 32) rc = launch_walk_kernel<32>(w, grid, st, h->walk_persist);
    else if (h->fte == 64) rc = launch_walk_kernel<64>(w, grid, st, h->walk_persist);
    else rc = launch_walk_kernel<128>(w, grid, st, h->walk_persist);
    if (rc) return rc;
    h->launches++;
    CK(cudaGetLastError());
    if (h->timing) {
        CK(cudaEventRecord(ev1, st));
        h->timing_events.emplace_back(ev0, ev1);
    }
    if (h->walk_tail && (h->world == 1 || h->peer_ok)) return MFVGP_OK;   // launch_walk_tail follows
    // refined dS integrals come back as (refined - fixed) in ds_ws and are applied below
    rc = launch_guard_refine(h, x, h->Nm, lvar, h->Ns, 1, grad ? h->ds_ws : nullptr,
                             h->fntiles, 1, st);
    if (rc) return rc;
    if (grad != nullptr) {
        ApplyRefineArgs ar;
        ar.flags = h->flags; ar.nflagged = h->refine_status + 1; ar.delta = h->ds_ws; ar.x = x;
        ar.gs = gs;
        ar.D = h->D; ar.Nm = h->Nm; ar.Ns = h->Ns; ar.n_begin = h->n_begin; ar.n_end = h->n_end;
        apply_refine_kernel<<<h->n_end - h->n_begin + 1, 256, 0, st>>>(ar);
        h->launches++;
        CK(cudaGetLastError());
    }
    return MFVGP_OK;
}

#define NCK(call)                                                             \
    do {                                                                      \
        int r_ = (call);                                                      \
        if (r_ != kNcclSuccess) {                                             \
            g_err = std::string(#call) + ": " + nccl().GetErrorString(r_);    \
            return MFVGP_ERR_CUDA;                                            \
        }                                                                     \
    } while (0)

// one exchange over peer memory (peer.cuh): every rank's record to every rank, the shared
// gradient columns to the neighbours; folded in rank order on arrival
static int peer_exchange(mfvgp_handle *h, const double *rec, int n_rec, int n_sum, double *out,
                         int32_t *status_out, double *grad, cudaStream_t st) {
    PeerArgs a;
    for (int r = 0; r < kPeerMaxWorld; ++r) a.box[r] = r < h->world ? h->peer_box[r] : nullptr;
    a.slot_doubles = h->peer_slot; a.epoch = ++h->xchg_epoch; a.ticket = h->peer_ticket;
    a.status = h->refine_status; a.rec = rec; a.out = out; a.status_out = status_out;
    a.grad = grad; a.n_rec = n_rec; a.n_sum = n_sum;
    a.D = h->D; a.Nm = h->Nm; a.Ns = h->Ns; a.n_begin = h->n_begin; a.n_end = h->n_end;
    a.rank = h->rank; a.world = h->world;
    const unsigned grid = grad ? (unsigned)((h->D + 255) / 256) : 1u;
    CK(launch_pdl(peer_publish_kernel, dim3(grid), 256, st, a));
    CK(launch_pdl(peer_combine_kernel, dim3(grid), 256, st, a));
    h->launches += 2;
    return MFVGP_OK;
}

// all-gather `n` doubles from every rank and fold them in rank order
static int combine_over_ranks(mfvgp_handle *h, const double *local, int n, int nsum,
                              double *out, int32_t *status, cudaStream_t st) {
    if (h->peer_ok) return peer_exchange(h, local, n, nsum, out, status, nullptr, st);
    NCK(nccl().AllGather(local, h->gather, n, kNcclFloat64, h->comm, st));
    combine_record_kernel<<<1, 32, 0, st>>>(h->gather, h->world, n, nsum, out, status);
    h->launches++;
    CK(cudaGetLastError());
    return MFVGP_OK;
}

// grad != nullptr on a sharded handle: ONE all-gather carries the scalar record and both shared
// columns of every rank (replaces send/recv + all-gather: the collectives' latency is what
// limits the 8-GPU point)
static int launch_final(mfvgp_handle *h, const double *apart, int napart, double *out,
                        int32_t *status, cudaStream_t st, double *grad = nullptr) {
    FinalArgs f;
    const bool multi = h->world > 1;
    f.esde_n = h->esde_n; f.flags = h->flags; f.apart = apart;
    f.out = multi ? h->rec_ws : out; f.status = status;
    f.refine_status = h->refine_status;
    f.nflagged = h->refine_status + 1;
    f.gbar = h->refine_gbar; f.ngbar = h->refine_grid / h->refine_group;
    f.info = nullptr;
    f.record9 = multi ? 1 : 0;
    f.n_begin = h->n_begin; f.n_end = h->n_end; f.D = napart;
    f.with_e0 = h->n_begin == 0;
    f.eobs_const = h->n_begin == 0 ? h->eobs_const : 0.0;
    if (h->inject_status) {              // tests: a status bit raised on this rank only
        or_status_kernel<<<1, 1, 0, st>>>(h->refine_status, h->inject_status);
        CK(cudaGetLastError());
    }
    if (napart > 2048) final_kernel_wide<<<1, 1024, 0, st>>>(f);
    else final_kernel<<<1, 256, 0, st>>>(f);
    h->launches++;
    CK(cudaGetLastError());
    if (multi && grad != nullptr && h->peer_ok) {
        int rc = peer_exchange(h, h->rec_ws, kRecordLen, kRecordLen, h->rec_ws, status, grad, st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out, h->rec_ws, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return MFVGP_OK;
    }
    if (multi && grad != nullptr && h->one_collective) {
        XchgArgs x;
        x.grad = grad; x.rec = h->rec_ws; x.send = h->xchg_send; x.gathered = h->xchg_gather;
        x.out = out; x.status = status;
        x.D = h->D; x.Nm = h->Nm; x.Ns = h->Ns; x.n_begin = h->n_begin; x.n_end = h->n_end;
        x.rank = h->rank; x.world = h->world;
        const int nb = (h->D + 255) / 256;
        pack_all_kernel<<<nb, 256, 0, st>>>(x);
        h->launches++;
        CK(cudaGetLastError());
        NCK(nccl().AllGather(h->xchg_send, h->xchg_gather, 16 + 4 * (size_t)h->D, kNcclFloat64,
                             h->comm, st));
        combine_all_kernel<<<nb, 256, 0, st>>>(x);
        h->launches++;
        CK(cudaGetLastError());
        return MFVGP_OK;
    }
    if (multi) {
        // scalar free energy: every rank ends up with the same F, E0, Esde, Eobs
        int rc = combine_over_ranks(h, h->rec_ws, kRecordLen, kRecordLen, h->rec_ws, status, st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out, h->rec_ws, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return MFVGP_OK;
}

// gradient of the column shared with each neighbour: exchange the two partial
// sums over NVLink and add (both sides then hold identical bits)
static int exchange_edges(mfvgp_handle *h, double *grad, cudaStream_t st) {
    EdgeArgs e;
    e.grad = grad; e.buf = h->edge_send; e.D = h->D; e.Nm = h->Nm; e.Ns = h->Ns;
    e.n_begin = h->n_begin; e.n_end = h->n_end;
    e.has_left = h->rank > 0; e.has_right = h->rank < h->world - 1;
    const int nb = (h->D + 255) / 256;
    pack_edges_kernel<<<nb, 256, 0, st>>>(e);
    h->launches++;
    CK(cudaGetLastError());
    const size_t cnt = 2 * (size_t)h->D;
    NCK(nccl().GroupStart());
    if (e.has_left) {
        NCK(nccl().Send(h->edge_send, cnt, kNcclFloat64, h->rank - 1, h->comm, st));
        NCK(nccl().Recv(h->edge_recv, cnt, kNcclFloat64, h->rank - 1, h->comm, st));
    }
    if (e.has_right) {
        NCK(nccl().Send(h->edge_send + cnt, cnt, kNcclFloat64, h->rank + 1, h->comm, st));
        NCK(nccl().Recv(h->edge_recv + cnt, cnt, kNcclFloat64, h->rank + 1, h->comm, st));
    }
    NCK(nccl().GroupEnd());
    e.buf = h->edge_recv;
    add_edges_kernel<<<nb, 256, 0, st>>>(e);
    h->launches++;
    CK(cudaGetLastError());
    return MFVGP_OK;
}

// walk path: everything behind the main kernel (walk_tail.cuh) -- guard + scalar sums + record
// and shared columns to the peers in one kernel, the adaptive path's kernels (no-ops unless an
// interval is flagged), and on sharded handles the kernel that waits for the peers and folds
static int launch_walk_tail(mfvgp_handle *h, const double *x, double *grad, double *out,
                            int32_t *status, cudaStream_t st) {
    const int nl = h->n_end - h->n_begin;
    const bool multi = h->world > 1;
    const double *lvar = x + (int64_t)h->D * h->Nm;
    double *gs = grad ? grad + (int64_t)h->D * h->Nm : nullptr;
    if (h->inject_status) {              // tests: a status bit raised on this rank only
        or_status_kernel<<<1, 1, 0, st>>>(h->refine_status, h->inject_status);
        CK(cudaGetLastError());
    }
    WalkLateArgs la;
    WalkTailArgs &t = la.t;
    t.g.part = h->part; t.g.obs_times = h->obs_times; t.g.esde_n = h->esde_n; t.g.flags = h->flags;
    t.g.n_begin = h->n_begin; t.g.n_end = h->n_end; t.g.ntiles = h->fntiles;
    t.g.nflagged = h->refine_status + 1;
    t.apart = h->fapart; t.napart = h->fntn * h->fntiles;
    t.tpart = h->tail_part; t.ticket = h->tickets + h->L + 3; t.late = h->tail_late;
    t.out = multi ? h->rec_ws : out; t.status = status; t.refine_status = h->refine_status;
    t.gbar = h->refine_gbar; t.ngbar = h->refine_grid / h->refine_group;
    t.with_e0 = h->n_begin == 0;
    t.eobs_const = h->n_begin == 0 ? h->eobs_const : 0.0;
    t.guard_ctas = (nl + 7) / 8;
    t.nchunk = h->peer_nchunk;
    PeerArgs &p = t.p;
    for (int r = 0; r < kPeerMaxWorld; ++r) p.box[r] = r < h->world ? h->peer_box[r] : nullptr;
    p.slot_doubles = h->peer_slot; p.epoch = multi ? ++h->xchg_epoch : 0; p.ticket = nullptr;
    p.status = h->refine_status; p.rec = h->rec_ws; p.out = h->rec_ws; p.status_out = status;
    p.grad = multi ? grad : nullptr; p.n_rec = kPeerRec; p.n_sum = kPeerRec;
    p.D = h->D; p.Nm = h->Nm; p.Ns = h->Ns; p.n_begin = h->n_begin; p.n_end = h->n_end;
    p.rank = h->rank; p.world = h->world;
    const int edge_ctas = (multi && grad) ? h->peer_nchunk : 0;
    CK(launch_pdl(walk_tail_kernel, dim3(t.guard_ctas + edge_ctas), 256, st, t));
    h->launches++;
    // adaptive path: refined dS integrals come back as (refined - fixed) in ds_ws
    int rc = launch_refine(h, x, h->Nm, lvar, h->Ns, 1, grad ? h->ds_ws : nullptr, 1, st);
    if (rc) return rc;
    if (grad != nullptr) {
        ApplyRefineArgs ar;
        ar.flags = h->flags; ar.nflagged = h->refine_status + 1; ar.delta = h->ds_ws; ar.x = x;
        ar.gs = gs;
        ar.D = h->D; ar.Nm = h->Nm; ar.Ns = h->Ns; ar.n_begin = h->n_begin; ar.n_end = h->n_end;
        CK(launch_pdl(apply_refine_kernel, dim3(nl + 1), 256, st, ar));
        h->launches++;
    }
    la.esde_n = h->esde_n; la.flags = h->flags;
    CK(launch_pdl(walk_late_kernel, dim3(1 + edge_ctas), 256, st, la));
    h->launches++;
    if (multi) {
        CK(launch_pdl(walk_combine_kernel, dim3(grad ? h->peer_nchunk : 1), 256, st, p,
                      h->peer_nchunk, out));
        h->launches++;
    }
    return MFVGP_OK;
}

extern "C" int mfvgp_comm_unique_id(void *id128_out) {
    ARG(id128_out != nullptr, "mfvgp_comm_unique_id: NULL argument");
    if (!nccl().load()) { g_err = "NCCL (libnccl.so.2) could not be loaded"; return MFVGP_ERR_CUDA; }
    ncclUniqueId id;
    NCK(nccl().GetUniqueId(&id));
    memcpy(id128_out, &id, sizeof(id));
    return MFVGP_OK;
}

extern "C" int mfvgp_comm_init(mfvgp_handle *h, int rank, int world, const void *id128) {
    ARG(h && id128 && world >= 1 && rank >= 0 && rank < world, "mfvgp_comm_init: bad argument");
    if (world == 1) return MFVGP_OK;
    if (!nccl().load()) { g_err = "NCCL (libnccl.so.2) could not be loaded"; return MFVGP_ERR_CUDA; }
    ARG((rank == 0) == (h->n_begin == 0) && (rank == world - 1) == (h->n_end == h->L),
        "mfvgp_comm_init: the handle's interval range does not match its rank");
    CK(use_device(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCK(nccl().CommInitRank(&h->comm, world, id, rank));
    h->rank = rank; h->world = world;
    if (const char *e = getenv("MFVGP_INJECT_STATUS")) {
        int bits = 0, at = -1;
        const int nf = sscanf(e, "%d@%d", &bits, &at);
        if (nf >= 1 && (nf == 1 || at == rank)) h->inject_status = bits;
    }
    {
        const char *e = getenv("MFVGP_XCHG");
        h->one_collective = !(e && std::string(e) == "split");
    }
    {
        // peer mailboxes (peer.cuh): allocate, exchange the IPC handles through the communicator,
        // map every peer's.  MFVGP_XCHG=nccl|split keeps the collectives; so does any failure
        // here (e.g. ranks that cannot map each other's memory), reported by mfvgp_describe.
        const char *e = getenv("MFVGP_XCHG");
        const bool want = world <= kPeerMaxWorld && !(e && (std::string(e) == "nccl" || std::string(e) == "split"));
        if (want) {
            h->peer_nchunk = (h->D + 255) / 256;
            h->peer_slot = ((int64_t)world * kPeerRec + 4 * (int64_t)h->D + world + 2
                            + 2 * (int64_t)h->peer_nchunk + 1) / 2 * 2;
            const size_t bytes = 2 * (size_t)h->peer_slot * sizeof(double);
            CK(cudaMalloc((void **)&h->mbox, bytes));
            CK(cudaMemset(h->mbox, 0, bytes));
            CK(cudaMalloc((void **)&h->peer_ticket, sizeof(unsigned)));
            CK(cudaMemset(h->peer_ticket, 0, sizeof(unsigned)));
            cudaIpcMemHandle_t mine;
            std::vector<cudaIpcMemHandle_t> all(world);
            bool ok = cudaIpcGetMemHandle(&mine, h->mbox) == cudaSuccess;
            if (!ok) cudaGetLastError();
            // 64 handle bytes + 1 "ok" word per rank, gathered as doubles
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
            double *hs = nullptr, *hr = nullptr;
            if (dev_alloc(&hs, 9) || dev_alloc(&hr, 9 * (size_t)world)) return MFVGP_ERR_CUDA;
            double pack[9];
            memcpy(pack, &mine, 64);
            pack[8] = ok ? 1.0 : 0.0;
            CK(cudaMemcpy(hs, pack, sizeof(pack), cudaMemcpyHostToDevice));
            CK(cudaDeviceSynchronize());
            NCK(nccl().AllGather(hs, hr, 9, kNcclFloat64, h->comm, nullptr));
            CK(cudaDeviceSynchronize());
            std::vector<double> got(9 * (size_t)world);
            CK(cudaMemcpy(got.data(), hr, got.size() * sizeof(double), cudaMemcpyDeviceToHost));
            cudaFree(hs);
            cudaFree(hr);
            for (int r = 0; r < world; ++r) ok = ok && got[9 * r + 8] == 1.0;
            for (int r = 0; r < world && ok; ++r) {
                if (r == rank) { h->peer_box[r] = h->mbox; continue; }
                cudaIpcMemHandle_t hd;
                memcpy(&hd, &got[9 * r], 64);
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                h->peer_box[r] = static_cast<double *>(ptr);
            }
            // every rank must take the same path: agree on the outcome
            double *fs = nullptr, *fr = nullptr;
            if (dev_alloc(&fs, 1) || dev_alloc(&fr, (size_t)world)) return MFVGP_ERR_CUDA;
            const double mineok = ok ? 1.0 : 0.0;
            CK(cudaMemcpy(fs, &mineok, sizeof(double), cudaMemcpyHostToDevice));
            CK(cudaDeviceSynchronize());
            NCK(nccl().AllGather(fs, fr, 1, kNcclFloat64, h->comm, nullptr));
            CK(cudaDeviceSynchronize());
            std::vector<double> oks(world);
            CK(cudaMemcpy(oks.data(), fr, world * sizeof(double), cudaMemcpyDeviceToHost));
            cudaFree(fs);
            cudaFree(fr);
            bool all_ok = true;
            for (int r = 0; r < world; ++r) all_ok = all_ok && oks[r] == 1.0;
            if (!all_ok) {
                for (int r = 0; r < world; ++r)
                    if (r != rank && h->peer_box[r]) { cudaIpcCloseMemHandle(h->peer_box[r]); h->peer_box[r] = nullptr; }
            }
            h->peer_ok = all_ok;
            if (all_ok) {
                PeerLayout pl;
                for (int r = 0; r < kPeerMaxWorld; ++r) pl.box[r] = r < world ? h->peer_box[r] : nullptr;
                pl.slot_doubles = h->peer_slot; pl.rank = rank; pl.world = world; pl.D = h->D;
                CK(cudaMalloc((void **)&h->peer_layout, sizeof(PeerLayout)));
                CK(cudaMemcpy(h->peer_layout, &pl, sizeof(PeerLayout), cudaMemcpyHostToDevice));
                CK(cudaDeviceSynchronize());
            }
        }
    }
    if (dev_alloc(&h->edge_send, 4 * (size_t)h->D) || dev_alloc(&h->edge_recv, 4 * (size_t)h->D) ||
        dev_alloc(&h->gather, 16 * (size_t)world) ||
        dev_alloc(&h->xchg_send, 16 + 4 * (size_t)h->D) ||
        dev_alloc(&h->xchg_gather, (size_t)world * (16 + 4 * (size_t)h->D)))
        return MFVGP_ERR_CUDA;
    return MFVGP_OK;
}

extern "C" int mfvgp_esde(mfvgp_handle *h, const double *mean_pts_d, int64_t ldm,
                          const double *vars_pts_d, int64_t lds, int want_grad, double *out_d,
                          double *dEsde_dm_d, double *dEsde_ds_d, int32_t *status_d,
                          void *cuda_stream) {
    ARG(h && mean_pts_d && vars_pts_d && out_d && status_d, "mfvgp_esde: NULL argument");
    ARG(!want_grad || (dEsde_dm_d && dEsde_ds_d), "mfvgp_esde: gradient buffers are NULL");
    if (!h->have_sde || !h->have_obs) {
        g_err = "mfvgp_esde: set_sde/set_obs not called";
        return MFVGP_ERR_STATE;
    }
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = launch_esde(h, mean_pts_d, ldm, vars_pts_d, lds, 0,
                         want_grad ? dEsde_dm_d : nullptr, want_grad ? dEsde_ds_d : nullptr, st);
    if (rc) return rc;
    // E_sde alone: out[0] = Esde (reported through the F slot as well)
    rc = launch_final(h, nullptr, 0, h->out_ws, status_d, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_d, h->out_ws + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
    return MFVGP_OK;
}

// optimistic: small problems only -- skip the refine_kernel launch.  The guard still flags the
// intervals and the status word still carries MFVGP_ST_REFINED, so a caller that reads the
// status before using the result (the host-buffer entries, the SCG loop) can evaluate
// optimistically and redo the rare flagged evaluation with the adaptive path.
static int ecost_impl(mfvgp_handle *h, const double *x_d, int want_grad, double *out_d,
                      double *grad_d, int32_t *status_d, void *cuda_stream, bool optimistic);

extern "C" int mfvgp_ecost(mfvgp_handle *h, const double *x_d, int want_grad, double *out_d,
                           double *grad_d, int32_t *status_d, void *cuda_stream) {
    return ecost_impl(h, x_d, want_grad, out_d, grad_d, status_d, cuda_stream, false);
}

static int ecost_impl(mfvgp_handle *h, const double *x_d, int want_grad, double *out_d,
                      double *grad_d, int32_t *status_d, void *cuda_stream, bool optimistic) {
    ARG(h && x_d && out_d && status_d, "mfvgp_ecost: NULL argument");
    ARG(!want_grad || grad_d, "mfvgp_ecost: gradient buffer is NULL");
    if (!h->have_sde || !h->have_obs || !h->have_prior) {
        g_err = "mfvgp_ecost: set_sde/set_prior/set_obs not called";
        return MFVGP_ERR_STATE;
    }
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const double *mean = x_d;
    const double *lvar = x_d + (int64_t)h->D * h->Nm;
    int rc;
    if (h->fused) {
        rc = launch_walk(h, x_d, want_grad ? grad_d : nullptr, st);
        if (rc) return rc;
        if (h->walk_tail && (h->world == 1 || h->peer_ok))
            return launch_walk_tail(h, x_d, want_grad ? grad_d : nullptr, out_d, status_d, st);
        if (h->world > 1 && want_grad && !h->one_collective && !h->peer_ok) {
            rc = exchange_edges(h, grad_d, st);
            if (rc) return rc;
        }
        return launch_final(h, h->fapart, h->fntn * h->fntiles, out_d, status_d, st,
                            want_grad ? grad_d : nullptr);
    }
    // optimistic L96 evaluations: no refine launch, and the guard moves from the main kernel's
    // last CTAs (ticket + fold on its critical path) into CTAs of the tail kernel
    const bool opt = optimistic && h->coop && h->world == 1;
    // Lorenz 96 on device pointers (nobody reads the status before using the result): two
    // passes.  The first is the optimistic one; the adaptive kernel and a second tail pass
    // follow in the stream and return at once unless the first pass flagged an interval.
    const bool two_pass = !optimistic && h->coop && h->world == 1 && h->system == MFVGP_SYS_L96 &&
                          h->guard_in_tail && h->two_pass;
    const bool guard_later = (opt || two_pass) && h->system == MFVGP_SYS_L96 && h->guard_in_tail;
    rc = launch_esde(h, mean, h->Nm, lvar, h->Ns, 1, want_grad ? h->dm_ws : nullptr,
                     want_grad ? h->ds_ws : nullptr, st, opt || two_pass, guard_later);
    if (rc) return rc;
    AssembleArgs a;
    a.x = x_d; a.dm = want_grad ? h->dm_ws : nullptr; a.ds = want_grad ? h->ds_ws : nullptr;
    a.mu0 = h->mu0; a.tau0 = h->tau0; a.obs_values = h->obs_values; a.obs_noise = h->obs_noise;
    a.obs_idx = h->obs_idx; a.grad = want_grad ? grad_d : nullptr; a.apart = h->apart;
    a.D = h->D; a.M = h->M; a.d = h->d; a.Nm = h->Nm; a.Ns = h->Ns;
    a.n_begin = h->n_begin; a.n_end = h->n_end; a.L = h->L; a.with_e0 = h->n_begin == 0;
    if (h->coop && h->world == 1) {
        // small problems: assembly and the final scalar sums in one launch
        TailArgs t;
        t.a = a;
        t.y_dense = h->y_dense; t.rinv_row = h->rinv_row;
        const int nb = h->tail_grid;
        t.f.esde_n = h->esde_n; t.f.flags = h->flags; t.f.apart = h->apart;
        t.f.out = out_d; t.f.status = status_d; t.f.refine_status = h->refine_status;
        t.f.nflagged = h->refine_status + 1;
        t.f.gbar = h->refine_gbar; t.f.ngbar = h->refine_grid / h->refine_group;
        t.f.info = h->refine_info;
        t.f.record9 = 0; t.f.n_begin = h->n_begin; t.f.n_end = h->n_end; t.f.D = nb;
        t.f.with_e0 = 1; t.f.eobs_const = h->eobs_const;
        t.ticket = h->tickets + h->L + 1;
        t.guard_ctas = 0;
        if (guard_later) {
            t.g.part = h->part; t.g.obs_times = h->obs_times; t.g.esde_n = h->esde_n;
            t.g.flags = h->flags; t.g.n_begin = h->n_begin; t.g.n_end = h->n_end;
            t.g.ntiles = h->ntiles; t.g.nflagged = h->refine_status + 1;
            const int per_cta = kTailThreads / 32, nl = h->n_end - h->n_begin;
            t.guard_ctas = (nl + per_cta - 1) / per_cta;
            if (t.guard_ctas > 64) t.guard_ctas = 64;
        }
        if (two_pass) t.f.late = h->tail_late;
        CK(launch_pdl(assemble_tail_kernel, dim3(nb + t.guard_ctas), kTailThreads, st, t));
        h->launches++;
        CK(cudaGetLastError());
        if (two_pass) {
            rc = launch_refine(h, mean, h->Nm, lvar, h->Ns, 1, want_grad ? h->ds_ws : nullptr, 0, st);
            if (rc) return rc;
            t.f.late = nullptr; t.f.late_clear = h->tail_late; t.run_if = h->tail_late;
            t.guard_ctas = 0;
            CK(launch_pdl(assemble_tail_kernel, dim3(nb), kTailThreads, st, t));
            h->launches++;
            CK(cudaGetLastError());
        }
        return MFVGP_OK;
    }
    assemble_kernel<<<h->D, 256, 0, st>>>(a);
    h->launches++;
    CK(cudaGetLastError());
    if (h->world > 1 && want_grad && !h->one_collective && !h->peer_ok) {
        rc = exchange_edges(h, grad_d, st);
        if (rc) return rc;
    }
    return launch_final(h, h->apart, h->D, out_d, status_d, st, want_grad ? grad_d : nullptr);
}

// ---- batched evaluation ---------------------------------------------------------------------
// B independent states of the SAME problem in two launches (small-problem path: DW, OU, L63,
// Lorenz 96 up to 37 888 cells): what scipy.optimize.check_grad behind find_minimum(check_gradients=True)
// (free_energy.py:805, :858) asks for -- len(x) + 1 evaluations -- and what fills the GPU at the
// north-star size, where ONE evaluation is 51 cells per SM.  Fixed rule + guard only: an entry
// whose guard flags an interval comes back with MFVGP_ST_REFINED and must be re-evaluated with
// mfvgp_ecost (adaptive stage); the Python wrapper does that.
extern "C" int mfvgp_ecost_batch(mfvgp_handle *h, int B, const double *x_d, int want_grad,
                                 double *out_d, double *grad_d, int32_t *status_d,
                                 void *cuda_stream) {
    ARG(h && x_d && out_d && status_d, "mfvgp_ecost_batch: NULL argument");
    ARG(B >= 1 && B <= 65535, "mfvgp_ecost_batch: 1 <= B <= 65535");
    ARG(!want_grad || grad_d, "mfvgp_ecost_batch: gradient buffer is NULL");
    if (!h->have_sde || !h->have_obs || !h->have_prior) {
        g_err = "mfvgp_ecost_batch: set_sde/set_prior/set_obs not called";
        return MFVGP_ERR_STATE;
    }
    ARG(h->coop && h->world == 1,
        "mfvgp_ecost_batch: small-problem path (cooperative kernels), one GPU");
    const bool l96 = h->system == MFVGP_SYS_L96;
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int D = h->D, L = h->L, nl = h->n_end - h->n_begin, nb = h->tail_grid;
    const int per_cta = kTailThreads / 32;
    int guard_ctas = (nl + per_cta - 1) / per_cta;
    if (guard_ctas > 64) guard_ctas = 64;
    if (!l96) guard_ctas = 0;          // DW / OU / L63: the guard runs inside the main kernel
    // which main kernel: the cooperative one (two lanes per cell, 26 dimensions per CTA) is built
    // for the latency of ONE evaluation; once the batch alone fills the GPU the one-thread-per-
    // cell split kernel does the same work in a third of the time.  MFVGP_BATCH_KERNEL=coop|split
    // ("lean": the same shape with the walk kernel's instruction count, l96_lean.cuh; "split":
    // the independently written esde_l96_split_body, kept as the cross-check)
    bool split = l96 && (int64_t)B * D * nl >= 148 * 2048;
    bool lean = true;
    if (l96) {
        const char *e = getenv("MFVGP_BATCH_KERNEL");
        if (e && e[0] == 'c') split = false;
        if (e && e[0] == 's') { split = true; lean = false; }
        if (e && e[0] == 'l') split = true;
    }
    // tile width with the fewest padded threads (D=500: 9 x 64 = 576 against 5 x 128 = 640)
    int te = ((D + 121) / 122) * 128 <= ((D + 57) / 58) * 64 ? 128 : 64;
    {
        const char *e = getenv("MFVGP_BATCH_TE");
        if (e && atoi(e) == 64) te = 64;
        if (e && atoi(e) == 128) te = 128;
    }
    const int ntiles = split ? (D + te - 7) / (te - 6) : h->ntiles;
    const int max_tiles = (D + 32 - 7) / (32 - 6) > h->ntiles ? (D + 32 - 7) / (32 - 6) : h->ntiles;
    BatchStrides bs;
    bs.x = (int64_t)D * (h->Nm + h->Ns);
    bs.dm = (int64_t)L * D * 4; bs.ds = (int64_t)L * D * 3;
    bs.part = (int64_t)nl * max_tiles * kPartStride;
    bs.apart = (int64_t)nb * 4; bs.esde_n = L; bs.out = 8;
    if (B > h->batch_cap) {
        void *old[] = {h->b_dm, h->b_ds, h->b_part, h->b_apart, h->b_esde_n, h->b_flags,
                       h->b_rstat, h->b_ticket};
        CK(cudaStreamSynchronize(st));
        for (void *p : old)
            if (p) cudaFree(p);
        h->batch_cap = 0;
        h->b_dm = h->b_ds = h->b_part = h->b_apart = h->b_esde_n = nullptr;
        h->b_flags = h->b_rstat = nullptr; h->b_ticket = nullptr;
        if (dev_alloc(&h->b_dm, (size_t)B * bs.dm) || dev_alloc(&h->b_ds, (size_t)B * bs.ds) ||
            dev_alloc(&h->b_part, (size_t)B * bs.part) || dev_alloc(&h->b_apart, (size_t)B * bs.apart) ||
            dev_alloc(&h->b_esde_n, (size_t)B * bs.esde_n))
            return MFVGP_ERR_CUDA;
        CK(cudaMalloc((void **)&h->b_flags, (size_t)B * L * sizeof(int32_t)));
        CK(cudaMalloc((void **)&h->b_rstat, (size_t)B * 2 * sizeof(int32_t)));
        CK(cudaMalloc((void **)&h->b_ticket, (size_t)B * sizeof(unsigned)));
        CK(cudaMemsetAsync(h->b_flags, 0, (size_t)B * L * sizeof(int32_t), st));
        CK(cudaMemsetAsync(h->b_rstat, 0, (size_t)B * 2 * sizeof(int32_t), st));
        CK(cudaMemsetAsync(h->b_ticket, 0, (size_t)B * sizeof(unsigned), st));
        h->batch_cap = B;
    }
    CoopArgs c;
    c.a.mean = x_d; c.a.vars = x_d + (int64_t)D * h->Nm; c.a.ldm = h->Nm; c.a.lds = h->Ns;
    c.a.obs_times = h->obs_times; c.a.sigma = h->sigma; c.a.theta = h->theta;
    c.a.dm_out = want_grad ? h->b_dm : nullptr; c.a.ds_out = want_grad ? h->b_ds : nullptr;
    c.a.part = h->b_part;
    c.a.D = D; c.a.n_begin = h->n_begin; c.a.n_end = h->n_end; c.a.ntiles = ntiles;
    c.a.log_vars = 1;
    c.g.part = h->b_part; c.g.obs_times = h->obs_times; c.g.esde_n = h->b_esde_n;
    c.g.flags = h->b_flags; c.g.n_begin = h->n_begin; c.g.n_end = h->n_end;
    c.g.ntiles = ntiles; c.g.nflagged = h->b_rstat + 1;
    c.tickets = h->tickets; c.theta = h->theta_host[0]; c.guard_later = 1;
    c.a.log_vars = 1;
    if (h->system == MFVGP_SYS_DW) {
        CK(launch_pdl(esde_small_batch_kernel<0>, dim3(nl, B), 64, st, c, bs, L));
    } else if (h->system == MFVGP_SYS_OU) {
        CK(launch_pdl(esde_small_batch_kernel<1>, dim3(nl, B), 64, st, c, bs, L));
    } else if (h->system == MFVGP_SYS_L63) {
        CK(launch_pdl(esde_small_batch_kernel<2>, dim3(nl, B), 64, st, c, bs, L));
    } else if (!split) {
        CK(launch_pdl(esde_l96_coop_batch_kernel, dim3(ntiles, nl, B), kCoopNT, st, c, bs));
    } else if (lean && te == 128) {
        esde_l96_lean_batch_kernel<128><<<dim3(nl * ntiles, B), 128, 0, st>>>(c.a, bs);
    } else if (lean) {
        esde_l96_lean_batch_kernel<64><<<dim3(nl * ntiles, B), 64, 0, st>>>(c.a, bs);
    } else if (te == 128) {
        esde_l96_split_batch_kernel<128><<<dim3(nl * ntiles, B), 128, 0, st>>>(c.a, bs);
    } else {
        esde_l96_split_batch_kernel<64><<<dim3(nl * ntiles, B), 64, 0, st>>>(c.a, bs);
    }
    CK(cudaGetLastError());
    h->launches++;
    TailArgs t;
    AssembleArgs &a = t.a;
    a.x = x_d; a.dm = c.a.dm_out; a.ds = c.a.ds_out;
    a.mu0 = h->mu0; a.tau0 = h->tau0; a.obs_values = h->obs_values; a.obs_noise = h->obs_noise;
    a.obs_idx = h->obs_idx; a.grad = want_grad ? grad_d : nullptr; a.apart = h->b_apart;
    a.D = D; a.M = h->M; a.d = h->d; a.Nm = h->Nm; a.Ns = h->Ns;
    a.n_begin = h->n_begin; a.n_end = h->n_end; a.L = L; a.with_e0 = 1;
    t.y_dense = h->y_dense; t.rinv_row = h->rinv_row;
    t.f.esde_n = h->b_esde_n; t.f.flags = h->b_flags; t.f.apart = h->b_apart;
    t.f.out = out_d; t.f.status = status_d; t.f.refine_status = h->b_rstat;
    t.f.nflagged = h->b_rstat + 1;
    t.f.gbar = h->refine_gbar; t.f.ngbar = 0; t.f.info = nullptr;
    // entry CTAs: one entry per thread beside the cooperative kernel (rows stay bit-identical
    // to single evaluations), eight per thread beside the split kernel (CTA count, not bytes,
    // is what bounds this kernel across a large batch; MFVGP_BATCH_TAIL_EPT)
    int ept = 8;
    {
        const char *e = getenv("MFVGP_BATCH_TAIL_EPT");
        if (e && atoi(e) >= 1 && atoi(e) <= 64) ept = atoi(e);
    }
    const int nbe = split ? (nb + ept - 1) / ept : nb;
    t.f.record9 = 0; t.f.n_begin = h->n_begin; t.f.n_end = h->n_end; t.f.D = nbe;
    t.f.with_e0 = 1; t.f.eobs_const = h->eobs_const;
    t.ticket = h->b_ticket;
    t.g = c.g;
    t.guard_ctas = guard_ctas;
    CK(launch_pdl(assemble_tail_batch_kernel, dim3(nbe + guard_ctas, B), kTailThreads, st, t, bs, L));
    h->launches++;
    CK(cudaGetLastError());
    return MFVGP_OK;
}

static int ensure_host_staging(mfvgp_handle *h) {
    const size_t nx = (size_t)h->D * (h->Nm + h->Ns);
    if (!h->x_dev) {
        // grad_dev: gradient followed by [F, E0, Esde, Eobs, status] (packed host entry)
        if (dev_alloc(&h->x_dev, nx) || dev_alloc(&h->grad_dev, nx + 8)) return MFVGP_ERR_CUDA;
    }
    return MFVGP_OK;
}

// ---- page-locked host pool ------------------------------------------------------------
namespace {
struct HostPool {
    std::mutex mu;
    std::multimap<int64_t, void *> free_blocks;      // size -> block
    std::map<void *, int64_t> live;                   // block -> size
    int64_t cached_bytes = 0;
};
HostPool &host_pool() {
    static HostPool *p = new HostPool();              // never destroyed: outlives the CUDA context
    return *p;
}
constexpr int64_t kHostPoolCap = 1LL << 30;           // keep at most 1 GiB of idle blocks
}  // namespace

extern "C" void *mfvgp_host_alloc(int64_t bytes) {
    if (bytes <= 0) bytes = 8;
    bytes = (bytes + 4095) / 4096 * 4096;
    HostPool &hp = host_pool();
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        auto it = hp.free_blocks.find(bytes);
        if (it != hp.free_blocks.end()) {
            void *p = it->second;
            hp.free_blocks.erase(it);
            hp.cached_bytes -= bytes;
            hp.live[p] = bytes;
            return p;
        }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        g_err = "mfvgp_host_alloc: cudaMallocHost failed";
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(hp.mu);
    hp.live[p] = bytes;
    return p;
}

extern "C" void mfvgp_host_free(void *p) {
    if (!p) return;
    HostPool &hp = host_pool();
    int64_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        auto it = hp.live.find(p);
        if (it == hp.live.end()) return;
        bytes = it->second;
        hp.live.erase(it);
        if (hp.cached_bytes + bytes <= kHostPoolCap) {
            hp.free_blocks.emplace(bytes, p);
            hp.cached_bytes += bytes;
            return;
        }
    }
    cudaFreeHost(p);
}

extern "C" int mfvgp_ecost_host_packed(mfvgp_handle *h, const double *x_h, int want_grad,
                                       double *buf_h) {
    ARG(h && x_h && buf_h, "mfvgp_ecost_host_packed: NULL argument");
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const size_t nx = (size_t)h->D * (h->Nm + h->Ns);
    CK(cudaMemcpyAsync(h->x_dev, x_h, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // Small-problem path: assemble_tail_kernel writes every gradient entry exactly once, in
    // index order, and never reads it back -- so it can store straight into the caller's
    // page-locked buffer over PCIe (zero-copy), together with the scalars: no D2H copy.
    bool direct = false;
    if (h->coop && h->world == 1) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, buf_h) == cudaSuccess)
            direct = attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
        else
            cudaGetLastError();
        if (direct) {
            double *buf_d = static_cast<double *>(attr.devicePointer);
            // optimistic (while flagged intervals are rare): no refine_kernel launch; the
            // status word comes back with the result
            const bool optimistic = h->opt_cooldown == 0;
            rc = ecost_impl(h, h->x_dev, want_grad, buf_d + nx, buf_d,
                            reinterpret_cast<int32_t *>(buf_d + nx + 4), h->stream, optimistic);
            if (rc) return rc;
            CK(cudaStreamSynchronize(h->stream));
            int32_t st0;
            memcpy(&st0, buf_h + nx + 4, sizeof(st0));
            if (st0 & MFVGP_ST_REFINED) h->opt_cooldown = 16;
            else if (h->opt_cooldown > 0) h->opt_cooldown--;
            if (!optimistic || !(st0 & MFVGP_ST_REFINED)) return MFVGP_OK;
            // the guard flagged an interval: evaluate again, this time with the adaptive path
            rc = ecost_impl(h, h->x_dev, want_grad, buf_d + nx, buf_d,
                            reinterpret_cast<int32_t *>(buf_d + nx + 4), h->stream, false);
            if (rc) return rc;
        }
    }
    if (!direct) {
        double *out_d = h->grad_dev + nx;
        rc = mfvgp_ecost(h, h->x_dev, want_grad, out_d, h->grad_dev,
                         reinterpret_cast<int32_t *>(out_d + 4), h->stream);
        if (rc) return rc;
        if (want_grad)
            CK(cudaMemcpyAsync(buf_h, h->grad_dev, (nx + 5) * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
        else
            CK(cudaMemcpyAsync(buf_h + nx, out_d, 5 * sizeof(double), cudaMemcpyDeviceToHost,
                               h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return MFVGP_OK;
}

extern "C" int mfvgp_ecost_host(mfvgp_handle *h, const double *x_h, int want_grad,
                                double *out_h, double *grad_h, int32_t *status_h) {
    ARG(h && x_h && out_h && status_h, "mfvgp_ecost_host: NULL argument");
    ARG(!want_grad || grad_h, "mfvgp_ecost_host: gradient buffer is NULL");
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const size_t nx = (size_t)h->D * (h->Nm + h->Ns);
    CK(cudaMemcpyAsync(h->x_dev, x_h, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = mfvgp_ecost(h, h->x_dev, want_grad, h->out_ws, h->grad_dev, h->status_ws, h->stream);
    if (rc) return rc;
    if (want_grad)
        CK(cudaMemcpyAsync(grad_h, h->grad_dev, nx * sizeof(double), cudaMemcpyDeviceToHost,
                           h->stream));
    CK(cudaMemcpyAsync(h->pin, h->out_ws, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->pin + 8, h->status_ws, sizeof(int32_t), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(out_h, h->pin, 4 * sizeof(double));
    memcpy(status_h, h->pin + 8, sizeof(int32_t));
    return MFVGP_OK;
}

extern "C" int mfvgp_esde_host(mfvgp_handle *h, const double *mean_pts_h,
                               const double *vars_pts_h, int want_grad, double *Esde_h,
                               double *dEsde_dm_h, double *dEsde_ds_h, int32_t *status_h) {
    ARG(h && mean_pts_h && vars_pts_h && Esde_h && status_h, "mfvgp_esde_host: NULL argument");
    ARG(!want_grad || (dEsde_dm_h && dEsde_ds_h), "mfvgp_esde_host: gradient buffers are NULL");
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const size_t nm = (size_t)h->D * h->Nm, ns = (size_t)h->D * h->Ns;
    CK(cudaMemcpyAsync(h->x_dev, mean_pts_h, nm * sizeof(double), cudaMemcpyHostToDevice,
                       h->stream));
    CK(cudaMemcpyAsync(h->x_dev + nm, vars_pts_h, ns * sizeof(double), cudaMemcpyHostToDevice,
                       h->stream));
    rc = mfvgp_esde(h, h->x_dev, h->Nm, h->x_dev + nm, h->Ns, want_grad, h->out_ws + 3,
                    h->dm_ws, h->ds_ws, h->status_ws, h->stream);
    if (rc) return rc;
    if (want_grad) {
        CK(cudaMemcpyAsync(dEsde_dm_h, h->dm_ws, (size_t)h->L * h->D * 4 * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(dEsde_ds_h, h->ds_ws, (size_t)h->L * h->D * 3 * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaMemcpyAsync(h->pin, h->out_ws + 3, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->pin + 8, h->status_ws, sizeof(int32_t), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *Esde_h = h->pin[0];
    memcpy(status_h, h->pin + 8, sizeof(int32_t));
    return MFVGP_OK;
}

// E_kl0 / E_obs as stand-alone calls on DEVICE buffers (free_energy.py:281-336, 567-649)
extern "C" int mfvgp_ekl0(mfvgp_handle *h, const double *m0_d, const double *s0_d, int want_grad,
                          double *E0_d, double *dm0_d, double *ds0_d, int32_t *status_d,
                          void *cuda_stream) {
    ARG(h && m0_d && s0_d && E0_d && status_d, "mfvgp_ekl0: NULL argument");
    ARG(!want_grad || (dm0_d && ds0_d), "mfvgp_ekl0: gradient buffers are NULL");
    if (!h->have_prior) { g_err = "mfvgp_ekl0: set_prior not called"; return MFVGP_ERR_STATE; }
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    ekl0_kernel<<<1, 256, 0, st>>>(m0_d, s0_d, h->mu0, h->tau0, h->D, E0_d,
                                   want_grad ? dm0_d : nullptr, want_grad ? ds0_d : nullptr,
                                   status_d);
    h->launches++;
    CK(cudaGetLastError());
    return MFVGP_OK;
}

extern "C" int mfvgp_eobs(mfvgp_handle *h, const double *mean_obs_d, const double *vars_obs_d,
                          int want_grad, double *Eobs_d, double *dm_d, double *ds_d,
                          int32_t *status_d, void *cuda_stream) {
    ARG(h && mean_obs_d && vars_obs_d && Eobs_d && status_d, "mfvgp_eobs: NULL argument");
    ARG(!want_grad || (dm_d && ds_d), "mfvgp_eobs: gradient buffers are NULL");
    if (!h->have_obs) { g_err = "mfvgp_eobs: set_obs not called"; return MFVGP_ERR_STATE; }
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    eobs_rows_kernel<<<h->d, 256, 0, st>>>(mean_obs_d, vars_obs_d, h->obs_values, h->obs_noise,
                                           h->d, h->M, h->apart, want_grad ? dm_d : nullptr,
                                           want_grad ? ds_d : nullptr);
    eobs_final_kernel<<<1, 256, 0, st>>>(h->apart, h->d, h->eobs_const, Eobs_d, status_d);
    h->launches += 2;
    CK(cudaGetLastError());
    return MFVGP_OK;
}

// E_kl0 / E_obs as stand-alone calls on host buffers (free_energy.py:281-336, 567-649)
extern "C" int mfvgp_ekl0_host(mfvgp_handle *h, const double *m0_h, const double *s0_h,
                               int want_grad, double *E0_h, double *dm0_h, double *ds0_h,
                               int32_t *status_h) {
    ARG(h && m0_h && s0_h && E0_h && status_h, "mfvgp_ekl0_host: NULL argument");
    ARG(!want_grad || (dm0_h && ds0_h), "mfvgp_ekl0_host: gradient buffers are NULL");
    if (!h->have_prior) { g_err = "mfvgp_ekl0_host: set_prior not called"; return MFVGP_ERR_STATE; }
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const int D = h->D;
    double *buf = h->x_dev, *gbuf = h->grad_dev;          // [m0 | s0], [dm0 | ds0]
    CK(cudaMemcpyAsync(buf, m0_h, D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(buf + D, s0_h, D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ekl0_kernel<<<1, 256, 0, h->stream>>>(buf, buf + D, h->mu0, h->tau0, D, h->out_ws,
                                          want_grad ? gbuf : nullptr,
                                          want_grad ? gbuf + D : nullptr, h->status_ws);
    h->launches++;
    CK(cudaGetLastError());
    if (want_grad) {
        CK(cudaMemcpyAsync(dm0_h, gbuf, D * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(ds0_h, gbuf + D, D * sizeof(double), cudaMemcpyDeviceToHost,
                           h->stream));
    }
    CK(cudaMemcpyAsync(h->pin, h->out_ws, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->pin + 8, h->status_ws, sizeof(int32_t), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *E0_h = h->pin[0];
    memcpy(status_h, h->pin + 8, sizeof(int32_t));
    return MFVGP_OK;
}

extern "C" int mfvgp_eobs_host(mfvgp_handle *h, const double *mean_obs_h,
                               const double *vars_obs_h, int want_grad, double *Eobs_h,
                               double *dm_h, double *ds_h, int32_t *status_h) {
    ARG(h && mean_obs_h && vars_obs_h && Eobs_h && status_h, "mfvgp_eobs_host: NULL argument");
    ARG(!want_grad || (dm_h && ds_h), "mfvgp_eobs_host: gradient buffers are NULL");
    if (!h->have_obs) { g_err = "mfvgp_eobs_host: set_obs not called"; return MFVGP_ERR_STATE; }
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const size_t dm_n = (size_t)h->d * h->M;               // <= D*Nm/3, fits the staging
    double *buf = h->x_dev, *gbuf = h->grad_dev;
    CK(cudaMemcpyAsync(buf, mean_obs_h, dm_n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(buf + dm_n, vars_obs_h, dm_n * sizeof(double), cudaMemcpyHostToDevice,
                       h->stream));
    eobs_rows_kernel<<<h->d, 256, 0, h->stream>>>(buf, buf + dm_n, h->obs_values, h->obs_noise,
                                                  h->d, h->M, h->apart,
                                                  want_grad ? gbuf : nullptr,
                                                  want_grad ? gbuf + dm_n : nullptr);
    eobs_final_kernel<<<1, 256, 0, h->stream>>>(h->apart, h->d, h->eobs_const, h->out_ws,
                                                h->status_ws);
    h->launches += 2;
    CK(cudaGetLastError());
    if (want_grad) {
        CK(cudaMemcpyAsync(dm_h, gbuf, dm_n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(ds_h, gbuf + dm_n, dm_n * sizeof(double), cudaMemcpyDeviceToHost,
                           h->stream));
    }
    CK(cudaMemcpyAsync(h->pin, h->out_ws, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->pin + 8, h->status_ws, sizeof(int32_t), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *Eobs_h = h->pin[0];
    memcpy(status_h, h->pin + 8, sizeof(int32_t));
    return MFVGP_OK;
}

extern "C" int mfvgp_refine_info(mfvgp_handle *h, int32_t *info_h) {
    ARG(h && info_h, "mfvgp_refine_info: NULL argument");
    CK(use_device(h->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(info_h, h->refine_info, 2 * h->L * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return MFVGP_OK;
}

extern "C" int mfvgp_timing(mfvgp_handle *h, int enable, double *total_ms, int64_t *count) {
    ARG(h != nullptr, "mfvgp_timing: NULL handle");
    CK(use_device(h->device));
    double tot = 0.0;
    int64_t cnt = 0;
    for (auto &pr : h->timing_events) {
        CK(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
        cnt++;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    h->timing_events.clear();
    if (total_ms) *total_ms = tot;
    if (count) *count = cnt;
    h->timing = enable != 0;
    return MFVGP_OK;
}

extern "C" int mfvgp_fp64_peak(int device, double *tflops_out) {
    ARG(tflops_out != nullptr, "mfvgp_fp64_peak: NULL argument");
    if (mfvgp_device_count() <= device) {
        g_err = "mfvgp_fp64_peak: no CUDA device";
        return MFVGP_ERR_CUDA;
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *buf = nullptr;
    CK(cudaMalloc((void **)&buf, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1.0e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1.0e-3) * 1.0e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops_out = best;
    return MFVGP_OK;
}

// ---- gradients with respect to theta and Sigma (eparam.cuh) ------------
extern "C" int mfvgp_eparam(mfvgp_handle *h, const double *x_d, double *dtheta_d,
                            double *dsigma_d, int32_t *status_d, void *cuda_stream) {
    ARG(h && x_d && dtheta_d && dsigma_d && status_d, "mfvgp_eparam: NULL argument");
    if (!h->have_sde || !h->have_obs) {
        g_err = "mfvgp_eparam: set_sde/set_obs not called";
        return MFVGP_ERR_STATE;
    }
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int nl = h->n_end - h->n_begin, D = h->D, nth = h->ntheta;
    const bool l96 = h->system == MFVGP_SYS_L96;
    // the dim tiling of the one-thread-per-cell kernels
    const int te = !l96 ? 0 : (D <= 26 ? 32 : ((int64_t)D * nl <= 148 * 256 ? 64 : 128));
    const int ntiles = l96 ? (D + te - 7) / (te - 6) : 1;
    const size_t n_sig = (size_t)nl * D, n_th = (size_t)nl * ntiles * kMaxTheta, n_out = nth + D;
    if (!h->eparam_ws &&
        dev_alloc(&h->eparam_ws, n_sig + n_th + n_out * (1 + (size_t)h->world)))
        return MFVGP_ERR_CUDA;
    EparamArgs a;
    a.x = x_d; a.obs_times = h->obs_times; a.sigma = h->sigma; a.theta = h->theta;
    a.part_sig = h->eparam_ws; a.part_th = h->eparam_ws + n_sig; a.status = status_d;
    a.D = D; a.Nm = h->Nm; a.Ns = h->Ns; a.n_begin = h->n_begin; a.n_end = h->n_end;
    a.ntiles = ntiles;
    double *out = h->eparam_ws + n_sig + n_th, *gathered = out + n_out;
    CK(cudaMemsetAsync(status_d, 0, sizeof(int32_t), st));
    switch (h->system) {
        case MFVGP_SYS_DW: eparam_small_kernel<0><<<nl, 64, 0, st>>>(a); break;
        case MFVGP_SYS_OU: eparam_small_kernel<1><<<nl, 64, 0, st>>>(a); break;
        case MFVGP_SYS_L63: eparam_small_kernel<2><<<nl, 64, 0, st>>>(a); break;
        default: {
            const unsigned grid = (unsigned)nl * ntiles;
            if (te == 32) eparam_l96_kernel<32><<<grid, 32, 0, st>>>(a);
            else if (te == 64) eparam_l96_kernel<64><<<grid, 64, 0, st>>>(a);
            else eparam_l96_kernel<128><<<grid, 128, 0, st>>>(a);
        }
    }
    CK(cudaGetLastError());
    eparam_reduce_kernel<<<(D + 255) / 256 + 1, 256, 0, st>>>(a.part_th, nl * ntiles, nth, a.part_sig,
                                                            nl, D, out, status_d);
    CK(cudaGetLastError());
    h->launches += 2;
    if (h->world > 1) {
        // every rank integrated its own intervals: all-gather, fold in rank order
        NCK(nccl().AllGather(out, gathered, n_out, kNcclFloat64, h->comm, st));
        eparam_combine_kernel<<<(int)((n_out + 255) / 256), 256, 0, st>>>(gathered, h->world,
                                                                         (int)n_out, out, status_d);
        CK(cudaGetLastError());
        h->launches++;
    }
    CK(cudaMemcpyAsync(dtheta_d, out, nth * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(dsigma_d, out + nth, D * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return MFVGP_OK;
}

extern "C" int mfvgp_eparam_host(mfvgp_handle *h, const double *x_h, double *dtheta_h,
                                 double *dsigma_h, int32_t *status_h) {
    ARG(h && x_h && dtheta_h && dsigma_h && status_h, "mfvgp_eparam_host: NULL argument");
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const size_t nx = (size_t)h->D * (h->Nm + h->Ns);
    CK(cudaMemcpyAsync(h->x_dev, x_h, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // results land behind the staged gradient buffer: [dtheta | dsigma]
    double *res = h->grad_dev;
    rc = mfvgp_eparam(h, h->x_dev, res, res + h->ntheta, h->status_ws, h->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dtheta_h, res, h->ntheta * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(dsigma_h, res + h->ntheta, h->D * sizeof(double), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaMemcpyAsync(h->pin + 8, h->status_ws, sizeof(int32_t), cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(status_h, h->pin + 8, sizeof(int32_t));
    return MFVGP_OK;
}

// ---- scaled conjugate gradient (scaled_cg.py:107-386) -----------------------
namespace {

// can this device co-schedule a 16-CTA cluster of the 512-thread SCG kernels?  (opt-in cluster
// size; asked once per device, the kernels are marked on the way)
static bool scg_cluster16_ok(int device) {
    static std::map<int, bool> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(device);
    if (it != cache.end()) return it->second;
    bool ok = !(getenv("MFVGP_SCG_CLUSTER16") && getenv("MFVGP_SCG_CLUSTER16")[0] == '0');
    const void *kerns[4] = {(const void *)scg_cl_direction_kernel<512>,
                            (const void *)scg_cl_theta_kernel<512>,
                            (const void *)scg_cl_stats_kernel<512>,
                            (const void *)scg_cl_stats_direction_kernel<512>};
    for (int i = 0; i < 4 && ok; ++i) {
        if (cudaFuncSetAttribute(kerns[i], cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
            cudaSuccess) { ok = false; break; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16); cfg.blockDim = dim3(512);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kerns[i], &cfg) != cudaSuccess ||
            nclusters < 1)
            ok = false;
    }
    if (!ok) cudaGetLastError();
    cache[device] = ok;
    return ok;
}

struct ScgRun {
    mfvgp_handle *h;
    cudaStream_t st;
    VecMap n;          // index map of this rank's part of x (named n: it replaces the length)
    int status_or = 0;
    bool fused = false;          // single GPU: folds in the producers, host-mapped record
    int cl_ctas = 0;             // > 0: the vector fits one thread-block cluster (scg.cuh)
    int cl_nt = 1024;            // ... of CTAs this wide (512 with 16-CTA clusters)
    // cluster path: the statistics kernel that closes a chain is held back until the host knows
    // whether the next chain's head is enqueued right behind it (then ONE launch does both)
    struct {
        bool on = false;
        const double *gnow = nullptr, *g = nullptr, *d = nullptr;
        FoldArgs f3;
    } pend;
    int flush_stats() {
        if (!pend.on) return MFVGP_OK;
        pend.on = false;
        CK(launch_cluster(cl_nt == 512 ? scg_cl_stats_kernel<512> : scg_cl_stats_kernel<1024>, cl_ctas, cl_nt, st, pend.gnow, pend.g, pend.d,
                          clvec(), pend.f3));
        h->launches++;
        return MFVGP_OK;
    }
    ClVec clvec(const double *x = nullptr, double *xt = nullptr) const {
        ClVec cv;
        cv.n = (int)n.n; cv.x = x; cv.xt = xt;
        return cv;
    }

    long long last_epoch = 0;    // epoch of the last publishing kernel enqueued
    // the host-mapped record has two slots (epoch & 1): a chain enqueued ahead of its
    // predecessor's outcome publishes while the host may still be reading the predecessor's
    double *rec_slot(long long epoch) const { return h->scg_hostrec + 32 * (epoch & 1); }
    FoldArgs fold_args(int mode = 0, double beta_in = 0.0) {
        FoldArgs fa;
        fa.mode = mode; fa.ctrl = h->scg_scal + 16; fa.beta_in = beta_in;
        fa.ticket = nullptr; fa.scal = h->scg_scal; fa.status = h->status_ws;
        fa.host_rec = h->scg_hostrec_dev;
        fa.host_flag = reinterpret_cast<volatile long long *>(h->scg_hostrec_dev + 24);
        fa.epoch = 0;
        fa.spec = 0; fa.decide = 0; fa.refined_mask = 0;
        fa.f_old_in = fa.x_tol = fa.f_tol = 0.0;
        fa.peer = nullptr; fa.peer_epoch = 0; fa.sync_status = h->refine_status;
        if (fused && h->world > 1) {              // every fold is one exchange between the ranks
            fa.peer = h->peer_layout;
            fa.peer_epoch = ++h->xchg_epoch;
        }
        if (fused) {
            fa.ticket = h->tickets + h->L;
            if (mode == 0 || mode == 3) {                             // only these publish
                fa.epoch = last_epoch = ++h->scg_epoch;
                fa.host_rec = h->scg_hostrec_dev + 32 * (fa.epoch & 1);
                fa.host_flag = reinterpret_cast<volatile long long *>(fa.host_rec + 24);
            }
        }
        return fa;
    }
    bool peer_rec = false;       // sharded over peer memory: folds publish to the host record too
    int fold() {
        if (fused) return MFVGP_OK;
        if (peer_rec) {
            // fold + send in one kernel, then the rank-ordered fold, which also writes the
            // host-mapped record: no memcpy, no stream synchronisation per read
            PeerArgs a;
            for (int r = 0; r < kPeerMaxWorld; ++r) a.box[r] = r < h->world ? h->peer_box[r] : nullptr;
            a.slot_doubles = h->peer_slot; a.epoch = ++h->xchg_epoch; a.ticket = h->peer_ticket;
            a.status = h->refine_status; a.rec = h->scg_scal; a.out = h->scg_scal;
            a.status_out = nullptr; a.grad = nullptr; a.n_rec = 4; a.n_sum = 3;
            a.D = h->D; a.Nm = h->Nm; a.Ns = h->Ns; a.n_begin = h->n_begin; a.n_end = h->n_end;
            a.rank = h->rank; a.world = h->world;
            a.fold_part = h->scg_part; a.fold_nblk = h->scg_nblk;
            last_epoch = ++h->scg_epoch;
            a.host_rec = h->scg_hostrec_dev + 32 * (last_epoch & 1);
            a.host_epoch = last_epoch;
            a.ecost_out = h->scg_scal + 4; a.ecost_status = h->status_ws;
            CK(launch_pdl(scg_peer_fold_publish_kernel, dim3(1), kRedBlock, st, a));
            CK(launch_pdl(peer_combine_kernel, dim3(1), 256, st, a));
            h->launches += 2;
            return MFVGP_OK;
        }
        scg_fold_kernel<<<1, kRedBlock, 0, st>>>(h->scg_part, h->scg_nblk, h->scg_scal);
        h->launches++;
        CK(cudaGetLastError());
        // sharded: the dot products of all ranks, folded in rank order (sum, sum, sum, max)
        if (h->world > 1) return combine_over_ranks(h, h->scg_scal, 4, 3, h->scg_scal, nullptr, st);
        return MFVGP_OK;
    }
    // waits until the kernel that publishes `epoch` has written its record
    double t_wait = 0.0, t_chain = 0.0;   // host seconds spent spinning / enqueueing (MFVGP_SCG_TRACE)
    long n_chain = 0, n_void = 0;
    int wait_record(long long epoch, const double **rec) {
        const auto t0 = std::chrono::steady_clock::now();
        int rc = wait_record_(epoch, rec);
        t_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    }
    int wait_record_(long long epoch, const double **rec) {
        const double *slot = rec_slot(epoch);
        volatile const long long *flag = reinterpret_cast<volatile const long long *>(slot + 24);
        long spins = 0;
        while (*flag != epoch) {
            if ((++spins & 0xfffff) == 0) {                 // a faulted kernel never publishes
                cudaError_t q = cudaStreamQuery(st);
                if (q != cudaSuccess && q != cudaErrorNotReady) {
                    g_err = std::string("mfvgp_scg: ") + cudaGetErrorString(q);
                    return MFVGP_ERR_CUDA;
                }
            }
        }
        __sync_synchronize();
        *rec = slot;
        return MFVGP_OK;
    }
    // reads scal[0..3] (fold output), scal[4..7] (E_cost output) and the status
    int read(double *s8) {
        if (fused || peer_rec) {
            // the last reduction kernel publishes the record and then the epoch
            const double *rec;
            int rc = wait_record(last_epoch, &rec);
            if (rc) return rc;
            memcpy(s8, rec, 8 * sizeof(double));
            status_or |= (int32_t)rec[8];
            return MFVGP_OK;
        }
        CK(cudaMemcpyAsync(h->pin, h->scg_scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h->pin + 8, h->status_ws, sizeof(int32_t), cudaMemcpyDeviceToHost,
                           st));
        CK(cudaStreamSynchronize(st));
        memcpy(s8, h->pin, 8 * sizeof(double));
        int32_t stt;
        memcpy(&stt, h->pin + 8, sizeof(int32_t));
        status_or |= stt;
        return MFVGP_OK;
    }
    // optimistic: without the refine_kernel launch (small problems); the caller checks
    // MFVGP_ST_REFINED in the status it reads back and redoes the evaluation if it is set
    int eval(const double *x, double *g, bool optimistic = false) {
        return ecost_impl(h, x, 1, h->scg_scal + 4, g, h->status_ws, st, optimistic);
    }
    int direction(double *d, const double *g, double gamma, int restart) {
        if (cl_ctas) {
            CK(launch_cluster(cl_nt == 512 ? scg_cl_direction_kernel<512> : scg_cl_direction_kernel<1024>, cl_ctas, cl_nt, st, d, g, gamma, restart,
                              clvec(), fold_args()));
            h->launches++;
            return MFVGP_OK;
        }
        CK(launch_pdl(scg_direction_kernel, dim3(h->scg_nblk), kRedBlock, st, d, g, gamma, restart,
                      n, h->scg_part, fold_args()));
        h->launches++;
        return fold();
    }
    int axpy(double *y, const double *x, const double *d, double a, const double *a_dev = nullptr,
             const double *void_flag = nullptr) {
        CK(launch_pdl(scg_axpy_kernel, dim3(h->scg_nblk), kRedBlock, st, y, x, d, a, a_dev,
                      void_flag, n));
        h->launches++;
        return MFVGP_OK;
    }
    // One successful iteration enqueued in one go (scaled_cg.py:192-247 + :361-369): direction
    // update, sigma step, theta, alpha step, trial statistics -- sigma and alpha stay on the
    // device (FoldArgs modes 1-3); the host reads ONE record at the end instead of three.
    // The stats kernel also takes the accept / continue decision (FoldArgs::decide), which lets
    // the host enqueue the chain of iteration j+1 BEFORE it has read the record of iteration j
    // (spec = true: gamma, beta and f_old come from the device, a void flag stops the chain
    // if iteration j ends any other way than "accepted, carry on").
    int chain(double *d, const double *g, double gamma, int restart, double beta, double *x,
              double *xt, double *gp, double *gnow, bool optimistic, bool spec, double f_old,
              double x_tol, double f_tol) {
        const auto t0 = std::chrono::steady_clock::now();
        int rc = chain_(d, g, gamma, restart, beta, x, xt, gp, gnow, optimistic, spec, f_old, x_tol,
                        f_tol);
        t_chain += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        n_chain++;
        return rc;
    }
    int chain_(double *d, const double *g, double gamma, int restart, double beta, double *x,
               double *xt, double *gp, double *gnow, bool optimistic, bool spec, double f_old,
               double x_tol, double f_tol) {
        const double *ctrl = h->scg_scal + 16;
        const double *vflag = spec ? ctrl + 10 : nullptr;
        if (cl_ctas) {
            // direction + sigma step, [evaluation], theta + alpha step, [evaluation]; the
            // statistics wait (pend) for the head of the next chain to join them, or flush_stats()
            FoldArgs f1 = fold_args(1), f2 = fold_args(2, beta);
            f1.spec = f2.spec = spec;
            if (spec && pend.on) {
                pend.on = false;
                CK(launch_cluster(cl_nt == 512 ? scg_cl_stats_direction_kernel<512> : scg_cl_stats_direction_kernel<1024>, cl_ctas, cl_nt, st, pend.gnow,
                                  pend.g, d, restart, clvec(x, xt), pend.f3, f1));
            } else {
                int rc = flush_stats();
                if (rc) return rc;
                CK(launch_cluster(cl_nt == 512 ? scg_cl_direction_kernel<512> : scg_cl_direction_kernel<1024>, cl_ctas, cl_nt, st, d, g, gamma,
                                  restart, clvec(x, xt), f1));
            }
            h->launches++;
            int rc = eval(xt, gp, optimistic);
            if (rc) return rc;
            CK(launch_cluster(cl_nt == 512 ? scg_cl_theta_kernel<512> : scg_cl_theta_kernel<1024>, cl_ctas, cl_nt, st, (const double *)d,
                              (const double *)gp, g, clvec(x, xt), f2));
            h->launches++;
            rc = eval(xt, gnow, optimistic);
            if (rc) return rc;
            pend.f3 = fold_args(3);
            pend.f3.spec = spec; pend.f3.decide = 1;
            pend.f3.refined_mask = optimistic ? MFVGP_ST_REFINED : 0;
            pend.f3.f_old_in = f_old; pend.f3.x_tol = x_tol; pend.f3.f_tol = f_tol;
            pend.gnow = gnow; pend.g = g; pend.d = d;
            pend.on = true;
            return MFVGP_OK;
        }
        FoldArgs f1 = fold_args(1);
        f1.spec = spec;
        CK(launch_pdl(scg_direction_kernel, dim3(h->scg_nblk), kRedBlock, st, d, g, gamma, restart,
                      n, h->scg_part, f1));
        h->launches++;
        int rc = axpy(xt, x, d, 0.0, ctrl + 0, vflag);
        if (rc) return rc;
        rc = eval(xt, gp, optimistic);
        if (rc) return rc;
        FoldArgs f2 = fold_args(2, beta);
        f2.spec = spec;
        CK(launch_pdl(scg_theta_kernel, dim3(h->scg_nblk), kRedBlock, st, (const double *)d,
                      (const double *)gp, g, n, h->scg_part, f2));
        h->launches++;
        rc = axpy(xt, x, d, 0.0, ctrl + 1, vflag);
        if (rc) return rc;
        rc = eval(xt, gnow, optimistic);
        if (rc) return rc;
        FoldArgs f3 = fold_args(3);
        f3.spec = spec; f3.decide = 1;
        f3.refined_mask = optimistic ? MFVGP_ST_REFINED : 0;
        f3.f_old_in = f_old; f3.x_tol = x_tol; f3.f_tol = f_tol;
        CK(launch_pdl(scg_stats_kernel, dim3(h->scg_nblk), kRedBlock, st, (const double *)gnow, g,
                      (const double *)d, n, h->scg_part, f3));
        h->launches++;
        return MFVGP_OK;
    }
    int theta(const double *d, const double *gp, const double *g) {
        if (cl_ctas) {
            CK(launch_cluster(cl_nt == 512 ? scg_cl_theta_kernel<512> : scg_cl_theta_kernel<1024>, cl_ctas, cl_nt, st, d, gp, g, clvec(),
                              fold_args()));
            h->launches++;
            return MFVGP_OK;
        }
        CK(launch_pdl(scg_theta_kernel, dim3(h->scg_nblk), kRedBlock, st, d, gp, g, n,
                      h->scg_part, fold_args()));
        h->launches++;
        return fold();
    }
    int stats(const double *g_now, const double *g_new, const double *d) {
        if (cl_ctas) {
            CK(launch_cluster(cl_nt == 512 ? scg_cl_stats_kernel<512> : scg_cl_stats_kernel<1024>, cl_ctas, cl_nt, st, g_now, g_new, d, clvec(),
                              fold_args()));
            h->launches++;
            return MFVGP_OK;
        }
        CK(launch_pdl(scg_stats_kernel, dim3(h->scg_nblk), kRedBlock, st, g_now, g_new, d, n,
                      h->scg_part, fold_args()));
        h->launches++;
        return fold();
    }
};

// copies the entries of a vector in the layout of x that this rank works on: the whole vector
// on one GPU, the rank's column blocks (one strided block each of the mean and the log-variance
// part) on a sharded handle -- the copies then shrink with the shard like everything else
static cudaError_t copy_active(double *dst, const double *src, const VecMap &vm, int64_t n,
                               cudaStream_t st) {
    if (vm.full) return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
    cudaError_t e = cudaMemcpy2DAsync(dst + vm.cm0, (size_t)vm.Nm * sizeof(double), src + vm.cm0,
                                      (size_t)vm.Nm * sizeof(double), (size_t)vm.wm * sizeof(double),
                                      (size_t)vm.D, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    return cudaMemcpy2DAsync(dst + vm.num_mp + vm.cs0, (size_t)vm.Ns * sizeof(double),
                             src + vm.num_mp + vm.cs0, (size_t)vm.Ns * sizeof(double),
                             (size_t)vm.ws * sizeof(double), (size_t)vm.D, cudaMemcpyDeviceToDevice, st);
}

#define RC(call)                \
    do {                        \
        int rc_ = (call);       \
        if (rc_) return rc_;    \
    } while (0)

}  // namespace

extern "C" int mfvgp_scg(mfvgp_handle *h, double *x_d, int max_it, double x_tol, double f_tol,
                         double *fx_hist_h, double *dfx_hist_h, double *result_h,
                         void *cuda_stream) {
    ARG(h && x_d && fx_hist_h && dfx_hist_h && result_h, "mfvgp_scg: NULL argument");
    ARG(max_it >= 1, "mfvgp_scg: max_it must be positive");
    CK(use_device(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int64_t n = (int64_t)h->D * (h->Nm + h->Ns);
    VecMap vm;
    vm.num_mp = (int64_t)h->D * h->Nm;
    vm.D = h->D; vm.Nm = h->Nm; vm.Ns = h->Ns;
    vm.full = h->world == 1 ? 1 : 0;
    {
        const bool last = h->n_end == h->L;
        vm.cm0 = 3 * h->n_begin; vm.cs0 = 2 * h->n_begin;
        vm.own_wm = (last ? h->Nm : 3 * h->n_end) - vm.cm0;
        vm.own_ws = (last ? h->Ns : 2 * h->n_end) - vm.cs0;
        vm.wm = vm.own_wm + (last ? 0 : 1);          // + ghost column shared with the next rank
        vm.ws = vm.own_ws + (last ? 0 : 1);
        vm.n = vm.full ? n : (int64_t)h->D * (vm.wm + vm.ws);
    }
    const int64_t n_total = n;       // SCG restarts after dim_x successes (scaled_cg.py:361)
    if (!h->scg_vec[0]) {
        for (int i = 0; i < 7; ++i)
            if (dev_alloc(&h->scg_vec[i], n)) return MFVGP_ERR_CUDA;
        int64_t nb = (n + kRedBlock * 4 - 1) / (kRedBlock * 4);
        h->scg_nblk = (int)(nb < 1 ? 1 : (nb > 1184 ? 1184 : nb));
        if (dev_alloc(&h->scg_part, (size_t)h->scg_nblk * kRedSlots) ||
            dev_alloc(&h->scg_scal, 32))
            return MFVGP_ERR_CUDA;
    }
    if (!h->scg_hostrec) {
        CK(cudaHostAlloc((void **)&h->scg_hostrec, 64 * sizeof(double), cudaHostAllocMapped));
        memset(h->scg_hostrec, 0, 64 * sizeof(double));
        CK(cudaHostGetDevicePointer((void **)&h->scg_hostrec_dev, h->scg_hostrec, 0));
    }
    ScgRun R{h, st, vm};
    {
        const char *e = getenv("MFVGP_SCG_FUSED");
        // sharded over peer memory: the same fused folds, each followed by an exchange in the
        // last block (MFVGP_SCG_FUSED=2 keeps round 2's first version: fold + send and
        // rank-ordered fold as two kernels per reduction, host computes the step lengths)
        const bool off = e && e[0] == '0', two_kernel = e && e[0] == '2';
        R.fused = !off && (h->world == 1 || (h->peer_ok && !two_kernel));
        R.peer_rec = h->world > 1 && h->peer_ok && two_kernel;
        const char *c = getenv("MFVGP_SCG_CLUSTER");
        if (R.fused && h->world == 1 && n <= kClMaxElems && !(c && c[0] == '0')) {
            const bool wide = scg_cluster16_ok(h->device);
            R.cl_nt = wide ? 512 : 1024;
            const int max_ctas = wide ? 16 : 8;
            int64_t want = (n + 4 * R.cl_nt - 1) / (4 * R.cl_nt);
            R.cl_ctas = (int)(want < 1 ? 1 : want > max_ctas ? max_ctas : want);
        }
    }
    double *x = h->scg_vec[0], *xt = h->scg_vec[1], *d = h->scg_vec[2];
    double *gnew = h->scg_vec[3], *gold = h->scg_vec[4], *gnow = h->scg_vec[5];
    double *gp = h->scg_vec[6];
    double s8[8];
    for (int i = 0; i < max_it; ++i) fx_hist_h[i] = dfx_hist_h[i] = 0.0;   // :123-124
    CK(copy_active(x, x_d, vm, n, st));
    if (!vm.full)   // the two iterates swap roles: both start from the caller's columns
        CK(copy_active(xt, x_d, vm, n, st));

    int nit = max_it, func_eval = 0;
    double fx = 0.0;
    // scaled_cg.py:145-175
    RC(R.eval(x, gnew));
    RC(R.stats(gnew, gnew, gnew));
    RC(R.read(s8));
    double f_now = s8[4], sa_new = s8[0];
    func_eval++;
    bool failed = (R.status_or & 7) != 0;
    CK(copy_active(gold, gnew, vm, n, st));
    double sa_old = sa_new, f_old = f_now;
    RC(R.direction(d, gnew, 0.0, 1));                            // d = -grad_new
    RC(R.read(s8));
    double mu = s8[0], kappa = s8[1], theta = 0.0, beta = 1.0;
    const double beta_min = 1.0e-15, beta_max = 1.0e100, eps = 2.220446049250313e-16;
    bool success = true, done = failed;
    int64_t count_success = 0;
    if (failed) { nit = 0; fx = f_now; }

    // optimistic evaluations exist on the small-problem path of one GPU only (ecost_impl); a
    // chain must not be voided for MFVGP_ST_REFINED when its evaluations did run the adaptive path
    const bool can_opt = h->coop && h->world == 1;
    bool chained = false;          // a speculative chain for this iteration is in flight
    bool chain_optimistic = false; // ... whose evaluations ran without the refine_kernel launch
    long long chain_epoch = 0;     // ... and publishes its record under this epoch
    // pipelining: while the chain of iteration j runs, the chain of iteration j+1 is enqueued
    // behind it on the assumption that j is accepted and carries on (the device takes the
    // decision and voids the chain otherwise); only after a run of clean accepted iterations
    bool ahead = false, ahead_optimistic = false, ahead_valid = false;
    long long ahead_epoch = 0;
    int clean_run = 0;
    // (sharded: no chain ahead of its predecessor's outcome -- a void chain would skip its
    // exchanges, and the mailboxes' two slots rely on every exchange epoch being used)
    bool pipeline = R.fused && h->world == 1;
    int clean_needed = 2;
    {
        const char *e = getenv("MFVGP_SCG_PIPELINE");       // 0: off, 2: after every chain (tests)
        if (e && e[0] == '0') pipeline = false;
        if (e && e[0] == '2') clean_needed = 0;
    }
    double alpha = 0.0;
    for (int j = 0; j < max_it && !done; ++j) {                   // :186
        bool have_stats = false;
        if (chained) {
            chained = false;
            ahead = false;
            if (pipeline && clean_run >= clean_needed && j + 1 < max_it) {
                // roles after an accepted iteration j: x <-> xt, (gold, gnew, gnow) rotate
                const int restart1 = count_success + 1 == n_total;
                ahead_optimistic = can_opt && h->opt_cooldown == 0;
                RC(R.chain(d, gnow, 0.0, restart1, 0.0, xt, x, gp, gold, ahead_optimistic, true,
                           0.0, x_tol, f_tol));
                ahead_epoch = R.last_epoch;
                ahead = true;
            } else {
                RC(R.flush_stats());
            }
            const double *rec;
            RC(R.wait_record(chain_epoch, &rec));
            double r16[16];
            memcpy(r16, rec, 16 * sizeof(double));
            ahead_valid = ahead && rec[17] != 0.0;
            const int flags = (int)r16[15];
            mu = r16[9]; kappa = r16[10];
            const bool flagged = (((int32_t)r16[11] | (int32_t)r16[8]) & MFVGP_ST_REFINED) != 0;
            if (flagged) h->opt_cooldown = 16;
            else if (h->opt_cooldown > 0) h->opt_cooldown--;
            if ((flags & 1) || ((flags & 2) == 0 && flagged && chain_optimistic)) {
                // mu >= 0 (:194-197), or the guard flagged an interval in one of the chain's
                // optimistic evaluations (no refine_kernel): rare; the chain's results are void,
                // redo the iteration on the step-by-step path below (which restarts the
                // direction first if mu >= 0 and evaluates with the adaptive path)
            } else if (flags & 2) {                               // :203
                fx = f_now; nit = j + 1; done = true; break;
            } else {
                func_eval++;
                R.status_or |= (int32_t)r16[11];
                if (R.status_or & 7) { fx = f_now; nit = j + 1; done = true; break; }
                theta = r16[12]; alpha = r16[13]; beta = r16[14];
                func_eval++;
                R.status_or |= (int32_t)r16[8];
                if (R.status_or & 7) { fx = f_now; nit = j + 1; done = true; break; }
                memcpy(s8, r16, 8 * sizeof(double));
                have_stats = true;
            }
        }
        if (!have_stats) {
        if (success) {
            if (mu >= 0.0) {                                      // :194-197
                RC(R.direction(d, gnew, 0.0, 1));
                RC(R.read(s8));
                mu = s8[0]; kappa = s8[1];
            }
            if (kappa < eps) { fx = f_now; nit = j + 1; done = true; break; }   // :203
            const double sigma = 1.0e-3 / sqrt(kappa);
            RC(R.axpy(xt, x, d, sigma));
            RC(R.eval(xt, gp));                                   // :222
            RC(R.theta(d, gp, gnew));
            RC(R.read(s8));
            func_eval++;
            if (R.status_or & 7) { fx = f_now; nit = j + 1; done = true; break; }
            theta = s8[0] / sigma;                                // :228
        }
        double delta = theta + beta * kappa;                      // :232-236
        if (delta <= 0.0) {
            delta = beta * kappa;
            beta = beta - theta / kappa;
        }
        alpha = -(mu / delta);
        RC(R.axpy(xt, x, d, alpha));
        RC(R.eval(xt, gnow));                                     // :247
        RC(R.stats(gnow, gnew, d));
        RC(R.read(s8));
        func_eval++;
        if (R.status_or & 7) { fx = f_now; nit = j + 1; done = true; break; }
        }
        const double f_new = s8[4], sa_now = s8[0], gg = s8[1], prnum = s8[2], maxd = s8[3];
        const double Delta = 2.0 * (f_new - f_old) / (alpha * mu);   // :253
        double total_grad;
        clean_run = (have_stats && Delta >= 0.0) ? clean_run + 1 : 0;
        if (Delta >= 0.0) {
            success = true;
            count_success++;
            f_now = f_new;
            double *t = x; x = xt; xt = t;                        // :266
            total_grad = sa_now;
        } else {
            success = false;
            f_now = f_old;
            total_grad = sa_old;                                  // :277 g_now <- grad_old
        }
        fx_hist_h[j] = f_now;                                     // :284-285
        dfx_hist_h[j] = total_grad;
        if (success) {
            if (fabs(alpha) * maxd <= x_tol && fabs(f_new - f_old) <= f_tol) {   // :306
                fx = f_new; nit = j + 1; done = true; break;
            }
            f_old = f_new;
            // :324-327 grad_old <- grad_new; (f_now, grad_new) = func(x).  x is the
            // point just evaluated and E_cost is deterministic, so the re-evaluation
            // returns (f_new, g_now) bit for bit: rotate the buffers instead.
            double *t = gold; gold = gnew; gnew = gnow; gnow = t;
            sa_old = sa_new; sa_new = sa_now;
            func_eval++;
            f_now = f_new;
            if (fabs(gg) <= 1.0e-8) { fx = f_now; nit = j + 1; done = true; break; }   // :333
        }
        if (Delta < 0.25) beta = fmin(4.0 * beta, beta_max);      // :350-356
        if (Delta > 0.75) beta = fmax(0.5 * beta, beta_min);
        // :361-369 -- the new direction; on one GPU it opens the speculative chain of the
        // next iteration (direction, sigma step, theta, alpha step, statistics: one host read)
        const bool spec = R.fused && success && j + 1 < max_it;
        if (ahead) {
            // the device took the decision above from the same numbers: it must agree
            ahead = false;
            if (ahead_valid != (spec && have_stats)) {
                g_err = "mfvgp_scg: host and device disagree on the outcome of an iteration";
                return MFVGP_ERR_STATE;
            }
            if (ahead_valid) {                                    // iteration j+1 is already running
                if (count_success == n_total) count_success = 0;
                chained = true;
                chain_optimistic = ahead_optimistic;
                chain_epoch = ahead_epoch;
                continue;
            }
            R.pend.on = false;      // the void chain's statistics would only publish "void"
        }
        if (count_success == n_total) {
            count_success = 0;
            if (spec) {
                chain_optimistic = can_opt && h->opt_cooldown == 0;
                RC(R.chain(d, gnew, 0.0, 1, beta, x, xt, gp, gnow, chain_optimistic, false, f_old,
                           x_tol, f_tol));
                chain_epoch = R.last_epoch;
                chained = true;
            } else {
                RC(R.direction(d, gnew, 0.0, 1));
                RC(R.read(s8));
                mu = s8[0]; kappa = s8[1];
            }
        } else if (success) {
            const double gamma = fmax(prnum / mu, 0.0);
            if (spec) {
                chain_optimistic = can_opt && h->opt_cooldown == 0;
                RC(R.chain(d, gnew, gamma, 0, beta, x, xt, gp, gnow, chain_optimistic, false,
                           f_old, x_tol, f_tol));
                chain_epoch = R.last_epoch;
                chained = true;
            } else {
                RC(R.direction(d, gnew, gamma, 0));
                RC(R.read(s8));
                mu = s8[0]; kappa = s8[1];
            }
        }
    }
    if (!done) fx = f_old;                                        // :379
    if (getenv("MFVGP_SCG_TRACE"))
        fprintf(stderr, "mfvgp_scg: nit=%d chains=%ld (%.1f us host enqueue each), record waits "
                        "%.1f us per iteration\n", nit, R.n_chain,
                R.n_chain ? 1e6 * R.t_chain / R.n_chain : 0.0, nit ? 1e6 * R.t_wait / nit : 0.0);
    CK(copy_active(x_d, x, vm, n, st));      // sharded: the rank's columns of the caller's x
    CK(cudaStreamSynchronize(st));
    result_h[0] = fx;
    result_h[1] = nit;
    result_h[2] = func_eval;
    result_h[3] = R.status_or;
    return MFVGP_OK;
}

extern "C" int mfvgp_scg_host(mfvgp_handle *h, double *x_h, int max_it, double x_tol,
                              double f_tol, double *fx_hist_h, double *dfx_hist_h,
                              double *result_h) {
    ARG(h && x_h, "mfvgp_scg_host: NULL argument");
    CK(use_device(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const size_t nx = (size_t)h->D * (h->Nm + h->Ns);
    CK(cudaMemcpyAsync(h->x_dev, x_h, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = mfvgp_scg(h, h->x_dev, max_it, x_tol, f_tol, fx_hist_h, dfx_hist_h, result_h,
                   h->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(x_h, h->x_dev, nx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MFVGP_OK;
}

// ---- posterior reconstruction -------------------------------------
extern "C" int mfvgp_reconstruct(int D, int M, int num, const double *Tx_d,
                                 const double *mean_pts_d, int64_t ldm, const double *vars_pts_d,
                                 int64_t lds, int log_vars, double *tt_d, double *mt_d,
                                 double *st_d, void *cuda_stream) {
    ARG(Tx_d && mean_pts_d && vars_pts_d && tt_d && mt_d && st_d, "mfvgp_reconstruct: NULL argument");
    ARG(D >= 1 && D <= 65535 * kReconRows && M >= 1 && num >= 1, "mfvgp_reconstruct: need 1 <= D <= 524280, M >= 1, num >= 1");
    ARG(ldm >= 3 * (int64_t)M + 4 && lds >= 2 * (int64_t)M + 3, "mfvgp_reconstruct: row pitch too small");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    double *vexp = nullptr;                    // variances of a log-variance input
    const int Ns = 2 * M + 3;
    if (log_vars) {
        // stream-ordered scratch; keep freed blocks in the pool instead of returning them to
        // the driver at every synchronisation (default threshold 0 makes each call re-map)
        static thread_local int pool_ready_dev = -1;
        int dev = 0;
        CK(cudaGetDevice(&dev));
        if (pool_ready_dev != dev) {
            cudaMemPool_t pool;
            CK(cudaDeviceGetDefaultMemPool(&pool, dev));
            uint64_t keep = ~0ull;
            CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            pool_ready_dev = dev;
        }
        CK(cudaMallocAsync((void **)&vexp, (size_t)D * Ns * sizeof(double), st));
        const int64_t ne = (int64_t)D * Ns;
        exp_rows_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(vars_pts_d, lds, Ns, D, vexp);
        CK(cudaGetLastError());
    }
    ReconArgs a;
    a.Tx = Tx_d; a.mean = mean_pts_d; a.ldm = ldm;
    a.vars = log_vars ? vexp : vars_pts_d; a.lds = log_vars ? Ns : lds;
    a.tt = tt_d; a.mt = mt_d; a.st = st_d;
    a.P = ((int64_t)M + 1) * num + 1;
    a.D = D; a.M = M; a.num = num;
    ARG(a.P < 4294967296LL, "mfvgp_reconstruct: too many samples (P must fit 32 bits)");
    const int64_t nbx = (a.P + 255) / 256;
    // 64 rows per CTA once that still leaves several CTAs per SM, fewer for small problems
    int groups = kReconGroups;
    while (groups > 1 && nbx * ((D + kReconRows * groups - 1) / (kReconRows * groups)) < 8 * 148) groups /= 2;
    const int rows_per_cta = kReconRows * groups;
    reconstruct_kernel<<<dim3((unsigned)nbx, (unsigned)((D + rows_per_cta - 1) / rows_per_cta)), 256, 0, st>>>(a, groups);
    CK(cudaGetLastError());
    if (vexp) CK(cudaFreeAsync(vexp, st));
    return MFVGP_OK;
}

extern "C" int mfvgp_reconstruct_host(int device, int D, int M, int num, const double *Tx_h,
                                      const double *mean_pts_h, const double *vars_pts_h,
                                      int log_vars, double *tt_h, double *mt_h, double *st_h) {
    ARG(Tx_h && mean_pts_h && vars_pts_h && tt_h && mt_h && st_h, "mfvgp_reconstruct_host: NULL argument");
    ARG(D >= 1 && M >= 1 && num >= 1, "mfvgp_reconstruct_host: bad size");
    if (mfvgp_device_count() <= device) {
        g_err = "mfvgp_reconstruct_host: no CUDA device (this library has no CPU fallback)";
        return MFVGP_ERR_CUDA;
    }
    CK(use_device(device));
    const int64_t Nm = 3 * (int64_t)M + 4, Ns = 2 * (int64_t)M + 3, P = ((int64_t)M + 1) * num + 1;
    double *buf = nullptr;
    auto pad4 = [](size_t n) { return (n + 3) / 4 * 4; };      // keep every block 32-byte aligned
    const size_t n_in = pad4((size_t)(M + 2)) + pad4((size_t)D * Nm) + pad4((size_t)D * Ns);
    const size_t n_out = pad4((size_t)P) + 2 * pad4((size_t)D * P);
    CK(cudaMalloc((void **)&buf, (n_in + n_out) * sizeof(double)));
    double *Tx = buf, *mp = Tx + pad4((size_t)(M + 2)), *sp = mp + pad4((size_t)D * Nm),
           *tt = sp + pad4((size_t)D * Ns), *mt = tt + pad4((size_t)P), *st = mt + pad4((size_t)D * P);
    int rc = MFVGP_OK;
    cudaError_t e = cudaMemcpy(Tx, Tx_h, (M + 2) * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(mp, mean_pts_h, (size_t)D * Nm * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(sp, vars_pts_h, (size_t)D * Ns * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = mfvgp_reconstruct(D, M, num, Tx, mp, Nm, sp, Ns, log_vars, tt, mt, st, nullptr);
        if (rc == MFVGP_OK) {
            e = cudaMemcpy(tt_h, tt, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(mt_h, mt, (size_t)D * P * sizeof(double), cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(st_h, st, (size_t)D * P * sizeof(double), cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(buf);
    if (e != cudaSuccess) { g_err = std::string("mfvgp_reconstruct_host: ") + cudaGetErrorString(e); return MFVGP_ERR_CUDA; }
    return rc;
}

// ---- initial state from a device-resident initial path -----------------
extern "C" int mfvgp_initial_state(int D, int M, int64_t T, const double *path_d,
                                   const int64_t *mp_idk_d, const double *u_d, double var_base,
                                   double var_spread, double *x_d, void *cuda_stream) {
    ARG(path_d && mp_idk_d && u_d && x_d, "mfvgp_initial_state: NULL argument");
    ARG(D >= 1 && M >= 1 && T >= 1, "mfvgp_initial_state: bad size");
    InitStateArgs a;
    a.path = path_d; a.mp_idk = mp_idk_d; a.u = u_d; a.x = x_d;
    a.var_base = var_base; a.var_spread = var_spread; a.T = T;
    a.D = D; a.Nm = 3 * M + 4; a.Ns = 2 * M + 3;
    const int64_t n = (int64_t)D * (a.Nm + a.Ns);
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    initial_state_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(a);
    CK(cudaGetLastError());
    return MFVGP_OK;
}

// ---- Lorenz 96 sample path ------------------------------------------
extern "C" int mfvgp_l96_trajectory(int D, double theta, double dt, int dim_t, int nburn,
                                    double delta_t, const double *ek_td_d, double *x_out_d,
                                    int32_t *status_h, void *cuda_stream) {
    ARG(ek_td_d && x_out_d && status_h, "mfvgp_l96_trajectory: NULL argument");
    ARG(D >= 4, "Lorenz96: Insufficient state vector dimensions");          // lorenz_96.py:60-63
    ARG(dim_t >= 1 && nburn >= 0, "mfvgp_l96_trajectory: bad step counts");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    TrajArgs a;
    a.ek = ek_td_d; a.x_out = x_out_d; a.scratch = nullptr;
    a.theta = theta; a.dt = dt; a.delta_t = delta_t;
    a.D = D; a.dim_t = dim_t; a.nburn = nburn;
    const size_t smem = 2 * (size_t)D * sizeof(double);
    a.use_smem = smem <= 200 * 1024;
    int threads = (D + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    int32_t *status_d = nullptr;
    CK(cudaMallocAsync((void **)&status_d, sizeof(int32_t), st));
    if (a.use_smem) {
        CK(cudaFuncSetAttribute(l96_trajectory_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
    } else {
        CK(cudaMallocAsync((void **)&a.scratch, smem, st));
    }
    a.status = status_d;
    l96_trajectory_kernel<<<1, threads, a.use_smem ? smem : 0, st>>>(a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(status_h, status_d, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFreeAsync(status_d, st));
    if (a.scratch) CK(cudaFreeAsync(a.scratch, st));
    return MFVGP_OK;
}

// ---- DW / OU / L63 sample paths, npaths independent ones per call ----------------------------
template <int SYS>
static void launch_small_trajectory(const SmallTrajArgs &a, cudaStream_t st) {
    small_trajectory_kernel<SYS><<<(a.npaths + 127) / 128, 128, 0, st>>>(a);
}

extern "C" int mfvgp_small_trajectory(int sys, const double *theta_h, int ntheta, double dt,
                                      int dim_t, int nburn, double delta_t, int npaths,
                                      const double *x0_d, const double *ek_d, double *x_out_d,
                                      int32_t *status_h, void *cuda_stream) {
    ARG(theta_h && x0_d && ek_d && x_out_d && status_h, "mfvgp_small_trajectory: NULL argument");
    ARG(sys == MFVGP_SYS_DW || sys == MFVGP_SYS_OU || sys == MFVGP_SYS_L63,
        "mfvgp_small_trajectory: system must be DW, OU or L63");
    const int want = sys == MFVGP_SYS_DW ? 1 : sys == MFVGP_SYS_OU ? 2 : 3;
    ARG(ntheta == want, "mfvgp_small_trajectory: wrong number of drift parameters");
    ARG(dim_t >= 1 && nburn >= 0 && npaths >= 1, "mfvgp_small_trajectory: bad counts");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    SmallTrajArgs a;
    a.x0 = x0_d; a.ek = ek_d; a.x_out = x_out_d;
    for (int k = 0; k < 3; ++k) a.theta[k] = k < ntheta ? theta_h[k] : 0.0;
    a.dt = dt; a.delta_t = delta_t; a.npaths = npaths; a.dim_t = dim_t; a.nburn = nburn;
    int32_t *status_d = nullptr;
    CK(cudaMallocAsync((void **)&status_d, (size_t)npaths * sizeof(int32_t), st));
    a.status = status_d;
    if (sys == MFVGP_SYS_DW) launch_small_trajectory<0>(a, st);
    else if (sys == MFVGP_SYS_OU) launch_small_trajectory<1>(a, st);
    else launch_small_trajectory<2>(a, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(status_h, status_d, (size_t)npaths * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFreeAsync(status_d, st));
    return MFVGP_OK;
}
