/*
 * mfvgp.h -- C ABI of the B200-native MeanFieldVarGP free-energy path.
 *
 * The reference (vrettasm/MeanFieldVarGP) is pure Python and has no FFI; its
 * boundary for this path is the class API of src/var_bayes/free_energy.py and
 * src/numerical/scaled_cg.py.  Each entry point below names the reference
 * interface it replaces (file:line under the reference repository).  All
 * arithmetic is FP64.  Matrices are row-major exactly as the reference lays
 * them out: mean points [D][3M+4], variance points [D][2M+3] (M = number of
 * observation times), x = [means | log-variances] (free_energy.py:678-684).
 *
 * Pointer kinds: names ending in _h are HOST pointers, names ending in _d are
 * DEVICE pointers on the handle's device.  The library owns only its
 * workspace handle; callers own every buffer they pass.  A handle is
 * re-entrant across handles, not thread-safe within one.  Every function
 * returns 0 on success, a negative MFVGP_ERR_* code otherwise;
 * mfvgp_last_error() gives the message.  There is no CPU fallback: without a
 * CUDA device every compute entry fails with MFVGP_ERR_CUDA.
 */
#ifndef MFVGP_H_
#define MFVGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFVGP_VERSION 100

/* drift plug-ins: src/dynamical_systems/{double_well,ornstein_uhlenbeck,
 * lorenz_63,lorenz_96}.py and their .sym expression files                  */
enum { MFVGP_SYS_DW = 0, MFVGP_SYS_OU = 1, MFVGP_SYS_L63 = 2, MFVGP_SYS_L96 = 3 };

enum {
    MFVGP_OK = 0,
    MFVGP_ERR_ARG = -1,      /* bad argument (dimension, NULL, system id)   */
    MFVGP_ERR_CUDA = -2,     /* CUDA runtime error / no device              */
    MFVGP_ERR_STATE = -3     /* problem data not set yet                    */
};

/* status bits written by the device (free_energy.py:313-316, 547-550,
 * 615-618: the reference raises RuntimeError on the first three)           */
enum {
    MFVGP_ST_ESDE_NONFINITE = 1,
    MFVGP_ST_E0_NONFINITE = 2,
    MFVGP_ST_EOBS_NONFINITE = 4,
    MFVGP_ST_REFINED = 8,        /* >=1 interval took the adaptive path      */
    MFVGP_ST_REFINE_LIMIT = 16,  /* adaptive path hit scipy's limit=100      */
    MFVGP_ST_SYNC_TIMEOUT = 32   /* internal: a CTA gave up waiting for its neighbour's mailbox
                                    (never expected; the gradient is then incomplete)        */
};

typedef struct mfvgp_handle mfvgp_handle;

int mfvgp_version(void);
const char *mfvgp_last_error(void);

/* number of CUDA devices visible (0 => every compute call fails loudly)    */
int mfvgp_device_count(void);

/*
 * Workspace for one FreeEnergy object: FreeEnergy.__init__
 * (free_energy.py:30-179).  D = state dimension (sigma.size, :95),
 * M = len(obs_times) (:150), d = observed dimensions (:136-147).
 * L96 needs D >= 4 (lorenz_96.py:60-63); DW/OU need D = 1; L63 D = 3.
 * interval range [n_begin, n_end) of the L = M-1 integrated intervals
 * (free_energy.py:490) is the shard this handle evaluates; pass 0, -1 for all.
 */
int mfvgp_create(mfvgp_handle **out, int system, int D, int M, int d,
                 int device, int n_begin, int n_end);
int mfvgp_destroy(mfvgp_handle *h);

/* drift / diffusion parameters: sde.theta, sde.sigma captured at
 * free_energy.py:76-77 and replaceable through the setters at :241-279     */
int mfvgp_set_sde(mfvgp_handle *h, const double *sigma_h /*[D]*/,
                  const double *theta_h, int ntheta);

/* prior moments (free_energy.py:68-72, setters :191-229)                   */
int mfvgp_set_prior(mfvgp_handle *h, const double *mu0_h /*[D]*/,
                    const double *tau0_h /*[D]*/);

/* observation model (free_energy.py:90-177): obs_mask NULL = all observed  */
int mfvgp_set_obs(mfvgp_handle *h, const double *obs_times_h /*[M]*/,
                  const double *obs_values_h /*[M][d]*/,
                  const double *obs_noise_h /*[d]*/,
                  const uint8_t *obs_mask_h /*[D] or NULL*/);

/*
 * FreeEnergy.E_sde(mean_pts, vars_pts, output_gradients)
 * (free_energy.py:461-565; single_interval :338-459; the quadrature is
 * scipy.integrate.quad_vec as called at :405-417).
 * out_d[0] = Esde.  dEsde_dm_d [L][D][4], dEsde_ds_d [L][D][3] (may be NULL
 * when want_grad = 0).  status_d[0] receives MFVGP_ST_* bits.
 */
int mfvgp_esde(mfvgp_handle *h, const double *mean_pts_d, int64_t ldm,
               const double *vars_pts_d, int64_t lds, int want_grad,
               double *out_d, double *dEsde_dm_d, double *dEsde_ds_d,
               int32_t *status_d, void *cuda_stream);

/*
 * FreeEnergy.E_cost(x, output_gradients) (free_energy.py:651-751), including
 * E_kl0 (:281-336) and E_obs (:567-649).  out_d[0..3] = F, E0, Esde, Eobs.
 * grad_d [D*(3M+4) + D*(2M+3)] (may be NULL when want_grad = 0).
 */
int mfvgp_ecost(mfvgp_handle *h, const double *x_d, int want_grad,
                double *out_d /*[4]*/, double *grad_d, int32_t *status_d,
                void *cuda_stream);

/* Same calls on HOST buffers (what a numpy caller of the reference passes):
 * the copies to and from the device happen inside, then the stream is
 * synchronised.  status_h receives the MFVGP_ST_* bits.                    */
int mfvgp_esde_host(mfvgp_handle *h, const double *mean_pts_h,
                    const double *vars_pts_h, int want_grad, double *Esde_h,
                    double *dEsde_dm_h, double *dEsde_ds_h, int32_t *status_h);
int mfvgp_ecost_host(mfvgp_handle *h, const double *x_h, int want_grad,
                     double *out_h /*[4]*/, double *grad_h, int32_t *status_h);

/*
 * E_cost on host buffers with ONE copy in each direction: buf_h holds
 * n + 8 doubles, n = D*(3M+4) + D*(2M+3).  On return buf_h[0..n) = gradient
 * (untouched when want_grad = 0), buf_h[n..n+4) = F, E0, Esde, Eobs and the
 * int32 at buf_h + n + 4 = MFVGP_ST_* bits.  Meant for page-locked buffers
 * from mfvgp_host_alloc (then both copies are plain DMA transfers).
 */
int mfvgp_ecost_host_packed(mfvgp_handle *h, const double *x_h, int want_grad,
                            double *buf_h /*[n + 8]*/);

/* Page-locked host memory from a small recycling pool (what numpy results of
 * FreeEnergy.E_cost live in, so the device->host copy is a direct DMA).      */
void *mfvgp_host_alloc(int64_t bytes);
void mfvgp_host_free(void *p);

/*
 * FreeEnergy.E_kl0(m0, s0, output_gradients) (free_energy.py:281-336) and
 * FreeEnergy.E_obs(mean_pts[d][M], vars_pts[d][M], output_gradients)
 * (:567-649) as stand-alone calls; E_cost computes both terms itself.
 * Returned energies carry the reference's 1/2 factor (:335, :648).  Device-pointer entries
 * (caller's stream, no synchronisation) and host-buffer entries (copies inside).
 */
int mfvgp_ekl0(mfvgp_handle *h, const double *m0_d /*[D]*/, const double *s0_d /*[D]*/,
               int want_grad, double *E0_d /*[1]*/, double *dE0_dm0_d /*[D]*/,
               double *dE0_ds0_d /*[D]*/, int32_t *status_d, void *cuda_stream);
int mfvgp_eobs(mfvgp_handle *h, const double *mean_obs_d /*[d][M]*/,
               const double *vars_obs_d /*[d][M]*/, int want_grad, double *Eobs_d /*[1]*/,
               double *dEobs_dm_d /*[d][M]*/, double *dEobs_ds_d /*[d][M]*/, int32_t *status_d,
               void *cuda_stream);
int mfvgp_ekl0_host(mfvgp_handle *h, const double *m0_h /*[D]*/,
                    const double *s0_h /*[D]*/, int want_grad, double *E0_h,
                    double *dE0_dm0_h /*[D]*/, double *dE0_ds0_h /*[D]*/,
                    int32_t *status_h);
int mfvgp_eobs_host(mfvgp_handle *h, const double *mean_obs_h /*[d][M]*/,
                    const double *vars_obs_h /*[d][M]*/, int want_grad,
                    double *Eobs_h, double *dEobs_dm_h /*[d][M]*/,
                    double *dEobs_ds_h /*[d][M]*/, int32_t *status_h);

/*
 * Gradients of the free energy with respect to the drift parameters theta and the diffusion
 * coefficients Sigma.  The reference has the hooks and no
 * implementation: the sde_theta / sde_sigma setters "used only if we optimize the drift /
 * diffusion parameters" (free_energy.py:241-279) and the closing note of E_cost (:745-748).
 * Only E_sde depends on them (E_kl0 :281-336 and E_obs :567-649 do not):
 *   dtheta [ntheta] = dEsde/dtheta_k = 1/2 sum_n sum_i 1/Sigma_i int_n dE_i/dtheta_k dt
 *   dsigma [D]      = dEsde/dSigma_i = 1/2 sum_n [-1/Sigma_i^2 int_n E_i dt
 *                                                  + 1/Sigma_i int_n dE_i/dSigma_i dt]
 * with the integrands E_i of energy_functions/ *.sym and the fixed 42-node Gauss-Kronrod rule
 * that quad_vec (free_energy.py:405-417) returns when it converges after its first bisection.
 * Where that rule is not converged (quad_vec would bisect further) a composite GK21 rule on 16
 * panels is used and MFVGP_ST_REFINED raised: the values then agree with quad_vec of the same
 * integrand to quad_vec's accuracy.  MFVGP_ST_ESDE_NONFINITE: a non-finite result.
 * x as in mfvgp_ecost.
 * On a sharded handle (mfvgp_comm_init) the call is collective and every rank receives the sums.
 */
int mfvgp_eparam(mfvgp_handle *h, const double *x_d, double *dtheta_d /*[ntheta]*/,
                 double *dsigma_d /*[D]*/, int32_t *status_d, void *cuda_stream);
int mfvgp_eparam_host(mfvgp_handle *h, const double *x_h, double *dtheta_h /*[ntheta]*/,
                      double *dsigma_h /*[D]*/, int32_t *status_h);

/*
 * SCG.minimize(x0) over E_cost (scaled_cg.py:107-386 as driven by
 * FreeEnergy.find_minimum, free_energy.py:815-830).  Iterates, gradients and
 * the line-search dot products stay on the device; x_d is updated in place.
 * fx_hist_h / dfx_hist_h [max_it] = SCG._stats["fx"], ["dfx"] (:284-285).
 * result_h[0] = fx returned, result_h[1] = nit, result_h[2] = func_eval,
 * result_h[3] = OR of status bits seen.
 */
int mfvgp_scg(mfvgp_handle *h, double *x_d, int max_it, double x_tol,
              double f_tol, double *fx_hist_h, double *dfx_hist_h,
              double *result_h /*[4]*/, void *cuda_stream);
int mfvgp_scg_host(mfvgp_handle *h, double *x_h, int max_it, double x_tol,
                   double f_tol, double *fx_hist_h, double *dfx_hist_h,
                   double *result_h /*[4]*/);

/*
 * Multi-GPU: one process (rank) per GPU, the L intervals sharded in contiguous
 * blocks (the reference's own unit of parallelism, free_energy.py:500-508).
 * Rank 0 calls mfvgp_comm_unique_id and distributes the 128 bytes (any host
 * channel); every rank then calls mfvgp_comm_init on a handle created with its
 * interval block [n_begin, n_end).  Afterwards mfvgp_ecost / mfvgp_scg are
 * collective calls: x_d / grad_d keep the full reference layout, each rank
 * reads and writes only its own columns (plus the one column it shares with
 * the next rank).  Per evaluation each rank's scalar record (F, E0, Esde, Eobs, the six
 * status bits) goes to every rank and its part of the gradient of a shared column to that
 * neighbour only.  mfvgp_comm_init maps every rank's mailbox into every other rank (CUDA IPC
 * handles exchanged through the communicator), after which the exchange is plain stores over
 * NVLink plus release / acquire flags inside the evaluation's own kernels; records are folded
 * in rank order, so every rank holds identical scalars and takes identical SCG branches.
 * If the ranks cannot map each other's memory, or with MFVGP_XCHG=nccl|split, the exchange is
 * one ncclAllGather of record + shared columns (or ncclSend/ncclRecv of the columns + an
 * all-gather of the record).  NCCL is bound at run time (dlopen).
 */
int mfvgp_comm_unique_id(void *id128_out /*128 bytes*/);
int mfvgp_comm_init(mfvgp_handle *h, int rank, int world, const void *id128);

/*
 * Posterior reconstruction on a fine time grid: what the
 * notebooks do after find_minimum with symbolics.get_local_polynomials
 * (src/numerical/symbolics.py:102-149; examples/example_500D.ipynb:537-575).
 * Tx = [t0, obs_times..., tf] (M+2 values, M+1 intervals).  Each interval is sampled at
 * `num` points of linspace(ti, tj, num, endpoint=False), plus tf at the end:
 * P = (M+1)*num + 1 samples.  mean_pts [D][ldm >= 3M+4], vars_pts [D][lds >= 2M+3]
 * (log-variances if log_vars != 0).  Outputs tt [P], mt [D][P], st [D][P].
 * _d: device pointers on the current device; _h: host buffers (copied inside).
 */
int mfvgp_reconstruct(int D, int M, int num, const double *Tx_d, const double *mean_pts_d,
                      int64_t ldm, const double *vars_pts_d, int64_t lds, int log_vars,
                      double *tt_d, double *mt_d, double *st_d, void *cuda_stream);
int mfvgp_reconstruct_host(int device, int D, int M, int num, const double *Tx_h,
                           const double *mean_pts_h, const double *vars_pts_h, int log_vars,
                           double *tt_h, double *mt_h, double *st_h);

/*
 * Initial state for find_minimum from an initial sample path that already lives on the device
 * (examples/example_500D.ipynb cells 17 and 19, :402-428): means = the
 * path at the mean support indices mp_idk [3M+4] (cell 17), variances = var_base + var_spread *
 * u [D][2M+3] (the notebook: 4 + 2 * np.random.rand), stored as logs.  path_d [D][T]; x_d is the
 * wire format of mfvgp_ecost (free_energy.py:678-684).
 */
int mfvgp_initial_state(int D, int M, int64_t T, const double *path_d, const int64_t *mp_idk_d,
                        const double *u_d, double var_base, double var_spread, double *x_d,
                        void *cuda_stream);

/*
 * Batched evaluation: B independent states x_d [B][n] of the SAME problem in two launches
 * (out_d [B][8]: F, E0, Esde, Eobs; grad_d [B][n]; status_d [B]).  Serves the n + 1 evaluations
 * scipy.optimize.check_grad makes behind find_minimum(check_gradients=True)
 * (src/var_bayes/free_energy.py:805, :858) and multi-start runs.  Small-problem path (DW, OU,
 * L63, Lorenz 96 up to 37 888 cells), one GPU.  Fixed rule + guard only: an entry whose status has
 * MFVGP_ST_REFINED set holds the fixed-rule value of a state on which quad_vec would bisect
 * further -- re-evaluate it with mfvgp_ecost.  Workspaces grow to the largest B seen.
 */
int mfvgp_ecost_batch(mfvgp_handle *h, int B, const double *x_d, int want_grad, double *out_d,
                      double *grad_d, int32_t *status_d, void *cuda_stream);

/*
 * Lorenz 96 sample path: Lorenz96.make_trajectory
 * (src/dynamical_systems/lorenz_96.py:103-185) -- burn-in of nburn Euler steps of size
 * delta_t from x = theta (+0.01 at D/2), then dim_t - 1 Euler-Maruyama steps of size dt
 * with the noise increments ek_td [dim_t][D] (time major; row 0 unused; the reference's
 * sqrt(sigma dt) * standard_normal((D, dim_t)), transposed).  x_out_d [dim_t][D].
 * Bit-identical to the reference's arithmetic (separately rounded multiplies and adds).
 * status_h: bit 0 = NaN after the burn-in, bit 1 = NaN on the path (the reference raises).
 */
int mfvgp_l96_trajectory(int D, double theta, double dt, int dim_t, int nburn, double delta_t,
                         const double *ek_td_d, double *x_out_d, int32_t *status_h,
                         void *cuda_stream);

/*
 * Double-well, Ornstein-Uhlenbeck and Lorenz 63 sample paths: DoubleWell.make_trajectory
 * (src/dynamical_systems/double_well.py:40-91), OrnsteinUhlenbeck.make_trajectory
 * (ornstein_uhlenbeck.py:38-79), Lorenz63.make_trajectory (lorenz_63.py:102-181).
 * npaths independent paths per call (one thread each): x0_d [npaths][dims] is the state
 * before the burn-in (nburn Euler steps of size delta_t; the reference: 5000 x 1e-3 for
 * L63, none for DW / OU), ek_d [npaths][dim_t][dims] the noise increments (row 0 unused),
 * x_out_d [npaths][dim_t][dims]; dims = 1 (DW, OU) or 3 (L63).  theta_h: DW (theta),
 * OU (theta, mu), L63 (sigma, rho, beta).  Bit-identical to the reference's arithmetic.
 * status_h [npaths]: bit 0 = NaN after the burn-in, bit 1 = NaN drift on the path.
 */
int mfvgp_small_trajectory(int sys, const double *theta_h, int ntheta, double dt, int dim_t,
                           int nburn, double delta_t, int npaths, const double *x0_d,
                           const double *ek_d, double *x_out_d, int32_t *status_h,
                           void *cuda_stream);

/* Which kernels this handle runs, as text ("system=L96 path=walk te=128 tn=32
 * ntiles=68 nruns=128 ..."): for benchmark reports and logs.  Returns the
 * length written (excluding the terminating 0).                            */
int mfvgp_describe(const mfvgp_handle *h, char *buf, int buflen);

/* counters for benchmarks: number of kernels this handle has launched      */
int64_t mfvgp_launch_count(const mfvgp_handle *h);

/* Diagnostics of the adaptive quadrature of the last evaluation:
 * info_h [L][2] = number of bisections quad_vec's loop performed for the
 * energy / dS integrand of each interval (0 = fixed 42-node rule was proven
 * sufficient; k bisections correspond to quad_vec neval = 21 + 42 k).      */
int mfvgp_refine_info(mfvgp_handle *h, int32_t *info_h /*[L][2]*/);

/* Per-launch CUDA-event timing of the dominant (Esde) kernel on the stream it
 * is launched on.  Returns the accumulated time and launch count since the
 * last call and then switches recording on (enable != 0) or off.           */
int mfvgp_timing(mfvgp_handle *h, int enable, double *total_ms, int64_t *count);

/* FP64 FMA micro-benchmark used as the roofline denominator (returns
 * achieved TFLOP/s on the handle's device, FMA = 2 FLOP)                   */
int mfvgp_fp64_peak(int device, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* MFVGP_H_ */
