/* pmp_b200.h — C-ABI of the B200-native PMP-MCMC hot path.
 *
 * The reference (guifengye1/PMP-MCMC) has no FFI; its seams are Python callables and one
 * CUDA kernel signature.  Every entry point below names the reference interface it replaces
 * (paths relative to the reference root):
 *
 *   sweep        log_likelihood_kernel<<<>>>           simple_net/MP_and_PMP_time_analysis/500/500_MP.cu:10-36
 *                (binary-table variant)                simple_net/MP_and_PMP_time_analysis/500/500_PMP.cu:10-33
 *                (general-table variant)               simple_net/MH_MP_PMP_Compare_convergence/conv_pmp.cu:10-36
 *                BayesNet.loglik                       simple_net/lb.py:103-108
 *                normal()/banana/normal(mu,cov)        simple_sampling/error/error.py:11-14, banana_data.ipynb cell 2,
 *                                                      complex_nets/correlation/com_dim.py:13-21
 *                loss(net) (FC MLP, CE/10)             complex_nets/Mnist/FC/PMP_FC.py:21-44
 *                d-dimensional logistic / Gaussian     (extension of lb.py:100-108)
 *   propose      host mt19937 generators               500_MP.cu:177-185, 500_PMP.cu:170-179, conv_pmp.cu:182-197,
 *                update()/tree loops                   lb.py:131-136,268-272,354-360
 *   accept       GMOptimizer.step                      lb.py:139-164
 *                preMOptimizer.step                    lb.py:206-258   (error.py:96-128, com_dim.py:39-78, PMP_FC.py:105-143)
 *                GMpreOptimizerV2.step                 lb.py:304-345   (error.py:151-183)
 *                MetropolisOptimizer.step              lb.py:55-73     (error.py:17-40, MH_FC.py:92-119)
 *                host exp/discrete_distribution        500_MP.cu:207-222
 *
 * Conventions: every function returns 0 on success or a negative pmp_status; the message of
 * the last failure on the calling thread is pmp_last_error().  A pmp_ctx owns its device
 * memory, one CUDA stream and (world_size > 1) one NCCL communicator; callers own every host
 * buffer.  A ctx is not thread-safe.  All work is enqueued on the ctx stream; only functions
 * documented as "blocking" synchronise with the host.  There is no CPU fallback: when no
 * CUDA device is usable pmp_create fails with PMP_ERR_CUDA.
 */
#ifndef PMP_B200_H
#define PMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMP_B200_ABI_VERSION 1

typedef struct pmp_ctx pmp_ctx;

typedef enum pmp_status {
    PMP_OK = 0,
    PMP_ERR_ARG = -1,     /* bad argument / state (e.g. accept before loglik) */
    PMP_ERR_CUDA = -2,    /* CUDA runtime error or no device                  */
    PMP_ERR_NCCL = -3,    /* NCCL missing or failed                           */
    PMP_ERR_ALLOC = -4,
    PMP_ERR_UNSUPPORTED = -5
} pmp_status;

/* Proposal-set shape.  P = number of candidate states evaluated per iteration.
 *   FLAT   : node 0 = current state, node i = node 0 + alpha*eps_i           P = b      (b = N+1)
 *   BINARY : node k+2^l = node k + alpha*eps, l<D, k<2^l                     P = 2^D
 *   BARY   : node k+b^l*(j+1) = node k + alpha*eps, l<D, j<b-1, k<b^l        P = b^D    (b = N+1) */
typedef enum pmp_tree { PMP_TREE_FLAT = 0, PMP_TREE_BINARY = 1, PMP_TREE_BARY = 2 } pmp_tree;

typedef enum pmp_target {
    PMP_TARGET_LINEAR_GAUSS = 0, /* y ~ N(b0 + b1 x, sigma^2), theta = (b0,b1,sigma); needs set_data_linear */
    PMP_TARGET_NORMAL1D = 1,     /* log N(theta; mu, sigma) (error.py:11-14), dim 1, params {mu, sigma}     */
    PMP_TARGET_BANANA = 2,       /* -x1^2/2 - (x2 - 2(x1^2-5))^2/2 (banana_data.ipynb cell 2), dim 2        */
    PMP_TARGET_STDNORMAL = 3,    /* log N(theta; 0, I_d) (com_dim.py:13-15 with mu=0, cov=I), any dim       */
    PMP_TARGET_FC = 4,           /* -CE(MLP 784-512-256-128-10)/loss_div (PMP_FC.py:21-44); needs set_data_fc */
    PMP_TARGET_EXTERNAL = 5,     /* log-targets supplied with pmp_write_logtarget (arbitrary loss(net) callables) */
    PMP_TARGET_GLM_LOGISTIC = 6, /* sum_i log sigmoid(s_i x_i.theta), s_i = 2 y_i - 1, theta in R^d; needs pmp_set_data_glm */
    PMP_TARGET_GLM_GAUSS = 7,    /* y_i ~ N(x_i.theta[0:d], theta[d]^2): d coefficients + sigma, dim = d + 1 (lb.py:100-108 with a d-vector covariate) */
    PMP_TARGET_CNN = 8           /* -CE(conv 1->10 5x5, pool, conv 10->20 3x3, 2000-500-10)/loss_div (PMP_CNN.py:22-52); needs pmp_set_data_cnn */
} pmp_target;

/* Acceptance rule: how the P log-targets become log-weights. */
typedef enum pmp_algo {
    PMP_ALGO_MH = 0,       /* P=2. accept node 1 iff u < exp(lt1 - lt0)                   lb.py:65-69           */
    PMP_ALGO_BARKER = 1,   /* P=2. accept node 1 iff u < w1/(w0+w1)                        error.py:29-35        */
    PMP_ALGO_MP = 2,       /* A_j = lt_j + sum_{k!=j} logK(j,k)                            lb.py:144-150, 500_MP.cu:22-31 */
    PMP_ALGO_PSP = 3,      /* binary tree Barker product, A_a = sum_c logsigmoid(...)      lb.py:216-240         */
    PMP_ALGO_PMP = 4,      /* general tree, per-level group softmax product                lb.py:315-330         */
    PMP_ALGO_TABLE = 5     /* A_p = lt_p + sum_d sum_s logK(from,to) (CUDA PMP variants)   500_PMP.cu:23-30, conv_pmp.cu:22-33 */
} pmp_algo;

/* Draw rule. */
typedef enum pmp_draw {
    PMP_DRAW_PYTHON = 0,   /* P inverse-CDF draws, searchsorted side='right'; next = draws[floor(u_pick*P)]  lb.py:154-163 */
    PMP_DRAW_CUDA = 1,     /* P draws, std::discrete_distribution (lower_bound); next = draws[0]            500_MP.cu:218-243 */
    PMP_DRAW_SINGLE = 2    /* one inverse-CDF draw (side='right'); next = that node                          PMP_FC.py:142-143 */
} pmp_draw;

/* flags for pmp_config.flags */
#define PMP_FLAG_QUIRK_LEVEL_MOD   1u  /* lb.py:330 / error.py:173 "% ((N+1)*(i+1))" typo instead of (N+1)**(i+1)  */
#define PMP_FLAG_QUIRK_TABLE_CONST 2u  /* 500_PMP.cu:198 short/type-punned table upload: transition term is constant */
#define PMP_FLAG_STANDARDIZE       4u  /* A=(A-mean)/std_unbiased before exp (PMP_FC.py:138-141, MP_FC.py:116-119)  */
#define PMP_FLAG_KERNEL_MEAN       8u  /* MP kernel term = sum_k mean_dim logK / P (MP_FC.py:107-114) instead of sum */
#define PMP_FLAG_NO_KERNEL_TERM   16u  /* drop the proposal-kernel term entirely (symmetric kernels in PSP cancel anyway) */
#define PMP_FLAG_UNIFORM_PROPOSAL 32u  /* increments alpha*(2u-1), u ~ U[0,1): random.uniform(-alpha, alpha) of SP (error.py:27) */

typedef struct pmp_config {
    int32_t tree;        /* pmp_tree   */
    int32_t b;           /* branching (N+1); FLAT: P = b; BINARY: ignored (2) */
    int32_t depth;       /* D; FLAT: ignored */
    int32_t dim;         /* parameter dimension */
    int32_t target;      /* pmp_target */
    int32_t algo;        /* pmp_algo   */
    int32_t draw;        /* pmp_draw   */
    uint32_t flags;
    float alpha;         /* proposal step (std of the additive normal) */
    float scale;         /* log-target = loglik / scale  (10, 1000, 2000 in the .cu files; n/50 in lb.py:108; loss_div for FC) */
    float kernel_sigma;  /* std used inside logK; the reference always uses 1; com_dim.py uses its global sigma */
    float target_p0;     /* NORMAL1D: mu   */
    float target_p1;     /* NORMAL1D: sigma */
    float mh_temperature;/* MH: ratio = exp(mh_temperature*(lt1-lt0)); 1 in lb.py, 10000 in MH_FC.py:99 (with scale=1) */
} pmp_config;

const char* pmp_last_error(void);
int pmp_abi_version(void);

/* nccl_unique_id: 128 bytes from ncclGetUniqueId on rank 0 (see pmp_nccl_unique_id), NULL when world_size == 1. */
int pmp_create(pmp_ctx** out, int device, int world_size, int rank, const void* nccl_unique_id);
int pmp_destroy(pmp_ctx* ctx);
int pmp_nccl_unique_id(void* out128);
int pmp_device_info(pmp_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

int pmp_configure(pmp_ctx* ctx, const pmp_config* cfg);   /* allocates proposal/weight/trace buffers for P nodes */
int pmp_num_nodes(pmp_ctx* ctx);                          /* P, or negative status */

/* Data of the linear-Gaussian model.  x,y are HOST pointers to this rank's shard (n_local points starting at
 * global index n_offset); copied to the device on the ctx stream (blocking).  Replaces cudaMemcpy x,y 500_MP.cu:150-151. */
int pmp_set_data_linear(pmp_ctx* ctx, const float* x, const float* y, int64_t n_local, int64_t n_offset, int64_t n_global);

/* Chain state (host pointers, dim floats). set_state also resets nothing else. Blocking. */
int pmp_set_state(pmp_ctx* ctx, const float* theta, int dim);
int pmp_get_state(pmp_ctx* ctx, float* theta, int dim);

/* Philox4x32-10 key and iteration counter (device-resident; advanced by pmp_accept). */
int pmp_seed(pmp_ctx* ctx, uint64_t seed, uint64_t iteration);
int pmp_get_iteration(pmp_ctx* ctx, uint64_t* iteration);   /* blocking */

/* Fill proposals[P,dim] (node 0 = current state) from the Philox stream of the current iteration. */
int pmp_propose(pmp_ctx* ctx);
/* Stream injection both ways (blocking): proposals are row-major [P, dim] float32. */
int pmp_read_proposals(pmp_ctx* ctx, float* out, int64_t count);
int pmp_write_proposals(pmp_ctx* ctx, const float* in, int64_t count);

/* The P x n sweep: log-target of every node (plus the cross-rank all-reduce when world_size > 1).
 * out_host (nullable): P doubles, log-target_p = loglik_p / scale (blocking only when non-NULL). */
int pmp_loglik(pmp_ctx* ctx, double* out_host);
int pmp_write_logtarget(pmp_ctx* ctx, const double* in, int64_t count);   /* PMP_TARGET_EXTERNAL / injection */

/* Acceptance: log-weights, categorical draw(s), state update, iteration += 1.
 * uniforms (nullable → device Philox): PYTHON: P+1 doubles (P draws then the pick); CUDA: P; SINGLE/MH/BARKER: 1.
 * idx_out (nullable): the draws (P, or 1), then blocking.  next_out (nullable): index of the new current state. */
int pmp_accept(pmp_ctx* ctx, const double* uniforms, int64_t n_uniforms, int32_t* idx_out, int32_t* next_out);
int pmp_read_logweights(pmp_ctx* ctx, double* out, int64_t count);   /* A_p of the last accept (blocking) */

/* Device-resident chain: iters x (propose → sweep → [allreduce] → accept) with no host round trip; blocking at the end
 * only if sync != 0.  Trace rows are kept in a device ring buffer sized by pmp_trace_config. */
#define PMP_TRACE_STATE   1u   /* new current state per iteration: [iters, dim] float                                  */
#define PMP_TRACE_NEXT    2u   /* index of the accepted node per iteration: [iters] int32                              */
#define PMP_TRACE_DRAWS   4u   /* all draws per iteration: [iters, P] int32                                            */
#define PMP_TRACE_SAMPLES 8u   /* parameters of all draws per iteration: [iters, P, dim] float (reference fit() traces) */
#define PMP_TRACE_LOGW   16u   /* log-weights per iteration: [iters, P] double                                         */
int pmp_trace_config(pmp_ctx* ctx, int64_t max_iters, uint32_t what);
int pmp_run(pmp_ctx* ctx, int64_t iters, int sync);
int pmp_sync(pmp_ctx* ctx);

/* Co-scheduled chains.  A single chain is a dependency loop (sweep on all SMs → acceptance on one → next sweep), so the sweep
 * SMs idle while their chain is being accepted.  pmp_run_multi runs n_ctx (<= 32) INDEPENDENT chains — one pmp_ctx each: own
 * state, Philox key, trace — in one cooperative kernel, sweeping chain B while chain A is accepted.  Each chain's results are
 * bit-identical to the same ctx run alone with pmp_run.  Requirements: one device, world_size 1, linear-Gaussian target, the
 * same tree / algo in every ctx, and every ctx sharing ctxs[0]'s device copy of the data (pmp_share_data: dst aliases src's
 * data without a copy; src must outlive dst's use of it).  This is the independent-repeats pattern of the reference's
 * experiments (error.py:191-213 runs 20 repeats; the ESS runs use one process per GPU). */
int pmp_share_data(pmp_ctx* dst, pmp_ctx* src);
/* world_size > 1: exchange of the per-node integer sums through NVLink peer memory INSIDE the chain kernel instead of an NCCL
 * call between kernels.  pmp_peer_exchange_handle allocates this rank's exchange buffer and returns its 64-byte CUDA IPC
 * handle; the caller gathers the handles of all ranks (any transport) and passes them, rank-major, to pmp_peer_exchange_attach.
 * With the peers attached pmp_run_multi runs the sharded chains in ONE cooperative kernel per GPU: sweep on the local shard,
 * partial sums stored straight into the peers' buffers, acceptance replicated.  Without it pmp_run_multi falls back to one
 * stream + NCCL communicator per chain.  Call both on every rank, for every ctx, in the same order. */
int pmp_peer_exchange_handle(pmp_ctx* ctx, void* out64);
int pmp_peer_exchange_attach(pmp_ctx* ctx, const void* handles, int n_ranks);
int pmp_run_multi(pmp_ctx** ctxs, int n_ctx, int64_t iters, int sync);
int pmp_run_multi_timed(pmp_ctx** ctxs, int n_ctx, int64_t iters, float* total_ms);   /* CUDA events around the joint kernel; blocking */
/* Copy out the rows recorded since the last pmp_trace_config / pmp_trace_reset (blocking); any pointer may be NULL. */
int pmp_read_trace(pmp_ctx* ctx, int64_t max_iters, float* state, int32_t* next, int32_t* draws, float* samples,
                   double* logw, int64_t* n_recorded);
int pmp_trace_reset(pmp_ctx* ctx);
/* Diagnostics of the recorded STATE trace, reduced on the device (blocking): per coordinate mean[dim], variance[dim] (1/n) and
 * autocovariances acov[(max_lag+1), dim] (acov[k,j] = 1/n sum_t (x_t - m)(x_{t+k} - m)); the mean squared jump distance; the fraction
 * of iterations whose accepted node is not node 0 (-1 without a NEXT trace).  ESS/s and MSJD/s are the reference's headline
 * comparison (README.md:56), computed offline there from the dumped sample files.  Any output pointer may be NULL. */
int pmp_trace_diagnostics(pmp_ctx* ctx, int max_lag, double* mean, double* var, double* acov, double* msjd, double* move_rate, int64_t* n_rows);

/* Timing helpers for bench.py: CUDA events on the ctx stream. pmp_run_timed = pmp_run + device time of the whole
 * region; sweep_ms (nullable) = summed device time of the sweep kernel alone over the region. */
int pmp_run_timed(pmp_ctx* ctx, int64_t iters, float* total_ms, float* sweep_ms);
/* reps back-to-back launches of the sweep kernel on the current proposals between two CUDA events (no gaps: the GPU is
 * the bottleneck); ms = total device time.  The partial sums are discarded. */
int pmp_time_sweep(pmp_ctx* ctx, int reps, float* ms);
int pmp_launch_count(pmp_ctx* ctx, int64_t* launches);   /* kernels launched by this ctx since creation */
int pmp_fp32_peak(pmp_ctx* ctx, int packed, double* tflops); /* FFMA (packed=0) / FFMA2 (packed=1) issue-rate microbenchmark */
int pmp_l2_flush(pmp_ctx* ctx);

/* Host-side evaluation of the library's counter-based streams (same bits as the device): uniforms in [0,1) with 53 bits,
 * standard normals in binary64.  stream: 0 proposal increments, 1 draw uniforms, 2 pick uniform, 3 chain initialisation. */
int pmp_stream_uniforms(uint64_t seed, uint64_t iteration, uint32_t stream, uint64_t idx0, int64_t count, double* out);
int pmp_stream_normals(uint64_t seed, uint64_t iteration, uint32_t stream, uint64_t idx0, int64_t count, double* out);

/* ---- Batched independent chains on analytic targets (error.py SP/MP/PSP/PMP, com_dim.py PMP, banana) -------------
 * n_chains chains, each running the configured tree/algo/draw on PMP_TARGET_{NORMAL1D,BANANA,STDNORMAL}; states and
 * samples are float32 on the device, weights in float64.  samples_out layout: [iters, P, dim, n_chains] (chain fastest). */
int pmp_chains_create(pmp_ctx* ctx, int64_t n_chains, const float* init_states /* [n_chains, dim] host, nullable → zeros */);
int pmp_chains_run(pmp_ctx* ctx, int64_t iters, int record_samples);
int pmp_chains_read_states(pmp_ctx* ctx, float* out /* [n_chains, dim] */);
int pmp_chains_read_samples(pmp_ctx* ctx, float* out, int64_t count);
int pmp_chains_run_timed(pmp_ctx* ctx, int64_t iters, int record_samples, float* total_ms);

/* ---- FC model (PMP_FC.py:21-44): X [n,784] float32 row-major, labels int64; theta layout = torch parameter order
 * fc1.weight[512,784], fc1.bias[512], fc2.weight[256,512], fc2.bias, fc3.weight[128,256], fc3.bias, fc4.weight[10,128], fc4.bias. */
int pmp_set_data_fc(pmp_ctx* ctx, const float* X, const int64_t* labels, int64_t n_local, int64_t n_offset, int64_t n_global);

/* ---- d-dimensional linear / logistic regression heads (PMP_TARGET_GLM_*): X [n, d] float32 row-major, y [n] float32 ({0,1} labels
 * for LOGISTIC, responses for GAUSS).  The P x n sweep is ONE tcgen05 GEMM [n, d] x [d, P] with a fused softplus / square epilogue
 * and a warp-shuffle per-node reduction; shards must start at multiples of 32 rows. */
int pmp_set_data_glm(pmp_ctx* ctx, const float* X, const float* y, int64_t n_local, int64_t n_offset, int64_t n_global, int d);

/* ---- CNN model (complex_nets/Mnist/CNN/PMP_CNN.py:22-44; the same Model in MP_CNN.py, MH_CNN.py): X [n,784] float32 row-major (28x28 images),
 * labels int64; theta layout = torch parameter order conv1.weight[10,1,5,5], conv1.bias[10], conv2.weight[20,10,3,3], conv2.bias[20],
 * fc1.weight[500,2000], fc1.bias[500], fc2.weight[10,500], fc2.bias[10] (1 007 590 floats, CNN_model.pkl).  Replaces the loop
 * `weights[all] = exp(-loss(proposal_nets[all]))` PMP_CNN.py:119-120 for PMP_TARGET_CNN; sharded by rows like the FC target. */
int pmp_set_data_cnn(pmp_ctx* ctx, const float* X, const int64_t* labels, int64_t n_local, int64_t n_offset, int64_t n_global);

/* ---- gradient (HMC) variants: complex_nets/Cifar-10/cifar_{SP,MP,PMP}hmc.py, "Bayesian Network Training"/main.py.
 * The potential's gradient is the caller's (autograd through an arbitrary network) and is handed over as a DEVICE pointer, like the log-targets of
 * PMP_TARGET_EXTERNAL; the leapfrog arithmetic, the kinetic energies and the acceptance run on the device.  All float pointers below are device pointers.
 *   begin: p0 = p_init (or p_scale * N(0,1) from the Philox stream (seed, iteration, momentum stream, stream_index*dim + i));  *ke_init = |p0|^2/2;
 *          p = p0 + sign*step*grad_parent/2;  theta_child = theta_parent + sign*step*p          (cifar_PMPhmc.py:128-147, cifar_MPhmc.py:104-125)
 *   end:   p += sign*step*grad_child/2;  *ke_final = |p|^2/2                                     (cifar_PMPhmc.py:148-162)
 *   accept: weights B of the P nodes and one categorical draw with the uniform u (inverse CDF in place of torch.multinomial).
 *          nets_loss[j] = -CrossEntropy of node j (host).  Tree rules: ke_out[c] / ke_in[c] = kinetic energy of the momentum drawn at the parent
 *          of node c for the edge to c / of the momentum that arrived at c (index 0 unused).  MP: ke_out[j] = |p_s[j]|^2/2.  SP: P = 2, ke_out = {K_0, K_1},
 *          index = 1 iff exp(temperature * (-(K_0 + nl_0) + (nl_1 + K_1))) > u (cifar_SPhmc.py:118-126, temperature 1000). */
typedef enum pmp_hmc_rule {
    PMP_HMC_RULE_SP = 0,          /* cifar_SPhmc.py:77-137 */
    PMP_HMC_RULE_MP = 1,          /* cifar_MPhmc.py:77-86  */
    PMP_HMC_RULE_TREE_CIFAR = 2,  /* cifar_PMPhmc.py:76-108: prod_c max(0, 1 - w_old/w_new) | min(1, w_new/w_old) */
    PMP_HMC_RULE_TREE_BNN = 3     /* main.py:67-103: the normalised pair w_new/(w_new + w_old) per level */
} pmp_hmc_rule;
int pmp_hmc_leapfrog_begin(pmp_ctx* ctx, const float* theta_parent, const float* grad_parent, float* theta_child, float* p_child, const float* p_init /* nullable */,
                           int64_t dim, float step, float sign, float p_scale, uint64_t stream_index, double* ke_init /* host */);
int pmp_hmc_leapfrog_end(pmp_ctx* ctx, float* p_child, const float* grad_child, int64_t dim, float step, float sign, double* ke_final /* host */);
int pmp_hmc_accept(pmp_ctx* ctx, int rule, int P, const double* nets_loss, const double* ke_out, const double* ke_in /* nullable for SP, MP */, double u,
                   double temperature, float* weights_out /* nullable, host [P] */, int32_t* index_out /* host */);

#ifdef __cplusplus
}
#endif
#endif /* PMP_B200_H */
