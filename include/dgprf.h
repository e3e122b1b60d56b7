/*
 * libdgprf -- C ABI of the B200-native DGP-RF SG-MCMC sampling hot path.
 *
 * The reference (shixinxing/DGP-RF-MCMC) is pure-Python TensorFlow eager code and
 * exposes NO FFI: its boundary is the Python class surface (kernels/, layers/,
 * likelihoods/, models/, utils.py).  This header is the C boundary those classes
 * bind in this build (ctypes, see dgp-rf-mcmc_b200/dgprf/_ffi.py); each entry point
 * cites the reference interface it replaces (path:line relative to the reference).
 *
 * Conventions
 *   - every function returns 0 on success, <0 = DGPRF_E*; never throws;
 *     dgprf_last_error() gives a thread-local message.
 *   - the CALLER OWNS ALL MEMORY (device pointers, fp32, contiguous row-major);
 *     scratch comes from dgprf_workspace_bytes() + a caller-supplied pointer.
 *   - every function only enqueues work on `stream` (a cudaStream_t passed as
 *     void*) and returns without synchronising; re-entrant across streams.
 *   - a leading "chain" dimension C batches independent chains / stored samples
 *     (one DGP_RF instance == one chain in the reference).  Chain strides are in
 *     floats; stride 0 shares the operand between chains.
 *   - no CPU fallback: without a CUDA device every compute entry point fails with
 *     DGPRF_ECUDA.
 *   - no model state is kept between calls.  What IS remembered, per thread, are plans
 *     that are pure functions of the arguments (the last workspace layout, the last
 *     step-kernel geometry and argument block), keyed by the model description, the
 *     batch size, the mode and a signature of the process environment.
 */
#ifndef DGPRF_H
#define DGPRF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGPRF_OK        0
#define DGPRF_EINVAL   -1   /* bad argument / unsupported shape            */
#define DGPRF_ECUDA    -2   /* CUDA runtime error (message in last_error)   */
#define DGPRF_EWORKSPACE -3 /* workspace too small                          */

#define DGPRF_MAX_LAYERS 8
#define DGPRF_MAX_SEGMENTS 64

#define DGPRF_KIND_RBF 0    /* layers/rf_layers.py:5  RBFLayer */
#define DGPRF_KIND_ARC 1    /* layers/rf_layers.py:51 ARCLayer */
#define DGPRF_LIK_GAUSSIAN 0 /* likelihoods/gaussian.py:6 */
#define DGPRF_LIK_SOFTMAX  1 /* likelihoods/softmax.py:4  */

/* precision of the two GEMMs of a layer (P = in*Omega, F = Phi*W) */
#define DGPRF_PREC_FP32  0  /* SIMT FFMA, fp32 everywhere (parity mode, rtol 1e-4)       */
#define DGPRF_PREC_TF32  1  /* tcgen05 kind::tf32, fp32 accumulate in TMEM                */

/* workspace modes */
#define DGPRF_MODE_EVAL   0 /* forward only, features not saved                           */
#define DGPRF_MODE_TRAIN  1 /* forward saves Phi; backward for W gradients                */
#define DGPRF_MODE_HYPER  2 /* TRAIN + kernel/likelihood hyper-parameter gradients        */

/* One [RF layer, GP layer] pair: layers/rf_layers.py:29-45 | :75-91 followed by
 * layers/GP_weight_layers.py:11-15.  Input of the RF layer is
 * concat([F_{l-1}[:, :d_prev], X[:, :d_x]])  (utils.py:42; layer 0: d_prev=0, d_x=d_in). */
typedef struct dgprf_layer {
    int32_t kind;            /* DGPRF_KIND_*                                               */
    int32_t d_prev;          /* columns taken from the previous GP layer output            */
    int32_t d_x;             /* columns taken from the model input X (concat tail)         */
    int32_t M;               /* out_feature of the RF layer; F = 2M (RBF) | M (ARC)        */
    int32_t g;               /* out_feature of the GP layer (n_gp[l])                      */
    int32_t has_mean;        /* set_nonzero_mean (rf_layers.py:24-27)                      */
    int64_t off_W;           /* offsets (floats) into the W buffer / the hyper buffer      */
    int64_t off_log_amp;
    int64_t off_log_inv_ls;
    int64_t off_mean;        /* valid iff has_mean                                         */
    const float* z;          /* [C?][d_prev+d_x, M] fixed N(0,1) draw (rf_layers.py:22)    */
    int64_t z_cs;            /* chain stride of z                                          */
} dgprf_layer;

/* A DGP_RF instance (models/dgp.py:9-52) or a batch of C of them. */
typedef struct dgprf_model {
    int32_t n_layers;
    int32_t likelihood;      /* DGPRF_LIK_*                                                */
    int32_t d_in;            /* width of X                                                 */
    int32_t d_out;           /* = g of the last layer                                      */
    int32_t n_chains;        /* C                                                          */
    int32_t precision;       /* DGPRF_PREC_*                                               */
    const float* w_base;     /* W_l lives at w_base + chain*w_cs + off_W                   */
    int64_t w_cs;
    const float* h_base;     /* hypers live at h_base + chain*h_cs + off_*                 */
    int64_t h_cs;
    int64_t off_lik_log_var; /* into the hyper buffer; <0: none (Softmax)                  */
    dgprf_layer layer[DGPRF_MAX_LAYERS];
} dgprf_model;

/* One parameter tensor inside the flat parameter buffer (a tf.Variable of the
 * reference, with the per-variable scalar mass `param.M`, models/dgp.py:235-237). */
typedef struct dgprf_segment {
    int64_t offset;          /* floats, multiple of 4                                      */
    int64_t length;          /* floats                                                     */
    float   mass;            /* param.M (1 = identity preconditioner)                      */
    int32_t flags;           /* bit0: add the N(0,1) prior gradient theta/N               */
} dgprf_segment;

const char* dgprf_last_error(void);
int dgprf_version(void);

/* ---- workspace ------------------------------------------------------------------------- */
/* Bytes of scratch a call with (model, B, mode) needs.  The workspace must be ZERO-FILLED once when it is allocated (it
 * holds the grid-barrier words of the fused step kernel and the tickets of the fused slab sums; every kernel leaves them
 * at zero again) and must not be shared by calls running concurrently. */
int dgprf_workspace_bytes(const dgprf_model* m, int B, int mode, size_t* bytes);

/* ---- model-level hot path ---------------------------------------------------------------*/
/* BNN_from_list(_input_cat).__call__  (utils.py:10-16, 32-44) through all layers.
 * X: [C?][B, d_in] (x_cs chain stride).  mode>=TRAIN saves Phi_l in the workspace for the
 * backward.  F_out (nullable): [C][B, d_out] final GP-layer output. */
int dgprf_forward(const dgprf_model* m, const float* X, int64_t x_cs, int B, int mode,
                  void* ws, size_t ws_bytes, float* F_out, void* stream);

/* likelihood.log_prob on the forward result held in the workspace
 * (likelihoods/gaussian.py:18-25, likelihoods/softmax.py:8-15) plus what the evaluation
 * methods need (models/regression_model.py:33-50: squared error = mean over D_out;
 * models/classification_model.py:17-30: argmax == label).
 * Y: [C?][B, d_out] (Gaussian) or [C?][B, 1] float labels (Softmax).
 * ll_rows [C][B], aux_rows [C][B] (se | correct 0/1), ll_sum [C] : all nullable.
 * inv_B > 0 additionally writes dU/dF = -(1/B) d ll/dF into the workspace for
 * dgprf_backward (the seed of tape.gradient, models/dgp.py:198). */
int dgprf_loglik(const dgprf_model* m, const float* Y, int64_t y_cs, int B, int mode,
                 void* ws, size_t ws_bytes, float* ll_rows, float* aux_rows, float* ll_sum,
                 float inv_B, void* stream);

/* Reverse pass of U (replaces tf.GradientTape.gradient, models/dgp.py:194-204).
 * Needs dgprf_forward(mode>=TRAIN) + dgprf_loglik(inv_B>0) on the same workspace.
 * Writes per-row-split partial gradients into the workspace; dgprf_grad_finalize or
 * dgprf_sgmcmc_update consume them.  mode==HYPER also produces the kernel / likelihood
 * hyper-parameter gradients (models/dgp.py:200-204, experiments/utils_training.py:341-354). */
int dgprf_backward(const dgprf_model* m, const float* X, int64_t x_cs, int B, int mode,
                   void* ws, size_t ws_bytes, void* stream);

/* Sum the partial gradients in fixed order into dense buffers laid out like the parameter
 * buffers: gW [C][w_len], gH [C][h_len] (gH nullable unless mode==HYPER).
 * prior_inv_N > 0 adds theta/N (gradient of -log N(theta;0,1)/N, models/dgp.py:171,178-180)
 * to W (and to hypers when prior_hyper != 0). */
int dgprf_grad_finalize(const dgprf_model* m, int B, int mode, void* ws, size_t ws_bytes,
                        float* gW, int64_t gw_cs, float* gH, int64_t gh_cs,
                        float prior_inv_N, int prior_hyper, void* stream);

/* dgprf_grad_finalize restricted to the W tensor of ONE layer (the slice [off_W, off_W + F*g) of gW): the data-parallel
 * split of a large minibatch (the 1/B mean of models/dgp.py:174 taken over the GLOBAL minibatch)
 * reduces a layer's slice across GPUs while the reverse pass of the layers below is still running. */
int dgprf_grad_finalize_layer(const dgprf_model* m, int layer, int B, int mode, void* ws, size_t ws_bytes,
                              float* gW, int64_t gw_cs, float prior_inv_N, void* stream);

/* Host hook of the layered reverse pass (dgprf_backward, and dgprf_gradients with allow_fused == 0): called on the calling
 * thread right after the kernels of layer `layer` have been enqueued, top layer first (the order tape.gradient walks the
 * layers, models/dgp.py:194-204).  Thread-local; hook == NULL clears it.  Nothing is synchronised: the hook orders its own
 * work against `stream` (event record + wait). */
typedef void (*dgprf_layer_hook)(int layer, void* user);
int dgprf_set_backward_hook(dgprf_layer_hook hook, void* user);

/* Two-shot all-reduce (sum) of the flat gradient of the data-parallel split over NVLink peer memory (csrc/k11_peer_allreduce.cu),
 * the library-free alternative to one NCCL all-reduce per step.  The 1/B mean of models/dgp.py:174 over a minibatch whose rows
 * are split over GPUs is the sum of the per-GPU data terms; every replica then applies the update of models/dgp.py:206-216.
 *   bufs[r]: rank r's symmetric allocation [2][n_pad] floats, mapped on every rank (half 0: r's own gradient, half 1: receives
 *            the reduced gradient); sigs[r]: rank r's zero-initialised signal pad (uint32 words; 2*world words are used from
 *            sig_word_offset on).  n_pad % (4*world) == 0.  epoch: 1, 2, 3, ... -- the same on every rank for the same step.
 * Collective: every rank of the group calls it once per step on its own stream; nothing blocks the host.  A rank whose peers
 * never arrive gives up after ~2 s and raises the sticky word dgprf_peer_allreduce_status returns and clears (0 = healthy;
 * else (phase + 1) << 8 | peer of a wait that timed out).  The status call synchronises the device. */
int dgprf_peer_allreduce(void* const* bufs, void* const* sigs, int rank, int world, int64_t n_pad,
                         unsigned int epoch, unsigned int sig_word_offset, void* stream);
int dgprf_peer_allreduce_status(unsigned int* status);

/* The whole gradient pass behind tf.GradientTape in one call (models/dgp.py:186-204 of sgmcmc_update, :246-251 of
 * precond_update, experiments/utils_training.py:341-354 of the M-step):
 *   dgprf_forward(mode) + dgprf_loglik(inv_B) + dgprf_backward + dgprf_grad_finalize.
 * With allow_fused != 0, mode == DGPRF_MODE_TRAIN and inv_B == 1/B (pass inv_B <= 0 for that default) a model the row-fused
 * step kernel takes runs forward, likelihood seed and backward in ONE launch (the kernel of dgprf_sgmcmc_step without its
 * update); every other case runs the layered sequence.  ll_sum [C] (nullable) receives sum_i log p(y_i | f_i). */
int dgprf_gradients(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y, int64_t y_cs, int B, int mode,
                    void* ws, size_t ws_bytes, float* gW, int64_t gw_cs, float* gH, int64_t gh_cs,
                    float prior_inv_N, int prior_hyper, float* ll_sum, float inv_B, int allow_fused, void* stream);

/* The update loop of sgmcmc_update (models/dgp.py:206-216) over a flat parameter buffer:
 *   h = sqrt(lr/N); [m <- N(0,1)]; m <- beta m - h N g + sqrt(2 (1-beta) T M) eps;
 *   theta <- theta + (h/M) m
 * theta/mom: [C][n] (chain stride cs); grad: n_part partial slabs, element (c,p,i) at
 * grad[c*grad_cs + p*part_stride + i], summed in order p=0..n_part-1.
 * Noise: Philox4x32-10 keyed by (seed, chain) with counter (element, step), Box-Muller;
 * eps_inject / mom_inject (nullable, laid out like theta) replace the generated draws --
 * the parity mode (tf.random.normal, models/dgp.py:210,212, is not reproducible). */
int dgprf_sgmcmc_update(float* theta, float* mom, int64_t cs, int64_t n, int n_chains,
                        const float* grad, int64_t grad_cs, int n_part, int64_t part_stride,
                        const dgprf_segment* segs, int n_seg,
                        float lr, float data_size, float momentum_decay, float temperature,
                        int resample_moments, uint64_t seed, uint64_t step,
                        const float* eps_inject, const float* mom_inject, void* stream);

/* One whole sgmcmc_update (models/dgp.py:184-216): forward, likelihood seed, backward and
 * the update of the W buffer (and the hyper buffer when full_bayesian) in one call.
 * theta_w/mom_w are the mutable views of m->w_base (same layout); segs_* describe them.
 * u_out (nullable) [C]: sum_i ll_i of the minibatch (the caller forms U). */
int dgprf_sgmcmc_step(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y,
                      int64_t y_cs, int B, int full_bayesian,
                      float* theta_w, float* mom_w, int64_t w_len,
                      const dgprf_segment* segs_w, int n_seg_w,
                      float* theta_h, float* mom_h, int64_t h_len,
                      const dgprf_segment* segs_h, int n_seg_h,
                      float lr, float data_size, float momentum_decay, float temperature,
                      int resample_moments, uint64_t seed, uint64_t step,
                      const float* eps_w, const float* res_w, const float* eps_h, const float* res_h,
                      void* ws, size_t ws_bytes, float* u_out, void* stream);

/* dgprf_sgmcmc_step for CUDA-graph capture (the sampler drivers replay one captured epoch per launch,
 * experiments/utils_training.py:41-66): every argument is baked into the captured kernel nodes, so the Philox
 * counter would repeat on replay -- here the noise is keyed by step + *step_base_dev, a device-resident
 * 64-bit base the caller advances between replays (no injected-noise pointers in this variant).  The call
 * only enqueues kernels on `stream` (no allocation, no synchronisation), so it may be issued while `stream`
 * is being captured. */
int dgprf_sgmcmc_step_graph(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y,
                            int64_t y_cs, int B, int full_bayesian,
                            float* theta_w, float* mom_w, int64_t w_len,
                            const dgprf_segment* segs_w, int n_seg_w,
                            float* theta_h, float* mom_h, int64_t h_len,
                            const dgprf_segment* segs_h, int n_seg_h,
                            float lr, float data_size, float momentum_decay, float temperature,
                            int resample_moments, uint64_t seed, uint64_t step, const uint64_t* step_base_dev,
                            void* ws, size_t ws_bytes, float* u_out, void* stream);

/* The same step driven from HOST minibatches (the reference's drivers hand numpy / tf.data batches to
 * sgmcmc_update, experiments/utils_training.py:45-61): X_host [B, d_in] and Y_host [B, d_out | 1] are copied
 * to the caller's device staging buffers X_dev / Y_dev with cudaMemcpyAsync on `stream` (pinned host memory
 * makes them truly asynchronous), the step runs, and sum_i ll_i is copied back to u_host (nullable; u_dev is
 * its device staging word).  One chain-shared minibatch (x_cs = y_cs = 0).  Nothing synchronises.
 * zero_copy == 1: the host buffers are pinned (page-locked, device-visible under UVA): the kernels read the
 * minibatch straight from host memory over PCIe/NVLink-C2C and write sum_i ll_i straight to u_host -- no copy
 * calls, but every CTA's reads are bus round trips inside the step (+9 us per step at BASELINE configs[1]).
 * zero_copy == 2 (pinned buffers; X_dev / Y_dev hold TWO minibatches): pipelined staging -- the H2D copies of
 * step n+1 run on a side stream under the kernels of step n (events order the two halves of the staging
 * buffers), the kernels read device memory, sum_i ll_i is still written in place to the pinned u_host.  The side
 * stream and its four events are created on first use, once per thread and device: the one exception to
 * "the library allocates nothing". */
int dgprf_sgmcmc_step_host(const dgprf_model* m, const float* X_host, const float* Y_host, int y_cols, int B,
                           float* X_dev, float* Y_dev, int zero_copy, int full_bayesian,
                           float* theta_w, float* mom_w, int64_t w_len,
                           const dgprf_segment* segs_w, int n_seg_w,
                           float* theta_h, float* mom_h, int64_t h_len,
                           const dgprf_segment* segs_h, int n_seg_h,
                           float lr, float data_size, float momentum_decay, float temperature,
                           int resample_moments, uint64_t seed, uint64_t step,
                           void* ws, size_t ws_bytes, float* u_dev, float* u_host, void* stream);

/* ---- reductions ------------------------------------------------------------------------ */
/* sum_i log N(x_i; 0, 1) over [C][n] -> out[C]  (prior_W, models/dgp.py:129-136; the
 * per-variable reduce_sum(log_gaussian) of :144-146,156,179-180). */
int dgprf_log_prior(const float* x, int64_t cs, int64_t n, int n_chains, float* out, void* stream);

/* Bayesian model average over stored samples (experiments/utils_training.py:79-85,160-166):
 *   out[0] = mean_n( logsumexp_s log_p[s,n] - log S_total )
 *   out[1] = sqrt(mean_{s,n} aux[s,n])   (aux_is_se) | mean_{s,n} aux   (accuracy)
 * lse_cols (nullable) [N] receives logsumexp_s per column so sharded sample sets can be
 * combined (logsumexp of per-rank logsumexps).  scratch: >= 2*ceil(N/256)+2 floats. */
int dgprf_predictive_reduce(const float* log_p, const float* aux, int S, int64_t N, int64_t ld,
                            float log_S_total, int aux_is_se, float* lse_cols, float* out,
                            float* scratch, void* stream);

/* keras Adam step on the hyper buffer (M-step, experiments/utils_training.py:358). */
int dgprf_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, int t, void* stream);

/* Welford update + mass estimate of precond_update (models/dgp.py:259-288). */
int dgprf_welford_update(const float* grad, float* mean, float* m2, int64_t n, int k, void* stream);
int dgprf_mass_estimate(const float* mean, const float* m2, const dgprf_segment* segs, int n_seg,
                        int K, int centered, float* mass_out, void* stream);

/* ---- stand-alone layer ops (the classes called one at a time) --------------------------- */
/* RBFLayer/ARCLayer.__call__ (layers/rf_layers.py:29-45, 75-91): Phi [B, F]. */
int dgprf_rf_features(int kind, const float* X, int B, int d, const float* z, const float* log_inv_ls,
                      const float* log_amp, const float* mean, int M, float* Phi, void* stream);
/* GPLayer.__call__ (layers/GP_weight_layers.py:11-15): out [B, g] = Phi [B, F] @ W [F, g]. */
int dgprf_gp_matmul(const float* Phi, const float* W, int B, int F, int g, float* out, void* stream);
/* Gaussian.log_prob / Softmax.log_prob / Softmax.predict_full on explicit tensors. */
int dgprf_gaussian_log_prob(const float* F, const float* Y, const float* lik_log_var, int B, int D,
                            float* out_rows, void* stream);
int dgprf_softmax_log_prob(const float* F, const float* Y, int B, int C, float* out_rows,
                           float* probs /*nullable [B,C]*/, void* stream);


/* Test hook: the N(0,1) stream the update kernel draws (Philox4x32-10 + Box-Muller) for
 * (seed, chain, step, stream_id: 0 eps / 1 resample of the W buffer, 2/3 of the hyper buffer). */
int dgprf_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t chain, uint64_t step,
                        int stream_id, void* stream);

/* Measurement hook (bench.py): between start and stop every kernel launched by this thread is
 * bracketed by CUDA events on its launch stream; stop synchronises them and returns up to
 * max_records (name[32], milliseconds) pairs in launch order.  Not for production loops. */
int dgprf_profile_start(void);
int dgprf_profile_stop(int max_records, char* names, float* ms, int* n_out);

#ifdef __cplusplus
}
#endif
#endif /* DGPRF_H */
