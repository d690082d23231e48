/*
 * noiseflow_b200 -- C ABI of the B200-native Noise Flow density / sampling engine.
 *
 * Drop-in boundary for the bijector-chain hot path of BorealisAI/noise_flow (pure TF1 Python; it has
 * no FFI of its own, so each entry point cites the Python interface it replaces; paths are relative
 * to the reference repository root).  Plain pointers and sizes only: no torch / CUDA types in the
 * signatures (streams travel as void*).  Unless a function name ends in _host, every data pointer
 * is a DEVICE pointer owned by the caller.  Width-4 models live in host memory only (parameters are passed
 * to each kernel launch by value).  Device memory the library does own, all bounded and released by
 * nf_model_destroy / nf_trainer_destroy: the folded parameter blob of a wide-net model (width != 4; a few hundred KB,
 * plus a stream-ordered temporary of the same size for partial-range / batch-statistics launches), the <= 16 KB
 * parameter images of the small-batch batch-statistics chain, the trainer's variables / Adam slots / workspace, and
 * the staging pool of the _host entry points (which take host buffers).
 *
 * All functions return 0 on success and a negative nf_status otherwise; nf_last_error() returns a
 * thread-local human-readable message.  Nothing here throws, exits or prints.  Entry points are
 * re-entrant: a finalized model may be used concurrently from many host threads (the reference is
 * driven by 16-32 Python threads sharing one session: train_noise_flow.py:38-47,
 * train_dncnn_noiseflow.py:195-198).  nf_model_set_* may run concurrently with launches (every launch
 * snapshots the parameter block under a lock and carries it by value); nf_model_add_* / finalize /
 * destroy must not race with anything on the same handle.
 *
 * Tensor layout: NHWC float32, patch = [32][32][4] (sidd patches: train_noise_flow.py:287-288).
 * Naming follows the reference: "inverse" = data -> latent (likelihood direction),
 * "forward" = latent -> data (sampling direction)  (noise_flow_model.py:394,430).
 */
#ifndef NOISEFLOW_B200_H
#define NOISEFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NF_ABI_VERSION 1

typedef enum nf_status {
    NF_OK = 0,
    NF_ERR_INVALID = -1,      /* bad argument */
    NF_ERR_UNSUPPORTED = -2,  /* shape / width / capacity outside what the sm_100a kernels are built for */
    NF_ERR_CUDA = -3,         /* CUDA runtime error (message carries cudaGetErrorString) */
    NF_ERR_STATE = -4         /* call order (e.g. launch before nf_model_finalize) */
} nf_status;

typedef struct nf_model nf_model;   /* opaque */

/* Reference-shaped parameters of one AffineCoupling + real_nvp_conv_template(width = 4):
 * borealisflows/layers.py:251-375 (coupling), :452-498 (template), :586-613 (conv2d),
 * :651-674 (conv2d_zeros), :378-401 (batch_norm).  Filters use TensorFlow's [kh][kw][in][out]
 * layout exactly as stored in the checkpoint. */
typedef struct nf_coupling_weights {
    const float* l1_w;      /* [3][3][2][4]  model/real_nvp_conv_template*/
    const float* l1_b;      /* [4] */
    const float* bn1_mean;  /* [4]  moving statistics (is_training == False path, layers.py:400) */
    const float* bn1_var;   /* [4] */
    const float* l2_w;      /* [1][1][4][4] */
    const float* l2_b;      /* [4] */
    const float* bn2_mean;  /* [4] */
    const float* bn2_var;   /* [4] */
    const float* last_w;    /* [3][3][5][4]  (input channel 4 = edge indicator, layers.py:555-583) */
    const float* last_b;    /* [4] */
    const float* last_logs; /* [4]  output is multiplied by exp(3 * logs), layers.py:671-673 */
    float rescaling_scale;  /* level0/bijector{i}/rescaling_scale0, layers.py:271-273 */
    float bn_eps;           /* 1e-4, layers.py:378 */
} nf_coupling_weights;

/* Scale-layer kinds: every AffineCouplingSdn* / Gain* / CamSdn template reduces to one of the two
 * (borealisflows/noise_flow_layers/*.py, cond_utils.py); the host computes the per-(camera, ISO)
 * scalars. */
#define NF_SCALE_SDN 1   /* scale = sqrt(a*y + b); table row = {a, b}     e.g. AffineCouplingSdnEx5.py */
#define NF_SCALE_GAIN 2  /* scale = g;             table row = {g, unused} e.g. AffineCouplingGainEx4.py */

/* ---- library ----------------------------------------------------------------------------------- */
int nf_abi_version(void);
const char* nf_last_error(void);
/* Fills in the SM count / max opt-in shared memory of the current device; negative if no usable GPU. */
int nf_device_info(int* sm_count, int* max_smem_optin, int* cc_major, int* cc_minor);

/* ---- model construction: mirrors NoiseFlow.noise_flow_arch (noise_flow_model.py:71-235) -------- */
/* x_shape must be 32x32x4; net_width (hps.width, sidd/ArgParser.py:43) 4 (the shipped configuration: fused warp-per-patch
 * kernel), 8 / 16 (CTA-per-patch CUDA-core kernel), 32 / 64 / 128 / 256 / 512 (tensor-core kernels, tcgen05; 32 also has
 * the CUDA-core kernel; 512 is the reference's default); anything else -> NF_ERR_UNSUPPORTED. */
int nf_model_create(int height, int width, int channels, int net_width, nf_model** out);
int nf_model_destroy(nf_model* m);
/* Conv2d1x1 (layers.py:74-145, bias=False): A / A_inv are [in][out] row-major as produced by
 * matrix_param_lu (matrix_param.py:130-138); log_abs_det = sum(log_S). */
int nf_model_add_conv1x1(nf_model* m, const float* A, const float* A_inv, float log_abs_det);
/* tfb.Permute(permutation) (noise_flow_model.py:80-84): forward y[..., i] = x[..., perm[i]]. */
int nf_model_add_permute(nf_model* m, const int32_t* perm);
int nf_model_add_affine_coupling(nf_model* m, const nf_coupling_weights* w);
/* The clean-image-conditioned couplings of the reference's legacy `revnet2d` models (noise_flow_model.py:237-392, reached
 * when hps.arch is unset).  mode 1 = AffineCouplingCondXY / CondXYG (noise_flow_layers/AffineCouplingCondXY.py:45-79): the net
 * sees concat(x0, yy) -- l1_w is [3][3][6][W] -- and transforms x1; mode 2 = AffineCouplingCondY / CondYG
 * (AffineCouplingCondY.py:44-72): the net sees yy only -- l1_w [3][3][4][W] -- and shifts / scales ALL four channels:
 * last_w [3][3][W+1][8], last_b [8], last_logs [8] (shift = outputs 0..3, log-scale = 4..7).  mode 0 = plain AffineCoupling.
 * The ISO-conditioned `G` variants (real_nvp_conv_template_iso, layers.py:501-547,616-648: W = B1 * iso + B2, b = C1 * iso + C2)
 * are the same couplings with the effective weights of the call's ISO (the reference feeds one ISO per minibatch): set them
 * with nf_model_set_cond_coupling.  Widths 4 / 8 / 16 / 32 (CUDA-core CTA-per-patch kernel); a model that has such a coupling
 * runs all its launches on that kernel. */
#define NF_COUPLING_X 0
#define NF_COUPLING_XY 1
#define NF_COUPLING_Y 2
int nf_model_add_cond_coupling(nf_model* m, int mode, const nf_coupling_weights* w);
int nf_model_set_cond_coupling(nf_model* m, int layer, const nf_coupling_weights* w);
/* table: [n_rows][2] floats; logdet_full_sum = 0 reproduces the Gain/GainEx1/GainEx3 quirk whose
 * log-det is log(scale) instead of 4096*log(scale) (AffineCouplingGain.py:86,96,111,125). */
int nf_model_add_scale(nf_model* m, int kind, int logdet_full_sum, const float* table, int n_rows);
int nf_model_finalize(nf_model* m);
int nf_model_num_layers(const nf_model* m);
/* In-place parameter updates of layer `layer` (index in add order), e.g. after an optimizer step. */
int nf_model_set_conv1x1(nf_model* m, int layer, const float* A, const float* A_inv, float log_abs_det);
int nf_model_set_affine_coupling(nf_model* m, int layer, const nf_coupling_weights* w);
int nf_model_set_scale(nf_model* m, int layer, const float* table, int n_rows);
/* Batched update: between begin and end the nf_model_set_* calls only store; end folds and uploads once (wide nets keep
 * their folded program in a device blob: the new one is uploaded beside the old one, swapped, and the old one retired
 * after the launches that read it). */
int nf_model_begin_update(nf_model* m);
int nf_model_end_update(nf_model* m);
/* Launch tuning: resident patches (warps) per CTA in [1, 16] and CTA count (0 = one per SM). */
int nf_model_set_launch(nf_model* m, int warps_per_cta, int num_ctas);
/* Width 4 -- which kernel runs full-chain / range calls (batch-statistics probes always use the direct-form all-fp32 kernel):
 *   0 (default) = 4: the all-fp32 vertical-Winograd kernel (csrc/nf_wino.cu: both 3x3 convolutions of every coupling net as
 *      F(2,3) along the image rows; 12.6 vs 10.5 M patches/s data -> latent, 11.8 vs 9.4 M latent -> data on one B200; differs
 *      from the direct form by fp32 rounding only).  It runs as one 16-warp CTA per SM, or as two 8-warp CTAs with half the
 *      tensor memory each when it draws its own noise (nf_sample with eps == NULL: +10 %); the environment variable
 *      NF_WINO_CTA_WARPS=8 / 16 forces either shape (an A/B switch, read once per process);
 *   5: the all-fp32 direct-form kernel (csrc/nf_kernels.cu) everywhere;
 *   2: the hybrid kernel (csrc/nf_hybrid.cu): conv-3 of every coupling net on the tensor cores (tcgen05.mma, fp16 hi/lo-split
 *      operands, fp32 accumulation in TMEM), everything else fp32 -- |dNLL| vs the fp32 kernels < 1e-6 nats/dim on the shipped
 *      model; 22 instead of 24 mantissa bits inside conv-3, which a stress model with O(1) random weights shows as 2e-5
 *      instead of 1e-5 relative error of a sampled patch -- hence opt-in;
 *   3: the hybrid kernel for the latent -> data direction only (7 % faster than the direct form there), the default otherwise;
 *   1: older experiment -- both 3x3 convolutions as bf16 hi/lo implicit GEMMs (csrc/nf_tc.cu), full-chain calls with explicit
 *      inputs only; slower than 0.
 * Widths 32 ... 512 -- the tensor-core kernels (csrc/nf_wide_tc.cu, nf_wide_tcs.cu: all three convolutions as tcgen05.mma
 * GEMMs, activations and accumulators in tensor memory) are the default; enable == 0 selects the CUDA-core kernel at
 * width 32 and is refused (NF_ERR_UNSUPPORTED) at 64 ... 512, which have no other kernel. */
int nf_model_set_tensor_cores(nf_model* m, int enable);

/* ---- hot path (device pointers) ---------------------------------------------------------------- */
/* rows: per-patch conditioning-table row (int32, device) or NULL -> default_row for every patch.
 * The reference feeds ONE (iso, cam) per minibatch (sidd/MiniBatchSampler.py:60-64). */

/* NoiseFlow._loss (noise_flow_model.py:458-480): nll[n] = -(sum ldj + log N(z; 0, I)),
 * sdz[n] = sqrt(var(z_n)) (the reference returns their batch mean: use nf_reduce_sums),
 * z (optional) = latent.  x, y: [n][32][32][4]. */
int nf_log_prob(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row,
                int64_t n, float* nll, float* sdz, float* z, void* stream);
/* NoiseFlow.inverse (noise_flow_model.py:394-428): z and per-patch objective increment logdet[n]. */
int nf_inverse(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row,
               int64_t n, float* z, float* logdet, void* stream);
/* NoiseFlow.forward (noise_flow_model.py:430-447): x = chain^-1(z); logdet optional. */
int nf_forward(const nf_model* m, const float* z, const float* y, const int32_t* rows, int32_t default_row,
               int64_t n, float* x, float* logdet, void* stream);
/* NoiseFlow.sample (noise_flow_model.py:449-456,499-504): z = eps * temp, x = forward(z).
 * eps == NULL draws eps in-kernel: Philox4x32-10, key = seed, counter = (pixel, patch, offset),
 * Box-Muller; patch index = patch_base + position in this call. */
int nf_sample(const nf_model* m, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
              float temp, const float* eps, uint64_t seed, uint64_t offset, uint64_t patch_base,
              float* x, void* stream);
/* Per-bijector access (_inverse_and_log_det_jacobian / _forward_and_log_det_jacobian of the
 * bijectors first..last-1 in add order; layers.py:132-140,333-375 and noise_flow_layers/*).
 * direction: 0 = inverse, 1 = forward. `in` and `out` may alias. */
int nf_run_layers(const nf_model* m, int first, int last, int direction, const float* in, const float* y,
                  const int32_t* rows, int32_t default_row, int64_t n, float* out, float* logdet, void* stream);
/* tf.reduce_mean pieces (noise_flow_model.py:478,484), deterministic fp64:
 * sums[0] = sum nll, sums[1] = sum sdz, sums[2] = n  (device double[3]; either input may be NULL). */
int nf_reduce_sums(const float* nll, const float* sdz, int64_t n, double* sums, void* stream);

/* The whole chain with BATCH-STATISTICS BatchNorm -- the reference's is_training == True path
 * (batch_norm, layers.py:388-398), which is what NoiseFlowWrapper.sample_noise_nf feeds
 * (NoiseFlowWrapper.py:85-86) and what train_thread runs (train_noise_flow.py:64-71).  Patches are not
 * independent in this mode: every coupling is probed twice (batch mean / population variance of its conv-1 and
 * conv-2 outputs) and then applied with those statistics -- layer by layer, or, for small batches, inside one
 * cooperative kernel (nf_model_set_bs_small).
 * direction 0: x -> z, outputs as nf_log_prob / nf_inverse; direction 1: z = in * temp (in == NULL: Philox
 * as nf_sample) -> x.  `out` ([n][32][32][4], required) doubles as the in-place state buffer; stats_ws is a
 * device double[8] scratch.  batch_stats_host (optional, host) receives [n_couplings][16] =
 * {mean1[4], var1[4], mean2[4], var2[4]} per coupling in add order so that the caller can apply the
 * moving-average update train_m -= 0.1 * (train_m - m) (layers.py:394-395).  Synchronises `stream`. */
int nf_chain_batch_stats(const nf_model* m, int direction, const float* in, const float* y, const int32_t* rows,
                         int32_t default_row, int64_t n, float temp, uint64_t seed, uint64_t offset,
                         uint64_t patch_base, float* out, float* logdet, float* nll, float* sdz, double* stats_ws,
                         float* batch_stats_host, void* stream);
/* 1 (default): nf_chain_batch_stats runs batches of up to 4096 patches (width-4 chains of [1x1 conv / permutation +] coupling
 * groups and scale layers) as ONE cooperative kernel with no host round trip -- one co-resident CTA per patch up to 296
 * patches on a B200 (the reference sampling script's call pattern is one patch per sess.run, sample_noise_flow.py:44,71),
 * the same kernel walking its patches grid-stride beyond (train_noise_flow.py:165-167 samples whole test minibatches in
 * this mode); 0: always layer by layer (two probe launches + a stream synchronisation each, per coupling). */
int nf_model_set_bs_small(nf_model* m, int enable);

/* ---- training step support: loss and its gradient (train_noise_flow.py:187-198, 50-77) -------------------- */
/* Gradient layout on the host: one block per bijector in add order, offsets[l] .. offsets[l+1] (doubles):
 *   conv1x1  : 16  d loss / d A[in][out]           (chain to the LU variables on the host)
 *   coupling : 285 [l_1/W 72][l_1/b 4][l_2/W 16][l_2/b 4][l_last/W 180][l_last/b 4][l_last/logs 4][rescaling_scale 1]
 *              in exactly the checkpoint's tensor layouts
 *   scale    : 2 * n_rows  d loss / d (a, b) (sdn) or d loss / d (g, -) (gain) per conditioning row
 *   permute  : 0 */
int nf_grad_layout(const nf_model* m, int64_t* offsets /* n_layers + 1 */);
int nf_train_workspace_floats(const nf_model* m, int64_t n, int64_t* n_floats);
/* loss = mean_n nll_n (NoiseFlow.loss, noise_flow_model.py:482-484) and d loss / d(every trainable variable).
 * batch_stats != 0: BatchNorm on batch statistics (is_training=True) incl. its backward reductions; 0: moving
 * statistics.  workspace: device floats (nf_train_workspace_floats); dscratch: device double[512]; grads_host:
 * host doubles (nf_grad_layout); batch_stats_host (optional) as nf_chain_batch_stats; sums_host (optional) as
 * nf_reduce_sums.  Activations are recomputed from the stored layer inputs; the gradients are accumulated in
 * fp64.  Synchronises `stream`. */
int nf_loss_and_grad(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row,
                     int64_t n, int batch_stats, float* workspace, double* dscratch, double* grads_host,
                     float* batch_stats_host, double* sums_host, void* stream);

/* ---- device-resident train step: sess.run([train_op, loss, sd_z]) without host round trips ---------------- */
/* The reference's train thread (train_noise_flow.py:50-77) runs Adam on `loss` with is_training=True
 * (:187-198).  nf_trainer keeps every TF variable, the Adam slots and the step counter in device memory and
 * runs the whole step there: LU assembly of the 1x1 matrices (matrix_param.py:117-130), the per-(camera, ISO)
 * scale tables (cond_utils.py:165-276,432-440), forward with batch-statistics BatchNorm (layers.py:388-398),
 * backward (three passes per coupling), the chain rules back to the LU / scale variables, Adam with
 * TensorFlow's update rule and the BatchNorm moving-average update (layers.py:394-395).  No call below
 * synchronises the stream except the _get_ functions.
 *
 * Variables live in ONE flat float array (order chosen by the caller); every op addresses its variables by
 * offset (-1 = absent).  Ops are listed in data -> latent order. */
typedef enum nf_train_op_kind { NF_TOP_COUPLING = 1, NF_TOP_SCALE = 2 } nf_train_op_kind;
typedef enum nf_train_token {       /* scale tokens whose chain rule is implemented on the device */
    NF_TOKEN_SDN4 = 4, NF_TOKEN_SDN5 = 5, NF_TOKEN_SDN6 = 6, NF_TOKEN_GAIN4 = 14
} nf_train_token;
typedef struct nf_train_op {
    int32_t kind;                 /* nf_train_op_kind */
    /* coupling: the 1x1 conv / permutation in front of it (noise_flow_model.py:79-104) */
    int32_t mix_kind;             /* 0 none, 1 Conv2d1x1 with LU variables, 2 fixed channel permutation */
    int32_t off_P, off_L, off_U, off_logS, off_signS;   /* [4][4], [6], [6], [4], [4] */
    int32_t perm[4];              /* mix_kind 2: out channel o takes in channel perm[o] (tfb.Permute) */
    /* coupling: real_nvp_conv_template variables in checkpoint layout, rescaling_scale0, BatchNorm moving stats */
    int32_t off_w1, off_b1, off_w2, off_b2, off_w3, off_b3, off_logs, off_scale;
    int32_t off_bn1_mean, off_bn1_var, off_bn2_mean, off_bn2_var;
    /* scale layer */
    int32_t token;                /* nf_train_token */
    int32_t off_beta1, off_beta2, off_gain_params, off_cam_params, off_gain_val;
    float c_i;                    /* hps.param_inits[0] (sdn5 / sdn6), cond_utils.py:207,244 */
} nf_train_op;

typedef struct nf_trainer nf_trainer;   /* opaque; owns device memory: variables, Adam slots, workspace */

/* tri_lower / tri_upper: for k in 0..5 the flat position r*4+c of L_vec[k] / U_vec[k] in the strict triangle
 * (matrix_param.py:31-57, TFP fill_triangular order).  trainable: one byte per variable element.
 * Device memory: about 16 KiB * max_batch * (n_ops + 5). */
int nf_trainer_create(const nf_train_op* ops, int n_ops, const float* vars_host, const uint8_t* trainable,
                      int64_t n_vars, int64_t max_batch, const int32_t* tri_lower, const int32_t* tri_upper,
                      float bn_eps, nf_trainer** out);
int nf_trainer_destroy(nf_trainer* t);
/* Length (doubles) of the reduce buffer: [d loss / d var (n_vars)] [sum nll, sum sd_z, n] [batch mean1[4],
 * var1[4], mean2[4], var2[4] per coupling].  Everything a data-parallel step has to sum over ranks, in one buffer. */
int nf_trainer_reduce_len(const nf_trainer* t, int64_t* n_doubles);
/* Enqueue loss + gradient of this rank's batch; fills reduce_buf (device doubles, caller-owned).  rows: per-patch
 * standard conditioning row (cam * 5 + iso index, 0..24) or NULL -> default_row.  batch_stats as nf_loss_and_grad. */
int nf_trainer_loss_and_grad(nf_trainer* t, const float* x, const float* y, const int32_t* rows, int32_t default_row,
                             int64_t n, int batch_stats, double* reduce_buf, void* stream);
/* Enqueue Adam (tf.train.AdamOptimizer semantics: lr_t = lr*sqrt(1-b2^t)/(1-b1^t), var -= lr_t*m/(sqrt(v)+eps))
 * on reduce_buf / world_size, and (update_bn != 0) the BatchNorm moving-average update towards the rank-averaged
 * batch statistics.  Advances the step counter. */
int nf_trainer_apply(nf_trainer* t, const double* reduce_buf, double lr, double beta1, double beta2, double eps,
                     int world_size, int update_bn, void* stream);
int nf_trainer_get_vars(nf_trainer* t, float* vars_host, void* stream);          /* synchronises */
int nf_trainer_set_vars(nf_trainer* t, const float* vars_host, void* stream);    /* synchronises */
int nf_trainer_launches_per_step(const nf_trainer* t, int batch_stats, int* n_launches);
/* Roofline aids for the fused step (bench.py): grid-wide barriers one loss+gradient evaluation crosses, and the measured
 * cost of one such barrier on a cooperative grid of n_ctas CTAs shaped like the step kernel (difference of two launches
 * of reps and 2*reps barriers; synchronises `stream`). */
int nf_trainer_barriers_per_step(const nf_trainer* t, int batch_stats, int* n_barriers);
int nf_probe_grid_barrier(int n_ctas, int reps, float* us_per_barrier, void* stream);
/* enable != 0 (default): nf_trainer_loss_and_grad stages x / y / rows into trainer-owned buffers and replays the
 * launch sequence as ONE CUDA graph (re-captured when n, default_row, batch_stats or reduce_buf change) on an
 * internal stream fenced against `stream` with events; 0: plain launches on `stream`. */
int nf_trainer_set_graph(nf_trainer* t, int enable);
/* Warps per patch-CTA of the coupling passes: 0 (default) = 8 (four image rows per warp: every weight read from shared
 * memory is applied to four rows); 8 / 16 force a shape (16 = two rows per warp, one CTA per SM). */
int nf_trainer_set_cta_warps(nf_trainer* t, int warps);
/* 1 (default): when every patch of the batch can own a co-resident CTA (<= 296 patches at 8 warps, <= 148 at 16 on a
 * B200) the whole loss + gradient evaluation is ONE cooperative kernel (grid barriers around the BatchNorm batch sums)
 * followed by the reduce and chain-rule kernels; larger batches, and enable = 0, use one launch per pass. */
int nf_trainer_set_fused(nf_trainer* t, int enable);

/* squeeze2d / unsqueeze2d (borealisflows/utils.py:30-86): bit-exact index permutation.
 * squeeze_type: 0 = 'chessboard' (also the unknown-type fallback), 1 = 'patch'.
 * H, W, C always describe the UN-squeezed tensor [n][H][W][C]. */
int nf_squeeze2d(const float* x, int64_t n, int H, int W, int C, int factor, int squeeze_type, float* out, void* stream);
int nf_unsqueeze2d(const float* x, int64_t n, int H, int W, int C, int factor, int squeeze_type, float* out, void* stream);

/* ---- evaluation metrics the reference computes right after the path ------------------------------ */
/* calc_baselines (sidd/PatchStatsCalculator.py:92-123): per-patch NLL of x under N(0, var_gauss) and under the
 * camera NLF N(0, y*nlf0 + nlf1); the reference reports their batch means next to the flow's NLL. */
int nf_baseline_nll(const float* x, const float* y, float nlf0, float nlf1, float var_gauss, int64_t n,
                    float* nll_gauss, float* nll_sdn, void* stream);
/* np.histogram(data, edges) as used by get_histogram (sidd/sidd_utils.py:1266-1274) for the marginal KL
 * divergence (:1044-1052, kl_div_3_data :1247-1263): counts[b] += #{edges[b] <= v < edges[b+1]} (last bin
 * closed), comparisons in double -> bit-identical to numpy.  edges: device double[n_bins + 1], counts: device
 * uint64[n_bins], accumulated (zero it first). */
int nf_histogram(const float* data, int64_t count, const double* edges, int n_bins, unsigned long long* counts,
                 void* stream);

/* ---- host-buffer entry points (what NoiseFlowWrapper.sample_noise_nf / sess.run replace) -------- */
/* Host pointers; copies are chunked (4096 patches) and pipelined over four internal streams.  Buffers from
 * nf_host_alloc (pinned) overlap copies with compute; pageable memory works but serialises.
 * sums (host double[3], optional) as nf_reduce_sums. rows_host may be NULL. */
int nf_log_prob_host(const nf_model* m, const float* x_host, const float* y_host, const int32_t* rows_host,
                     int32_t default_row, int64_t n, float* nll_host, float* sdz_host, float* z_host, double* sums_host);
int nf_sample_host(const nf_model* m, const float* y_host, const int32_t* rows_host, int32_t default_row, int64_t n,
                   float temp, const float* eps_host, uint64_t seed, uint64_t offset, float* x_host);
int nf_host_alloc(void** ptr, size_t bytes);   /* cudaHostAlloc (pinned) */
int nf_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* NOISEFLOW_B200_H */
