/* mimo_b200.h -- C ABI of libmimo_b200.so (hand-written sm_100a CUDA kernels for the MIMO U-Net hot path).
 *
 * Drop-in boundary: the reference (antonbaumann/MIMO-Unet) has no FFI of its own; its hot path is
 * the torch.nn call sites listed next to each entry point below (paths relative to the reference checkout). A
 * binding replaces those call sites by passing raw device pointers and the current CUDA stream.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types cross the boundary.
 *   - every entry point returns 0 (MIMO_OK) or a negative mimo_status; the message is available from
 *     mimo_last_error() (thread-local). No exceptions, no exit(), no CPU fallback.
 *   - the library never allocates or frees device memory and never synchronises: all buffers (inputs,
 *     outputs, workspaces) are owned by the caller; every call only enqueues work on `stream`
 *     (a cudaStream_t passed as void*), so calls are safe inside CUDA-graph capture.
 *   - activations are NHWC bf16, optionally with a 1-pixel reflect halo and as a channel slice of a wider
 *     buffer: see mimo_act_t.
 */
#ifndef MIMO_B200_H
#define MIMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMO_B200_VERSION 100

typedef enum {
  MIMO_OK = 0,
  MIMO_ERR_ARG = -1,     /* bad shape / argument */
  MIMO_ERR_ALIGN = -2,   /* pointer or pitch alignment */
  MIMO_ERR_CUDA = -3,    /* CUDA runtime / driver error (launch, tensor-map encode, ...) */
  MIMO_ERR_ARCH = -4,    /* device is not sm_100 */
  MIMO_ERR_STATE = -5    /* plan not bound / call order */
} mimo_status;

/* NHWC bf16 activation view.
 * element (n,h,w,c) at ptr[((n*hb + h+o)*wb + w+o)*cpitch + c_off + c] (bf16 elements), hb = h_+(pad?2:0),
 * wb = w_+(pad?2:0), o = (pad==1) */
typedef struct {
  void* ptr;   /* base of the whole buffer (16-byte aligned) */
  int n, h, w; /* logical (un-haloed) extent */
  int pad;     /* 0 = dense; 1 = one-pixel reflect halo stored around every image (buffer [n][h+2][w+2], interior at
                * (1,1)); 2 = zero tail: buffer [n][h+2][w+2], interior at (0,0), the 2 extra columns / rows stay zero
                * (gradient buffers read by the flat dgrad / wgrad kernels) */
  int cpitch;  /* channels per pixel in memory, multiple of 8 */
  int c_off;   /* first channel of the view */
  int c;       /* channels of the view */
} mimo_act_t;

int mimo_version(void);
const char* mimo_last_error(void);
/* 0 if the current device is sm_100 (B200), MIMO_ERR_ARCH otherwise */
int mimo_check_device(void);

/* ------------------------------------------------------------------ layout ------------------------------- */
/* fp32 NCHW image (element (b,c,h,w) at x[b*sb + c*sc + h*W + w]) -> bf16 NHWC view with reflect halo.
 * gather (device int64[n], may be NULL) remaps the batch index: folds apply_input_transform's
 * index_select+stack (mimo/models/utils.py:38-41) into the load. */
int mimo_pack_input(const float* x, long long sb, long long sc, const long long* gather, mimo_act_t out, void* stream);
/* OIHW fp32 conv weight -> bf16 [9][cout][cin_pitch] (fprop) and flipped/transposed [9][cin][cout_pitch] (dgrad,
 * may be NULL). Pitches are multiples of 8. */
int mimo_weight_pack(const float* w_oihw, int cout, int cin, void* w_fprop, int cin_pitch, void* w_dgrad, int cout_pitch,
                     void* stream);

/* ------------------------------------------------------------------ convolution -------------------------- */
/* nn.Conv2d(k=3, padding=1, padding_mode="reflect") forward, bias-free (components.py:23,26) -- tcgen05/TMA
 * implicit GEMM. mode 0: `in` is a pad==1 view, output domain h x w. mode 1 (input gradient): `in` is the
 * pad==0 output gradient, w_packed the dgrad pack, output domain (h+2) x (w+2) (padded-domain gradient whose
 * halo is folded back by mimo_grad_gather). out: bf16 [n][oh][ow][out_cpitch].
 * stat_sum/stat_sq (may be NULL): fp32 [mimo_conv3x3_m_tiles()][out_cpitch] per-tile partial sums of the stored
 * outputs and their squares (training-mode BatchNorm statistics). bias (may be NULL) / relu: fused epilogue. */
int mimo_conv3x3_m_tiles(int n, int out_h, int out_w);
int mimo_conv3x3(mimo_act_t in, int mode, const void* w_packed, int cout, int cin_pitch, void* out, int out_cpitch,
                 float* stat_sum, float* stat_sq, const float* bias, int relu, void* stream);
/* weight gradient (cuDNN wgrad in the reference's autograd): dy pad==0 view, x pad==1 view of the conv input.
 * dw_packed: fp32 [9][cout][cin_pitch] scratch (zeroed inside); grad_oihw receives (or accumulates) the OIHW grad. */
int mimo_conv3x3_wgrad(mimo_act_t dy, mimo_act_t x, float* dw_packed, int cin_pitch, float* grad_oihw, int accumulate,
                       void* stream);

/* ------------------------------------------------------------------ BatchNorm / ReLU / pool / upsample ---- */
/* nn.BatchNorm2d training statistics (components.py:24,27): reduces the conv partials, writes scale/shift
 * (z = scale*y + shift), the saved mean / invstd, updates running_mean/var (momentum 0.1, unbiased var;
 * conv_bias is added to the running mean because the stored conv output is bias-free) and num_batches_tracked. */
int mimo_bn_finalize(const float* stat_sum, const float* stat_sq, int tiles, int cpitch, int c, double count,
                     const float* gamma, const float* beta, const float* conv_bias, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float momentum, float eps, float* scale,
                     float* shift, float* save_mean, float* save_invstd, void* stream);
/* eval-mode affine from the running statistics */
int mimo_bn_eval_affine(int c, const float* gamma, const float* beta, const float* conv_bias, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, float* save_mean,
                        float* save_invstd, void* stream);
/* BN apply + ReLU (+ Dropout2d keep-scale [n][c], may be NULL) (+ MaxPool2d(2) into `pool`, may be NULL), writing
 * the reflect halo of every output (components.py:24-29,48). y: raw conv output [n][h][w][y_cpitch]. */
int mimo_bn_relu_apply(const void* y, int y_cpitch, const float* scale, const float* shift, const float* drop,
                       mimo_act_t out, const mimo_act_t* pool, void* stream);
/* nn.MaxPool2d(2[, return_indices]) (components.py:48); idx_nchw (may be NULL): int64 [n][c][h/2][w/2], h*W+w */
int mimo_maxpool2x2(mimo_act_t in, mimo_act_t out, long long* idx_nchw, void* stream);
/* nn.Upsample(x2, bilinear, align_corners=True) + F.pad to the skip size (components.py:78,112-115) into `out` */
int mimo_upsample_bilinear2x(mimo_act_t in, mimo_act_t out, void* stream);
int mimo_upsample_bilinear2x_bwd(mimo_act_t g_out, mimo_act_t g_in, int accumulate, void* stream);
/* The whole `Up` front end of the decoders in one pass (components.py:110-119: up(x1), F.pad to the skip size, cat([x2, x1], 1)):
 * `out` (pad 1, c_off 0, c = skip.c + in.c) receives whole pixel lines: channels [0, skip.c) copied from the DENSE (pad 0) skip tensor,
 * [skip.c, skip.c + in.c) = bilinear x2 (align_corners) of `in`, zero outside the up-sampled area, pad channels of the pitch = 0, and
 * the reflect halo. MIMO_ERR_ARG when the geometry is not supported (out.cpitch > 256, unaligned views): callers then use
 * mimo_upsample_bilinear2x into the slice. */
int mimo_upsample_concat(mimo_act_t in, mimo_act_t skip, mimo_act_t out, void* stream);
/* nn.MaxUnpool2d(2) (components.py:87): `in` pooled map, idx_nchw int64 [n][c][h][w] (flat h*W+w of the output) */
int mimo_maxunpool2x2(mimo_act_t in, const long long* idx_nchw, mimo_act_t out, void* stream);
/* nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2) (components.py:96-98): w fp32 [cin][cout][2][2] */
int mimo_convtranspose2x2(mimo_act_t in, const float* w, const float* bias, mimo_act_t out, void* stream);
/* NHWC bf16 view -> contiguous fp32 NCHW (component-level API boundary) */
int mimo_unpack_nchw(mimo_act_t in, float* out, void* stream);
/* G = fold_reflect(dpad) [+ maxpool backward of gpool through act]; any of dpad / (gpool, act) may be NULL */
int mimo_grad_gather(const mimo_act_t* dpad, const mimo_act_t* gpool, const mimo_act_t* act, mimo_act_t g_out,
                     int accumulate, void* stream);
/* BN + ReLU (+dropout) backward: dy, dgamma, dbeta (dbias_conv == 0 in training). part: fp32 scratch of
 * mimo_bn_bwd_scratch_floats(c) floats; s1s2: fp32 [2][c]. dy: whole-buffer view (c_off 0), pad 0 (dense) or
 * pad 2 (zero tail, the layout the flat dgrad / wgrad kernels read); only interior pixels are written. */
size_t mimo_bn_bwd_scratch_floats(int c);
int mimo_bn_relu_bwd(mimo_act_t g, const void* y, int y_cpitch, const float* scale, const float* shift,
                     const float* save_mean, const float* save_invstd, const float* drop, int training, float* part,
                     float* s1s2, float* dgamma, float* dbeta, float* dbias, int accumulate, mimo_act_t dy, void* stream);

/* Same, with the upstream gradient given as the padded-domain gradient `dpad` [N][H+2][W+2] of the following reflect-padded
 * convolution (autograd of F.pad(mode="reflect") feeding conv2 of DoubleConv, components.py:23-27): G = fold_reflect(dpad) is
 * formed inside the two passes when the operands are dense; otherwise it is materialised in `g_scratch` first. */
int mimo_bn_relu_bwd_folded(mimo_act_t dpad, mimo_act_t g_scratch, const void* y, int y_cpitch, const float* scale,
                            const float* shift, const float* save_mean, const float* save_invstd, const float* drop,
                            int training, float* part, float* s1s2, float* dgamma, float* dbeta, float* dbias,
                            int accumulate, mimo_act_t dy, void* stream);

/* Host-only (no device work): the stream-K schedule of the weight-gradient kernel for a layer with `cout` x `cin` channels
 * over `positions` = N*(H+2)*(W+2) grid positions on `sms` SMs. segments == NULL: returns the grid size. Otherwise writes
 * up to max_segments entries {co tile, first ci chunk, ci chunks (1|2), kh, first k-block, end k-block} of CTA `cta` and
 * returns how many segments that CTA has. Every (co tile, ci chunk, kh, k-block) is covered exactly once over all CTAs. */
int mimo_wgrad_streamk_schedule(int cout, int cin, long long positions, int sms, int cta, int* segments, int max_segments);

/* nn.Dropout (element-wise; reference model.py:239 center_dropout, model.py:294 final_dropouts): a *= keep * scale in place on
 * the interior of the view; keep = dense bf16 0/1 [N][H][W][mask_cpitch], scale = 1/(1-p). Backward = same call on the gradient. */
int mimo_mask_mul(mimo_act_t a, const void* keep, int mask_cpitch, float scale, void* stream);

/* ------------------------------------------------------------------ heads / loss / aggregation ------------ */
/* OutConv 1x1 (components.py:123-129): out fp32 planes, element (n,k,h,w) at out[n*out_bstride + k*h*w + ...] */
int mimo_head1x1(mimo_act_t feat, const float* w, const float* bias, int k, float* out, long long out_bstride, void* stream);
size_t mimo_head1x1_bwd_scratch_floats(int k, int c);
int mimo_head1x1_bwd(mimo_act_t feat, const float* w, int k, const float* dout, long long out_bstride,
                     const float* grad_scale, mimo_act_t g_feat, float* part, float* dw, float* db, int accumulate,
                     void* stream);

/* LaplaceNLL.forward (mimo/losses.py:132-164) over [rows][cols] fp32 with per-tensor row strides.
 * out_elem (may be NULL): elementwise loss [rows*cols]; out_mean (may be NULL): scalar mean, needs `part`
 * (mimo_laplace_scratch_floats() floats). */
size_t mimo_laplace_scratch_floats(void);
int mimo_laplace_nll_fwd(const float* mu, long long mu_rs, const float* log_s, long long ls_rs, const float* y,
                         long long y_rs, const float* mask, long long m_rs, long long rows, long long cols, float eps_min,
                         float eps_max, float* out_elem, float* part, float* out_mean, void* stream);
/* gradients w.r.t. mu / log_s (contiguous outputs); upstream is elementwise [rows*cols] or a device scalar */
int mimo_laplace_nll_bwd(const float* mu, long long mu_rs, const float* log_s, long long ls_rs, const float* y,
                         long long y_rs, const float* mask, long long m_rs, long long rows, long long cols, float eps_min,
                         float eps_max, const float* upstream, int upstream_is_scalar, float upstream_scale, float* g_mu,
                         float* g_log_s, void* stream);

/* LossBuffer (mimo/models/mimo_components/loss_buffer.py:18-74) kept on the device */
size_t mimo_lossbuffer_bytes(int subnetworks, int buffer_size);
int mimo_lossbuffer_init(void* state, int subnetworks, int buffer_size, float temperature, void* stream);
int mimo_lossbuffer_get_weights(const void* state, float* weights, void* stream);
int mimo_lossbuffer_add(void* state, const float* loss, void* stream);

/* _calculate_train_loss + loss.backward seed (mimo/models/mimo_unet.py:223-247,138) in one pass:
 * out fp32 [B][S][2C][H][W]; y element (b,s,j) at y[src_b*y_bs + s*y_ss + j], src_b = gather[s*B+b] if gather;
 * weights come from lb_state (read before the new loss is added), from fixed_w, or are 1.
 * Writes loss[S], weights[S], weighted[1] = mean_s(w_s*loss_s) and (if dout) d weighted / d out. */
size_t mimo_laplace_train_scratch_floats(int batch, int subnetworks, int c, long long hw);
int mimo_laplace_nll_train(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask,
                           long long m_bs, long long m_ss, const long long* gather, int batch, int subnetworks, int c,
                           long long hw, float eps_min, float eps_max, void* lb_state, const float* fixed_w,
                           int update_buffer, float* dout, float* part, float* loss, float* weights, float* weighted,
                           void* stream);
/* Same pass, plus the per-step regression metrics the reference computes with four torchmetrics reductions
 * (mimo/metrics.py:22-34, called from mimo/models/mimo_unet.py:135,172): metrics[4] = {mae, mse, rmse, r2} of the
 * location channels (mu vs y) over all B*S*C*H*W elements; the mask does not enter (as in the reference).
 * `part` must hold mimo_laplace_train_metrics_scratch_floats() floats. */
size_t mimo_laplace_train_metrics_scratch_floats(int batch, int subnetworks, int c, long long hw);
int mimo_laplace_nll_train_metrics(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask,
                                   long long m_bs, long long m_ss, const long long* gather, int batch, int subnetworks, int c,
                                   long long hw, float eps_min, float eps_max, void* lb_state, const float* fixed_w,
                                   int update_buffer, float* dout, float* part, float* loss, float* weights, float* weighted,
                                   float* metrics, void* stream);
/* GaussianNLL (mimo/losses.py:39-79; selected by UncertaintyLoss.from_name("gaussian_nll"), mimo/losses.py:29-36):
 * log(v) + (mu - y)^2 / v with v = clamp(exp(log_var), eps_min, eps_max), same clamp-gradient semantics and the same
 * argument meaning as the mimo_laplace_nll_* entry points. */
int mimo_gaussian_nll_fwd(const float* mu, long long mu_rs, const float* log_var, long long lv_rs, const float* y, long long y_rs,
                          const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                          float* out_elem, float* part, float* out_mean, void* stream);
int mimo_gaussian_nll_bwd(const float* mu, long long mu_rs, const float* log_var, long long lv_rs, const float* y, long long y_rs,
                          const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                          const float* upstream, int upstream_is_scalar, float upstream_scale, float* g_mu, float* g_log_var,
                          void* stream);
int mimo_gaussian_nll_train_metrics(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask,
                                    long long m_bs, long long m_ss, const long long* gather, int batch, int subnetworks, int c,
                                    long long hw, float eps_min, float eps_max, void* lb_state, const float* fixed_w,
                                    int update_buffer, float* dout, float* part, float* loss, float* weights, float* weighted,
                                    float* metrics, void* stream);
/* Deep evidential regression baseline (M = 1, out_channels = 4): the softplus head of EvidentialUnetModel.forward
 * (mimo/models/evidential_unet.py:85-96: raw (mu, log v, log alpha, log beta) -> (mu, v, alpha, beta)) and EvidentialLoss
 * (mimo/losses.py:195-271), each one elementwise pass forward and backward. raw/out/params/g_*: fp32 [batch][4][hw];
 * y, mask, out_elem: fp32 [batch][hw]; upstream as in mimo_laplace_nll_bwd; part: mimo_laplace_scratch_floats() floats. */
int mimo_evidential_head(const float* raw, float* out, long long batch, long long hw, void* stream);
int mimo_evidential_head_bwd(const float* raw, const float* g_out, float* g_raw, long long batch, long long hw, void* stream);
int mimo_evidential_loss_fwd(const float* params, const float* y, const float* mask, long long batch, long long hw,
                             float* out_elem, float* part, float* out_mean, void* stream);
int mimo_evidential_loss_bwd(const float* params, const float* y, const float* mask, long long batch, long long hw,
                             const float* upstream, int upstream_is_scalar, float upstream_scale, float* g_params, void* stream);
int mimo_scale_by_scalar(float* x, long long n, const float* scalar, void* stream);

/* compute_uncertainties (mimo/models/utils.py:76-101): element (b,s,j) at p[b*bs + s*ss + j], j < inner */
int mimo_ensemble_aggregate(const float* p1, long long p1_bs, long long p1_ss, const float* p2, long long p2_bs,
                            long long p2_ss, int batch, int members, long long inner, float* mean, float* aleatoric_var,
                            float* epistemic_var, void* stream);

/* MimoUnetModel.validation_step math (mimo/models/mimo_unet.py:146-183 with LaplaceNLL, losses.py:132-192) in ONE pass over
 * (p1, p2, label): element (b, s, j) of p1 / p2 at p[b*bs + s*ss + j], j < inner = C*H*W; label / mask (mask may be NULL): the
 * UN-repeated [batch][inner] tensors (repeat_subnetworks shows every subnetwork the same label). Writes the four [batch][inner] maps
 * (ensemble mean, aleatoric std, epistemic std, mean - label) and `scalars`: val_loss[members], val_loss_combined, mae, mse, rmse, r2
 * of (mean, label) (mimo/metrics.py:22-34), mean clip(aleatoric_std, 0, 5), mean clip(epistemic_std, 0, 5)
 * (members + 7 floats). scratch: mimo_validation_scratch_floats(members) floats. members <= 16. */
int mimo_validation_scratch_floats(int members);
int mimo_validation_laplace(const float* p1, const float* p2, long long bs, long long ss, const float* label, const float* mask,
                            int batch, int members, long long inner, float eps_min, float eps_max, float* mean, float* aleatoric_std,
                            float* epistemic_std, float* err, float* scratch, float* scalars, void* stream);

/* ------------------------------------------------------------------ whole-network executor ---------------- */
/* MimoUNet.forward / autograd backward (mimo/models/mimo_components/model.py:94-117), bilinear path. */
typedef struct {
  int in_channels, out_channels, num_subnetworks, filter_base_count;
  int batch, height, width;
} mimo_unet_config_t;
typedef struct mimo_unet_plan mimo_unet_plan_t;

int mimo_unet_plan_create(const mimo_unet_config_t* cfg, mimo_unet_plan_t** plan);
void mimo_unet_plan_destroy(mimo_unet_plan_t* plan);
size_t mimo_unet_workspace_bytes(const mimo_unet_plan_t* plan);
/* number of state_dict entries expected by mimo_unet_bind, in MimoUNet.state_dict() order */
int mimo_unet_num_state(const mimo_unet_plan_t* plan);
int mimo_unet_num_double_convs(const mimo_unet_plan_t* plan);
/* channels of the Dropout2d mask of double conv i (masks are [batch][channels] fp32 keep-scales) */
int mimo_unet_dropout_channels(const mimo_unet_plan_t* plan, int i);
/* state[i]: device pointer of state_dict entry i; grads[i]: fp32 gradient destination or NULL */
int mimo_unet_bind(mimo_unet_plan_t* plan, void* workspace, size_t workspace_bytes, void* const* state, void* const* grads,
                   int n);
/* x fp32 [B][S][Cin][H][W] when gather == NULL; with gather (device int64 [S][B]) x is the un-shuffled batch
 * [B][Cin][H][W] and subnetwork s reads image gather[s*B + b] (apply_input_transform, mimo/models/utils.py:27-41).
 * drop_masks[i]: device fp32 [B][channels] keep-scales of double conv i or NULL (array may be NULL altogether);
 * out fp32 [B][S][Cout][H][W]. training: batch statistics (+ running-stat update) vs running statistics. */
int mimo_unet_forward(mimo_unet_plan_t* plan, const float* x, const long long* gather, int training,
                      const float* const* drop_masks, float* out, void* stream);
/* dout fp32 like out; grad_scale: device scalar multiplier or NULL; dx: fp32 like x or NULL (input gradient,
 * not supported together with gather). Parameter gradients are WRITTEN (accumulate==0) or ADDED to grads[]. */
int mimo_unet_backward(mimo_unet_plan_t* plan, const float* dout, const float* grad_scale, float* dx, int accumulate,
                       void* stream);
/* Element-wise dropout for the NEXT forward/backward pair (reference model.py:239,294): keep masks as for mimo_mask_mul,
 * center_keep [N][H/16][W/16][round_up(8*f*S, 8)] applied to the core centre, final_keep[s] [N][H][W][round_up(f, 8)] applied
 * to the decoder features in front of head s; NULL entries = no dropout. Masks stay caller-owned until backward has run. */
int mimo_unet_set_elementwise_dropout(mimo_unet_plan_t* plan, const void* center_keep, float center_scale,
                                      const void* const* final_keep, float final_scale);
/* Inference mode of the following forward calls with training == 0 (EnsembleModule.forward, mimo/models/ensemble.py:76-115;
 * validation_step, mimo/models/mimo_unet.py:146-183, under no_grad): with on != 0 the eval-mode BatchNorm affine, the ReLU and
 * the Dropout2d factors of every DoubleConv (mimo/models/mimo_components/components.py:23-29) are applied in the convolution
 * epilogues, which write straight into the consumer's haloed buffer; no raw convolution output is kept, so
 * mimo_unet_backward is refused after such a forward. Off by default (eval-mode forward + backward, e.g. FGSM in
 * scripts/test/test_nyuv2_depth.py:26-90, keeps working). */
int mimo_unet_set_inference_fusion(mimo_unet_plan_t* plan, int on);
/* Overlap hook for the data-parallel gradient all-reduce: events[4] are caller-owned cudaEvent_t (or NULL
 * to disable). mimo_unet_backward records events[k] on its stream as soon as the parameter gradients of stage k are
 * final: 0 = decoders + heads, 1 = core up path, 2 = core down path, 3 = encoders (end of backward). The state entries
 * of stage k are the contiguous range [mimo_unet_backward_stage_first_state(plan, k), first state of stage k-1) (stage 0
 * runs to the end), i.e. stages are contiguous slices of a flat gradient buffer laid out in state order. */
int mimo_unet_set_backward_events(mimo_unet_plan_t* plan, void* const* events);
int mimo_unet_backward_stage_first_state(const mimo_unet_plan_t* plan, int stage);
/* test hook: geometry of a named intermediate ("<node>.<buf>", e.g. "core.down2.c1.y") inside the workspace.
 * kind: 0 = bf16 activation view, 1 = fp32 vector of view->c floats. */
int mimo_unet_debug_view(const mimo_unet_plan_t* plan, const char* name, mimo_act_t* view, int* kind);
/* Channel layout of the subnetwork stack cat(x2_0, ..., x2_{S-1}) (model.py:113) at the head of the core's input and of core.up3's
 * concat buffer: n_slices slices of slice_len channels, slice_stride apart (8-channel aligned; gap channels are zero). The views
 * "core.down2.in", "core.up3.in" and their ".c1.dpad" gradients of mimo_unet_debug_view are in this physical order.
 * n_slices == 0: no gaps. */
int mimo_unet_stack_layout(const mimo_unet_plan_t* plan, int* slice_len, int* slice_stride, int* n_slices);
/* number of kernels the last forward / backward enqueued (bench.py's gpu_launches) */
int mimo_unet_last_launches(const mimo_unet_plan_t* plan);
/* CUDA graphs of the executor's fixed launch sequences (built after two eager calls per plan; env MIMO_GRAPH=0 disables):
 * bit 0 forward body, bits 1..4 the four backward stages, bit 8 = a capture failed and the plan stays eager. */
int mimo_unet_graph_state(const mimo_unet_plan_t* plan);
/* optional per-launch CUDA-event timing by kernel class (records events on the caller's stream; read synchronises) */
int mimo_unet_profile_classes(void);
const char* mimo_unet_profile_class_name(int i);
int mimo_unet_profile_enable(mimo_unet_plan_t* plan, int on);
int mimo_unet_profile_read(mimo_unet_plan_t* plan, float* ms_by_class, int* count_by_class);
/* per-launch variant: fills up to max_n entries (ms, class, tag = 2*double_conv_index + second_conv or -1) in
 * enqueue order and returns how many; resets the recording like mimo_unet_profile_read. */
int mimo_unet_profile_read_launches(mimo_unet_plan_t* plan, int max_n, float* ms, int* cls, int* tag);
/* same, plus the tensor-core kernel each conv launch dispatched to (0 = not a conv launch); names via mimo_conv_kernel_name */
int mimo_unet_profile_read_launches_ex(mimo_unet_plan_t* plan, int max_n, float* ms, int* cls, int* tag, int* kernel);
const char* mimo_conv_kernel_name(int id);
/* state_dict prefix of double conv i ("core.down2", "decoder.up4s.1", ...) */
const char* mimo_unet_node_name(const mimo_unet_plan_t* plan, int i);

#ifdef __cplusplus
}
#endif
#endif /* MIMO_B200_H */
