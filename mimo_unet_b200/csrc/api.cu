// extern "C" surface of libmimo_b200.so for the individual operators (the whole-network executor's entry
// points live in engine.cu). Thin argument adapters over ops.h -- see include/mimo_b200.h for the contract.
#include "common.cuh"
#include "ops.h"

using namespace mimo;

extern "C" {

int mimo_version(void) { return MIMO_B200_VERSION; }
const char* mimo_last_error(void) { return last_error(); }

int mimo_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  MIMO_CUDA(cudaGetDevice(&dev));
  MIMO_CUDA(cudaGetDeviceProperties(&prop, dev));
  MIMO_CHECK(prop.major == 10 && prop.minor == 0, MIMO_ERR_ARCH, "device %s is sm_%d%d; this library is built for sm_100a only",
             prop.name, prop.major, prop.minor);
  return MIMO_OK;
}

int mimo_pack_input(const float* x, long long sb, long long sc, const long long* gather, mimo_act_t out, void* stream) {
  MIMO_CHECK(x && out.ptr, MIMO_ERR_ARG, "pack_input: null pointer");
  return pack_input_launch(x, sb, sc, gather, make_view(out), (cudaStream_t)stream);
}

int mimo_weight_pack(const float* w, int cout, int cin, void* wf, int cin_pitch, void* wd, int cout_pitch, void* stream) {
  MIMO_CHECK(w && wf, MIMO_ERR_ARG, "weight_pack: null pointer");
  MIMO_CHECK(cin_pitch % 8 == 0 && cin_pitch >= cin && (!wd || (cout_pitch % 8 == 0 && cout_pitch >= cout)), MIMO_ERR_ALIGN,
             "weight_pack: pitches must be multiples of 8 and cover the channel counts");
  return weight_pack_launch(w, cout, cin, (bf16*)wf, cin_pitch, (bf16*)wd, cout_pitch, (cudaStream_t)stream);
}

int mimo_conv3x3_m_tiles(int n, int out_h, int out_w) { (void)n; (void)out_h; (void)out_w; return conv3x3_stat_rows(); }

int mimo_conv3x3(mimo_act_t in, int mode, const void* w_packed, int cout, int cin_pitch, void* out, int out_cpitch, float* stat_sum,
                 float* stat_sq, const float* bias, int relu, void* stream) {
  MIMO_CHECK(in.ptr && w_packed && out, MIMO_ERR_ARG, "conv3x3: null pointer");
  return conv3x3_launch(make_view(in), mode, (const bf16*)w_packed, cout, cin_pitch, (bf16*)out, out_cpitch, stat_sum, stat_sq, bias,
                        relu, (cudaStream_t)stream);
}

int mimo_conv3x3_wgrad(mimo_act_t dy, mimo_act_t x, float* dw_packed, int cin_pitch, float* grad_oihw, int accumulate, void* stream) {
  MIMO_CHECK(dy.ptr && x.ptr && dw_packed && grad_oihw, MIMO_ERR_ARG, "wgrad: null pointer");
  int rc = conv3x3_wgrad_launch(make_view(dy), make_view(x), dw_packed, cin_pitch, (cudaStream_t)stream);
  if (rc) return rc;
  return wgrad_unpack_launch(dw_packed, grad_oihw, dy.c, x.c, cin_pitch, 1.f, accumulate, (cudaStream_t)stream);
}

int mimo_bn_finalize(const float* stat_sum, const float* stat_sq, int tiles, int cpitch, int c, double count, const float* gamma,
                     const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                     long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift, float* save_mean,
                     float* save_invstd, void* stream) {
  MIMO_CHECK(stat_sum && stat_sq && gamma && beta && scale && shift && save_mean && save_invstd, MIMO_ERR_ARG, "bn_finalize: null pointer");
  return bn_finalize_launch(stat_sum, stat_sq, tiles, cpitch, c, count, gamma, beta, conv_bias, running_mean, running_var,
                            num_batches_tracked, momentum, eps, scale, shift, save_mean, save_invstd, (cudaStream_t)stream);
}

int mimo_bn_eval_affine(int c, const float* gamma, const float* beta, const float* conv_bias, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                        void* stream) {
  return bn_eval_affine_launch(c, gamma, beta, conv_bias, running_mean, running_var, eps, scale, shift, save_mean, save_invstd,
                               (cudaStream_t)stream);
}

int mimo_bn_relu_apply(const void* y, int y_cpitch, const float* scale, const float* shift, const float* drop, mimo_act_t out,
                       const mimo_act_t* pool, void* stream) {
  MIMO_CHECK(y && scale && shift && out.ptr, MIMO_ERR_ARG, "bn_relu_apply: null pointer");
  if (pool) {
    const ActView pv = make_view(*pool);
    return bn_relu_apply_launch((const bf16*)y, y_cpitch, scale, shift, drop, make_view(out), &pv, (cudaStream_t)stream);
  }
  return bn_relu_apply_launch((const bf16*)y, y_cpitch, scale, shift, drop, make_view(out), nullptr, (cudaStream_t)stream);
}

int mimo_maxpool2x2(mimo_act_t in, mimo_act_t out, long long* idx_nchw, void* stream) {
  MIMO_CHECK(in.ptr && out.ptr, MIMO_ERR_ARG, "maxpool: null pointer");
  return maxpool_launch(make_view(in), make_view(out), idx_nchw, (cudaStream_t)stream);
}

int mimo_upsample_bilinear2x(mimo_act_t in, mimo_act_t out, void* stream) {
  MIMO_CHECK(in.ptr && out.ptr, MIMO_ERR_ARG, "upsample: null pointer");
  return upsample_launch(make_view(in), make_view(out), (cudaStream_t)stream);
}

int mimo_upsample_concat(mimo_act_t in, mimo_act_t skip, mimo_act_t out, void* stream) {
  MIMO_CHECK(in.ptr && skip.ptr && out.ptr, MIMO_ERR_ARG, "upsample_concat: null pointer");
  return upsample_concat_launch(make_view(in), make_view(skip), make_view(out), (cudaStream_t)stream);
}

int mimo_maxunpool2x2(mimo_act_t in, const long long* idx_nchw, mimo_act_t out, void* stream) {
  MIMO_CHECK(in.ptr && idx_nchw && out.ptr, MIMO_ERR_ARG, "maxunpool: null pointer");
  return maxunpool_launch(make_view(in), idx_nchw, make_view(out), (cudaStream_t)stream);
}

int mimo_convtranspose2x2(mimo_act_t in, const float* w, const float* bias, mimo_act_t out, void* stream) {
  MIMO_CHECK(in.ptr && w && out.ptr, MIMO_ERR_ARG, "convtranspose2x2: null pointer");
  return convtranspose2x2_launch(make_view(in), w, bias, make_view(out), (cudaStream_t)stream);
}

int mimo_unpack_nchw(mimo_act_t in, float* out, void* stream) {
  MIMO_CHECK(in.ptr && out, MIMO_ERR_ARG, "unpack_nchw: null pointer");
  return unpack_nchw_launch(make_view(in), out, (cudaStream_t)stream);
}

int mimo_upsample_bilinear2x_bwd(mimo_act_t g_out, mimo_act_t g_in, int accumulate, void* stream) {
  MIMO_CHECK(g_out.ptr && g_in.ptr, MIMO_ERR_ARG, "upsample_bwd: null pointer");
  return upsample_bwd_launch(make_view(g_out), make_view(g_in), accumulate, (cudaStream_t)stream);
}

int mimo_grad_gather(const mimo_act_t* dpad, const mimo_act_t* gpool, const mimo_act_t* act, mimo_act_t g_out, int accumulate,
                     void* stream) {
  MIMO_CHECK(g_out.ptr && (dpad || gpool), MIMO_ERR_ARG, "grad_gather: nothing to gather");
  ActView a, b, c;
  if (dpad) a = make_view(*dpad);
  if (gpool) b = make_view(*gpool);
  if (act) c = make_view(*act);
  return grad_gather_launch(dpad ? &a : nullptr, gpool ? &b : nullptr, act ? &c : nullptr, make_view(g_out), accumulate,
                            (cudaStream_t)stream);
}

size_t mimo_bn_bwd_scratch_floats(int c) { return (size_t)bn_bwd_parts(c) * 2 * c; }

int mimo_bn_relu_bwd(mimo_act_t g, const void* y, int y_cpitch, const float* scale, const float* shift, const float* save_mean,
                     const float* save_invstd, const float* drop, int training, float* part, float* s1s2, float* dgamma, float* dbeta,
                     float* dbias, int accumulate, mimo_act_t dy, void* stream) {
  MIMO_CHECK(g.ptr && y && scale && shift && save_mean && save_invstd && part && s1s2 && dy.ptr, MIMO_ERR_ARG, "bn_relu_bwd: null pointer");
  return bn_bwd_launch(make_view(g), (const bf16*)y, y_cpitch, scale, shift, save_mean, save_invstd, drop, training, part, s1s2, dgamma,
                       dbeta, dbias, 1.f, accumulate, make_view(dy), (cudaStream_t)stream);
}

int mimo_bn_relu_bwd_folded(mimo_act_t dpad, mimo_act_t g_scratch, const void* y, int y_cpitch, const float* scale, const float* shift,
                            const float* save_mean, const float* save_invstd, const float* drop, int training, float* part, float* s1s2,
                            float* dgamma, float* dbeta, float* dbias, int accumulate, mimo_act_t dy, void* stream) {
  MIMO_CHECK(dpad.ptr && g_scratch.ptr && y && scale && shift && save_mean && save_invstd && part && s1s2 && dy.ptr, MIMO_ERR_ARG,
             "bn_relu_bwd_folded: null pointer");
  const ActView d = make_view(dpad);
  return bn_bwd_launch(make_view(g_scratch), (const bf16*)y, y_cpitch, scale, shift, save_mean, save_invstd, drop, training, part, s1s2,
                       dgamma, dbeta, dbias, 1.f, accumulate, make_view(dy), (cudaStream_t)stream, &d);
}

int mimo_wgrad_streamk_schedule(int cout, int cin, long long positions, int sms, int cta, int* segments, int max_segments) {
  return conv3x3_wgrad_flatk_schedule(cout, cin, positions, sms, cta, segments, max_segments);
}

int mimo_mask_mul(mimo_act_t a, const void* keep, int mask_cpitch, float scale, void* stream) {
  MIMO_CHECK(a.ptr && keep, MIMO_ERR_ARG, "mask_mul: null pointer");
  return mask_mul_launch(make_view(a), (const bf16*)keep, mask_cpitch, scale, (cudaStream_t)stream);
}

int mimo_head1x1(mimo_act_t feat, const float* w, const float* bias, int k, float* out, long long out_bstride, void* stream) {
  MIMO_CHECK(feat.ptr && w && bias && out, MIMO_ERR_ARG, "head1x1: null pointer");
  return head_fwd_launch(make_view(feat), w, bias, k, out, out_bstride, (cudaStream_t)stream);
}

size_t mimo_head1x1_bwd_scratch_floats(int k, int c) { return (size_t)head_bwd_parts() * (k * c + k); }

int mimo_head1x1_bwd(mimo_act_t feat, const float* w, int k, const float* dout, long long out_bstride, const float* grad_scale,
                     mimo_act_t g_feat, float* part, float* dw, float* db, int accumulate, void* stream) {
  MIMO_CHECK(feat.ptr && w && dout && g_feat.ptr && part && dw && db, MIMO_ERR_ARG, "head1x1_bwd: null pointer");
  return head_bwd_launch(make_view(feat), w, k, dout, out_bstride, grad_scale, make_view(g_feat), part, dw, db, accumulate,
                         (cudaStream_t)stream);
}

size_t mimo_laplace_scratch_floats(void) { return (size_t)laplace_parts(); }

int mimo_laplace_nll_fwd(const float* mu, long long mu_rs, const float* log_s, long long ls_rs, const float* y, long long y_rs,
                         const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                         float* out_elem, float* part, float* out_mean, void* stream) {
  MIMO_CHECK(mu && log_s && y && (out_elem || out_mean), MIMO_ERR_ARG, "laplace_nll_fwd: null pointer");
  if (rows * cols == 0) return MIMO_OK;
  return laplace_fwd_launch(mu, mu_rs, log_s, ls_rs, y, y_rs, mask, m_rs, rows, cols, eps_min, eps_max, out_elem, part, out_mean,
                            (cudaStream_t)stream);
}

int mimo_laplace_nll_bwd(const float* mu, long long mu_rs, const float* log_s, long long ls_rs, const float* y, long long y_rs,
                         const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                         const float* upstream, int upstream_is_scalar, float upstream_scale, float* g_mu, float* g_log_s,
                         void* stream) {
  MIMO_CHECK(mu && log_s && y && upstream && g_mu && g_log_s, MIMO_ERR_ARG, "laplace_nll_bwd: null pointer");
  if (rows * cols == 0) return MIMO_OK;
  return laplace_bwd_launch(mu, mu_rs, log_s, ls_rs, y, y_rs, mask, m_rs, rows, cols, eps_min, eps_max, upstream, upstream_is_scalar,
                            upstream_scale, g_mu, g_log_s, (cudaStream_t)stream);
}

int mimo_gaussian_nll_fwd(const float* mu, long long mu_rs, const float* log_var, long long lv_rs, const float* y, long long y_rs,
                          const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                          float* out_elem, float* part, float* out_mean, void* stream) {
  MIMO_CHECK(mu && log_var && y && (out_elem || out_mean), MIMO_ERR_ARG, "gaussian_nll_fwd: null pointer");
  if (rows * cols == 0) return MIMO_OK;
  return laplace_fwd_launch(mu, mu_rs, log_var, lv_rs, y, y_rs, mask, m_rs, rows, cols, eps_min, eps_max, out_elem, part, out_mean,
                            (cudaStream_t)stream, 1);
}
int mimo_gaussian_nll_bwd(const float* mu, long long mu_rs, const float* log_var, long long lv_rs, const float* y, long long y_rs,
                          const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                          const float* upstream, int upstream_is_scalar, float upstream_scale, float* g_mu, float* g_log_var,
                          void* stream) {
  MIMO_CHECK(mu && log_var && y && upstream && g_mu && g_log_var, MIMO_ERR_ARG, "gaussian_nll_bwd: null pointer");
  if (rows * cols == 0) return MIMO_OK;
  return laplace_bwd_launch(mu, mu_rs, log_var, lv_rs, y, y_rs, mask, m_rs, rows, cols, eps_min, eps_max, upstream, upstream_is_scalar,
                            upstream_scale, g_mu, g_log_var, (cudaStream_t)stream, 1);
}

int mimo_evidential_head(const float* raw, float* out, long long batch, long long hw, void* stream) {
  MIMO_CHECK(raw && out, MIMO_ERR_ARG, "evidential_head: null pointer");
  if (batch * hw == 0) return MIMO_OK;
  return evidential_head_launch(raw, out, batch, hw, (cudaStream_t)stream);
}
int mimo_evidential_head_bwd(const float* raw, const float* g_out, float* g_raw, long long batch, long long hw, void* stream) {
  MIMO_CHECK(raw && g_out && g_raw, MIMO_ERR_ARG, "evidential_head_bwd: null pointer");
  if (batch * hw == 0) return MIMO_OK;
  return evidential_head_bwd_launch(raw, g_out, g_raw, batch, hw, (cudaStream_t)stream);
}
int mimo_evidential_loss_fwd(const float* params, const float* y, const float* mask, long long batch, long long hw, float* out_elem,
                             float* part, float* out_mean, void* stream) {
  MIMO_CHECK(params && y && (out_elem || out_mean), MIMO_ERR_ARG, "evidential_loss_fwd: null pointer");
  if (batch * hw == 0) return MIMO_OK;
  return evidential_loss_fwd_launch(params, y, mask, batch, hw, out_elem, part, out_mean, (cudaStream_t)stream);
}
int mimo_evidential_loss_bwd(const float* params, const float* y, const float* mask, long long batch, long long hw,
                             const float* upstream, int upstream_is_scalar, float upstream_scale, float* g_params, void* stream) {
  MIMO_CHECK(params && y && upstream && g_params, MIMO_ERR_ARG, "evidential_loss_bwd: null pointer");
  if (batch * hw == 0) return MIMO_OK;
  return evidential_loss_bwd_launch(params, y, mask, batch, hw, upstream, upstream_is_scalar, upstream_scale, g_params, (cudaStream_t)stream);
}

size_t mimo_lossbuffer_bytes(int subnetworks, int buffer_size) { return lossbuffer_bytes(subnetworks, buffer_size); }
int mimo_lossbuffer_init(void* state, int subnetworks, int buffer_size, float temperature, void* stream) {
  MIMO_CHECK(state, MIMO_ERR_ARG, "lossbuffer_init: null pointer");
  return lossbuffer_init_launch(state, subnetworks, buffer_size, temperature, (cudaStream_t)stream);
}
int mimo_lossbuffer_get_weights(const void* state, float* weights, void* stream) {
  MIMO_CHECK(state && weights, MIMO_ERR_ARG, "lossbuffer_get_weights: null pointer");
  return lossbuffer_weights_launch(state, weights, (cudaStream_t)stream);
}
int mimo_lossbuffer_add(void* state, const float* loss, void* stream) {
  MIMO_CHECK(state && loss, MIMO_ERR_ARG, "lossbuffer_add: null pointer");
  return lossbuffer_add_launch(state, loss, (cudaStream_t)stream);
}

size_t mimo_laplace_train_scratch_floats(int batch, int subnetworks, int c, long long hw) {
  return (size_t)batch * subnetworks * laplace_train_blocks((long long)c * hw);
}

int mimo_laplace_nll_train(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask, long long m_bs,
                           long long m_ss, const long long* gather, int batch, int subnetworks, int c, long long hw, float eps_min,
                           float eps_max, void* lb_state, const float* fixed_w, int update_buffer, float* dout, float* part,
                           float* loss, float* weights, float* weighted, void* stream) {
  MIMO_CHECK(out && y && part && loss, MIMO_ERR_ARG, "laplace_nll_train: null pointer");
  return laplace_train_launch(out, y, y_bs, y_ss, mask, m_bs, m_ss, gather, batch, subnetworks, c, hw, eps_min, eps_max, lb_state,
                              fixed_w, update_buffer, dout, part, loss, weights, weighted, (cudaStream_t)stream);
}

int mimo_laplace_nll_train_metrics(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask, long long m_bs,
                                   long long m_ss, const long long* gather, int batch, int subnetworks, int c, long long hw,
                                   float eps_min, float eps_max, void* lb_state, const float* fixed_w, int update_buffer, float* dout,
                                   float* part, float* loss, float* weights, float* weighted, float* metrics, void* stream) {
  MIMO_CHECK(out && y && part && loss && metrics, MIMO_ERR_ARG, "laplace_nll_train_metrics: null pointer");
  // scratch: [B*S*nblk] loss partials followed by [B*S*nblk][4] metric partials (mimo_laplace_train_metrics_scratch_floats)
  float* mpart = part + mimo_laplace_train_scratch_floats(batch, subnetworks, c, hw);
  return laplace_train_launch(out, y, y_bs, y_ss, mask, m_bs, m_ss, gather, batch, subnetworks, c, hw, eps_min, eps_max, lb_state,
                              fixed_w, update_buffer, dout, part, loss, weights, weighted, (cudaStream_t)stream, mpart, metrics);
}
size_t mimo_laplace_train_metrics_scratch_floats(int batch, int subnetworks, int c, long long hw) {
  return 5 * mimo_laplace_train_scratch_floats(batch, subnetworks, c, hw);
}
int mimo_gaussian_nll_train_metrics(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask, long long m_bs,
                                    long long m_ss, const long long* gather, int batch, int subnetworks, int c, long long hw,
                                    float eps_min, float eps_max, void* lb_state, const float* fixed_w, int update_buffer, float* dout,
                                    float* part, float* loss, float* weights, float* weighted, float* metrics, void* stream) {
  MIMO_CHECK(out && y && part && loss && metrics, MIMO_ERR_ARG, "gaussian_nll_train_metrics: null pointer");
  float* mpart = part + mimo_laplace_train_scratch_floats(batch, subnetworks, c, hw);
  return laplace_train_launch(out, y, y_bs, y_ss, mask, m_bs, m_ss, gather, batch, subnetworks, c, hw, eps_min, eps_max, lb_state,
                              fixed_w, update_buffer, dout, part, loss, weights, weighted, (cudaStream_t)stream, mpart, metrics, 1);
}

int mimo_scale_by_scalar(float* x, long long n, const float* scalar, void* stream) {
  MIMO_CHECK(x && scalar, MIMO_ERR_ARG, "scale_by_scalar: null pointer");
  return scale_by_scalar_launch(x, n, scalar, (cudaStream_t)stream);
}

int mimo_ensemble_aggregate(const float* p1, long long p1_bs, long long p1_ss, const float* p2, long long p2_bs, long long p2_ss,
                            int batch, int members, long long inner, float* mean, float* aleatoric_var, float* epistemic_var,
                            void* stream) {
  MIMO_CHECK(p1 && p2 && mean && aleatoric_var && epistemic_var, MIMO_ERR_ARG, "ensemble_aggregate: null pointer");
  return aggregate_launch(p1, p1_bs, p1_ss, p2, p2_bs, p2_ss, batch, members, inner, mean, aleatoric_var, epistemic_var,
                          (cudaStream_t)stream);
}

int mimo_validation_scratch_floats(int members) { return validation_scratch_floats(members); }

int mimo_validation_laplace(const float* p1, const float* p2, long long bs, long long ss, const float* label, const float* mask,
                            int batch, int members, long long inner, float eps_min, float eps_max, float* mean, float* aleatoric_std,
                            float* epistemic_std, float* err, float* scratch, float* scalars, void* stream) {
  MIMO_CHECK(p1 && p2 && label && mean && aleatoric_std && epistemic_std && err && scratch && scalars, MIMO_ERR_ARG,
             "validation_laplace: null pointer");
  return validation_laplace_launch(p1, p2, bs, ss, label, mask, batch, members, inner, eps_min, eps_max, mean, aleatoric_std,
                                   epistemic_std, err, scratch, scalars, (cudaStream_t)stream);
}

}  // extern "C"
