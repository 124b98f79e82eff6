// Internal C++ launch API shared by the C-ABI wrappers (api.cu) and the U-Net executor (engine.cu).
// Every function enqueues work on `stream`, never synchronises, never allocates device memory.
#pragma once
#include "common.cuh"
#include "conv_epilogue.cuh"

namespace mimo {

// ---- tensor-core convolutions (conv_igemm.cu / conv_wgrad.cu) ----
int conv3x3_block_n(int cout);
int conv3x3_stat_rows();
bool conv3x3_flat_ok(const ActView& in, int mode, int cout);
int conv3x3_flat_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                        float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse = nullptr);
// fuse != nullptr (inference): out = interior/first channel of the CONSUMER's haloed buffer, out_cpitch its pixel pitch; the
// epilogue applies v * scale + shift, ReLU and the Dropout2d factors; only kernels with that epilogue are eligible
// (conv3x3_fuse_ok), the halo ring is filled afterwards by halo_fill_launch.
int conv3x3_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                   float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse = nullptr);
bool conv3x3_fuse_ok(const ActView& in, int cout);
bool conv3x3_flatk_ok(const ActView& in, int mode, int cout);
int conv3x3_flatk_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                         float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream);
bool conv3x3_flat2_ok(const ActView& in, int mode, int cout);
int conv3x3_flat2_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                         float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream);
bool conv3x3_c2_ok(const ActView& in, int mode, int cout);
int conv3x3_c2_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                      float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse = nullptr,
                      const ActView* in2 = nullptr);
// in2 != nullptr: VIRTUAL CONCAT -- the conv input is cat([in, in2], channels) without the concatenated tensor existing: the K
// loop reads the 64-channel chunks of `in` and then those of `in2` through separate tensor maps; wpacked must use the chunk-aligned
// channel order (WeightPackJob::sl = ChannelSlices{in.C, round_up(in.C, 64), 1}), cin_pitch accordingly.
// pre_zeroed: the caller has already cleared dw_packed on `stream` (the executor clears all layers with one memset)
// CUDA-core kernel of the first layer (<= 4 input channels, <= 32 output channels): conv_thin.cu
bool conv3x3_thin_ok(const ActView& in, int mode, int cout);
int conv3x3_thin_launch(const ActView& in, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch, float* stat_sum,
                        float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse = nullptr);
int conv3x3_wgrad_launch(const ActView& dy, const ActView& x, float* dw_packed, int cin_pitch, cudaStream_t stream, bool pre_zeroed = false);
bool conv3x3_wgrad_flat_ok(const ActView& dy, const ActView& x);
int conv3x3_wgrad_flat_launch(const ActView& dy, const ActView& x, float* dw_packed, int cin_pitch, cudaStream_t stream, bool pre_zeroed = false);
bool conv3x3_wgrad_flatk_ok(const ActView& dy, const ActView& x);
int conv3x3_wgrad_flatk_launch(const ActView& dy, const ActView& x, float* dw_packed, int cin_pitch, cudaStream_t stream, bool pre_zeroed = false);
// batched variants: one launch for many layers (the per-layer kernels are a few microseconds each; their launch gaps
// cost more than their work)
// Input-channel layout of a packed weight (ChannelSlices): the first sl_n * sl_len source channels are sl_n slices of sl_len channels
// that sit sl_stride apart in the packed / physical order (zeros in the gaps), the remaining channels follow at sl_n * sl_stride.
// sl_n == 0: identity. Two users: the chunk-aligned order of a virtual concat (one slice of in.C channels padded to a multiple of 64,
// conv3x3_c2_launch with a second input view) and the 8-channel-aligned subnetwork slices of the stacked buffers (S slices of 2f
// channels, stride round_up(2f, 8)).
struct ChannelSlices {
  int sl_len, sl_stride, sl_n;
  __host__ __device__ int phys_count(int cin) const { return sl_n == 0 ? cin : sl_n * sl_stride + (cin - sl_n * sl_len); }
  // physical position -> source channel, or -1 for a gap
  __host__ __device__ int source(int cp, int cin) const {
    if (sl_n == 0) return cp < cin ? cp : -1;
    if (cp < sl_n * sl_stride) { const int s = cp / sl_stride, r = cp - s * sl_stride; return r < sl_len ? s * sl_len + r : -1; }
    const int ci = sl_n * sl_len + (cp - sl_n * sl_stride);
    return ci < cin ? ci : -1;
  }
  // source channel -> physical position
  __host__ __device__ int position(int ci) const {
    if (sl_n == 0 || ci >= sl_n * sl_len) return sl_n == 0 ? ci : sl_n * sl_stride + (ci - sl_n * sl_len);
    const int s = ci / sl_len;
    return s * sl_stride + (ci - s * sl_len);
  }
};
// wf: [9][cout][cin_pitch] over PHYSICAL input positions; wd (may be null): [9][phys_count(cin)][cout_pitch], rows at physical positions
struct WeightPackJob { const float* w; bf16* wf; bf16* wd; int cout, cin, cin_pitch, cout_pitch; ChannelSlices sl; };
// packed: fp32 [9][cout][cin_pitch] over physical input positions -> grad OIHW [cout][cin][3][3]
struct WgradUnpackJob { const float* packed; float* grad; int cout, cin, cin_pitch; ChannelSlices sl; };
int weight_pack_batched_launch(const WeightPackJob* jobs, int n, cudaStream_t st);
int wgrad_unpack_batched_launch(const WgradUnpackJob* jobs, int n, float scale, int accumulate, cudaStream_t st);
int conv3x3_wgrad_flatk_schedule(int cout, int cin, long long total_pos, int sms, int cta, int* out, int max_segs);
int wgrad_unpack_launch(const float* packed, float* grad_oihw, int cout, int cin, int cin_pitch, float scale, int accumulate,
                        cudaStream_t stream);

// ---- NHWC elementwise / reduction kernels (elementwise.cu) ----
int pack_input_launch(const float* x, long long sb, long long sc, const long long* gather, const ActView& o, cudaStream_t st);
int weight_pack_launch(const float* w, int cout, int cin, bf16* wf, int cin_pitch, bf16* wd, int cout_pitch, cudaStream_t st);
int bn_finalize_launch(const float* psum, const float* psq, int tiles, int cpitch, int C, double count, const float* gamma,
                       const float* beta, const float* conv_bias, float* rm, float* rv, long long* nbt, float momentum, float eps,
                       float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st);
int bn_eval_affine_launch(int C, const float* gamma, const float* beta, const float* conv_bias, const float* rm, const float* rv,
                          float eps, float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st);
int bn_relu_apply_launch(const bf16* y, int ycp, const float* scale, const float* shift, const float* drop, const ActView& o,
                         const ActView* pool, cudaStream_t st);
int maxpool_launch(const ActView& in, const ActView& o, long long* idx_nchw, cudaStream_t st);
int upsample_launch(const ActView& in, const ActView& o, cudaStream_t st);
bool upsample_concat_ok(const ActView& in, const ActView& skip, const ActView& o);
int upsample_concat_launch(const ActView& in, const ActView& skip, const ActView& o, cudaStream_t st);
int upsample_bwd_launch(const ActView& gdst, const ActView& gsrc, int accumulate, cudaStream_t st);
int grad_gather_launch(const ActView* dpad, const ActView* gpool, const ActView* act, const ActView& gout, int accumulate,
                       cudaStream_t st, const ActView* dpad2 = nullptr);
int halo_fill_launch(const ActView& v, cudaStream_t st);
int mask_mul_launch(const ActView& a, const bf16* keep, int mask_cp, float scale, cudaStream_t st);
int maxunpool_launch(const ActView& in, const long long* idx_nchw, const ActView& o, cudaStream_t st);
int convtranspose2x2_launch(const ActView& in, const float* wt, const float* bias, const ActView& o, cudaStream_t st);
int unpack_nchw_launch(const ActView& in, float* out, cudaStream_t st);
int bn_bwd_parts(int C);
int bn_bwd_launch(const ActView& G, const bf16* y, int ycp, const float* scale, const float* shift, const float* mean,
                  const float* invstd, const float* drop, int training, float* part, float* s1s2, float* dgamma, float* dbeta,
                  float* dbias, float grad_scale, int accumulate, const ActView& dy, cudaStream_t st, const ActView* fold_src = nullptr);
// fold_src != nullptr: the upstream gradient is fold_reflect(*fold_src) (padded-domain gradient of the next conv); it is fused
// into the two passes when the operands are dense, otherwise G is filled by grad_gather first (G is scratch in both cases).

// ---- heads, loss, loss buffer, aggregation (head_loss.cu) ----
int head_fwd_launch(const ActView& f, const float* W, const float* bias, int K, float* out, long long out_bstride, cudaStream_t st);
int head_bwd_parts();
int head_bwd_launch(const ActView& f, const float* W, int K, const float* dout, long long out_bstride, const float* grad_scale,
                    const ActView& G, float* part, float* dW, float* db, int accumulate, cudaStream_t st);
int laplace_parts();
int laplace_fwd_launch(const float* mu, long long mu_rs, const float* ls, long long ls_rs, const float* y, long long y_rs,
                       const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                       float* out_elem, float* part, float* out_mean, cudaStream_t st, int kind = 0 /* 0 Laplace, 1 Gaussian */);
int laplace_bwd_launch(const float* mu, long long mu_rs, const float* ls, long long ls_rs, const float* y, long long y_rs,
                       const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                       const float* up, int up_is_scalar, float up_scale, float* g_mu, float* g_ls, cudaStream_t st, int kind = 0);
int evidential_head_launch(const float* raw, float* out, long long B, long long HW, cudaStream_t st);
int evidential_head_bwd_launch(const float* raw, const float* g_out, float* g_raw, long long B, long long HW, cudaStream_t st);
int evidential_loss_fwd_launch(const float* par, const float* y, const float* mask, long long B, long long HW, float* out_elem,
                               float* part, float* out_mean, cudaStream_t st);
int evidential_loss_bwd_launch(const float* par, const float* y, const float* mask, long long B, long long HW, const float* up,
                               int up_is_scalar, float up_scale, float* g_par, cudaStream_t st);
size_t lossbuffer_bytes(int S, int size);
int lossbuffer_init_launch(void* state, int S, int size, float T, cudaStream_t st);
int lossbuffer_weights_launch(const void* state, float* w, cudaStream_t st);
int lossbuffer_add_launch(void* state, const float* loss, cudaStream_t st);
int laplace_train_blocks(long long n);
int laplace_train_launch(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask, long long m_bs,
                         long long m_ss, const long long* gather, int B, int S, int C, long long HW, float eps_min, float eps_max,
                         void* lb_state, const float* fixed_w, int update_buffer, float* dout, float* part, float* loss,
                         float* weights, float* weighted, cudaStream_t st, float* mpart = nullptr, float* metrics = nullptr, int kind = 0);
int validation_scratch_floats(int S);
int validation_laplace_launch(const float* p1, const float* p2, long long p_bs, long long p_ss, const float* y, const float* mask, int B,
                              int S, long long inner, float eps_min, float eps_max, float* mean, float* alea_std, float* epi_std,
                              float* err, float* scratch, float* scalars, cudaStream_t st);
int scale_by_scalar_launch(float* x, long long n, const float* s, cudaStream_t st);
int aggregate_launch(const float* p1, long long p1_bs, long long p1_ss, const float* p2, long long p2_bs, long long p2_ss, int B,
                     int S, long long inner, float* mean, float* alea, float* epi, cudaStream_t st);

}  // namespace mimo
