// Whole-network executor for the MIMO U-Net (bilinear path): builds the layer graph once, lays every
// activation / gradient / scratch buffer out in ONE caller-provided workspace (no allocation, no sync), and
// replays forward / backward as a fixed sequence of kernel launches on the caller's stream.
//
// Topology restated from the reference (mimo/models/mimo_components/model.py:94-117,150-175,232-243,285-297):
//   per subnetwork s:  x[:,s] -> DoubleConv(Cin->f) = x1_s ; MaxPool ; DoubleConv(f->2f) = x2_s
//   xc = cat_s(x2_s)  ->  down2, down3, down4 (MaxPool + DoubleConv)  ->  up1, up2, up3 (bilinear x2, cat skip)
//   per subnetwork s:  up(u3) cat x1_s -> DoubleConv -> 1x1 head
// Concats are never materialised as copies: producers write straight into channel slices of the consumer's
// (reflect-haloed) input buffer.
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace mimo {
namespace {

struct Arena {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    return a;
  }
};

struct Buf {  // activation-like buffer inside the workspace
  size_t off = 0;
  int N = 0, H = 0, W = 0, pad = 0, cpitch = 0;
  size_t bytes() const { return (size_t)N * (H + (pad ? 2 : 0)) * (W + (pad ? 2 : 0)) * cpitch * sizeof(bf16); }
};

struct View {  // channel slice of a Buf
  int buf = -1;
  int c_off = 0, C = 0;
};

struct ConvL {
  int cin = 0, cout = 0, cin_p = 0, cout_p = 0, N = 0, H = 0, W = 0;
  int state0 = -1;  // index of "<prefix>.weight" in the state list; +1 bias, +2 bn.w, +3 bn.b, +4 rm, +5 rv, +6 nbt
  int m_tiles = 0;
  size_t wf = 0, wd = 0, psum = 0, psq = 0, vec = 0 /* scale,shift,mean,invstd,s1,s2: 6*cout_p floats */, bnpart = 0, dwp = 0;
  int y = -1, dy = -1, dpad = -1;  // Buf ids: raw output, its gradient, padded-domain input gradient
  size_t wf_v = 0; int cin_pv = 0; // inference-only weight pack in the chunk-aligned channel order of a virtual concat (decoders)
  int cin_x = 0;                   // PHYSICAL input channels (>= cin): the input buffer may keep its leading subnetwork slices 8-aligned
  ChannelSlices sl{0, 0, 0};       // ... with this layout (sl_n == 0: cin_x == cin)
};

struct Node {  // one DoubleConv
  std::string name;
  ConvL c1, c2;
  View in;        // whole buffer (c_off == 0)
  int a1 = -1;    // Buf id of the intermediate activation
  View out;       // destination slice
  View pool;      // pooled destination slice (buf == -1: none)
  int g1 = -1;    // Buf id: gradient w.r.t. a1 (unpadded)
  View g2;        // gradient w.r.t. out (unpadded); may be a slice of a wider gradient buffer
  bool need_in_grad = true;
};

}  // namespace
}  // namespace mimo

using namespace mimo;

struct mimo_unet_plan {
  mimo_unet_config_t cfg;
  std::vector<Buf> bufs;
  std::vector<Node> nodes;  // canonical double-conv order == state_dict order
  std::map<std::string, int> node_by_name;
  // node indices
  std::vector<int> enc_in, enc_down, dec;
  int down2 = -1, down3 = -1, down4 = -1, up1 = -1, up2 = -1, up3 = -1;
  // extra buffers
  std::vector<int> xin, dcat, p1, gp1;   // per subnetwork
  std::vector<int> x1d;                  // per subnetwork: DENSE copy of the encoder output x1_s (training / unfused path): the decoder's
                                         // concat kernel merges it with the up-sampled core output into whole lines of dcat[s]
  bool last_dense_skip = false;          // the last forward wrote x1_s into x1d (backward reads the pooling winners from there)
  int cat3 = -1, cat2 = -1, cat1 = -1, pxc = -1, px3 = -1, px4 = -1, x5 = -1, u1 = -1, u2 = -1, u3 = -1;
  int stack_c = 0;   // physical channels of the subnetwork stack at the head of cat3 / pxc (S * round_up(2f, 8), or c)
  int upd = -1;   // inference: the up-sampled core output shared by all decoders (virtual concat), -1 when not used
  int g_xc = -1, g_x3 = -1, g_x4 = -1, g_x5 = -1, g_u1 = -1, g_u2 = -1, g_u3 = -1;
  int gp_xc = -1, gp_x3 = -1, gp_x4 = -1;  // pooled-map gradients
  int tmp0 = -1, tmp1 = -1, tmp2 = -1, tmp3 = -1;  // folded upsample-branch gradients per level
  std::vector<int> g_feat, g_x1;         // per subnetwork
  size_t head_part = 0;
  size_t dwp_begin = 0, dwp_end = 0;
  // CUDA graphs of the fixed launch sequences (forward body; the four backward stages). Only launches whose arguments are
  // workspace / bound-state pointers are captured; the kernels that touch caller tensors (input packing, heads) stay eager,
  // so a graph stays valid for as long as the binding does.
  struct GraphSlot { cudaGraphExec_t exec = nullptr; unsigned long long key = ~0ull; int launches = 0; int captures = 0; };
  GraphSlot g_fwd, g_bwd[4];
  cudaStream_t cap_stream = nullptr;   // private stream the graphs are captured on
  int graph_mode = 1;          // env MIMO_GRAPH (0 disables)
  int eval_fuse = 1;           // env MIMO_EVAL_FUSE (0: eval mode keeps the unfused conv -> y -> bn_relu_apply path)
  bool fuse_next = false;      // mimo_unet_set_inference_fusion: the caller promises not to call backward on the next forwards
  bool last_fused = false;     // the last forward took the fused inference path (it keeps no raw conv outputs: no backward)
  bool graph_failed = false;   // a capture failed once: stay eager
  int fwd_calls = 0, bwd_calls = 0;
  std::vector<WgradUnpackJob> unpack_jobs;   // pending weight-gradient transposes of the current backward stage
  int head_state0 = -1;
  int n_state = 0;
  size_t ws_bytes = 0;
  // binding
  uint8_t* ws = nullptr;
  std::vector<void*> state, grads;
  bool bound = false;
  bool last_training = false;
  bool have_forward = false;
  bool dy_tails_zeroed = false;  // the zero tails of the dy buffers are cleared once per binding
  // optional caller-owned events recorded during backward when a group of parameter gradients is final:
  // 0 decoders + heads, 1 core up path, 2 core down path, 3 encoders (= end of backward)
  cudaEvent_t stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  const float* const* last_masks = nullptr;
  std::vector<const float*> masks_copy;
  int launches = 0;
  // element-wise dropout (nn.Dropout) keep masks of the next forward/backward pair: centre of the core, one per decoder
  const bf16* center_keep = nullptr;
  float center_scale = 1.f;
  std::vector<const bf16*> final_keep;
  float final_scale = 1.f;
  int Hs[5], Ws[5];
  // optional per-launch CUDA-event profiling (bench.py roofline numbers)
  bool prof = false;
  std::vector<cudaEvent_t> ev;   // pairs
  std::vector<int> ev_cls, ev_tag, ev_kid;   // class, 2*node+conv tag, tensor-core kernel id (common.cuh note_kernel)
  int ev_used = 0;
  int cur_tag = -1;              // 2*node + (second conv) of the launch being enqueued, -1 outside a DoubleConv
};

namespace mimo {
namespace eng {

inline int p8(int c) { return round_up(c, 8); }

int add_buf(mimo_unet_plan* P, Arena& A, int N, int H, int W, int pad, int C) {
  Buf b;
  b.N = N; b.H = H; b.W = W; b.pad = pad; b.cpitch = p8(C);
  b.off = A.take(b.bytes());
  P->bufs.push_back(b);
  return (int)P->bufs.size() - 1;
}

ActView view_of(const mimo_unet_plan* P, int buf, int c_off, int C) {
  const Buf& b = P->bufs[buf];
  ActView v;
  v.base = reinterpret_cast<bf16*>(P->ws + b.off);
  v.N = b.N; v.H = b.H; v.W = b.W; v.pad = b.pad; v.cpitch = b.cpitch; v.c_off = c_off; v.C = C;
  return v;
}
ActView view_of(const mimo_unet_plan* P, const View& v) { return view_of(P, v.buf, v.c_off, v.C); }

void setup_conv(mimo_unet_plan* P, Arena& A, ConvL& c, int cin, int cout, int N, int H, int W, int& state_cursor,
                ChannelSlices sl = ChannelSlices{0, 0, 0}) {
  c.cin = cin; c.cout = cout; c.sl = sl; c.cin_x = sl.phys_count(cin); c.cin_p = p8(c.cin_x); c.cout_p = p8(cout); c.N = N; c.H = H; c.W = W;
  c.state0 = state_cursor;
  state_cursor += 7;
  c.m_tiles = conv3x3_stat_rows();
  c.wf = A.take((size_t)9 * cout * c.cin_p * sizeof(bf16));
  c.wd = A.take((size_t)9 * c.cin_x * c.cout_p * sizeof(bf16));
  c.psum = A.take((size_t)c.m_tiles * c.cout_p * sizeof(float));
  c.psq = A.take((size_t)c.m_tiles * c.cout_p * sizeof(float));
  c.vec = A.take((size_t)6 * c.cout_p * sizeof(float));
  c.bnpart = A.take((size_t)bn_bwd_parts(cout) * 2 * cout * sizeof(float));
  c.y = add_buf(P, A, N, H, W, 0, cout);
  c.dy = add_buf(P, A, N, H, W, 2, cout);  // zero-tail layout: read by the flat dgrad / wgrad kernels
  c.dpad = add_buf(P, A, N, H + 2, W + 2, 0, c.cin_x);
}

int add_node(mimo_unet_plan* P, Arena& A, const std::string& name, View in, int cin, int cmid, int cout, View out, View pool,
             int level, int& state_cursor, ChannelSlices in_slices = ChannelSlices{0, 0, 0}) {
  Node n;
  n.name = name;
  const int N = P->cfg.batch, H = P->Hs[level], W = P->Ws[level];
  setup_conv(P, A, n.c1, cin, cmid, N, H, W, state_cursor, in_slices);
  setup_conv(P, A, n.c2, cmid, cout, N, H, W, state_cursor);
  n.in = in;
  n.a1 = add_buf(P, A, N, H, W, 1, cmid);
  n.out = out;
  n.pool = pool;
  n.g1 = add_buf(P, A, N, H, W, 0, cmid);
  P->nodes.push_back(n);
  P->node_by_name[name] = (int)P->nodes.size() - 1;
  return (int)P->nodes.size() - 1;
}

enum ProfClass { kPackW = 0, kPackIn, kConvFprop, kBnFinalize, kBnApply, kUpsample, kHeadFwd, kHeadBwd, kBnBwd, kConvWgrad,
                 kWgradUnpack, kConvDgrad, kGradGather, kUpsampleBwd, kOther, kNumProfClasses };
static const char* kProfNames[kNumProfClasses] = {"weight_pack", "pack_input", "conv_fprop", "bn_finalize", "bn_relu_apply", "upsample",
                                                  "head_fwd", "head_bwd", "bn_relu_bwd", "conv_wgrad", "wgrad_unpack", "conv_dgrad",
                                                  "grad_gather", "upsample_bwd", "other"};

inline void prof_begin(mimo_unet_plan* P, int cls, cudaStream_t st) {
  if (!P->prof || (size_t)(2 * P->ev_used + 1) >= P->ev.size()) return;
  cudaEventRecord(P->ev[2 * P->ev_used], st);
  P->ev_cls[P->ev_used] = cls;
  P->ev_tag[P->ev_used] = P->cur_tag;
  note_kernel(0);
}
inline void prof_end(mimo_unet_plan* P, cudaStream_t st) {
  if (!P->prof || (size_t)(2 * P->ev_used + 1) >= P->ev.size()) return;
  cudaEventRecord(P->ev[2 * P->ev_used + 1], st);
  P->ev_kid[P->ev_used] = last_kernel();
  ++P->ev_used;
}

#define RUN(cls, expr)              \
  do {                              \
    prof_begin(P, (cls), st);       \
    int _rc = (expr);               \
    prof_end(P, st);                \
    if (_rc != MIMO_OK) return _rc; \
    ++P->launches;                  \
  } while (0)

void drop_graphs(mimo_unet_plan* P) {
  if (P->g_fwd.exec) { cudaGraphExecDestroy(P->g_fwd.exec); P->g_fwd.exec = nullptr; }
  for (auto& g : P->g_bwd)
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
}

// Runs `body` (a fixed sequence of launches on `st`) through a cached CUDA graph: replay when the key matches, otherwise
// capture + instantiate + launch. Any capture problem (stream already capturing, unsupported call) falls back to the
// plain launches for good.
template <class F>
int run_graphed(mimo_unet_plan* P, mimo_unet_plan::GraphSlot& slot, unsigned long long key, bool allow, cudaStream_t& st, F&& body) {
  if (!allow) return body();
  if (slot.exec && slot.key == key) {
    MIMO_CUDA(cudaGraphLaunch(slot.exec, st));
    P->launches += slot.launches;
    return MIMO_OK;
  }
  if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
  // a caller that keeps alternating keys (BatchNorm mode / accumulate flag every call) would re-capture every time, which is
  // slower than plain launches: give up on graphs for this plan after a few dozen captures of the same sequence
  if (++slot.captures > 32) { P->graph_failed = true; return body(); }
  // `st` is the variable the body's launches read (captured by reference). The capture runs on a private stream -- the
  // caller's stream is usually the legacy default stream, which cannot capture, and nothing executes during capture anyway
  // -- and the instantiated graph is launched on the caller's stream.
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return body(); }
  if (!P->cap_stream && cudaStreamCreateWithFlags(&P->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    P->graph_failed = true;
    return body();
  }
  if (cudaStreamBeginCapture(P->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    P->graph_failed = true;
    return body();
  }
  const cudaStream_t user_stream = st;
  st = P->cap_stream;
  const int l0 = P->launches;
  const int rc = body();
  st = user_stream;
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(P->cap_stream, &g);
  if (rc != MIMO_OK || e != cudaSuccess || g == nullptr) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    P->graph_failed = true;
    P->launches = l0;
    return rc != MIMO_OK ? rc : body();
  }
  const cudaError_t ei = cudaGraphInstantiate(&slot.exec, g, 0);
  cudaGraphDestroy(g);
  if (ei != cudaSuccess) {
    slot.exec = nullptr;
    cudaGetLastError();
    P->graph_failed = true;
    P->launches = l0;
    return body();
  }
  slot.key = key;
  slot.launches = P->launches - l0;
  MIMO_CUDA(cudaGraphLaunch(slot.exec, st));
  return MIMO_OK;
}

float* fptr(const mimo_unet_plan* P, size_t off) { return reinterpret_cast<float*>(P->ws + off); }
bf16* bptr(const mimo_unet_plan* P, size_t off) { return reinterpret_cast<bf16*>(P->ws + off); }

int conv_bn_forward(mimo_unet_plan* P, ConvL& c, const ActView& in, const ActView& out, const ActView* pool, const float* drop,
                    bool training, cudaStream_t st, const ActView* in2 = nullptr) {
  bf16* y = reinterpret_cast<bf16*>(P->ws + P->bufs[c.y].off);
  float* vec = fptr(P, c.vec);
  float *scale = vec, *shift = vec + c.cout_p, *mean = vec + 2 * c.cout_p, *invstd = vec + 3 * c.cout_p;
  // Inference (BatchNorm with running statistics): the affine is known BEFORE the convolution, so it is folded into the conv
  // epilogue together with the ReLU and the Dropout2d factors (reference components.py:23-29 in eval mode, ensemble.py:54-66
  // for MC dropout). The epilogue writes the activation straight into the interior of the consumer's haloed buffer: no raw
  // output y, no bn_relu_apply pass; the halo ring and the pooled copy are produced by two small kernels. The first conv
  // channels past C of the last 8-channel group are written as zeros: they are pad channels or belong to a concat slice that a
  // LATER kernel of the forward pass writes (up-sampling into dcat, the next subnetwork's slice of cat3).
  if (!training && P->eval_fuse && P->fuse_next && out.pad == 1 && (out.c_off & 7) == 0 && out.c_off + round_up(c.cout, 8) <= out.cpitch &&
      (in2 != nullptr ? conv3x3_c2_ok(in, 0, c.cout) : conv3x3_fuse_ok(in, c.cout))) {
    const float* bias0 = (const float*)P->state[c.state0 + 1];
    RUN(kBnFinalize, bn_eval_affine_launch(c.cout, (const float*)P->state[c.state0 + 2], (const float*)P->state[c.state0 + 3], bias0,
                                           (const float*)P->state[c.state0 + 4], (const float*)P->state[c.state0 + 5], 1e-5f, scale, shift,
                                           mean, invstd, st));
    ConvFuse fz{scale, shift, drop, 1, 1};
    if (in2 != nullptr)   // virtual concat (decoders): `in` is the skip slice, *in2 the shared up-sampled core output
      RUN(kConvFprop, conv3x3_c2_launch(in, 0, bptr(P, c.wf_v), c.cout, c.cin_pv, out.base + out.c_off, out.cpitch, nullptr, nullptr, nullptr,
                                        0, st, &fz, in2));
    else
      RUN(kConvFprop, conv3x3_launch(in, 0, bptr(P, c.wf), c.cout, c.cin_p, out.base + out.c_off, out.cpitch, nullptr, nullptr, nullptr, 0,
                                     st, &fz));
    RUN(kBnApply, halo_fill_launch(out, st));
    if (pool) RUN(kBnApply, maxpool_launch(out, *pool, nullptr, st));
    return MIMO_OK;
  }
  RUN(kConvFprop, conv3x3_launch(in, 0, bptr(P, c.wf), c.cout, c.cin_p, y, c.cout_p, training ? fptr(P, c.psum) : nullptr,
                     training ? fptr(P, c.psq) : nullptr, nullptr, 0, st));
  const float* bias = (const float*)P->state[c.state0 + 1];
  const float* gamma = (const float*)P->state[c.state0 + 2];
  const float* beta = (const float*)P->state[c.state0 + 3];
  float* rm = (float*)P->state[c.state0 + 4];
  float* rv = (float*)P->state[c.state0 + 5];
  long long* nbt = (long long*)P->state[c.state0 + 6];
  if (training) {
    RUN(kBnFinalize, bn_finalize_launch(fptr(P, c.psum), fptr(P, c.psq), c.m_tiles, c.cout_p, c.cout, (double)c.N * c.H * c.W, gamma, beta, bias,
                           rm, rv, nbt, 0.1f, 1e-5f, scale, shift, mean, invstd, st));
  } else {
    RUN(kBnFinalize, bn_eval_affine_launch(c.cout, gamma, beta, bias, rm, rv, 1e-5f, scale, shift, mean, invstd, st));
  }
  RUN(kBnApply, bn_relu_apply_launch(y, c.cout_p, scale, shift, drop, out, pool, st));
  return MIMO_OK;
}

int node_forward(mimo_unet_plan* P, int ni, bool training, const float* drop, cudaStream_t st, const ActView* in1 = nullptr,
                 const ActView* in2 = nullptr, const ActView* out_override = nullptr) {
  Node& n = P->nodes[ni];
  const ActView in = in1 ? *in1 : view_of(P, n.in);
  const ActView a1 = view_of(P, n.a1, 0, n.c1.cout);
  const ActView out = out_override ? *out_override : view_of(P, n.out);
  P->cur_tag = 2 * ni;
  int rc = conv_bn_forward(P, n.c1, in, a1, nullptr, nullptr, training, st, in2);
  if (rc) return rc;
  P->cur_tag = 2 * ni + 1;
  if (n.pool.buf >= 0) {
    const ActView pool = view_of(P, n.pool);
    rc = conv_bn_forward(P, n.c2, a1, out, &pool, drop, training, st);
  } else {
    rc = conv_bn_forward(P, n.c2, a1, out, nullptr, drop, training, st);
  }
  P->cur_tag = -1;
  return rc;
}

int conv_bn_backward(mimo_unet_plan* P, ConvL& c, const ActView& G, const ActView& in, const float* drop, bool training,
                     bool need_in_grad, int accumulate, cudaStream_t st, const ActView* fold_src = nullptr) {
  float* vec = fptr(P, c.vec);
  float *scale = vec, *shift = vec + c.cout_p, *mean = vec + 2 * c.cout_p, *invstd = vec + 3 * c.cout_p, *s1s2 = vec + 4 * c.cout_p;
  const bf16* y = bptr(P, P->bufs[c.y].off);
  const ActView dyv = view_of(P, c.dy, 0, c.cout);
  RUN(kBnBwd, bn_bwd_launch(G, y, c.cout_p, scale, shift, mean, invstd, drop, training ? 1 : 0, fptr(P, c.bnpart), s1s2,
                    (float*)P->grads[c.state0 + 2], (float*)P->grads[c.state0 + 3], (float*)P->grads[c.state0 + 1], 1.f, accumulate,
                    dyv, st, fold_src));
  ++P->launches; ++P->launches;  // bn_bwd is three kernels
  if (P->grads[c.state0] != nullptr) {
    RUN(kConvWgrad, conv3x3_wgrad_launch(dyv, in, fptr(P, c.dwp), c.cin_p, st, true));
    P->unpack_jobs.push_back({fptr(P, c.dwp), (float*)P->grads[c.state0], c.cout, c.cin, c.cin_p, c.sl});  // flushed per stage
  }
  if (need_in_grad) {
    bf16* dpad = bptr(P, P->bufs[c.dpad].off);
    RUN(kConvDgrad, conv3x3_launch(dyv, 1, bptr(P, c.wd), c.cin_x, c.cout_p, dpad, c.cin_p, nullptr, nullptr, nullptr, 0, st));
  }
  return MIMO_OK;
}

int node_backward(mimo_unet_plan* P, int ni, bool training, const float* drop, int accumulate, cudaStream_t st) {
  Node& n = P->nodes[ni];
  const ActView G2 = view_of(P, n.g2);
  const ActView a1 = view_of(P, n.a1, 0, n.c1.cout);
  P->cur_tag = 2 * ni + 1;
  int rc = conv_bn_backward(P, n.c2, G2, a1, drop, training, true, accumulate, st);
  if (rc) return rc;
  const ActView dpad2 = view_of(P, n.c2.dpad, 0, n.c2.cin);
  const ActView G1 = view_of(P, n.g1, 0, n.c1.cout);
  P->cur_tag = 2 * ni;
  // G1 = fold_reflect(dpad2) is formed inside the BN-backward passes of c1 (bn_bwd_launch falls back to a grad_gather
  // launch into G1 when the operands are not dense)
  const ActView in = view_of(P, n.in);
  rc = conv_bn_backward(P, n.c1, G1, in, nullptr, training, n.need_in_grad, accumulate, st, &dpad2);
  P->cur_tag = -1;
  return rc;
}

__global__ void unpack_input_grad_kernel(ActView dpad /* (H+2)x(W+2) domain */, int H, int W, int C, float* dx, long long sb, long long sc) {
  const long long total = (long long)dpad.N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
    int hs[3] = {h + 1, (h == 1) ? 0 : -1, (h == H - 2) ? H + 1 : -1};
    int ws[3] = {w + 1, (w == 1) ? 0 : -1, (w == W - 2) ? W + 1 : -1};
    for (int c = 0; c < C; ++c) {
      float a = 0.f;
      for (int p = 0; p < 3; ++p) {
        if (hs[p] < 0) continue;
        for (int q = 0; q < 3; ++q) {
          if (ws[q] < 0) continue;
          a += __bfloat162float(dpad.base[dpad.pix(n, hs[p], ws[q]) + c]);
        }
      }
      dx[n * sb + c * sc + (long long)h * W + w] = a;
    }
  }
}

}  // namespace eng
}  // namespace mimo
using namespace mimo::eng;

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

int mimo_unet_plan_create(const mimo_unet_config_t* cfg, mimo_unet_plan_t** out) {
  MIMO_CHECK(cfg && out, MIMO_ERR_ARG, "plan_create: null argument");
  MIMO_CHECK(cfg->num_subnetworks >= 1 && cfg->num_subnetworks <= 16, MIMO_ERR_ARG, "plan_create: num_subnetworks must be in [1,16]");
  MIMO_CHECK(cfg->in_channels >= 1 && cfg->out_channels >= 1 && cfg->out_channels <= 8, MIMO_ERR_ARG, "plan_create: out_channels must be in [1,8]");
  MIMO_CHECK(cfg->filter_base_count >= 1 && cfg->batch >= 1, MIMO_ERR_ARG, "plan_create: bad filter_base_count / batch");
  MIMO_CHECK(cfg->height / 16 >= 2 && cfg->width / 16 >= 2, MIMO_ERR_ARG,
             "plan_create: H and W must be >= 32 (reflect padding needs every feature map >= 2 px), got %dx%d", cfg->height, cfg->width);
  mimo_unet_plan* P = new mimo_unet_plan();
  P->cfg = *cfg;
  { const char* e = getenv("MIMO_GRAPH"); P->graph_mode = e ? atoi(e) : 1; }
  { const char* e = getenv("MIMO_EVAL_FUSE"); P->eval_fuse = e ? atoi(e) : 1; }
  const int S = cfg->num_subnetworks, f = cfg->filter_base_count, N = cfg->batch, Cin = cfg->in_channels;
  P->Hs[0] = cfg->height; P->Ws[0] = cfg->width;
  for (int l = 1; l < 5; ++l) { P->Hs[l] = P->Hs[l - 1] / 2; P->Ws[l] = P->Ws[l - 1] / 2; }
  const int c = 2 * f * S;
  Arena A;
  auto buf = [&](int level, int pad, int C) { return add_buf(P, A, N, P->Hs[level], P->Ws[level], pad, C); };
  for (int s = 0; s < S; ++s) {
    P->xin.push_back(buf(0, 1, Cin));
    P->dcat.push_back(buf(0, 1, f + c / 2));
    P->p1.push_back(buf(1, 1, f));
    P->gp1.push_back(buf(1, 0, f));
    P->g_feat.push_back(buf(0, 0, f));
    P->g_x1.push_back(buf(0, 0, f));
    P->x1d.push_back(buf(0, 0, f));
  }
  // The subnetwork stack [x2_0 | x2_1 | ...] (2f channels each) leads cat3 and pxc. With 2f % 8 != 0 (f = 21: 42) every slice but the
  // first would start at a channel offset that is not 16-byte aligned: scalar stores in bn_relu_apply, unaligned loads in the BatchNorm
  // backward and no fused inference epilogue for those subnetworks. So the slices sit round_up(2f, 8) apart (gap channels stay
  // zero), and the two convolutions that read the stack (core.down2.c1, core.up3.c1) get weights packed in that order (ChannelSlices).
  static const int aligned_on = getenv("MIMO_ALIGNED_STACK") ? atoi(getenv("MIMO_ALIGNED_STACK")) : 1;
  const int sp = (aligned_on && S > 1) ? round_up(2 * f, 8) : 2 * f;   // slice stride
  const int cs = S * sp;                                                // physical width of the stack (c = S * 2f logical)
  const ChannelSlices stack = sp != 2 * f ? ChannelSlices{2 * f, sp, S} : ChannelSlices{0, 0, 0};
  P->stack_c = cs;
  P->cat3 = buf(1, 1, cs + c); P->pxc = buf(2, 1, cs);
  P->cat2 = buf(2, 1, 4 * c); P->px3 = buf(3, 1, 2 * c);
  P->cat1 = buf(3, 1, 8 * c); P->px4 = buf(4, 1, 4 * c);
  P->x5 = buf(4, 1, 4 * c); P->u1 = buf(3, 1, 2 * c); P->u2 = buf(2, 1, c); P->u3 = buf(1, 1, c / 2);
  P->g_xc = buf(1, 0, cs); P->g_x3 = buf(2, 0, 2 * c); P->g_x4 = buf(3, 0, 4 * c); P->g_x5 = buf(4, 0, 4 * c);
  P->g_u1 = buf(3, 0, 2 * c); P->g_u2 = buf(2, 0, c); P->g_u3 = buf(1, 0, c / 2);
  P->gp_xc = buf(2, 0, cs); P->gp_x3 = buf(3, 0, 2 * c); P->gp_x4 = buf(4, 0, 4 * c);
  P->tmp0 = buf(0, 0, c / 2); P->tmp1 = buf(1, 0, c); P->tmp2 = buf(2, 0, 2 * c); P->tmp3 = buf(3, 0, 4 * c);

  int cur = 0;
  auto V = [](int b, int off, int C) { View v; v.buf = b; v.c_off = off; v.C = C; return v; };
  const View none = V(-1, 0, 0);
  // canonical (state_dict) order: encoder.in_convs.*, encoder.down1s.*, core.*, decoder.up4s.*, decoder.outcs.*
  for (int s = 0; s < S; ++s) {
    int ni = add_node(P, A, "encoder.in_convs." + std::to_string(s), V(P->xin[s], 0, Cin), Cin, f, f, V(P->dcat[s], 0, f),
                      V(P->p1[s], 0, f), 0, cur);
    P->nodes[ni].need_in_grad = false;
    P->nodes[ni].g2 = V(P->g_x1[s], 0, f);
    P->enc_in.push_back(ni);
  }
  for (int s = 0; s < S; ++s) {
    int ni = add_node(P, A, "encoder.down1s." + std::to_string(s), V(P->p1[s], 0, f), f, 2 * f, 2 * f, V(P->cat3, sp * s, 2 * f),
                      V(P->pxc, sp * s, 2 * f), 1, cur);
    P->nodes[ni].g2 = V(P->g_xc, sp * s, 2 * f);
    P->enc_down.push_back(ni);
  }
  P->down2 = add_node(P, A, "core.down2", V(P->pxc, 0, cs), c, 2 * c, 2 * c, V(P->cat2, 0, 2 * c), V(P->px3, 0, 2 * c), 2, cur, stack);
  P->nodes[P->down2].g2 = V(P->g_x3, 0, 2 * c);
  P->down3 = add_node(P, A, "core.down3", V(P->px3, 0, 2 * c), 2 * c, 4 * c, 4 * c, V(P->cat1, 0, 4 * c), V(P->px4, 0, 4 * c), 3, cur);
  P->nodes[P->down3].g2 = V(P->g_x4, 0, 4 * c);
  P->down4 = add_node(P, A, "core.down4", V(P->px4, 0, 4 * c), 4 * c, 4 * c, 4 * c, V(P->x5, 0, 4 * c), none, 4, cur);
  P->nodes[P->down4].g2 = V(P->g_x5, 0, 4 * c);
  P->up1 = add_node(P, A, "core.up1", V(P->cat1, 0, 8 * c), 8 * c, 4 * c, 2 * c, V(P->u1, 0, 2 * c), none, 3, cur);
  P->nodes[P->up1].g2 = V(P->g_u1, 0, 2 * c);
  P->up2 = add_node(P, A, "core.up2", V(P->cat2, 0, 4 * c), 4 * c, 2 * c, c, V(P->u2, 0, c), none, 2, cur);
  P->nodes[P->up2].g2 = V(P->g_u2, 0, c);
  P->up3 = add_node(P, A, "core.up3", V(P->cat3, 0, cs + c), 2 * c, c, c / 2, V(P->u3, 0, c / 2), none, 1, cur, stack);
  P->nodes[P->up3].g2 = V(P->g_u3, 0, c / 2);
  const int d = c / 2 + f;
  for (int s = 0; s < S; ++s) {
    // the decoder feature map is only read by the 1x1 head: it still gets a halo for uniformity
    int featb = buf(0, 1, f);
    int ni = add_node(P, A, "decoder.up4s." + std::to_string(s), V(P->dcat[s], 0, d), d, d / 2, f, V(featb, 0, f), none, 0, cur);
    P->nodes[ni].g2 = V(P->g_feat[s], 0, f);
    P->dec.push_back(ni);
  }
  // Inference with more than 64 concat channels per decoder (M >= 3 at f = 21): the decoders' first conv reads cat([x1_s, up(u3)])
  // as a VIRTUAL concat (conv_c2.cu): u3 is up-sampled ONCE into `upd` instead of once per subnetwork into every dcat[s].
  if (f + c / 2 > 64) {
    P->upd = buf(0, 1, c / 2);
    for (int s = 0; s < S; ++s) {
      ConvL& cl = P->nodes[P->dec[s]].c1;
      cl.cin_pv = round_up(round_up(f, 64) + c / 2, 8);
      cl.wf_v = A.take((size_t)9 * cl.cout * cl.cin_pv * sizeof(bf16));
    }
  }
  P->head_state0 = cur;
  cur += 2 * S;
  P->n_state = cur;
  // packed fp32 weight gradients of all layers in ONE contiguous range: cleared by a single memset per backward pass
  P->dwp_begin = A.take(0);
  for (auto& n : P->nodes)
    for (ConvL* cl : {&n.c1, &n.c2}) cl->dwp = A.take((size_t)9 * cl->cout * cl->cin_p * sizeof(float));
  P->dwp_end = A.off;
  P->head_part = A.take((size_t)head_bwd_parts() * (cfg->out_channels * f + cfg->out_channels) * sizeof(float));
  P->ws_bytes = A.off + 256;
  *out = P;
  return MIMO_OK;
}

void mimo_unet_plan_destroy(mimo_unet_plan_t* plan) {
  if (plan) {
    for (auto& e : plan->ev) cudaEventDestroy(e);
    drop_graphs(plan);
    if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
  }
  delete plan;
}
size_t mimo_unet_workspace_bytes(const mimo_unet_plan_t* plan) { return plan ? plan->ws_bytes : 0; }
int mimo_unet_num_state(const mimo_unet_plan_t* plan) { return plan ? plan->n_state : 0; }
int mimo_unet_num_double_convs(const mimo_unet_plan_t* plan) { return plan ? (int)plan->nodes.size() : 0; }
int mimo_unet_dropout_channels(const mimo_unet_plan_t* plan, int i) {
  if (!plan || i < 0 || i >= (int)plan->nodes.size()) return 0;
  return plan->nodes[i].c2.cout;
}
int mimo_unet_last_launches(const mimo_unet_plan_t* plan) { return plan ? plan->launches : 0; }
int mimo_unet_graph_state(const mimo_unet_plan_t* plan) {
  if (!plan) return 0;
  int m = plan->g_fwd.exec ? 1 : 0;
  for (int k = 0; k < 4; ++k) m |= plan->g_bwd[k].exec ? (2 << k) : 0;
  if (plan->graph_failed) m |= 0x100;
  return m;
}

int mimo_unet_profile_classes(void) { return kNumProfClasses; }
const char* mimo_unet_profile_class_name(int i) { return (i >= 0 && i < kNumProfClasses) ? kProfNames[i] : ""; }
int mimo_unet_profile_enable(mimo_unet_plan_t* P, int on) {
  MIMO_CHECK(P, MIMO_ERR_ARG, "profile_enable: null plan");
  if (on && P->ev.empty()) {
    P->ev.resize(2 * 4096);
    P->ev_cls.assign(4096, 0);
    P->ev_tag.assign(4096, -1);
    P->ev_kid.assign(4096, 0);
    for (auto& e : P->ev) MIMO_CUDA(cudaEventCreate(&e));
  }
  P->prof = on != 0;
  P->ev_used = 0;
  return MIMO_OK;
}
int mimo_unet_profile_read(mimo_unet_plan_t* P, float* ms_by_class, int* count_by_class) {
  MIMO_CHECK(P && ms_by_class && count_by_class, MIMO_ERR_ARG, "profile_read: null argument");
  MIMO_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < kNumProfClasses; ++i) { ms_by_class[i] = 0.f; count_by_class[i] = 0; }
  for (int i = 0; i < P->ev_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, P->ev[2 * i], P->ev[2 * i + 1]) == cudaSuccess) {
      ms_by_class[P->ev_cls[i]] += ms;
      count_by_class[P->ev_cls[i]] += 1;
    }
  }
  P->ev_used = 0;
  return MIMO_OK;
}

int mimo_unet_profile_read_launches(mimo_unet_plan_t* P, int max_n, float* ms, int* cls, int* tag) {
  MIMO_CHECK(P && ms && cls && tag, MIMO_ERR_ARG, "profile_read_launches: null argument");
  MIMO_CUDA(cudaDeviceSynchronize());
  int n = P->ev_used < max_n ? P->ev_used : max_n;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, P->ev[2 * i], P->ev[2 * i + 1]);
    ms[i] = t; cls[i] = P->ev_cls[i]; tag[i] = P->ev_tag[i];
  }
  P->ev_used = 0;
  return n;
}
int mimo_unet_profile_read_launches_ex(mimo_unet_plan_t* P, int max_n, float* ms, int* cls, int* tag, int* kernel) {
  MIMO_CHECK(P && ms && cls && tag && kernel, MIMO_ERR_ARG, "profile_read_launches_ex: null argument");
  MIMO_CUDA(cudaDeviceSynchronize());
  int n = P->ev_used < max_n ? P->ev_used : max_n;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, P->ev[2 * i], P->ev[2 * i + 1]);
    ms[i] = t; cls[i] = P->ev_cls[i]; tag[i] = P->ev_tag[i]; kernel[i] = P->ev_kid[i];
  }
  P->ev_used = 0;
  return n;
}
const char* mimo_conv_kernel_name(int id) { return conv_kernel_name(id); }
const char* mimo_unet_node_name(const mimo_unet_plan_t* P, int i) {
  return (P && i >= 0 && i < (int)P->nodes.size()) ? P->nodes[i].name.c_str() : "";
}

int mimo_unet_bind(mimo_unet_plan_t* P, void* workspace, size_t workspace_bytes, void* const* state, void* const* grads, int n) {
  MIMO_CHECK(P && workspace && state, MIMO_ERR_ARG, "bind: null argument");
  MIMO_CHECK(n == P->n_state, MIMO_ERR_ARG, "bind: expected %d state entries, got %d", P->n_state, n);
  MIMO_CHECK(workspace_bytes >= P->ws_bytes, MIMO_ERR_ARG, "bind: workspace too small (%zu < %zu)", workspace_bytes, P->ws_bytes);
  MIMO_CHECK(((uintptr_t)workspace % 256) == 0, MIMO_ERR_ALIGN, "bind: workspace must be 256-byte aligned");
  P->ws = (uint8_t*)workspace;
  P->state.assign(state, state + n);
  if (grads) P->grads.assign(grads, grads + n);
  else P->grads.assign(n, nullptr);
  P->bound = true;
  P->have_forward = false;
  P->dy_tails_zeroed = false;
  drop_graphs(P);   // the graphs embed the previous binding's pointers
  return MIMO_OK;
}

int mimo_unet_forward(mimo_unet_plan_t* P, const float* x, const long long* gather, int training, const float* const* drop_masks,
                      float* out, void* stream) {
  MIMO_CHECK(P && P->bound, MIMO_ERR_STATE, "forward: plan is not bound");
  MIMO_CHECK(x && out, MIMO_ERR_ARG, "forward: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const mimo_unet_config_t& cfg = P->cfg;
  const int S = cfg.num_subnetworks, f = cfg.filter_base_count, Cin = cfg.in_channels, B = cfg.batch;
  const long long HW = (long long)cfg.height * cfg.width;
  const int c = 2 * f * S;
  P->launches = 0;
  P->masks_copy.assign(P->nodes.size(), nullptr);
  if (drop_masks)
    for (size_t i = 0; i < P->nodes.size(); ++i) P->masks_copy[i] = drop_masks[i];
  auto mask = [&](int ni) { return P->masks_copy[ni]; };
  const bool tr = training != 0;

  // caller tensors are only touched by eager launches: input packing here, the 1x1 heads after the body
  for (int s = 0; s < S; ++s) {
    const ActView xin = view_of(P, P->xin[s], 0, Cin);
    // with a gather table x is the un-shuffled batch [B][Cin][H][W] shared by all subnetworks
    if (gather) RUN(kPackIn, pack_input_launch(x, (long long)Cin * HW, HW, gather + (long long)s * B, xin, st));
    else RUN(kPackIn, pack_input_launch(x + (long long)s * Cin * HW, (long long)S * Cin * HW, HW, nullptr, xin, st));
  }
  // Training / unfused path: the encoder output x1_s goes to a dense buffer and reaches the decoder's concat buffer through the
  // concat kernel (whole-line writes). The fused inference path keeps writing x1_s straight into dcat[s] from the conv epilogue.
  static const int dense_skip_on = getenv("MIMO_DENSE_SKIP") ? atoi(getenv("MIMO_DENSE_SKIP")) : 1;
  const bool dense_skip = dense_skip_on && !(!tr && P->eval_fuse && P->fuse_next) &&
                          upsample_concat_ok(view_of(P, P->u3, 0, c / 2), view_of(P, P->x1d[0], 0, f), view_of(P, P->dcat[0], 0, f + c / 2));
  const bool virt = !tr && P->eval_fuse && P->fuse_next && P->upd >= 0 &&
                    conv3x3_c2_ok(view_of(P, P->dcat[0], 0, f), 0, P->nodes[P->dec[0]].c1.cout);
  auto body = [&]() -> int {
    {  // bf16 weight packs of every layer (fprop layout + flipped dgrad layout): one launch
      std::vector<WeightPackJob> jobs;
      for (auto& n : P->nodes)
        for (ConvL* cl : {&n.c1, &n.c2})
          jobs.push_back({(const float*)P->state[cl->state0], bptr(P, cl->wf), bptr(P, cl->wd), cl->cout, cl->cin, cl->cin_p, cl->cout_p, cl->sl});
      if (virt)   // chunk-aligned packs of the decoders' first conv: [x1_s (f, padded to 64) | up (c/2)]
        for (int s2 = 0; s2 < S; ++s2) {
          ConvL& cl = P->nodes[P->dec[s2]].c1;
          jobs.push_back({(const float*)P->state[cl.state0], bptr(P, cl.wf_v), nullptr, cl.cout, cl.cin, cl.cin_pv, cl.cout_p,
                          ChannelSlices{f, round_up(f, 64), 1}});
        }
      RUN(kPackW, weight_pack_batched_launch(jobs.data(), (int)jobs.size(), st));
    }
    int rc;
    for (int s = 0; s < S; ++s) {
      if (dense_skip) {
        const ActView x1 = view_of(P, P->x1d[s], 0, f);
        if ((rc = node_forward(P, P->enc_in[s], tr, mask(P->enc_in[s]), st, nullptr, nullptr, &x1))) return rc;
      } else {
        if ((rc = node_forward(P, P->enc_in[s], tr, mask(P->enc_in[s]), st))) return rc;
      }
      if ((rc = node_forward(P, P->enc_down[s], tr, mask(P->enc_down[s]), st))) return rc;
    }
    if ((rc = node_forward(P, P->down2, tr, mask(P->down2), st))) return rc;
    if ((rc = node_forward(P, P->down3, tr, mask(P->down3), st))) return rc;
    if ((rc = node_forward(P, P->down4, tr, mask(P->down4), st))) return rc;
    if (P->center_keep)  // x_drop = center_dropout(x5), model.py:239
      RUN(kOther, mask_mul_launch(view_of(P, P->x5, 0, 4 * c), P->center_keep, p8(4 * c), P->center_scale, st));
    RUN(kUpsample, upsample_launch(view_of(P, P->x5, 0, 4 * c), view_of(P, P->cat1, 4 * c, 4 * c), st));
    if ((rc = node_forward(P, P->up1, tr, mask(P->up1), st))) return rc;
    RUN(kUpsample, upsample_launch(view_of(P, P->u1, 0, 2 * c), view_of(P, P->cat2, 2 * c, 2 * c), st));
    if ((rc = node_forward(P, P->up2, tr, mask(P->up2), st))) return rc;
    RUN(kUpsample, upsample_launch(view_of(P, P->u2, 0, c), view_of(P, P->cat3, P->stack_c, c), st));
    if ((rc = node_forward(P, P->up3, tr, mask(P->up3), st))) return rc;
    if (virt) RUN(kUpsample, upsample_launch(view_of(P, P->u3, 0, c / 2), view_of(P, P->upd, 0, c / 2), st));
    for (int s = 0; s < S; ++s) {
      if (virt) {
        const ActView skip = view_of(P, P->dcat[s], 0, f), up = view_of(P, P->upd, 0, c / 2);
        if ((rc = node_forward(P, P->dec[s], tr, mask(P->dec[s]), st, &skip, &up))) return rc;
      } else {
        if (dense_skip)   // skip + up-sampled core output -> whole 128-byte lines of the concat buffer
          RUN(kUpsample, upsample_concat_launch(view_of(P, P->u3, 0, c / 2), view_of(P, P->x1d[s], 0, f), view_of(P, P->dcat[s], 0, f + c / 2), st));
        else
          RUN(kUpsample, upsample_launch(view_of(P, P->u3, 0, c / 2), view_of(P, P->dcat[s], f, c / 2), st));
        if ((rc = node_forward(P, P->dec[s], tr, mask(P->dec[s]), st))) return rc;
      }
      if (s < (int)P->final_keep.size() && P->final_keep[s])  // x_i = final_dropouts[i](x_i), model.py:294
        RUN(kOther, mask_mul_launch(view_of(P, P->nodes[P->dec[s]].out), P->final_keep[s], p8(f), P->final_scale, st));
    }
    return MIMO_OK;
  };
  ++P->fwd_calls;
  bool elem_drop = P->center_keep != nullptr;
  for (const bf16* m : P->final_keep) elem_drop = elem_drop || (m != nullptr);
  const bool allow = P->graph_mode != 0 && !P->graph_failed && !P->prof && drop_masks == nullptr && !elem_drop && P->fwd_calls > 2;
  int rc = run_graphed(P, P->g_fwd, (tr ? 1ull : 0ull) | ((!tr && P->eval_fuse && P->fuse_next) ? 2ull : 0ull), allow, st, body);
  if (rc) return rc;
  const int K = cfg.out_channels;
  for (int s = 0; s < S; ++s) {
    const ActView feat = view_of(P, P->nodes[P->dec[s]].out);
    RUN(kHeadFwd, head_fwd_launch(feat, (const float*)P->state[P->head_state0 + 2 * s], (const float*)P->state[P->head_state0 + 2 * s + 1], K,
                        out + (long long)s * K * HW, (long long)S * K * HW, st));
  }
  P->last_training = tr;
  P->last_fused = !tr && P->eval_fuse && P->fuse_next;
  P->last_dense_skip = dense_skip;
  P->have_forward = true;
  return MIMO_OK;
}

int mimo_unet_backward(mimo_unet_plan_t* P, const float* dout, const float* grad_scale, float* dx, int accumulate, void* stream) {
  MIMO_CHECK(P && P->bound, MIMO_ERR_STATE, "backward: plan is not bound");
  MIMO_CHECK(P->have_forward, MIMO_ERR_STATE, "backward: no forward pass to differentiate");
  MIMO_CHECK(!P->last_fused, MIMO_ERR_STATE, "backward: the last forward ran with the fused inference epilogues (mimo_unet_set_inference_fusion), "
                                             "which keep no raw convolution outputs; run the forward without it to differentiate");
  MIMO_CHECK(dout, MIMO_ERR_ARG, "backward: null dout");
  cudaStream_t st = (cudaStream_t)stream;
  const mimo_unet_config_t& cfg = P->cfg;
  const int S = cfg.num_subnetworks, f = cfg.filter_base_count, Cin = cfg.in_channels, K = cfg.out_channels;
  const long long HW = (long long)cfg.height * cfg.width;
  const int c = 2 * f * S;
  const bool tr = P->last_training;
  P->launches = 0;
  auto mask = [&](int ni) { return P->masks_copy[ni]; };
  for (int s = 0; s < S; ++s) P->nodes[P->enc_in[s]].need_in_grad = (dx != nullptr);
  int rc;
  if (!P->dy_tails_zeroed) {
    // the interior of every dy buffer is rewritten each step; the 2-pixel zero tail only needs clearing once
    for (auto& n : P->nodes)
      for (ConvL* c : {&n.c1, &n.c2}) MIMO_CUDA(cudaMemsetAsync(P->ws + P->bufs[c->dy].off, 0, P->bufs[c->dy].bytes(), st));
    P->dy_tails_zeroed = true;
  }
  // ---- heads (read the caller's dout): eager ----
  for (int s = 0; s < S; ++s) {
    Node& n = P->nodes[P->dec[s]];
    const ActView feat = view_of(P, n.out);
    const ActView G = view_of(P, n.g2);
    RUN(kHeadBwd, head_bwd_launch(feat, (const float*)P->state[P->head_state0 + 2 * s], K, dout + (long long)s * K * HW, (long long)S * K * HW,
                        grad_scale, G, fptr(P, P->head_part), (float*)P->grads[P->head_state0 + 2 * s],
                        (float*)P->grads[P->head_state0 + 2 * s + 1], accumulate, st));
    ++P->launches;
    if (s < (int)P->final_keep.size() && P->final_keep[s])
      RUN(kOther, mask_mul_launch(G, P->final_keep[s], p8(f), P->final_scale, st));
  }
  // packed [9][cout][cin] -> OIHW .grad of every layer of a stage in one launch, at the end of the stage
  auto flush_unpack = [&]() -> int {
    if (P->unpack_jobs.empty()) return MIMO_OK;
    RUN(kWgradUnpack, wgrad_unpack_batched_launch(P->unpack_jobs.data(), (int)P->unpack_jobs.size(), 1.f, accumulate, st));
    P->unpack_jobs.clear();
    return MIMO_OK;
  };
  // The four stages (decoders -> core up path -> core down path -> encoders) are four fixed launch sequences over the
  // workspace: each one is replayed from its own CUDA graph; the caller's stage events are recorded eagerly in between
  // (ordinary stream semantics for the overlapped gradient all-reduce).
  auto stage0 = [&]() -> int {
    int rc;
    MIMO_CUDA(cudaMemsetAsync(P->ws + P->dwp_begin, 0, P->dwp_end - P->dwp_begin, st));
    P->unpack_jobs.clear();
    for (int s = 0; s < S; ++s)
      if ((rc = node_backward(P, P->dec[s], tr, mask(P->dec[s]), accumulate, st))) return rc;
    // gradient of the shared up-sampled core output: fold the [f, f + c/2) slice of every decoder into ONE buffer (the bilinear
    // backward is linear, so it runs once on the sum instead of once per subnetwork); two decoders per launch
    for (int s = 0; s < S; s += 2) {
      const ActView dp = view_of(P, P->nodes[P->dec[s]].c1.dpad, f, c / 2);
      const ActView t0 = view_of(P, P->tmp0, 0, c / 2);
      if (s + 1 < S) {
        const ActView dp2 = view_of(P, P->nodes[P->dec[s + 1]].c1.dpad, f, c / 2);
        RUN(kGradGather, grad_gather_launch(&dp, nullptr, nullptr, t0, s > 0 ? 1 : 0, st, &dp2));
      } else {
        RUN(kGradGather, grad_gather_launch(&dp, nullptr, nullptr, t0, s > 0 ? 1 : 0, st));
      }
    }
    RUN(kUpsampleBwd, upsample_bwd_launch(view_of(P, P->tmp0, 0, c / 2), view_of(P, P->g_u3, 0, c / 2), 0, st));
    return flush_unpack();
  };
  auto stage1 = [&]() -> int {
    int rc;
    P->unpack_jobs.clear();
    if ((rc = node_backward(P, P->up3, tr, mask(P->up3), accumulate, st))) return rc;
    {
      const ActView dp = view_of(P, P->nodes[P->up3].c1.dpad, P->stack_c, c);   // the up-sampled part follows the (aligned) stack
      const ActView t = view_of(P, P->tmp1, 0, c);
      RUN(kGradGather, grad_gather_launch(&dp, nullptr, nullptr, t, 0, st));
      RUN(kUpsampleBwd, upsample_bwd_launch(t, view_of(P, P->g_u2, 0, c), 0, st));
    }
    if ((rc = node_backward(P, P->up2, tr, mask(P->up2), accumulate, st))) return rc;
    {
      const ActView dp = view_of(P, P->nodes[P->up2].c1.dpad, 2 * c, 2 * c);
      const ActView t = view_of(P, P->tmp2, 0, 2 * c);
      RUN(kGradGather, grad_gather_launch(&dp, nullptr, nullptr, t, 0, st));
      RUN(kUpsampleBwd, upsample_bwd_launch(t, view_of(P, P->g_u1, 0, 2 * c), 0, st));
    }
    if ((rc = node_backward(P, P->up1, tr, mask(P->up1), accumulate, st))) return rc;
    {
      const ActView dp = view_of(P, P->nodes[P->up1].c1.dpad, 4 * c, 4 * c);
      const ActView t = view_of(P, P->tmp3, 0, 4 * c);
      RUN(kGradGather, grad_gather_launch(&dp, nullptr, nullptr, t, 0, st));
      RUN(kUpsampleBwd, upsample_bwd_launch(t, view_of(P, P->g_x5, 0, 4 * c), 0, st));
    }
    return flush_unpack();
  };
  // core down path: skip gradient (fold of the concat slice) + max-pool backward
  auto stage2 = [&]() -> int {
    int rc;
    P->unpack_jobs.clear();
    if (P->center_keep) RUN(kOther, mask_mul_launch(view_of(P, P->g_x5, 0, 4 * c), P->center_keep, p8(4 * c), P->center_scale, st));
    if ((rc = node_backward(P, P->down4, tr, mask(P->down4), accumulate, st))) return rc;
    {
      const ActView dpp = view_of(P, P->nodes[P->down4].c1.dpad, 0, 4 * c);
      const ActView gp = view_of(P, P->gp_x4, 0, 4 * c);
      RUN(kGradGather, grad_gather_launch(&dpp, nullptr, nullptr, gp, 0, st));
      const ActView dskip = view_of(P, P->nodes[P->up1].c1.dpad, 0, 4 * c);
      const ActView act = view_of(P, P->cat1, 0, 4 * c);
      RUN(kGradGather, grad_gather_launch(&dskip, &gp, &act, view_of(P, P->g_x4, 0, 4 * c), 0, st));
    }
    if ((rc = node_backward(P, P->down3, tr, mask(P->down3), accumulate, st))) return rc;
    {
      const ActView dpp = view_of(P, P->nodes[P->down3].c1.dpad, 0, 2 * c);
      const ActView gp = view_of(P, P->gp_x3, 0, 2 * c);
      RUN(kGradGather, grad_gather_launch(&dpp, nullptr, nullptr, gp, 0, st));
      const ActView dskip = view_of(P, P->nodes[P->up2].c1.dpad, 0, 2 * c);
      const ActView act = view_of(P, P->cat2, 0, 2 * c);
      RUN(kGradGather, grad_gather_launch(&dskip, &gp, &act, view_of(P, P->g_x3, 0, 2 * c), 0, st));
    }
    if ((rc = node_backward(P, P->down2, tr, mask(P->down2), accumulate, st))) return rc;
    {
      // the whole physical stack (gap channels carry zeros on every operand: zero weight rows, zero-initialised buffers)
      const int cs = P->stack_c;
      const ActView dpp = view_of(P, P->nodes[P->down2].c1.dpad, 0, cs);
      const ActView gp = view_of(P, P->gp_xc, 0, cs);
      RUN(kGradGather, grad_gather_launch(&dpp, nullptr, nullptr, gp, 0, st));
      const ActView dskip = view_of(P, P->nodes[P->up3].c1.dpad, 0, cs);
      const ActView act = view_of(P, P->cat3, 0, cs);
      RUN(kGradGather, grad_gather_launch(&dskip, &gp, &act, view_of(P, P->g_xc, 0, cs), 0, st));
    }
    return flush_unpack();
  };
  auto stage3 = [&]() -> int {
    int rc;
    P->unpack_jobs.clear();
    for (int s = 0; s < S; ++s) {
      if ((rc = node_backward(P, P->enc_down[s], tr, mask(P->enc_down[s]), accumulate, st))) return rc;
      const ActView dpp = view_of(P, P->nodes[P->enc_down[s]].c1.dpad, 0, f);
      const ActView gp = view_of(P, P->gp1[s], 0, f);
      RUN(kGradGather, grad_gather_launch(&dpp, nullptr, nullptr, gp, 0, st));
      const ActView dskip = view_of(P, P->nodes[P->dec[s]].c1.dpad, 0, f);
      const ActView act = P->last_dense_skip ? view_of(P, P->x1d[s], 0, f) : view_of(P, P->dcat[s], 0, f);
      RUN(kGradGather, grad_gather_launch(&dskip, &gp, &act, view_of(P, P->g_x1[s], 0, f), 0, st));
      if ((rc = node_backward(P, P->enc_in[s], tr, mask(P->enc_in[s]), accumulate, st))) return rc;
      if (dx) {
        const ActView dp = view_of(P, P->nodes[P->enc_in[s]].c1.dpad, 0, Cin);
        const long long total = (long long)cfg.batch * HW;
        int grid = (int)((total + 255) / 256);
        if (grid > num_sms() * 16) grid = num_sms() * 16;
        prof_begin(P, kOther, st);
        unpack_input_grad_kernel<<<grid, 256, 0, st>>>(dp, cfg.height, cfg.width, Cin, dx + (long long)s * Cin * HW, (long long)S * Cin * HW, HW);
        prof_end(P, st);
        MIMO_LAUNCH_CHECK();
        ++P->launches;
      }
    }
    return flush_unpack();
  };
  ++P->bwd_calls;
  bool has_mask = false;
  for (const float* m : P->masks_copy) has_mask = has_mask || (m != nullptr);
  has_mask = has_mask || P->center_keep != nullptr;
  for (const bf16* m : P->final_keep) has_mask = has_mask || (m != nullptr);
  const bool allow = P->graph_mode != 0 && !P->graph_failed && !P->prof && !has_mask && dx == nullptr && P->bwd_calls > 2;
  const unsigned long long key = (tr ? 1ull : 0ull) | (accumulate ? 2ull : 0ull) | (P->last_dense_skip ? 4ull : 0ull);
  if ((rc = run_graphed(P, P->g_bwd[0], key, allow, st, stage0))) return rc;
  if (P->stage_ev[0]) MIMO_CUDA(cudaEventRecord(P->stage_ev[0], st));
  if ((rc = run_graphed(P, P->g_bwd[1], key, allow, st, stage1))) return rc;
  if (P->stage_ev[1]) MIMO_CUDA(cudaEventRecord(P->stage_ev[1], st));
  if ((rc = run_graphed(P, P->g_bwd[2], key, allow, st, stage2))) return rc;
  if (P->stage_ev[2]) MIMO_CUDA(cudaEventRecord(P->stage_ev[2], st));
  if ((rc = run_graphed(P, P->g_bwd[3], key, allow, st, stage3))) return rc;
  if (P->stage_ev[3]) MIMO_CUDA(cudaEventRecord(P->stage_ev[3], st));
  return MIMO_OK;
}

int mimo_unet_set_elementwise_dropout(mimo_unet_plan_t* P, const void* center_keep, float center_scale, const void* const* final_keep,
                                      float final_scale) {
  MIMO_CHECK(P, MIMO_ERR_ARG, "set_elementwise_dropout: null plan");
  P->center_keep = (const bf16*)center_keep;
  P->center_scale = center_scale;
  P->final_keep.assign(P->cfg.num_subnetworks, nullptr);
  if (final_keep)
    for (int s = 0; s < P->cfg.num_subnetworks; ++s) P->final_keep[s] = (const bf16*)final_keep[s];
  P->final_scale = final_scale;
  return MIMO_OK;
}

int mimo_unet_set_inference_fusion(mimo_unet_plan_t* P, int on) {
  MIMO_CHECK(P, MIMO_ERR_ARG, "set_inference_fusion: null plan");
  P->fuse_next = on != 0;
  return MIMO_OK;
}

int mimo_unet_set_backward_events(mimo_unet_plan_t* P, void* const* events) {
  MIMO_CHECK(P, MIMO_ERR_ARG, "set_backward_events: null plan");
  for (int i = 0; i < 4; ++i) P->stage_ev[i] = events ? (cudaEvent_t)events[i] : nullptr;
  return MIMO_OK;
}

int mimo_unet_backward_stage_first_state(const mimo_unet_plan_t* P, int stage) {
  if (!P || stage < 0 || stage > 3) return -1;
  // state order == flat order: encoders | core.down2..4 | core.up1..3 | decoders + heads
  if (stage == 3) return 0;
  if (stage == 2) return P->nodes[P->down2].c1.state0;
  if (stage == 1) return P->nodes[P->up1].c1.state0;
  return P->nodes[P->dec[0]].c1.state0;
}

int mimo_unet_stack_layout(const mimo_unet_plan_t* P, int* slice_len, int* slice_stride, int* n_slices) {
  MIMO_CHECK(P && slice_len && slice_stride && n_slices, MIMO_ERR_ARG, "stack_layout: null argument");
  const ChannelSlices& sl = P->nodes[P->down2].c1.sl;
  *slice_len = sl.sl_len; *slice_stride = sl.sl_stride; *n_slices = sl.sl_n;
  return MIMO_OK;
}

int mimo_unet_debug_view(const mimo_unet_plan_t* P, const char* name, mimo_act_t* view, int* kind) {
  MIMO_CHECK(P && name && view && kind, MIMO_ERR_ARG, "debug_view: null argument");
  MIMO_CHECK(P->bound, MIMO_ERR_STATE, "debug_view: plan is not bound");
  const std::string full(name);
  const size_t dot = full.rfind('.');
  MIMO_CHECK(dot != std::string::npos, MIMO_ERR_ARG, "debug_view: bad name %s", name);
  std::string node = full.substr(0, dot), what = full.substr(dot + 1);
  const ConvL* cl = nullptr;
  // "<node>.c1.y" style names carry one more component
  if (what == "y" || what == "dy" || what == "dpad" || what == "scale" || what == "shift" || what == "mean" || what == "invstd") {
    const size_t dot2 = node.rfind('.');
    MIMO_CHECK(dot2 != std::string::npos, MIMO_ERR_ARG, "debug_view: bad name %s", name);
    const std::string which = node.substr(dot2 + 1);
    node = node.substr(0, dot2);
    auto it = P->node_by_name.find(node);
    MIMO_CHECK(it != P->node_by_name.end(), MIMO_ERR_ARG, "debug_view: unknown node %s", node.c_str());
    cl = which == "c1" ? &P->nodes[it->second].c1 : &P->nodes[it->second].c2;
  }
  auto fill = [&](const ActView& v) {
    view->ptr = v.base; view->n = v.N; view->h = v.H; view->w = v.W; view->pad = v.pad; view->cpitch = v.cpitch; view->c_off = v.c_off; view->c = v.C;
  };
  *kind = 0;
  if (cl) {
    if (what == "y") fill(view_of(P, cl->y, 0, cl->cout));
    else if (what == "dy") fill(view_of(P, cl->dy, 0, cl->cout));
    else if (what == "dpad") fill(view_of(P, cl->dpad, 0, cl->cin_x));
    else {
      const int k = what == "scale" ? 0 : what == "shift" ? 1 : what == "mean" ? 2 : 3;
      *kind = 1;
      view->ptr = (void*)(fptr(P, cl->vec) + k * cl->cout_p);
      view->n = view->h = view->w = 1; view->pad = 0; view->cpitch = cl->cout_p; view->c_off = 0; view->c = cl->cout;
    }
    return MIMO_OK;
  }
  auto it = P->node_by_name.find(node);
  MIMO_CHECK(it != P->node_by_name.end(), MIMO_ERR_ARG, "debug_view: unknown node %s", node.c_str());
  const Node& n = P->nodes[it->second];
  if (what == "in") fill(view_of(P, n.in));
  else if (what == "a1") fill(view_of(P, n.a1, 0, n.c1.cout));
  else if (what == "out") fill(view_of(P, n.out));
  else if (what == "pool") { MIMO_CHECK(n.pool.buf >= 0, MIMO_ERR_ARG, "debug_view: node has no pooled output"); fill(view_of(P, n.pool)); }
  else if (what == "g1") fill(view_of(P, n.g1, 0, n.c1.cout));
  else if (what == "g2") fill(view_of(P, n.g2));
  else { set_error("debug_view: unknown buffer %s", what.c_str()); return MIMO_ERR_ARG; }
  return MIMO_OK;
}

}  // extern "C"
