// Whole-model plan: ResUNet30.forward (eval, 1 channel in / out) as a fixed list of kernel launches over a
// caller-provided workspace.  Reference: models/resunet.py:522-595 (ResUNet30_Base.forward) and :640-653.
//
// Data layout in HBM (all inside the workspace; `P_k` = pixels of level k, level 0 = T' x F' = padded frames x
// n_fft/2 bins, levels 1..5 halve both axes, level 6 halves frequency once more):
//   mag / cos / sin       (B, T, F) fp32            K1 output, K5 input
//   shift                 (B, 8256) fp32            FiLM + folded-BN activation shifts (K2 output)
//   x_raw[k] / x_act[k]   (B, H_k, W_k, cin_k)      encoder block input: raw fp16 (shortcut operand) / activated bf16
//   a2[k]                 (B, H_k, W_k, cout_k)     activated output of a block's first conv (bf16), shared by the
//                                                  encoder and decoder block of the same level
//   cat_raw[k]/cat_act[k] (B, H_k, W_k, 2*cout_k)   decoder concat buffer: [upsampled | encoder skip] (fp16 / bf16)
//   d_act[k]              (B, H_k, W_k, c)          activated decoder-block output = input of the next transposed conv
//   feat                  (B, 3, T', F') fp32       after_conv output (mask features), planar
#include <string.h>

#include <new>
#include <vector>

#include "conv.cuh"
#include "lass_internal.cuh"

namespace lass {

cudaError_t launch_film(const float* cond, const float* W, const float* bias, float* shift, int B, int K, int J,
                        cudaStream_t stream);

namespace {

const int kEncCin[7] = {32, 32, 64, 128, 256, 384, 384};
const int kEncCout[7] = {32, 64, 128, 256, 384, 384, 384};
const int kDecCin[6] = {384, 384, 384, 256, 128, 64};
const int kDecCout[6] = {384, 384, 256, 128, 64, 32};

struct Sites {
  int off[32];
  int rows;
  Sites() {
    int o = 0;
    for (int k = 0; k < 7; ++k) {
      off[2 * k] = o;
      o += kEncCin[k];
      off[2 * k + 1] = o;
      o += kEncCout[k];
    }
    for (int j = 0; j < 6; ++j) {
      off[14 + 3 * j] = o;
      o += kDecCin[j];
      off[14 + 3 * j + 1] = o;
      o += 2 * kDecCout[j];
      off[14 + 3 * j + 2] = o;
      o += kDecCout[j];
    }
    rows = o;
  }
};
const Sites& sites() {
  static Sites s;
  return s;
}

struct Buf {
  char* ptr = nullptr;
  int dims[4] = {0, 0, 0, 0};
  int elem = 2;
  size_t bytes() const { return (size_t)dims[0] * dims[1] * dims[2] * dims[3] * elem; }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace lass

using namespace lass;

struct lass_plan {
  lass_resunet30_weights w;
  int B, L, T, F, Tp, Fp;
  int H[7], W[7];
  char* ws;
  size_t ws_bytes;
  void* stft_ws;
  Buf mag, cosb, sinb, shift, feat;
  Buf x_raw[7], x_act[7], a2[7], cat_raw[6], cat_act[6], d_act[7];
  std::vector<ConvPrepared*> convs;
  std::vector<double> conv_flops;
  double conv_flops_total = 0;
};

namespace {

struct Geometry {
  int T, F, Tp, Fp, H[7], W[7];
};

int make_geometry(int L, int n_fft, int hop, Geometry* g) {
  if (n_fft < 256 || (n_fft & (n_fft - 1)) || hop <= 0 || hop % 8 || L <= n_fft / 2) return LASS_ERR_ARG;
  g->T = L / hop + 1;
  g->F = n_fft / 2 + 1;
  g->Tp = (g->T + 31) / 32 * 32;
  g->Fp = n_fft / 2;
  g->H[0] = g->Tp;
  g->W[0] = g->Fp;
  for (int k = 1; k <= 5; ++k) {
    g->H[k] = g->H[k - 1] / 2;
    g->W[k] = g->W[k - 1] / 2;
  }
  g->H[6] = g->H[5];
  g->W[6] = g->W[5] / 2;
  return 0;
}

// Lays the buffers out in the workspace; with ws == nullptr only computes the size.
size_t layout(lass_plan* p, int B, const Geometry& g, int n_fft, int hop, int L, char* ws) {
  size_t off = 0;
  auto place = [&](Buf* b, int d0, int d1, int d2, int d3, int elem) {
    off = align_up(off, 1024);
    if (b) {
      b->ptr = ws ? ws + off : nullptr;
      b->dims[0] = d0;
      b->dims[1] = d1;
      b->dims[2] = d2;
      b->dims[3] = d3;
      b->elem = elem;
    }
    off += (size_t)d0 * d1 * d2 * d3 * elem;
  };
  off = align_up(off, 1024);
  if (p) p->stft_ws = ws ? ws + off : nullptr;
  off += stft_workspace_bytes(B, L, n_fft, hop);
  place(p ? &p->mag : nullptr, B, 1, g.T, g.F, 4);
  place(p ? &p->cosb : nullptr, B, 1, g.T, g.F, 4);
  place(p ? &p->sinb : nullptr, B, 1, g.T, g.F, 4);
  place(p ? &p->shift : nullptr, B, 1, 1, sites().rows, 4);
  place(p ? &p->feat : nullptr, B, 3, g.Tp, g.Fp, 4);
  for (int k = 0; k < 7; ++k) {
    if (k > 0) {   // x_raw[0] / x_act[0] are never materialised (regenerated from the magnitude inside encoder_block1's convs)
      place(p ? &p->x_raw[k] : nullptr, B, g.H[k], g.W[k], kEncCin[k], 2);
      place(p ? &p->x_act[k] : nullptr, B, g.H[k], g.W[k], kEncCin[k], 2);
    }
    place(p ? &p->a2[k] : nullptr, B, g.H[k], g.W[k], kEncCout[k], 2);
  }
  for (int k = 0; k < 6; ++k) {
    place(p ? &p->cat_raw[k] : nullptr, B, g.H[k], g.W[k], 2 * kEncCout[k], 2);
    place(p ? &p->cat_act[k] : nullptr, B, g.H[k], g.W[k], 2 * kEncCout[k], 2);
  }
  // d_act[k]: activated input of the transposed conv that produces level k-1... indexed by the level it lives on
  // d_act[6] = conv_block7a output (384 ch); d_act[k] (k = 5..1) = decoder block output on level k
  place(p ? &p->d_act[6] : nullptr, B, g.H[6], g.W[6], 384, 2);
  for (int j = 0; j < 5; ++j) {
    const int lvl = 5 - j;
    place(p ? &p->d_act[lvl] : nullptr, B, g.H[lvl], g.W[lvl], kDecCout[j], 2);
  }
  return align_up(off, 1024);
}

ConvOut out_spec(const Buf& b, int coff, bool fp16, const lass_plan* p, int site, int site_coff) {
  ConvOut o;
  memset(&o, 0, sizeof(o));
  o.ptr = b.ptr;
  o.cstride = b.dims[3];
  o.coff = coff;
  o.fp16 = fp16 ? 1 : 0;
  if (site >= 0) {
    const int row = sites().off[site] + site_coff;
    o.scale = p->w.act_scale + row;
    o.shift = reinterpret_cast<const float*>(p->shift.ptr) + row;
    o.shift_bstride = sites().rows;
  }
  return o;
}

ConvSegment seg_spec(const Buf& b, int cin, int taps, bool fp16, const void* weights) {
  ConvSegment s;
  memset(&s, 0, sizeof(s));
  s.src = b.ptr;
  s.src_cstride = b.dims[3];
  s.src_coff = 0;
  s.cin = cin;
  s.kc = (cin % 64 == 0) ? 64 : 32;
  s.taps = taps;
  s.fp16 = fp16 ? 1 : 0;
  s.weights = weights;
  return s;
}

ConvLaunch base_launch(const lass_plan* p, int lvl, int ncols) {
  ConvLaunch l;
  memset(&l, 0, sizeof(l));
  l.B = p->B;
  l.H = p->H[lvl];
  l.W = p->W[lvl];
  l.ncols = ncols;
  l.up_h = l.up_w = 1;
  l.group_c = ncols;
  l.pool_h = l.pool_w = 1;
  return l;
}

int add_conv(lass_plan* p, const ConvLaunch& l) {
  ConvPrepared* cp = nullptr;
  int e = conv_prepare(l, &cp);
  if (e) return e;
  p->convs.push_back(cp);
  p->conv_flops.push_back(conv_flops(l));
  p->conv_flops_total += conv_flops(l);
  return 0;
}

int build_launches(lass_plan* p) {
  int e;
  // ---------------- encoder ----------------
  for (int k = 0; k < 7; ++k) {
    const int cin = kEncCin[k], cout = kEncCout[k];
    if (!p->w.enc[k].conv1_w || !p->w.enc[k].conv2_w || (k > 0 && !p->w.enc[k].sc_w))
      return set_error(LASS_ERR_ARG, "plan: encoder block %d weights missing", k);
    {  // conv1: act(x) -> a2 = lrelu(bn2(.) + beta2)
      ConvLaunch l = base_launch(p, k, cout);
      l.nseg = 1;
      l.seg[0] = seg_spec(p->x_act[k], cin, 9, false, p->w.enc[k].conv1_w);
      if (k == 0) {
        // encoder_block1's input is lrelu(bn1(pre_conv(bn0(mag))) + beta1), a function of ONE scalar per pixel: the conv's
        // producer warps generate the activated 32-channel operand tile in shared memory from the magnitude, so neither
        // x_act[0] (33.5 MB per clip written and read back) nor a pre_conv kernel exists
        l.seg[0].src = nullptr;
        l.seg[0].src_cstride = cin;
        l.gen_src = reinterpret_cast<const float*>(p->mag.ptr);
        l.gen_in_scale = p->w.bn0_scale;
        l.gen_in_shift = p->w.bn0_shift;
        l.gen_w = p->w.pre_w;
        l.gen_b = p->w.pre_b;
        l.gen_scale = p->w.act_scale + sites().off[0];
        l.gen_shift = reinterpret_cast<const float*>(p->shift.ptr) + sites().off[0];
        l.gen_shift_bstride = sites().rows;
        l.gen_T = p->T;
        l.gen_F = p->F;
      }
      l.full_act = out_spec(p->a2[k], 0, false, p, 2 * k + 1, 0);
      if ((e = add_conv(p, l))) return e;
    }
    {  // conv2 + shortcut(raw x): block output
      ConvLaunch l = base_launch(p, k, cout);
      l.seg[0] = seg_spec(p->a2[k], cout, 9, false, p->w.enc[k].conv2_w);
      if (k == 0) {
        // encoder_block1 has no shortcut conv and its input is pre_conv(bn0(mag)): the identity residual is regenerated
        // in the epilogue from the 1-channel magnitude (exact fp32), so x_raw[0] is never written or read
        l.nseg = 1;
        l.resid_src = reinterpret_cast<const float*>(p->mag.ptr);
        l.resid_in_scale = p->w.bn0_scale;
        l.resid_in_shift = p->w.bn0_shift;
        l.resid_w = p->w.pre_w;
        l.resid_b = p->w.pre_b;
        l.resid_T = p->T;
        l.resid_F = p->F;
      } else {
        l.nseg = 2;
        l.seg[1] = seg_spec(p->x_raw[k], cin, 1, true, p->w.enc[k].sc_w);
        l.bias = p->w.enc[k].sc_b;
      }
      if (k < 6) {
        const int j = 5 - k;  // decoder block that consumes this skip (its output lives on level k)
        l.full_raw = out_spec(p->cat_raw[k], cout, true, p, -1, 0);
        l.full_act = out_spec(p->cat_act[k], cout, false, p, 14 + 3 * j + 1, cout);
        l.pool_h = (k < 5) ? 2 : 1;
        l.pool_w = 2;
        l.pool_raw = out_spec(p->x_raw[k + 1], 0, true, p, -1, 0);
        l.pool_act = out_spec(p->x_act[k + 1], 0, false, p, 2 * (k + 1), 0);
      } else {
        // conv_block7a: pooling (1,1) is the identity and its `encoder` output is unused (models/resunet.py:562)
        l.full_act = out_spec(p->d_act[6], 0, false, p, 14, 0);
      }
      if ((e = add_conv(p, l))) return e;
    }
  }
  // ---------------- decoder ----------------
  for (int j = 0; j < 6; ++j) {
    const int cin = kDecCin[j], cout = kDecCout[j];
    const int lin = 6 - j, lo = 5 - j;
    const int uh = (j == 0) ? 1 : 2, uw = 2;
    if (!p->w.dec[j].up_w || !p->w.dec[j].conv1_w || !p->w.dec[j].conv2_w || !p->w.dec[j].sc_w)
      return set_error(LASS_ERR_ARG, "plan: decoder block %d weights missing", j);
    {  // transposed conv -> first half of the concat buffers
      ConvLaunch l = base_launch(p, lin, uh * uw * cout);
      l.nseg = 1;
      l.seg[0] = seg_spec(p->d_act[lin], cin, 1, false, p->w.dec[j].up_w);
      l.up_h = uh;
      l.up_w = uw;
      l.group_c = cout;
      l.full_raw = out_spec(p->cat_raw[lo], 0, true, p, -1, 0);
      l.full_act = out_spec(p->cat_act[lo], 0, false, p, 14 + 3 * j + 1, 0);
      if ((e = add_conv(p, l))) return e;
    }
    {  // conv_block2.conv1 over the concat
      ConvLaunch l = base_launch(p, lo, cout);
      l.nseg = 1;
      l.seg[0] = seg_spec(p->cat_act[lo], 2 * cout, 9, false, p->w.dec[j].conv1_w);
      l.full_act = out_spec(p->a2[lo], 0, false, p, 14 + 3 * j + 2, 0);
      if ((e = add_conv(p, l))) return e;
    }
    {  // conv_block2.conv2 + shortcut(raw concat)
      ConvLaunch l = base_launch(p, lo, cout);
      l.nseg = 2;
      l.seg[0] = seg_spec(p->a2[lo], cout, 9, false, p->w.dec[j].conv2_w);
      l.seg[1] = seg_spec(p->cat_raw[lo], 2 * cout, 1, true, p->w.dec[j].sc_w);
      l.bias = p->w.dec[j].sc_b;
      if (j < 5) {
        l.full_act = out_spec(p->d_act[lo], 0, false, p, 14 + 3 * (j + 1), 0);
      } else {
        l.after_w = p->w.after_w;
        l.after_b = p->w.after_b;
        l.feat = reinterpret_cast<float*>(p->feat.ptr);
      }
      if ((e = add_conv(p, l))) return e;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int lass_film(const float* condition, const float* film_w, const float* film_b, int B, int condition_size, int J,
              float* shift_out, void* stream) {
  if (!condition || !film_w || !film_b || !shift_out) return set_error(LASS_ERR_ARG, "lass_film: null pointer");
  if (B <= 0 || J <= 0 || condition_size <= 0 || condition_size > 1024)
    return set_error(LASS_ERR_ARG, "lass_film: bad shape B=%d K=%d J=%d", B, condition_size, J);
  return set_cuda_error(launch_film(condition, film_w, film_b, shift_out, B, condition_size, J, (cudaStream_t)stream),
                        "film launch");
}

int lass_resunet30_film_rows(void) { return sites().rows; }

int lass_resunet30_film_offset(int site) {
  if (site < 0 || site >= 32) return -1;
  return sites().off[site];
}

size_t lass_resunet30_workspace_bytes(int B, int L, int n_fft, int hop) {
  Geometry g;
  if (B <= 0 || make_geometry(L, n_fft, hop, &g)) return 0;
  return layout(nullptr, B, g, n_fft, hop, L, nullptr);
}

int lass_resunet30_plan_create(const lass_resunet30_weights* wh, int B, int L, void* workspace,
                               size_t workspace_bytes, lass_plan** plan_out) {
  if (!wh || !workspace || !plan_out) return set_error(LASS_ERR_ARG, "plan_create: null pointer");
  *plan_out = nullptr;
  Geometry g;
  if (B <= 0 || make_geometry(L, wh->n_fft, wh->hop, &g))
    return set_error(LASS_ERR_ARG, "plan_create: bad geometry B=%d L=%d n_fft=%d hop=%d", B, L, wh->n_fft, wh->hop);
  if (wh->film_rows != sites().rows)
    return set_error(LASS_ERR_ARG, "plan_create: film_rows %d != %d", wh->film_rows, sites().rows);
  if (wh->condition_size <= 0 || wh->condition_size > 1024) return set_error(LASS_ERR_ARG, "plan_create: condition_size");
  if (!wh->stft_basis_hi || !wh->stft_basis_lo || !wh->istft_window || !wh->istft_twiddle || !wh->bn0_scale ||
      !wh->bn0_shift || !wh->pre_w || !wh->pre_b || !wh->film_w || !wh->film_b || !wh->act_scale || !wh->after_w ||
      !wh->after_b)
    return set_error(LASS_ERR_ARG, "plan_create: missing weight pointer");
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return set_error(LASS_ERR_ARG, "plan_create: workspace not 1 KiB aligned");
  const size_t need = layout(nullptr, B, g, wh->n_fft, wh->hop, L, nullptr);
  if (workspace_bytes < need) return set_error(LASS_ERR_WORKSPACE, "plan_create: workspace %zu < %zu", workspace_bytes, need);
  lass_plan* p = new (std::nothrow) lass_plan();
  if (!p) return set_error(LASS_ERR_ARG, "plan_create: out of host memory");
  p->w = *wh;
  p->B = B;
  p->L = L;
  p->T = g.T;
  p->F = g.F;
  p->Tp = g.Tp;
  p->Fp = g.Fp;
  for (int k = 0; k < 7; ++k) {
    p->H[k] = g.H[k];
    p->W[k] = g.W[k];
  }
  p->ws = reinterpret_cast<char*>(workspace);
  p->ws_bytes = workspace_bytes;
  layout(p, B, g, wh->n_fft, wh->hop, L, p->ws);
  int e = build_launches(p);
  if (e) {
    lass_resunet30_plan_destroy(p);
    return e;
  }
  *plan_out = p;
  return 0;
}

int lass_resunet30_forward(lass_plan* p, const float* mixture, const float* condition, const float* shift_override,
                           float* waveform, int stft_precision_mode, void* stream_v) {
  return lass_resunet30_forward_stages(p, LASS_STAGE_ALL, mixture, condition, shift_override, waveform,
                                       stft_precision_mode, stream_v);
}

double lass_resunet30_unet_flops(const lass_plan* p) { return p ? p->conv_flops_total : 0.0; }

int lass_resunet30_forward_stages(lass_plan* p, int stage_mask, const float* mixture, const float* condition,
                                  const float* shift_override, float* waveform, int stft_precision_mode,
                                  void* stream_v) {
  if (!p) return set_error(LASS_ERR_ARG, "forward: null plan");
  if ((stage_mask & LASS_STAGE_FRONT) && (!mixture || (!condition && !shift_override)))
    return set_error(LASS_ERR_ARG, "forward: null input pointer");
  if ((stage_mask & LASS_STAGE_BACK) && !waveform) return set_error(LASS_ERR_ARG, "forward: null output pointer");
  cudaStream_t stream = (cudaStream_t)stream_v;
  int e;
  float* mag = reinterpret_cast<float*>(p->mag.ptr);
  float* cs = reinterpret_cast<float*>(p->cosb.ptr);
  float* sn = reinterpret_cast<float*>(p->sinb.ptr);
  float* shift = reinterpret_cast<float*>(p->shift.ptr);
  if (stage_mask & LASS_STAGE_FRONT) {
  // K1: STFT -> mag / cos / sin
  if ((e = launch_stft(mixture, p->B, p->L, p->w.n_fft, p->w.hop, p->w.stft_basis_hi, p->w.stft_basis_lo, mag, cs, sn,
                       stft_precision_mode, 0, p->stft_ws, stream)))
    return e;
  // K2: FiLM + folded BN shifts (or the caller's precomputed table)
  if (shift_override) {
    if ((e = set_cuda_error(cudaMemcpyAsync(shift, shift_override, (size_t)p->B * sites().rows * sizeof(float),
                                            cudaMemcpyDeviceToDevice, stream),
                            "shift table copy")))
      return e;
  } else if ((e = set_cuda_error(launch_film(condition, p->w.film_w, p->w.film_b, shift, p->B, p->w.condition_size,
                                             sites().rows, stream),
                                 "film launch")))
    return e;
  }
  // K3 / K4: the UNet
  if (stage_mask & LASS_STAGE_UNET)
    for (size_t i = 0; i < p->convs.size(); ++i)
      if ((e = conv_run(p->convs[i], stream))) return e;
  if (!(stage_mask & LASS_STAGE_BACK)) return 0;
  // K5: mask + iSTFT
  const long long plane = (long long)p->Tp * p->Fp;
  return set_cuda_error(launch_mask_istft(reinterpret_cast<const float*>(p->feat.ptr), 3 * plane, plane, p->Fp, p->Fp, mag,
                                          cs, sn, p->w.istft_window, p->w.istft_twiddle, waveform, p->B, p->T, p->F,
                                          p->w.n_fft, p->w.hop, p->L, stream),
                        "mask_istft launch");
}

int lass_resunet30_num_launches(const lass_plan* p) { return p ? (int)p->convs.size() + 4 : 0; }

int lass_debug_time_unet_launches(lass_plan* p, float* ms_out, double* flops_out, int capacity, void* stream_v) {
  if (!p || !ms_out) return set_error(LASS_ERR_ARG, "time_unet_launches: null pointer");
  const int n = (int)p->convs.size();
  if (capacity < n) return set_error(LASS_ERR_ARG, "time_unet_launches: capacity %d < %d launches", capacity, n);
  cudaStream_t stream = (cudaStream_t)stream_v;
  std::vector<cudaEvent_t> ev(n + 1);
  for (int i = 0; i <= n; ++i)
    if (cudaEventCreate(&ev[i]) != cudaSuccess) return set_cuda_error(cudaGetLastError(), "event create");
  int e = 0;
  cudaEventRecord(ev[0], stream);
  for (int i = 0; i < n && !e; ++i) {
    e = conv_run(p->convs[i], stream);
    cudaEventRecord(ev[i + 1], stream);
  }
  cudaError_t ce = cudaStreamSynchronize(stream);
  for (int i = 0; i < n; ++i) {
    ms_out[i] = 0.0f;
    if (!e && ce == cudaSuccess) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
    if (flops_out) flops_out[i] = p->conv_flops[i];
  }
  for (int i = 0; i <= n; ++i) cudaEventDestroy(ev[i]);
  if (e) return e;
  if (ce != cudaSuccess) return set_cuda_error(ce, "time_unet_launches sync");
  return n;
}

void* lass_resunet30_buffer(const lass_plan* p, const char* name, int dims[4], int* elem_bytes) {
  if (!p || !name) return nullptr;
  const Buf* b = nullptr;
  int idx = -1;
  const size_t n = strlen(name);
  if (n > 0 && name[n - 1] >= '0' && name[n - 1] <= '9') idx = name[n - 1] - '0';
  if (!strcmp(name, "mag")) b = &p->mag;
  else if (!strcmp(name, "cos")) b = &p->cosb;
  else if (!strcmp(name, "sin")) b = &p->sinb;
  else if (!strcmp(name, "shift")) b = &p->shift;
  else if (!strcmp(name, "feat")) b = &p->feat;
  else if (!strncmp(name, "x_raw", 5) && idx >= 0 && idx < 7) b = &p->x_raw[idx];
  else if (!strncmp(name, "x_act", 5) && idx >= 0 && idx < 7) b = &p->x_act[idx];
  else if (!strncmp(name, "a2_", 3) && idx >= 0 && idx < 7) b = &p->a2[idx];
  else if (!strncmp(name, "cat_raw", 7) && idx >= 0 && idx < 6) b = &p->cat_raw[idx];
  else if (!strncmp(name, "cat_act", 7) && idx >= 0 && idx < 6) b = &p->cat_act[idx];
  else if (!strncmp(name, "d_act", 5) && idx >= 1 && idx < 7) b = &p->d_act[idx];
  if (!b || !b->ptr) return nullptr;
  if (dims)
    for (int i = 0; i < 4; ++i) dims[i] = b->dims[i];
  if (elem_bytes) *elem_bytes = b->elem;
  return b->ptr;
}

void lass_resunet30_plan_destroy(lass_plan* p) {
  if (!p) return;
  for (size_t i = 0; i < p->convs.size(); ++i) conv_free(p->convs[i]);
  delete p;
}

}  // extern "C"
