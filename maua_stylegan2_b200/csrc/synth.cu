// Whole-forward C ABI: one call = one Generator.forward of the synthesis network (models/stylegan2.py:537-576) on the
// tensor-core path — style prologue, 2*log2(size)-3 modulated convs with fused epilogues, blur, ToRGB chain, uint8 pack.
//
// The per-operator entry points (maua_modconv_tc, maua_blur_act_nhwc, ...) stay the primitives; this file only sequences
// them, which is what maua_stylegan2_b200/synthesis.py does from Python with ~50 ctypes calls per batch.  A host
// language without a Python interpreter (the reference's native-extension boundary, SURVEY.md §8(b) last cell) binds
//     maua_synth_create -> maua_synth_prepare (once per weight version) -> maua_synth_bind (once per batch size /
//     workspace) -> maua_synth_forward (every batch; capturable into a CUDA graph) -> maua_synth_destroy.
// Nothing here allocates device memory: packed weights live in the caller's `plan` buffer, styles / activations in the
// caller's `workspace` (sizes from maua_synth_plan_bytes / maua_synth_workspace_bytes).
// Not covered (use the per-operator ABI): network bends between layers, returning the activation maps, LatentInput.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

namespace maua {
namespace {

constexpr size_t ALIGN = 1024;
constexpr size_t SPLITK_BYTES = 64u << 20;
inline size_t up_to(size_t v, size_t a = ALIGN) { return (v + a - 1) / a * a; }

struct LayerState {
  MauaSynthLayer l;
  int fmt;               // 3 = bf16 (hi, lo), 2 = fp16 plane, 1 = bf16 hi only  (== n_products of maua_modconv_tc)
  int in_h, in_w, out_h, out_w;
  bool has_rgb, fuse_rgb;
  // plan buffer
  float* wsq;
  void *w_hi, *w_lo;
  // workspace (bound)
  float *s, *d, *s_norm, *rgb_s;
  int job, rgb_job;
};

}  // namespace
}  // namespace maua

struct MauaSynth {
  MauaSynthDesc desc;
  std::vector<maua::LayerState> layers;
  int n_jobs = 0;
  size_t plan_bytes = 0;
  bool prepared = false;
  // binding
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int batch = 0;
  MauaStyleJob* jobs_dev = nullptr;
  float* latent_t = nullptr;
  void* splitk = nullptr;
  uint8_t* arena[2] = {nullptr, nullptr};
  size_t arena_bytes = 0;
  float* image[2] = {nullptr, nullptr};
  // ToRGB branch: the skip-connection chain (ToRGB of layer i + upsampled image of layer i-2) only meets the conv chain
  // again at the last layer, so it runs on a side stream (fork / join with events: capturable, the graph gets a parallel
  // branch) and fills the tails of the persistent conv kernels instead of serialising ~9 latency-bound launches.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_main = nullptr, ev_side = nullptr;
  int side_device = -1;
  bool overlap = true;
};

namespace maua {
namespace {

bool tc_ok(int cin, int cout) { return cin % 32 == 0 && cout % 16 == 0 && cout >= 16; }

// bytes one layer needs in its arena at `batch` (outputs + intermediates)
size_t layer_arena_bytes(const LayerState& ls, const LayerState* nxt, int batch) {
  const size_t px = (size_t)batch * ls.out_h * ls.out_w;
  size_t n = 0;
  if (nxt) n += up_to(px * ls.l.cout * (nxt->fmt == 2 ? 2 : 4));             // consumer operand planes
  if (ls.l.up) n += up_to((size_t)batch * (2 * ls.in_h + 1) * (2 * ls.in_w + 1) * ls.l.cout * 4);  // u
  if (ls.has_rgb && !ls.fuse_rgb) n += up_to(px * ls.l.cout * 4);            // fp32 NCHW map for the ToRGB kernel
  if (ls.fuse_rgb) n += up_to(px * 3 * 4) + up_to((size_t)batch * 3 * ls.l.cout * 4);
  return n;
}

}  // namespace
}  // namespace maua

extern "C" {

int maua_synth_create(const MauaSynthDesc* desc, MauaSynth** out) {
  using namespace maua;
  MAUA_CHECK_ARG(desc && out && desc->layers && desc->n_layers >= 1 && desc->const_input, "synth_create: bad descriptor");
  MAUA_CHECK_ARG(desc->style_dim >= 1 && desc->style_dim <= 1024 && desc->n_latent >= 1, "synth_create: bad style layout");
  MAUA_CHECK_ARG(desc->in_h >= 1 && desc->in_w >= 1, "synth_create: bad input size");
  MAUA_CHECK_ARG(desc->precision >= 0 && desc->precision <= 2, "synth_create: precision must be 0 (bf16x3), 1 (mixed), 2 (bf16)");
  MauaSynth* h = new (std::nothrow) MauaSynth();
  MAUA_CHECK_ARG(h != nullptr, "synth_create: out of host memory");
  h->desc = *desc;
  h->desc.layers = nullptr;
  int hh = desc->in_h, ww = desc->in_w, res = 2;
  size_t off = 0;
  for (int i = 0; i < desc->n_layers; ++i) {
    LayerState ls{};
    ls.l = desc->layers[i];
    const MauaSynthLayer& l = ls.l;
    if (!(l.conv_weight && l.mod_weight && l.mod_bias && l.act_bias && l.noise_weight) || !tc_ok(l.cin, l.cout) ||
        (l.up && !l.blur_kernel) || l.latent_index < 0 || l.latent_index >= desc->n_latent) {
      delete h;
      set_error("synth_create: layer %d is incomplete or not supported by the tensor-core path (Cin %% 32, Cout %% 16)", i);
      return MAUA_E_UNSUPPORTED;
    }
    ls.in_h = hh; ls.in_w = ww;
    if (l.up) { hh *= 2; ww *= 2; }
    ls.out_h = hh; ls.out_w = ww;
    res *= (l.up || i == 0) ? 2 : 1;
    const bool big = (ls.in_h + (l.up ? 1 : 0) >= 64) && (ls.in_w + (l.up ? 1 : 0) >= 32);
    ls.fmt = desc->precision == 2 ? 1 : ((desc->precision == 1 && res >= desc->f16_min_res && big) ? 2 : 3);
    ls.has_rgb = l.rgb_weight != nullptr && desc->min_rgb_size <= res;
    if (ls.has_rgb && !(l.rgb_mod_weight && l.rgb_mod_bias && l.rgb_bias && l.rgb_latent_index >= 0 &&
                        l.rgb_latent_index < desc->n_latent)) {
      delete h;
      set_error("synth_create: layer %d has an incomplete ToRGB", i);
      return MAUA_E_ARG;
    }
    ls.fuse_rgb = ls.has_rgb && !l.up && l.cout <= 128 && ls.in_h >= 64 && ls.in_w >= 32;
    ls.job = h->n_jobs++;
    ls.rgb_job = ls.has_rgb ? h->n_jobs++ : -1;
    // plan buffer layout: Wsq | w_hi | w_lo
    ls.wsq = reinterpret_cast<float*>(off);
    off += up_to((size_t)l.cout * l.cin * 4);
    ls.w_hi = reinterpret_cast<void*>(off);
    off += up_to((size_t)9 * l.cout * l.cin * 2);
    ls.w_lo = reinterpret_cast<void*>(off);
    off += up_to((size_t)9 * l.cout * l.cin * 2);
    h->layers.push_back(ls);
  }
  h->plan_bytes = off;
  *out = h;
  return MAUA_OK;
}

void maua_synth_destroy(MauaSynth* h) {
  if (!h) return;
  if (h->side) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(h->side_device);
    cudaStreamSynchronize(h->side);
    cudaEventDestroy(h->ev_main);
    cudaEventDestroy(h->ev_side);
    cudaStreamDestroy(h->side);
    cudaSetDevice(cur);
  }
  delete h;
}

size_t maua_synth_plan_bytes(const MauaSynth* h) { return h ? h->plan_bytes + maua::ALIGN : 0; }

int maua_synth_prepare(MauaSynth* h, void* plan, size_t plan_bytes, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(h && plan && plan_bytes >= h->plan_bytes + ALIGN, "synth_prepare: plan buffer too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(up_to(reinterpret_cast<size_t>(plan)));
  for (auto& ls : h->layers) {
    if (!h->prepared) {  // first call: offsets -> pointers
      ls.wsq = reinterpret_cast<float*>(base + reinterpret_cast<size_t>(ls.wsq));
      ls.w_hi = base + reinterpret_cast<size_t>(ls.w_hi);
      ls.w_lo = base + reinterpret_cast<size_t>(ls.w_lo);
    }
    const float scale = (float)(1.0 / std::sqrt((double)(ls.l.cin * 9)));   // == Python's 1 / math.sqrt(cin * k * k)
    int rc = maua_weight_sq_f32(ls.l.conv_weight, ls.wsq, ls.l.cout, ls.l.cin, 3, scale, stream);
    if (rc != MAUA_OK) return rc;
    rc = ls.fmt == 2 ? maua_pack_weight_f16x2(ls.l.conv_weight, ls.w_hi, ls.w_lo, ls.l.cout, ls.l.cin, 3, scale, stream)
                     : maua_pack_weight_bf16x2(ls.l.conv_weight, ls.w_hi, ls.w_lo, ls.l.cout, ls.l.cin, 3, scale, stream);
    if (rc != MAUA_OK) return rc;
  }
  h->prepared = true;
  return MAUA_OK;
}

size_t maua_synth_workspace_bytes(const MauaSynth* h, int batch) {
  using namespace maua;
  if (!h || batch < 1) return 0;
  size_t n = ALIGN + up_to(sizeof(MauaStyleJob) * h->n_jobs) + SPLITK_BYTES;
  n += up_to((size_t)batch * h->desc.n_latent * h->desc.style_dim * 4);
  size_t arena = up_to((size_t)batch * h->desc.in_h * h->desc.in_w * h->layers[0].l.cin * 4);
  size_t img = 0;
  for (size_t i = 0; i < h->layers.size(); ++i) {
    const auto& ls = h->layers[i];
    n += up_to((size_t)batch * ls.l.cin * 4) * 2 + up_to((size_t)batch * ls.l.cout * 4);   // s, s_norm, d
    if (ls.has_rgb) n += up_to((size_t)batch * ls.l.cout * 4);                             // ToRGB style
    const size_t a = layer_arena_bytes(ls, i + 1 < h->layers.size() ? &h->layers[i + 1] : nullptr, batch);
    if (a > arena) arena = a;
    if (ls.has_rgb) img = up_to((size_t)batch * 3 * ls.out_h * ls.out_w * 4);
  }
  return n + 2 * arena + 2 * img;
}

int maua_synth_bind(MauaSynth* h, void* workspace, size_t workspace_bytes, int batch, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(h && workspace && batch >= 1, "synth_bind: bad arguments");
  MAUA_CHECK_ARG(h->prepared, "synth_bind: call maua_synth_prepare first");
  MAUA_CHECK_ARG(workspace_bytes >= maua_synth_workspace_bytes(h, batch), "synth_bind: workspace too small");
  uint8_t* p = reinterpret_cast<uint8_t*>(up_to(reinterpret_cast<size_t>(workspace)));
  auto take = [&](size_t bytes) { uint8_t* r = p; p += up_to(bytes); return r; };
  h->jobs_dev = reinterpret_cast<MauaStyleJob*>(take(sizeof(MauaStyleJob) * h->n_jobs));
  h->splitk = take(SPLITK_BYTES);
  h->latent_t = reinterpret_cast<float*>(take((size_t)batch * h->desc.n_latent * h->desc.style_dim * 4));
  std::vector<MauaStyleJob> jobs(h->n_jobs);
  size_t arena = up_to((size_t)batch * h->desc.in_h * h->desc.in_w * h->layers[0].l.cin * 4), img = 0;
  for (size_t i = 0; i < h->layers.size(); ++i) {
    auto& ls = h->layers[i];
    ls.s = reinterpret_cast<float*>(take((size_t)batch * ls.l.cin * 4));
    ls.s_norm = reinterpret_cast<float*>(take((size_t)batch * ls.l.cin * 4));
    ls.d = reinterpret_cast<float*>(take((size_t)batch * ls.l.cout * 4));
    MauaStyleJob& j = jobs[ls.job];
    memset(&j, 0, sizeof(j));
    j.mod_w = ls.l.mod_weight; j.mod_b = ls.l.mod_bias; j.wsq = ls.wsq;
    j.s_out = ls.s; j.d_out = ls.d; j.s_norm_out = ls.fmt == 2 ? ls.s_norm : nullptr;
    j.cin = ls.l.cin; j.cout = ls.l.cout; j.latent_index = ls.l.latent_index;
    if (ls.has_rgb) {
      ls.rgb_s = reinterpret_cast<float*>(take((size_t)batch * ls.l.cout * 4));
      MauaStyleJob& r = jobs[ls.rgb_job];
      memset(&r, 0, sizeof(r));
      r.mod_w = ls.l.rgb_mod_weight; r.mod_b = ls.l.rgb_mod_bias; r.wsq = nullptr;
      r.s_out = ls.rgb_s; r.cin = ls.l.cout; r.cout = 3; r.latent_index = ls.l.rgb_latent_index;
      img = up_to((size_t)batch * 3 * ls.out_h * ls.out_w * 4);
    }
    const size_t a = layer_arena_bytes(ls, i + 1 < h->layers.size() ? &h->layers[i + 1] : nullptr, batch);
    if (a > arena) arena = a;
  }
  h->arena_bytes = arena;
  h->arena[0] = take(arena);
  h->arena[1] = take(arena);
  h->image[0] = reinterpret_cast<float*>(take(img));
  h->image[1] = reinterpret_cast<float*>(take(img));
  cudaStream_t st = as_stream(stream);
  // (synchronous w.r.t. the host buffer: `jobs` is pageable memory, which is why bind() must stay outside graph capture)
  MAUA_CHECK_CUDA(cudaMemcpyAsync(h->jobs_dev, jobs.data(), sizeof(MauaStyleJob) * h->n_jobs, cudaMemcpyHostToDevice, st));
  MAUA_CHECK_CUDA(cudaMemsetAsync(h->splitk, 0, SPLITK_BYTES, st));
  MAUA_CHECK_CUDA(cudaStreamSynchronize(st));
  if (!h->side) {   // (created here, not in forward(): bind is the call that is documented as not capturable)
    const char* e = getenv("MAUA_SYNTH_OVERLAP");
    h->overlap = !(e && atoi(e) == 0);
    int lo = 0, hi = 0;
    MAUA_CHECK_CUDA(cudaGetDevice(&h->side_device));
    MAUA_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // highest priority: its few short blocks take the first SMs a draining conv kernel frees
    MAUA_CHECK_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
    MAUA_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    MAUA_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
  }
  h->ws = workspace;
  h->ws_bytes = workspace_bytes;
  h->batch = batch;
  return MAUA_OK;
}

int maua_synth_forward(MauaSynth* h, const float* latent, int latent_rows, const float* const* noise,
                       const long long* noise_bstride, const float* mean_latent, const float* psi, float psi_scalar,
                       int batch, float* out_rgb, uint8_t* out_u8, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(h && latent && (out_rgb || out_u8), "synth_forward: bad arguments");
  MAUA_CHECK_ARG(h->ws && batch == h->batch, "synth_forward: call maua_synth_bind for this batch size first");
  MAUA_CHECK_ARG(latent_rows >= h->desc.n_latent, "synth_forward: latents have fewer rows than n_latent");
  const MauaSynthDesc& D = h->desc;
  int rc = maua_style_prologue_f32(h->jobs_dev, h->n_jobs, latent, mean_latent, psi, psi_scalar, h->latent_t, batch,
                                   latent_rows, D.style_dim, stream);
  if (rc != MAUA_OK) return rc;

  const void *x_hi = nullptr, *x_lo = nullptr;   // operand planes of the current layer's input
  float* image = nullptr;
  int image_slot = 0;
  const size_t n_layers = h->layers.size();
  cudaStream_t main_st = as_stream(stream);
  int side_reads_arena = -1;   // arena the newest ToRGB kernel on the side stream still reads (-1: side stream joined)
  auto join_side = [&]() -> int {
    if (side_reads_arena >= 0) {
      MAUA_CHECK_CUDA(cudaStreamWaitEvent(main_st, h->ev_side, 0));
      side_reads_arena = -1;
    }
    return MAUA_OK;
  };
  for (size_t li = 0; li < n_layers; ++li) {
    const LayerState& ls = h->layers[li];
    // this layer overwrites arena[li & 1]: a ToRGB kernel of layer li-2 that still reads its map there must be done
    if (side_reads_arena == (int)(li & 1) && (rc = join_side()) != MAUA_OK) return rc;
    const LayerState* nxt = li + 1 < n_layers ? &h->layers[li + 1] : nullptr;
    const MauaSynthLayer& l = ls.l;
    uint8_t* a = h->arena[li & 1];
    auto take = [&](size_t bytes) { uint8_t* r = a; a += up_to(bytes); return r; };
    const float* s_in = ls.fmt == 2 ? ls.s_norm : ls.s;
    if (li == 0) {  // constant input (one sample, broadcast) * s -> operand planes, in the OTHER arena
      uint8_t* b = h->arena[1];
      const size_t plane = (size_t)batch * ls.in_h * ls.in_w * l.cin * 2;
      if (ls.fmt == 2) {
        rc = maua_modulate_f16_nhwc(D.const_input, 0, s_in, b, batch, l.cin, ls.in_h, ls.in_w, stream);
        x_hi = b; x_lo = nullptr;
      } else {
        rc = maua_modulate_split_nhwc(D.const_input, 0, s_in, b, b + up_to(plane), batch, l.cin, ls.in_h, ls.in_w, stream);
        x_hi = b; x_lo = b + up_to(plane);
      }
      if (rc != MAUA_OK) return rc;
    }
    const size_t px = (size_t)batch * ls.out_h * ls.out_w;
    MauaConvEpilogue ep;
    memset(&ep, 0, sizeof(ep));
    const float* nz = noise ? noise[li] : nullptr;
    long long nbs = (nz && noise_bstride) ? noise_bstride[li] : 0;
    if (!nz) { nz = l.noise_buffer; nbs = 0; }
    MAUA_CHECK_ARG(nz != nullptr, "synth_forward: layer %d has neither per-frame noise nor a noise buffer", (int)li);
    ep.noise = nz; ep.noise_weight = l.noise_weight; ep.noise_bstride = nbs;
    ep.bias = l.act_bias;
    ep.slope = 0.2f; ep.act_scale = 1.4142135623730951f; ep.activate = 1;
    ep.workspace = h->splitk; ep.workspace_bytes = (long long)SPLITK_BYTES;
    void *o_hi = nullptr, *o_lo = nullptr;
    if (nxt) {
      ep.s_next = nxt->fmt == 2 ? nxt->s_norm : nxt->s;
      ep.out_fmt = nxt->fmt == 2 ? 1 : 0;
      const size_t plane = px * l.cout * 2;
      o_hi = take(nxt->fmt == 2 ? plane : 2 * plane);
      o_lo = nxt->fmt == 2 ? nullptr : reinterpret_cast<uint8_t*>(o_hi) + plane;
      ep.out_hi = o_hi; ep.out_lo = o_lo;
    }
    float* y = nullptr;
    if (ls.has_rgb && !ls.fuse_rgb) {
      y = reinterpret_cast<float*>(take(px * l.cout * 4));
      ep.out_f32_nchw = y;
    }
    float* partial = nullptr;
    if (ls.fuse_rgb) {
      partial = reinterpret_cast<float*>(take(px * 3 * 4));
      float* wr = reinterpret_cast<float*>(take((size_t)batch * 3 * l.cout * 4));
      rc = maua_rgb_weights_f32(l.rgb_weight, ls.rgb_s, wr, batch, l.cout, (float)(1.0 / std::sqrt((double)l.cout)), stream);
      if (rc != MAUA_OK) return rc;
      ep.rgb_w = wr; ep.rgb_out = partial;
    }
    if (!l.up) {
      ep.d = ls.d;
      rc = maua_modconv_tc(x_hi, x_lo, ls.w_hi, ls.w_lo, &ep, batch, l.cin, l.cout, ls.in_h, ls.in_w, 0, ls.fmt, stream);
      if (rc != MAUA_OK) return rc;
    } else {
      float* u = reinterpret_cast<float*>(take((size_t)batch * (2 * ls.in_h + 1) * (2 * ls.in_w + 1) * l.cout * 4));
      MauaConvEpilogue er;
      memset(&er, 0, sizeof(er));
      er.out_raw_nhwc = u; er.activate = 0;   // raw phases; the demodulation rides in blur_act (commutes with the FIR)
      ep.d = ls.d;
      er.workspace = ep.workspace; er.workspace_bytes = ep.workspace_bytes;
      rc = maua_modconv_tc(x_hi, x_lo, ls.w_hi, ls.w_lo, &er, batch, l.cin, l.cout, ls.in_h, ls.in_w, 1, ls.fmt, stream);
      if (rc != MAUA_OK) return rc;
      rc = maua_blur_act_nhwc(u, l.blur_kernel, &ep, batch, l.cout, 2 * ls.in_h + 1, 2 * ls.in_w + 1, stream);
      if (rc != MAUA_OK) return rc;
    }
    x_hi = o_hi; x_lo = o_lo;
    if (ls.has_rgb) {
      const bool last_rgb = [&] { for (size_t k = li + 1; k < n_layers; ++k) if (h->layers[k].has_rgb) return false; return true; }();
      float* dst = (last_rgb && out_rgb) ? out_rgb : h->image[image_slot];
      const float* upk = image ? l.rgb_up_kernel : nullptr;
      MAUA_CHECK_ARG(!image || upk, "synth_forward: layer %d needs rgb_up_kernel for its skip connection", (int)li);
      const bool bytes_only = last_rgb && out_u8 && !out_rgb && ls.fuse_rgb && ls.out_w % 4 == 0;
      // every ToRGB but the last goes to the side stream, behind this layer's conv; the last one joins the branches
      const bool on_side = h->overlap && !last_rgb;
      void* rgb_stream = stream;
      if (on_side) {
        MAUA_CHECK_CUDA(cudaEventRecord(h->ev_main, main_st));
        MAUA_CHECK_CUDA(cudaStreamWaitEvent(h->side, h->ev_main, 0));
        rgb_stream = h->side;
      } else if ((rc = join_side()) != MAUA_OK) {
        return rc;
      }
      stream = rgb_stream;
      if (bytes_only)   // last ToRGB straight to uint8 NHWC: the full-resolution fp32 image is never written
        rc = maua_rgb_finish_u8(partial, l.rgb_bias, image, upk, out_u8, batch, ls.out_h, ls.out_w, stream);
      else if (ls.fuse_rgb)
        rc = maua_rgb_finish_f32(partial, l.rgb_bias, image, upk, dst, batch, ls.out_h, ls.out_w, stream);
      else
        rc = maua_torgb_f32(y, l.rgb_weight, ls.rgb_s, l.rgb_bias, image, upk, dst, batch, l.cout, ls.out_h, ls.out_w,
                            (float)(1.0 / std::sqrt((double)l.cout)), stream);
      stream = main_st;
      if (rc != MAUA_OK) return rc;
      if (on_side) {
        MAUA_CHECK_CUDA(cudaEventRecord(h->ev_side, h->side));
        side_reads_arena = (int)(li & 1);
      }
      image = dst;
      image_slot ^= 1;
      if (last_rgb && out_u8 && !bytes_only) {
        rc = maua_rgb_to_u8_nhwc(image, out_u8, batch, ls.out_h, ls.out_w, stream);
        if (rc != MAUA_OK) return rc;
      }
    }
  }
  if ((rc = join_side()) != MAUA_OK) return rc;
  MAUA_CHECK_ARG(image != nullptr, "synth_forward: the network has no ToRGB layer");
  return MAUA_OK;
}

const float* maua_synth_truncated_latents(const MauaSynth* h) { return h ? h->latent_t : nullptr; }

}  // extern "C"
