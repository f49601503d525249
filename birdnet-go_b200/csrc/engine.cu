// engine.cu — weight upload, workspaces and the per-micro-batch kernel chain.
//
// Execution model: a batch of B independent 3 s chunks is cut into micro-batches of `micro_`
// chunks; one micro-batch runs the whole kernel chain before the next starts, so every
// producer->consumer activation pair (<= 2.4 MB per chunk) stays resident in the 126 MB L2 and HBM
// sees little more than the PCM in and the logits out (SURVEY.md §7 "L2-resident batch tiling").
// Host-buffer calls overlap the H2D copy of micro-batch i+1 (copy stream) with the compute of
// micro-batch i (compute stream).
#include "engine.h"
#include "mbconv_tc.h"
#include "mbconv2.h"
#include "pw2.h"

#include <mutex>

#include <math.h>
#include <string.h>

#include <algorithm>

namespace bnb {

void frontend_set_attributes();      // frontend.cu
void pw_tc_set_attributes();         // pw_tc.cu
void mbconv_tc_set_attributes();     // mbconv_tc.cu

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: do it once per device ordinal, under a lock
// (a process may hold classifiers on several GPUs).  The caller has made `device` current.
namespace {
std::mutex g_prep_mu;
bool g_prep_done[64] = {};
}  // namespace
// test hooks only: after cudaDeviceReset() the per-device attributes are gone
void tc_forget_devices() { std::lock_guard<std::mutex> lk(g_prep_mu); for (bool& b : g_prep_done) b = false; }

void tc_prepare_device(int device) {
  bool* done = g_prep_done;
  std::lock_guard<std::mutex> lk(g_prep_mu);
  if (device < 0 || device >= 64) throw std::invalid_argument("device ordinal out of range");
  if (done[device]) return;
  frontend_set_attributes(); pw_tc_set_attributes(); mbconv_tc_set_attributes(); mb2_set_attributes(); pw2_set_attributes();
  done[device] = true;
}

namespace {
constexpr double kPi = 3.14159265358979323846;
size_t fmt_bytes(int fmt) { return fmt == BNB_PCM_S16 ? 2 : 4; }
}  // namespace

// ------------------------------------------------------------------------------------------------
const float* Engine::up(const float* host, size_t n) { return up_t<float>(host, n); }

template <class T>
const T* Engine::up_t(const T* host, size_t n) {
  void* d = nullptr;
  BNB_CUDA(cudaMalloc(&d, std::max<size_t>(n * sizeof(T), 16)));
  allocs_.push_back(d);
  BNB_CUDA(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return static_cast<const T*>(d);
}

Engine::Engine(const void* tflite, size_t len, const bnb_options& opts) {
  // a constructor that throws never runs the destructor: release streams / events / device memory here
  try { init(tflite, len, opts); } catch (...) { release(); throw; }
}

void Engine::init(const void* tflite, size_t len, const bnb_options& opts) {
  int dev = opts.device;
  if (dev < 0) BNB_CUDA(cudaGetDevice(&dev));
  BNB_CUDA(cudaSetDevice(dev));
  device_ = dev;
  cudaDeviceProp prop{};
  BNB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) throw std::runtime_error(std::string("device '") + prop.name + "' is not compute capability 10.x (sm_100a kernels only)");
  device_name_ = std::string("CUDA:") + std::to_string(dev) + " " + prop.name;
  tc_prepare_device(dev);
  max_batch_ = opts.max_batch > 0 ? opts.max_batch : 256;
  micro_ = opts.micro_batch > 0 ? opts.micro_batch : 128;   // r02 sweep (profiles/r02_bench_lines.txt): 64 -> 59.4k, 128 -> 63.2k, 256 -> 63.5k chunks/s but no H2D overlap inside a 256-chunk call
  n_lanes_ = opts.lanes > 0 ? std::min<int>(opts.lanes, kMaxLanes) : 2;
  use_graphs_ = opts.use_graphs != 0;
  fused_ = !(getenv("BNB_FUSED") && atoi(getenv("BNB_FUSED")) == 0);
  fused_force_ = getenv("BNB_FUSED") && atoi(getenv("BNB_FUSED")) == 2;   // also in keep-intermediates mode (debug)
  if (micro_ > max_batch_) micro_ = max_batch_;
  precision_ = opts.precision == BNB_PRECISION_DEFAULT ? BNB_PRECISION_F16X3 : opts.precision;
  precision_name_ = precision_ == BNB_PRECISION_F16X3 ? "FP16x3(tcgen05)+FP32" : "FP32";
  // F16X3: round-2 kernels on fp16 hi/lo activation planes (mbconv2.cu + pw2.cu); BNB_V2=0 keeps the round-1 fp32-activation chain
  v2_ = precision_ == BNB_PRECISION_F16X3 && !(getenv("BNB_V2") && atoi(getenv("BNB_V2")) == 0);

  TfModel model = parse_tflite(tflite, len);
  NetPlan P = build_plan(model);
  // ---- kernel-specific geometry the sm_100a kernels are specialised for ---------------------------
  const FrontendPlan& F = P.fe;
  if (F.n_samples != 144000 || F.spec[0].frame_len != 2048 || F.spec[0].hop != 278 || F.spec[1].frame_len != 1024 ||
      F.spec[1].hop != 280 || F.spec[0].n_frames != 511 || F.spec[0].n_mel != 96)
    throw unsupported_model("frontend geometry differs from BirdNET v2.4 (144000 samples, 2048/278 + 1024/280 frames, 96 mel)");
  if (P.stem.conv.kh != 4 || P.stem.conv.kw != 8 || P.stem.conv.cin != 2 || P.stem.conv.cout != 24 || P.stem.stride_h != 2 ||
      P.stem.stride_w != 2 || P.stem.out_w != 256 || P.mix.conv.cout != 24)
    throw unsupported_model("stem geometry differs from BirdNET v2.4 (4x8/s2, 2->24, 1x1 48->24)");
  if (P.post.out_h != 1) throw unsupported_model("post conv must reduce the mel axis to 1");
  for (const BlockPlan& b : P.blocks)
    if (b.has_se) {
      const MbGeom mg = mbconv_geometry(b.in_h, b.in_w, b.out_h, b.out_w, b.stride, b.cin, b.cexp, 0, kMaxDwParts);
      if (b.out_h > kMaxDwParts || mg.th == 0) throw unsupported_model("squeeze-excite block needs more partial-sum slots than the buffer holds");
    }
  for (const BlockPlan& b : P.blocks)
    if (b.cin % 4 || b.cexp % 4 || b.cout % 4 || (b.has_se && (b.cexp > 1536 || b.cse > 64))) throw unsupported_model("block channel counts");
  n_species_ = P.n_species(); n_samples_ = F.n_samples; emb_dim_ = P.emb_dim();

  BNB_CUDA(cudaStreamCreateWithFlags(&compute_, cudaStreamNonBlocking));
  BNB_CUDA(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking));
  for (int i = 0; i < n_lanes_; ++i) {
    BNB_CUDA(cudaStreamCreateWithFlags(&lanes_[i].stream, cudaStreamNonBlocking));
    BNB_CUDA(cudaEventCreateWithFlags(&lanes_[i].done, cudaEventDisableTiming));
  }
  BNB_CUDA(cudaEventCreateWithFlags(&ev_in_, cudaEventDisableTiming));
  BNB_CUDA(cudaEventCreateWithFlags(&ev_small_, cudaEventDisableTiming));
  upload_weights(P);
  alloc_workspace();
}

Engine::~Engine() { release(); }

void Engine::release() noexcept {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (int i = 0; i < n_lanes_; ++i) { if (lanes_[i].stream) cudaStreamDestroy(lanes_[i].stream); if (lanes_[i].done) cudaEventDestroy(lanes_[i].done); }
  if (ev_in_) cudaEventDestroy(ev_in_);
  if (ev_small_) cudaEventDestroy(ev_small_);
  for (void* p : allocs_) cudaFree(p);
  for (auto& kv : keep_bufs_) cudaFree(kv.second.first);
  for (auto& kv : keep_imgs_) cudaFree(kv.second.first);
  for (Slot& S : slots_) {
    if (S.h_in) cudaFreeHost(S.h_in);
    if (S.h_out) cudaFreeHost(S.h_out);
    for (cudaEvent_t e : S.ev_h2d) cudaEventDestroy(e);
    if (S.start) cudaEventDestroy(S.start);
    if (S.done) cudaEventDestroy(S.done);
  }
  if (d_det_) cudaFree(d_det_);
  if (h_det_) cudaFreeHost(h_det_);
  for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second.exec);
  for (cudaEvent_t e : prof_ev_) cudaEventDestroy(e);
  if (compute_) cudaStreamDestroy(compute_);
  if (copy_) cudaStreamDestroy(copy_);
}

// ------------------------------------------------------------------------------------------------
void Engine::upload_weights(const NetPlan& P) {
  // ---- frontend tables -----------------------------------------------------------------------------
  const FrontendPlan& F = P.fe;
  fe_.n_samples = F.n_samples; fe_.n_frames = F.spec[0].n_frames; fe_.n_mel = F.spec[0].n_mel;
  fe_.eps = F.eps; fe_.center = F.center; fe_.gain = F.gain;
  for (int s = 0; s < 2; ++s) {
    const SpecPlan& S = F.spec[s];
    fe_.pow_exp[s] = S.pow_exp; fe_.bn_scale[s] = F.bn_scale[s]; fe_.bn_shift[s] = F.bn_shift[s];
    fe_.win[s] = up(S.window, (size_t)S.frame_len);
    const int N = S.frame_len / 2;                  // complex FFT length: 1024 / 512
    const int n2 = N / 32;                          // first-pass FFT size: 32 / 16
    std::vector<float2> tw((size_t)n2 * 32), post((size_t)N);
    for (int k2 = 0; k2 < n2; ++k2)
      for (int l = 0; l < 32; ++l) {
        const double a = -2.0 * kPi * (double)(l * k2) / (double)N;
        tw[(size_t)k2 * 32 + l] = make_float2((float)cos(a), (float)sin(a));
      }
    for (int k = 0; k < N; ++k) {
      const double a = kPi * (double)k / (double)N;
      post[k] = make_float2((float)(0.5 * cos(a)), (float)(0.5 * sin(a)));
    }
    // Re DFT of the window itself, from the file's fp32 constants, in double: W[k] = sum_n w[n] cos(2 pi k n / L)
    std::vector<float> wdft((size_t)N);
    {
      const int L = S.frame_len;
      std::vector<double> ct((size_t)L);
      for (int i = 0; i < L; ++i) ct[i] = cos(2.0 * kPi * (double)i / (double)L);
      for (int k = 0; k < N; ++k) {
        double acc = 0.0;
        for (int n = 0; n < L; ++n) acc += (double)S.window[n] * ct[(size_t)(((long long)k * n) % L)];
        wdft[k] = (float)acc;
      }
    }
    fe_.win_dft[s] = up(wdft.data(), wdft.size());
    fe_.tw[s] = up_t<float2>(tw.data(), tw.size());
    fe_.post[s] = up_t<float2>(post.data(), post.size());
    // sparse mel rows: contiguous non-zero run per band
    std::vector<int> start(S.n_mel, 0), cnt(S.n_mel, 0);
    int stride = 1, max_bin = 0;
    for (int m = 0; m < S.n_mel; ++m) {
      int lo = -1, hi = -1;
      for (int k = 0; k < S.n_bins; ++k) if (S.mel[(size_t)m * S.n_bins + k] != 0.f) { if (lo < 0) lo = k; hi = k; }
      if (lo >= 0) { start[m] = lo; cnt[m] = hi - lo + 1; stride = std::max(stride, cnt[m]); max_bin = std::max(max_bin, hi); }
    }
    std::vector<float> w((size_t)S.n_mel * stride, 0.f);
    for (int m = 0; m < S.n_mel; ++m) for (int i = 0; i < cnt[m]; ++i) w[(size_t)m * stride + i] = S.mel[(size_t)m * S.n_bins + start[m] + i];
    const int group = (s == 0) ? 32 : 16;
    fe_.nk1[s] = max_bin / group + 1;
    if ((s == 0 && fe_.nk1[0] > 4) || (s == 1 && fe_.nk1[1] > 32) || max_bin >= N)
      throw unsupported_model("mel matrix reaches beyond the spectrum range the frontend kernel computes");
    fe_.mel_stride[s] = stride;
    fe_.mel_start[s] = up_t<int>(start.data(), start.size());
    fe_.mel_cnt[s] = up_t<int>(cnt.data(), cnt.size());
    fe_.mel_w[s] = up(w.data(), w.size());
  }
  fe_tensor_ = F.out_tensor;
  // ---- stem + pool mix ---------------------------------------------------------------------------------
  {
    const ConvW& c = P.stem.conv;   // OHWI [24][4][8][2] -> [kh][kw][ci][co]
    std::vector<float> w((size_t)c.kh * c.kw * c.cin * c.cout);
    for (int o = 0; o < c.cout; ++o) for (int kh = 0; kh < c.kh; ++kh) for (int kw = 0; kw < c.kw; ++kw) for (int ci = 0; ci < c.cin; ++ci)
      w[(((size_t)kh * c.kw + kw) * c.cin + ci) * c.cout + o] = c.w[(((size_t)o * c.kh + kh) * c.kw + kw) * c.cin + ci];
    stem_.w_stem = up(w.data(), w.size());
    std::vector<float> zb(c.cout, 0.f);
    stem_.b_stem = up(c.b ? c.b : zb.data(), (size_t)c.cout);
    const ConvW& m = P.mix.conv;    // [24][48]; normalise concat order to (max | avg)
    std::vector<float> wm((size_t)m.cout * m.cin);
    const int half = m.cin / 2;
    for (int o = 0; o < m.cout; ++o) for (int ci = 0; ci < m.cin; ++ci) {
      int src = ci;
      if (!P.mix.max_first) src = (ci < half) ? ci + half : ci - half;
      wm[(size_t)o * m.cin + ci] = m.w[(size_t)o * m.cin + src];
    }
    stem_.w_mix = up(wm.data(), wm.size());
    std::vector<float> zb2(m.cout, 0.f);
    stem_.b_mix = up(m.b ? m.b : zb2.data(), (size_t)m.cout);
    stem_.in_h = P.stem.in_h; stem_.in_w = P.stem.in_w; stem_.out_h = P.stem.out_h; stem_.out_w = P.stem.out_w;
    stem_.pad_t = P.stem.pad_t; stem_.pad_l = P.stem.pad_l;
    stem_tensor_ = P.stem.out_tensor; mix_tensor_ = P.mix.out_tensor;
  }
  // ---- blocks ----------------------------------------------------------------------------------------------
  const bool tc = (precision_ == BNB_PRECISION_F16X3);
  auto upconv = [&](const ConvW& c, size_t n, bool gemm = false) {
    DevConv d; d.w = up(c.w, n);
    std::vector<float> z((size_t)c.cout + 64, 0.f);      // +64: the tensor-core epilogue stages bias strips a little past n_pad
    if (c.b) memcpy(z.data(), c.b, (size_t)c.cout * sizeof(float));
    d.b = up(z.data(), z.size());
    if (gemm && tc) {                                    // fp16 hi/lo split + swizzled smem image for pw_tc.cu
      std::vector<uint8_t> img;
      d.tc = pw_tc_prepare(c.w, c.cout, (int)(n / (size_t)c.cout), &img);
      d.tc_img = up_t<uint8_t>(img.data(), img.size());
    }
    return d;
  };
  for (const BlockPlan& b : P.blocks) {
    DevBlock d; d.g = b;
    d.expand = upconv(b.expand, (size_t)b.cexp * b.cin, true);
    d.dw = upconv(b.dw, (size_t)9 * b.cexp);
    if (b.has_se) {
      d.se1 = upconv(b.se1, (size_t)b.cse * b.cexp);
      ConvW t = b.se2;                                   // second SE FC stored transposed [Cse][C] for coalesced reads
      std::vector<float> wt((size_t)b.cse * b.cexp);
      for (int c = 0; c < b.cexp; ++c) for (int j = 0; j < b.cse; ++j) wt[(size_t)j * b.cexp + c] = b.se2.w[(size_t)c * b.cse + j];
      t.w = wt.data();
      d.se2 = upconv(t, (size_t)b.cexp * b.cse);
    }
    d.proj = upconv(b.proj, (size_t)b.cout * b.cexp, true);
    if (v2_) {
      static const bool sw128_only = getenv("BNB_MB2_SW128") && atoi(getenv("BNB_MB2_SW128")) != 0;
      d.mb2 = mb2_plan(b.in_h, b.in_w, b.out_h, b.out_w, b.stride, b.cin, b.cexp, !sw128_only);
      if (!d.mb2.ok) throw unsupported_model("MBConv block geometry does not fit the fused expand+depthwise kernel");
      if (b.has_se && d.mb2.tiles_h * d.mb2.tiles_w > kMaxDwParts) throw unsupported_model("squeeze-excite block needs more partial-sum slots than the buffer holds");
      std::vector<uint8_t> img;
      mb2_prepare_weights(d.mb2, b.expand.w, &img);
      d.mb2_img = up_t<uint8_t>(img.data(), img.size());
      std::vector<float> be((size_t)d.mb2.n_units * 128 + 64, 0.f);
      if (b.expand.b) memcpy(be.data(), b.expand.b, (size_t)b.cexp * sizeof(float));
      d.mb2_bias = up(be.data(), be.size());
      d.in_patch = mb2_patch_layout(d.mb2, b.in_h, b.in_w);
      const PatchTables T = patch_build_tables(d.in_patch);                   // pixel -> (tile, row) lookup tables for the producers' epilogues
      d.in_patch.dst_tbl = reinterpret_cast<const uint4*>(up_t<uint32_t>(T.dst.data(), T.dst.size()));
      d.in_patch.res_tbl = up_t<uint32_t>(T.res.data(), T.res.size());
    }
    blocks_.push_back(d);
  }
  // ---- post + head ------------------------------------------------------------------------------------------
  post_g_ = P.post;
  post_mul_ = up(P.post.mul, (size_t)P.post.conv.cin);
  post_add_ = up(P.post.add, (size_t)P.post.conv.cin);
  post_conv_ = upconv(P.post.conv, (size_t)P.post.conv.cout * P.post.conv.kh * P.post.conv.kw * P.post.conv.cin, true);
  fc_ = upconv(P.head.fc, (size_t)P.head.fc.cout * P.head.fc.cin, true);
  logits_tensor_ = P.head.out_tensor;
}

void Engine::alloc_workspace() {
  auto dmalloc = [&](size_t floats) { void* p = nullptr; BNB_CUDA(cudaMalloc(&p, std::max<size_t>(floats, 4) * sizeof(float))); allocs_.push_back(p); return static_cast<float*>(p); };
  // Phase split: blocks whose input map has >= 256 pixels per chunk run per micro-batch (their activations are MBs per
  // chunk and must stay in L2); the rest (<= 96 pixels per chunk) run once over the whole batch so that their GEMMs have
  // enough rows to fill 148 SMs.
  split_ = (int)blocks_.size();
  for (size_t i = 0; i < blocks_.size(); ++i) if (blocks_[i].g.in_h * blocks_[i].g.in_w < 256) { split_ = (int)i; break; }
  if (split_ == 0) split_ = 1;
  const size_t mb = (size_t)micro_, bb = (size_t)max_batch_;
  auto alloc_work = [&](Work& w, size_t cap_n, int lo, int hi, size_t cx) {
    w.cap_n = cap_n;
    size_t ce = 0, cd = 0, cg = 0;
    for (int i = lo; i < hi; ++i) {
      const BlockPlan& g = blocks_[i].g;
      ce = std::max(ce, (size_t)g.in_h * g.in_w * g.cexp);
      cd = std::max(cd, (size_t)g.out_h * g.out_w * g.cexp);
      cx = std::max(cx, (size_t)g.out_h * g.out_w * g.cout);
      cx = std::max(cx, (size_t)g.in_h * g.in_w * g.cin);
      cg = std::max(cg, (size_t)g.cexp);
    }
    w.x0 = dmalloc(cap_n * cx); w.x1 = dmalloc(cap_n * cx);
    w.e = dmalloc(cap_n * ce); w.d = dmalloc(cap_n * cd); w.g = dmalloc(cap_n * cg);
    w.sep = dmalloc(cap_n * kMaxDwParts * cg);   // SE partial sums: <= kMaxDwParts parts per chunk (rows, or fused-kernel tiles)
  };
  // the same for the plane path: x images hold block inputs / outputs (PatchTiles of the consuming block), d the depthwise output
  // (RowTiles).  Image buffers are zero-filled once: bytes no kernel writes must stay finite (layouts.h).
  auto alloc_work2 = [&](Work2& w, size_t cap_n, int lo, int hi) {
    size_t cx = 0, cd = 0, cg = 0;
    for (int i = lo; i < hi; ++i) {
      const BlockPlan& g = blocks_[i].g;
      cx = std::max(cx, blocks_[i].in_patch.bytes(1));
      if (i + 1 < (int)blocks_.size()) cx = std::max(cx, blocks_[i + 1].in_patch.bytes(1));
      cd = std::max(cd, (size_t)RowTiles::make(g.cexp).tile_bytes * (((size_t)cap_n * g.out_h * g.out_w + 127) / 128));
      cg = std::max(cg, (size_t)g.cexp);
    }
    w.x_bytes = cx; w.d_bytes = cd;
    w.x0 = alloc_img(cap_n * cx); w.x1 = alloc_img(cap_n * cx); w.d = alloc_img(cd);
    w.g = dmalloc(cap_n * cg);
    w.sep = dmalloc(cap_n * kMaxDwParts * cg);
  };
  for (int l = 0; l < n_lanes_; ++l) {
    if (v2_) alloc_work2(lanes_[l].w2, mb, 0, split_);
    else alloc_work(lanes_[l].w, mb, 0, split_, (size_t)stem_.out_h * (stem_.out_w / 2) * 24);
    lanes_[l].partial = dmalloc(mb * kMinMaxParts * 2);
    lanes_[l].fe = dmalloc(mb * fe_.n_mel * fe_.n_frames * 2);
  }
  const BlockPlan& gs = blocks_[split_ - 1].g;
  ws_pc_ = dmalloc(bb * post_g_.out_w * post_g_.conv.cout);
  ws_emb_ = dmalloc(bb * emb_dim_);
  const size_t im2col_sz = (size_t)post_g_.out_w * post_g_.conv.kh * post_g_.conv.kw * post_g_.conv.cin;
  if (v2_) {
    alloc_work2(work2_back_, bb, split_, (int)blocks_.size());
    if (split_ >= (int)blocks_.size()) throw unsupported_model("the network has no small-map blocks for the whole-batch phase");
    mid2_bytes_ = blocks_[split_].in_patch.bytes(1);                      // the split-point tensor = input image of block `split_`
    mid2_ = alloc_img(bb * mid2_bytes_); mid2_base_ = mid2_;
    const BlockPlan& gl = blocks_.back().g;
    last2_ = alloc_planes(bb * gl.out_h * gl.out_w * plane_pitch(gl.cout));   // the last block's output: plain planes for the im2col kernel
    im2col2_ = alloc_img(RowTiles::make((int)im2col_sz).bytes((long long)bb * post_g_.out_w));
    emb2_ = alloc_img(RowTiles::make(emb_dim_).bytes((long long)bb));
  } else {
    alloc_work(work_back_, bb, split_, (int)blocks_.size(), 0);
    mid_sz_ = (size_t)gs.out_h * gs.out_w * gs.cout;
    ws_mid_ = dmalloc(bb * mid_sz_); ws_mid_base_ = ws_mid_;
    ws_im2col_ = dmalloc(bb * im2col_sz);
  }
}

float* Engine::scratch(int tensor_id, float* normal, size_t per_chunk, int n) {
  if (!keep_ || tensor_id < 0) return normal;
  auto it = keep_bufs_.find(tensor_id);
  const size_t need = per_chunk * (size_t)std::max(micro_, n);
  if (it == keep_bufs_.end() || it->second.second < need) {
    if (it != keep_bufs_.end()) cudaFree(it->second.first);
    void* p = nullptr;
    BNB_CUDA(cudaMalloc(&p, need * sizeof(float)));
    keep_bufs_[tensor_id] = {static_cast<float*>(p), need};
    return static_cast<float*>(p);
  }
  return it->second.first;
}

Planes Engine::alloc_planes(size_t elems) {
  Planes p;
  void* q = nullptr;
  BNB_CUDA(cudaMalloc(&q, std::max<size_t>(elems, 8) * sizeof(__half))); allocs_.push_back(q); p.h = static_cast<__half*>(q);
  BNB_CUDA(cudaMalloc(&q, std::max<size_t>(elems, 8) * sizeof(__half))); allocs_.push_back(q); p.l = static_cast<__half*>(q);
  return p;
}

uint8_t* Engine::alloc_img(size_t bytes) {
  void* q = nullptr;
  bytes = std::max<size_t>(bytes, 1024);
  BNB_CUDA(cudaMalloc(&q, bytes)); allocs_.push_back(q);
  BNB_CUDA(cudaMemset(q, 0, bytes));
  return static_cast<uint8_t*>(q);
}

uint8_t* Engine::scratch_img(int tensor_id, uint8_t* normal, size_t bytes_per_chunk, int n) {
  if (!keep_ || tensor_id < 0) return normal;
  auto it = keep_imgs_.find(tensor_id);
  const size_t need = bytes_per_chunk * (size_t)std::max(micro_, n) + 65536;
  if (it == keep_imgs_.end() || it->second.second < need) {
    if (it != keep_imgs_.end()) cudaFree(it->second.first);
    void* q = nullptr;
    BNB_CUDA(cudaMalloc(&q, need));
    BNB_CUDA(cudaMemset(q, 0, need));
    keep_imgs_[tensor_id] = {static_cast<uint8_t*>(q), need};
    return static_cast<uint8_t*>(q);
  }
  return it->second.first;
}

// ------------------------------------------------------------------------------------------------
void Engine::pw(const PwArgs& a, const DevConv& c, int cat, cudaStream_t s) {
  ProfScope ps(this, cat, s);
  if (c.tc_img) launch_pw_tc(c.tc, a, c.tc_img, s, lc_);
  else launch_pw_conv(a, s, lc_);
}

float* Engine::run_blocks(int lo, int hi, float* cur, int n, Work& w, cudaStream_t s) {
  float* nxt_normal = (cur == w.x0) ? w.x1 : w.x0;
  for (int bi = lo; bi < hi; ++bi) {
    const DevBlock& b = blocks_[bi];
    const BlockPlan& g = b.g;
    const int hw_in = g.in_h * g.in_w, hw_out = g.out_h * g.out_w;
    float* d = scratch(g.dw_tensor, w.d, (size_t)hw_out * g.cexp, n);
    int se_parts = 0;
    // fused front half: wins on the large-map blocks of the front phase and on the stride-2 block of the back phase; the
    // stride-1 small-map blocks of the back phase measure the same either way (run 26) and keep the two-kernel chain
    const int mb_cap = g.has_se ? kMaxDwParts : 0;
    // the tile search sees the NOMINAL launch size (micro-batch in the front phase, max batch in the back phase), never the
    // actual one: the geometry fixes the order of the SE partial sums, and results must not depend on batch composition
    const MbGeom mg = mbconv_geometry(g.in_h, g.in_w, g.out_h, g.out_w, g.stride, g.cin, g.cexp, bi < split_ ? micro_ : max_batch_, mb_cap);
    const bool fits = mg.th > 0 && mg.smem_bytes <= kMbSmemLimit;        // K > 128 does not fit the shared-memory plan
    if (fused_ && fits && b.expand.tc_img && ((!keep_ && (bi < split_ || g.stride == 2)) || fused_force_)) {
      // expand + SiLU + depthwise + SiLU + SE sums in one tcgen05 kernel: the expanded tensor is never materialised
      se_parts = mg.tiles_h * mg.tiles_w;
      MbLaunch ml{};
      ml.x = cur; ml.Wimg = b.expand.tc_img; ml.bias_e = b.expand.b; ml.w_dw = b.dw.w; ml.bias_dw = b.dw.b; ml.D = d;
      ml.partial = g.has_se ? w.sep : nullptr;
      ml.B = n; ml.H = g.in_h; ml.W = g.in_w; ml.Cin = g.cin; ml.C = g.cexp; ml.Ho = g.out_h; ml.Wo = g.out_w; ml.stride = g.stride; ml.max_tiles = mb_cap; ml.B_nominal = bi < split_ ? micro_ : max_batch_;
      { ProfScope ps(this, C_PW_EXPAND, s); launch_mbconv_tc(ml, s, lc_); }
    } else {
      float* e = scratch(g.exp_tensor, w.e, (size_t)hw_in * g.cexp, n);
      PwArgs ex{};
      ex.A = cur; ex.W = b.expand.w; ex.bias = b.expand.b; ex.C = e; ex.M = n * hw_in; ex.N = g.cexp; ex.K = g.cin;
      ex.rows_per_chunk = hw_in; ex.act = ACT_SILU; ex.a_mode = A_PLAIN;
      pw(ex, b.expand, C_PW_EXPAND, s);
      record(g.exp_tensor, e, (size_t)hw_in * g.cexp, n);
      DwArgs dw{};
      dw.in = e; dw.w = b.dw.w; dw.bias = b.dw.b; dw.out = d; dw.B = n; dw.H = g.in_h; dw.W = g.in_w; dw.C = g.cexp;
      dw.stride = g.stride; dw.Ho = g.out_h; dw.Wo = g.out_w;
      dw.parts = std::min(kMaxDwParts, dw_parts(n, g.out_h, g.out_w, g.cexp, g.in_w, g.stride));
      dw.partial = g.has_se ? w.sep : nullptr;
      se_parts = dw.parts;
      { ProfScope ps(this, C_DW, s); launch_dw_conv(dw, s, lc_); }
    }
    record(g.dw_tensor, d, (size_t)hw_out * g.cexp, n);
    float* gate = nullptr;
    if (g.has_se) {
      gate = scratch(g.gate_tensor, w.g, (size_t)g.cexp, n);
      SeArgs se{w.sep, b.se1.w, b.se1.b, b.se2.w, b.se2.b, gate, n, hw_out, g.cexp, g.cse, se_parts};
      { ProfScope ps(this, C_SE, s); launch_se_gate(se, s, lc_); }
      record(g.gate_tensor, gate, (size_t)g.cexp, n);
    }
    float* out = scratch(g.out_tensor, nxt_normal, (size_t)hw_out * g.cout, n);
    PwArgs pj{};
    pj.A = d; pj.W = b.proj.w; pj.bias = b.proj.b; pj.C = out; pj.M = n * hw_out; pj.N = g.cout; pj.K = g.cexp;
    pj.rows_per_chunk = hw_out; pj.act = ACT_NONE; pj.a_mode = A_PLAIN; pj.gate = gate; pj.residual = g.residual ? cur : nullptr;
    pw(pj, b.proj, C_PW_PROJECT, s);
    record(g.out_tensor, out, (size_t)hw_out * g.cout, n);
    if (!keep_) nxt_normal = cur;      // ping-pong (in keep mode `out` is a private buffer)
    cur = out;
  }
  return cur;
}

// F16X3 path: every block = mbconv2 (expand + SiLU + depthwise + SiLU + SE sums) -> [SE gate] -> pw2 (project, + gate, + residual).
// Activations travel as fp16 hi/lo images pre-tiled for their consumer (layouts.h): block inputs as PatchTiles, depthwise
// outputs as RowTiles; every operand tile is one cp.async.bulk.
void Engine::run_blocks2(int lo, int hi, const uint8_t* cur, int n, Work2& w, cudaStream_t s, uint8_t* final_img, Planes final_plain) {
  uint8_t* nxt_normal = (cur == w.x0) ? w.x1 : w.x0;
  const int nb = (int)blocks_.size();
  for (int bi = lo; bi < hi; ++bi) {
    const DevBlock& b = blocks_[bi];
    const BlockPlan& g = b.g;
    const int hw_out = g.out_h * g.out_w;
    const RowTiles dt = RowTiles::make(g.cexp);
    uint8_t* d = scratch_img(g.dw_tensor, w.d, dt.bytes(hw_out) , n);
    Mb2Launch ml{};
    ml.x_img = cur; ml.Wimg = b.mb2_img; ml.bias_e = b.mb2_bias; ml.w_dw = b.dw.w; ml.bias_dw = b.dw.b; ml.d_img = d;
    ml.partial = g.has_se ? w.sep : nullptr;
    ml.B = n; ml.H = g.in_h; ml.W = g.in_w; ml.Ho = g.out_h; ml.Wo = g.out_w;
    { ProfScope ps(this, C_PW_EXPAND, s); launch_mbconv2(b.mb2, ml, s, lc_); }
    record_rows(g.dw_tensor, d, hw_out, g.cexp, n);
    float* gate = nullptr;
    if (g.has_se) {
      gate = scratch(g.gate_tensor, w.g, (size_t)g.cexp, n);
      SeArgs se{w.sep, b.se1.w, b.se1.b, b.se2.w, b.se2.b, gate, n, hw_out, g.cexp, g.cse, b.mb2.tiles_h * b.mb2.tiles_w};
      { ProfScope ps(this, C_SE, s); launch_se_gate(se, s, lc_); }
      record(g.gate_tensor, gate, (size_t)g.cexp, n);
    }
    Pw2Launch pj{};
    pj.a_img = d; pj.Wimg = b.proj.tc_img; pj.bias = b.proj.b; pj.gate = gate;
    if (g.residual) { pj.r_img = cur; pj.r_patch = b.in_patch; }
    pj.M = n * hw_out; pj.N = g.cout; pj.K = g.cexp; pj.rows_per_chunk = hw_out; pj.act = ACT_NONE;
    const uint8_t* out_img = nullptr;
    if (bi + 1 < nb) {                       // the next block's input image
      const PatchTiles& nt = blocks_[bi + 1].in_patch;
      uint8_t* o = (bi == hi - 1 && final_img && !keep_) ? final_img : scratch_img(g.out_tensor, nxt_normal, nt.bytes(1), n);
      pj.o_img = o; pj.o_patch = nt; out_img = o;
      { ProfScope ps(this, C_PW_PROJECT, s); launch_pw2(b.proj.tc, pj, s, lc_); }
      record_patch(g.out_tensor, o, nt, n);
      if (bi == hi - 1 && final_img && keep_) BNB_CUDA(cudaMemcpyAsync(final_img, o, nt.bytes(n), cudaMemcpyDeviceToDevice, s));
    } else {                                 // last block of the network: plain planes for the im2col kernel
      pj.oh = final_plain.h; pj.ol = final_plain.l; pj.o_pitch = plane_pitch(g.cout);
      { ProfScope ps(this, C_PW_PROJECT, s); launch_pw2(b.proj.tc, pj, s, lc_); }
      record_planes(g.out_tensor, final_plain, hw_out, g.cout, n);
    }
    if (!keep_) nxt_normal = (cur == w.x0 || cur == w.x1) ? const_cast<uint8_t*>(cur) : ((out_img == w.x0) ? w.x1 : w.x0);
    cur = out_img;
  }
}

void Engine::run_back2(int n, float* d_logits, float* d_emb, cudaStream_t s) {
  run_blocks2(split_, (int)blocks_.size(), mid2_, n, work2_back_, s, nullptr, last2_);
  const PostPlan& q = post_g_;
  const int kpost = q.conv.kh * q.conv.kw * q.conv.cin;
  { ProfScope ps(this, C_POST_CONV, s); launch_post_prep2(last2_.h, last2_.l, post_mul_, post_add_, im2col2_, n, q.conv.kh, q.conv.kw, q.in_w, q.out_w, q.conv.cin, s, lc_); }
  float* pc = scratch(q.conv_tensor, ws_pc_, (size_t)q.out_w * q.conv.cout, n);
  Pw2Launch pa{};
  pa.a_img = im2col2_; pa.Wimg = post_conv_.tc_img; pa.bias = post_conv_.b;
  pa.out32 = pc; pa.M = n * q.out_w; pa.N = q.conv.cout; pa.K = kpost; pa.rows_per_chunk = q.out_w; pa.act = ACT_RELU;
  { ProfScope ps(this, C_POST_CONV, s); launch_pw2(post_conv_.tc, pa, s, lc_); }
  record(q.conv_tensor, pc, (size_t)q.out_w * q.conv.cout, n);
  float* emb = d_emb ? d_emb : ws_emb_;
  { ProfScope ps(this, C_ROW_MEAN, s); launch_row_mean2(pc, emb, emb2_, n, q.out_w, q.conv.cout, s, lc_); }
  record(q.emb_tensor, emb, (size_t)emb_dim_, n);
  Pw2Launch fa{};
  fa.a_img = emb2_; fa.Wimg = fc_.tc_img; fa.bias = fc_.b; fa.out32 = d_logits;
  fa.M = n; fa.N = n_species_; fa.K = emb_dim_; fa.rows_per_chunk = 1; fa.act = ACT_NONE;
  { ProfScope ps(this, C_FC, s); launch_pw2(fc_.tc, fa, s, lc_); }
  record(logits_tensor_, d_logits, (size_t)n_species_, n);
}

// frontend -> stem -> blocks [0, split_) for one micro-batch; result (the split-point tensor) goes to slot `chunk0` of the mid buffer
void Engine::run_front(const void* d_pcm, int fmt, int n, int chunk0, Lane& L, cudaStream_t s) {
  Work& w = L.w;
  float* ws_partial_ = L.partial; float* ws_fe_ = L.fe;
  { ProfScope ps(this, C_MINMAX, s); launch_minmax(d_pcm, fmt, n, n_samples_, ws_partial_, s, lc_); }
  const size_t fe_sz = (size_t)fe_.n_mel * fe_.n_frames * 2;
  float* fe_out = scratch(fe_tensor_, ws_fe_, fe_sz, n);
  { ProfScope ps(this, C_FRONTEND, s); launch_frontend(fe_, d_pcm, fmt, n, ws_partial_, fe_out, s, lc_); }
  record(fe_tensor_, fe_out, fe_sz, n);
  const size_t stem_sz = (size_t)stem_.out_h * stem_.out_w * 24, mix_sz = stem_sz / 2;
  float* stem_dump = keep_ ? scratch(stem_tensor_, nullptr, stem_sz, n) : nullptr;
  if (v2_) {
    const PatchTiles& p0 = blocks_[0].in_patch;
    uint8_t* cur = scratch_img(mix_tensor_, L.w2.x0, p0.bytes(1), n);
    { ProfScope ps(this, C_STEM_MIX, s); launch_stem_mix(stem_, fe_out, stem_dump, nullptr, n, s, lc_, cur, &p0); }
    if (stem_dump) record(stem_tensor_, stem_dump, stem_sz, n);
    record_patch(mix_tensor_, cur, p0, n);
    run_blocks2(0, split_, cur, n, L.w2, s, mid2_ + (size_t)chunk0 * mid2_bytes_, Planes());
    return;
  }
  float* cur = scratch(mix_tensor_, w.x0, mix_sz, n);
  { ProfScope ps(this, C_STEM_MIX, s); launch_stem_mix(stem_, fe_out, stem_dump, cur, n, s, lc_); }
  if (stem_dump) record(stem_tensor_, stem_dump, stem_sz, n);
  record(mix_tensor_, cur, mix_sz, n);
  float* out = run_blocks(0, split_, cur, n, w, s);
  BNB_CUDA(cudaMemcpyAsync(ws_mid_ + (size_t)chunk0 * mid_sz_, out, (size_t)n * mid_sz_ * sizeof(float), cudaMemcpyDeviceToDevice, s));
}

// blocks [split_, end) -> post conv -> embedding -> FC head over `n` chunks at once
void Engine::run_back(const float* mid, int n, float* d_logits, float* d_emb, cudaStream_t s) {
  if (v2_) { run_back2(n, d_logits, d_emb, s); return; }
  Work& w = work_back_;
  BNB_CUDA(cudaMemcpyAsync(w.x0, mid, (size_t)n * mid_sz_ * sizeof(float), cudaMemcpyDeviceToDevice, s));
  float* cur = run_blocks(split_, (int)blocks_.size(), w.x0, n, w, s);
  const PostPlan& q = post_g_;
  float* pc = scratch(q.conv_tensor, ws_pc_, (size_t)q.out_w * q.conv.cout, n);
  PwArgs pa{};
  pa.A = cur; pa.W = post_conv_.w; pa.bias = post_conv_.b; pa.C = pc; pa.M = n * q.out_w; pa.N = q.conv.cout;
  pa.K = q.conv.kh * q.conv.kw * q.conv.cin; pa.rows_per_chunk = q.out_w; pa.act = ACT_RELU; pa.a_mode = A_CONV3X3_ROW;
  pa.a_mul = post_mul_; pa.a_add = post_add_; pa.a_ch = q.conv.cin; pa.in_w = q.in_w; pa.out_w = q.out_w; pa.cin = q.conv.cin; pa.kw = q.conv.kw;
  if (post_conv_.tc_img) {      // tensor-core GEMM wants a plain matrix: materialise relu(affine) + im2col (40 KB/chunk)
    { ProfScope ps(this, C_POST_CONV, s); launch_post_prep(cur, post_mul_, post_add_, ws_im2col_, n, q.conv.kh, q.conv.kw, q.in_w, q.out_w, q.conv.cin, s, lc_); }
    pa.A = ws_im2col_; pa.a_mode = A_PLAIN; pa.a_mul = nullptr; pa.a_add = nullptr;
  }
  pw(pa, post_conv_, C_POST_CONV, s);
  record(q.conv_tensor, pc, (size_t)q.out_w * q.conv.cout, n);
  float* emb = d_emb ? d_emb : ws_emb_;
  { ProfScope ps(this, C_ROW_MEAN, s); launch_row_mean(pc, emb, n, q.out_w, q.conv.cout, s, lc_); }
  record(q.emb_tensor, emb, (size_t)emb_dim_, n);
  PwArgs fa{};
  fa.A = emb; fa.W = fc_.w; fa.bias = fc_.b; fa.C = d_logits; fa.M = n; fa.N = n_species_; fa.K = emb_dim_;
  fa.rows_per_chunk = 1; fa.act = ACT_NONE; fa.a_mode = A_PLAIN;
  pw(fa, fc_, C_FC, s);
  record(logits_tensor_, d_logits, (size_t)n_species_, n);
}

void Engine::predict_device(const void* d_pcm, int fmt, int B, float* d_logits, float* d_emb, cudaStream_t s) {
  BNB_CUDA(cudaSetDevice(device_));
  if (!s) s = compute_;
  ws_mid_ = ws_mid_base_; mid2_ = mid2_base_;
  BNB_CUDA(cudaStreamWaitEvent(s, ev_small_, 0));
  views_.clear();
  const size_t cb = (size_t)n_samples_ * fmt_bytes(fmt);
  const int lanes = keep_ ? 1 : n_lanes_;
  for (int j = 0; j < B; j += max_batch_) {                       // back-phase capacity
    const int nb = std::min(max_batch_, B - j);
    // fan the micro-batches out over the lane streams, join before the back phase
    BNB_CUDA(cudaEventRecord(ev_in_, s));
    for (int l = 0; l < lanes; ++l) BNB_CUDA(cudaStreamWaitEvent(lanes_[l].stream, ev_in_, 0));
    int mi = 0;
    for (int i = 0; i < nb; i += micro_, ++mi) {
      const int n = std::min(micro_, nb - i);
      Lane& L = lanes_[mi % lanes];
      run_front(static_cast<const char*>(d_pcm) + (size_t)(j + i) * cb, fmt, n, i, L, L.stream);
    }
    for (int l = 0; l < lanes; ++l) { BNB_CUDA(cudaEventRecord(lanes_[l].done, lanes_[l].stream)); BNB_CUDA(cudaStreamWaitEvent(s, lanes_[l].done, 0)); }
    run_back(ws_mid_, nb, d_logits + (size_t)j * n_species_, d_emb ? d_emb + (size_t)j * emb_dim_ : nullptr, s);
  }
}

void Engine::analyze_device(const void* d_pcm, int fmt, int B, float sensitivity, int k, int32_t* d_idx, float* d_conf,
                            float* d_logits_or_null, cudaStream_t s) {
  BNB_CUDA(cudaSetDevice(device_));
  if (!s) s = compute_;
  float* lg = d_logits_or_null;
  if (!lg) { ensure_host_staging(); lg = slots_[0].d_logits; }
  predict_device(d_pcm, fmt, B, lg, nullptr, s);
  { ProfScope ps(this, C_TOPK, s); launch_sigmoid_topk(lg, B, n_species_, sensitivity, k, d_idx, d_conf, s, lc_); }
}

// ------------------------------------------------------------------------------------------------
// Host-buffer entry points.  Two submission slots (device input, device outputs, pinned staging, split-point buffer each) so
// that ONE caller can keep two batches in flight: the host-to-device copy of batch i+1 (copy stream) and its front phase
// (lane streams) overlap the back phase of batch i (compute stream).  The synchronous calls are submit + wait.
void Engine::ensure_host_staging() {
  if (slots_[0].d_in) return;
  const size_t mbs = (size_t)max_batch_;
  auto dm = [&](size_t bytes) { void* p = nullptr; BNB_CUDA(cudaMalloc(&p, bytes)); allocs_.push_back(p); return p; };
  topk_cap_ = 64;
  const int n_micro = ceil_div(max_batch_, micro_);
  for (int i = 0; i < 2; ++i) {
    Slot& S = slots_[i];
    S.d_in = dm(mbs * n_samples_ * 4);
    S.d_logits = static_cast<float*>(dm(mbs * n_species_ * 4));
    S.d_emb = static_cast<float*>(dm(mbs * emb_dim_ * 4));
    S.d_idx = static_cast<int32_t*>(dm(mbs * topk_cap_ * 4));
    S.d_conf = static_cast<float*>(dm(mbs * topk_cap_ * 4));
    S.h_in_bytes = mbs * n_samples_ * 4;
    S.off_emb = mbs * (size_t)n_species_ * 4; S.off_idx = S.off_emb + mbs * (size_t)emb_dim_ * 4; S.off_conf = S.off_idx + mbs * topk_cap_ * 4;
    BNB_CUDA(cudaHostAlloc(&S.h_in, S.h_in_bytes, cudaHostAllocDefault));
    BNB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&S.h_out), S.off_conf + mbs * topk_cap_ * 4, cudaHostAllocDefault));
    S.ev_h2d.resize(n_micro + 4);      // + the ramp of small micro-batches at the head of a host call
    for (auto& e : S.ev_h2d) BNB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    BNB_CUDA(cudaEventCreate(&S.start));
    BNB_CUDA(cudaEventCreate(&S.done));
    if (i == 0) { S.mid = ws_mid_base_; S.mid2 = mid2_base_; }             // slot 0 shares the split-point buffer with the device entry points
    else if (v2_) S.mid2 = alloc_img(mbs * mid2_bytes_);
    else { void* p = dm(mbs * mid_sz_ * sizeof(float)); S.mid = static_cast<float*>(p); }
  }
}

namespace {
bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
}  // namespace

void Engine::predict_host(const void* pcm, int fmt, int B, float* logits, float* emb) {
  wait_host(submit_host(pcm, fmt, B, 0.f, 0, nullptr, nullptr, logits, emb));
}

void Engine::analyze_host(const void* pcm, int fmt, int B, float sensitivity, int k, int32_t* idx, float* conf, float* logits) {
  wait_host(submit_host(pcm, fmt, B, sensitivity, k, idx, conf, logits, nullptr));
}

// the kernel chain of one small batch (B <= micro batch) on ONE stream: what a CUDA graph captures
void Engine::small_batch_chain(Slot& S, int fmt, int B, float sensitivity, int k, cudaStream_t s) {
  ws_mid_ = S.mid; mid2_ = S.mid2;
  run_front(S.d_in, fmt, B, 0, lanes_[0], s);
  run_back(ws_mid_, B, S.d_logits, S.d_emb, s);
  if (k > 0) { ProfScope ps(this, C_TOPK, s); launch_sigmoid_topk(S.d_logits, B, n_species_, sensitivity, k, S.d_idx, S.d_conf, s, lc_); }
}

int Engine::submit_host(const void* pcm, int fmt, int B, float sensitivity, int k, int32_t* idx, float* conf, float* logits, float* emb) {
  BNB_CUDA(cudaSetDevice(device_));
  ensure_host_staging();
  if (k > topk_cap_) throw std::invalid_argument("k exceeds the top-k capacity (64)");
  const int si = next_slot_;
  Slot& S = slots_[si];
  if (S.busy) throw std::invalid_argument("two submissions are already outstanding: wait for the older ticket first");
  const size_t cb = (size_t)n_samples_ * fmt_bytes(fmt);
  const bool src_pinned = is_pinned(pcm);
  views_.clear();
  ws_mid_ = S.mid; mid2_ = S.mid2;
  S.B = B; S.k = k; S.u_idx = idx; S.u_conf = conf; S.u_logits = logits; S.u_emb = emb;
  const int lanes = keep_ ? 1 : n_lanes_;
  const bool small = B <= micro_ && !keep_;
  if (small) {
    // ---- one micro-batch: everything on the compute stream, replayed from a CUDA graph when enabled (batch-1 latency) ----
    const char* src = static_cast<const char*>(pcm);
    if (!src_pinned) { memcpy(S.h_in, src, (size_t)B * cb); src = static_cast<const char*>(S.h_in); }
    BNB_CUDA(cudaEventRecord(S.start, compute_));
    for (int l = 0; l < lanes; ++l) BNB_CUDA(cudaStreamWaitEvent(compute_, lanes_[l].done, 0));        // lane 0's workspaces are free
    BNB_CUDA(cudaMemcpyAsync(S.d_in, src, (size_t)B * cb, cudaMemcpyHostToDevice, compute_));
    if (use_graphs_ && !profiling_) {
      uint32_t sbits; memcpy(&sbits, &sensitivity, 4);
      const GraphKey key{si, fmt, B, k, sbits};
      auto it = graphs_.find(key);
      if (it == graphs_.end()) {
        const long long l0 = lc_.n;
        cudaGraph_t g = nullptr;
        BNB_CUDA(cudaStreamBeginCapture(compute_, cudaStreamCaptureModeThreadLocal));
        try { small_batch_chain(S, fmt, B, sensitivity, k, compute_); }
        catch (...) { cudaStreamEndCapture(compute_, &g); if (g) cudaGraphDestroy(g); throw; }
        BNB_CUDA(cudaStreamEndCapture(compute_, &g));
        GraphEntry ge; ge.kernels = lc_.n - l0;
        cudaError_t ie = cudaGraphInstantiate(&ge.exec, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) throw cuda_error(ie, "cudaGraphInstantiate", __FILE__, __LINE__);
        lc_.n = l0;
        if (graphs_.size() > 64) { for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second.exec); graphs_.clear(); }
        it = graphs_.emplace(key, ge).first;
      }
      BNB_CUDA(cudaGraphLaunch(it->second.exec, compute_));
      lc_.n += it->second.kernels;
    } else {
      small_batch_chain(S, fmt, B, sensitivity, k, compute_);
    }
    BNB_CUDA(cudaEventRecord(ev_small_, compute_));                      // lane 0's workspaces were used on the compute stream
  } else {
    BNB_CUDA(cudaEventRecord(S.start, compute_));
    for (int l = 0; l < lanes; ++l) BNB_CUDA(cudaStreamWaitEvent(lanes_[l].stream, ev_small_, 0));
    // H2D per micro-batch on the copy stream; a lane starts a micro-batch as soon as its samples have landed.
    // The first copy cannot overlap compute of THIS batch, so the head ramps up: micro/4, micro/4, micro/2, then full
    // micro-batches (chunks are independent: any split gives the same bits).
    int mi = 0;
    for (int i = 0; i < B; ++mi) {
      int n = micro_;
      static const int ramp = getenv("BNB_RAMP") ? atoi(getenv("BNB_RAMP")) : 1;     // 0 = none, 1 = /4 /4 /2, 2 = /2 (tuning knob)
      // no ramp when another submission is still in flight: this call's copies already overlap THAT call's kernels, and the small
      // head micro-batches would only cost kernel efficiency (r02: e2e 0.92 -> see profiles/README.md)
      if (micro_ >= 16 && B > micro_ && !slots_[si ^ 1].busy) {
        if (ramp == 1) n = mi < 2 ? micro_ / 4 : (mi == 2 ? micro_ / 2 : micro_);
        else if (ramp == 2) n = mi == 0 ? micro_ / 2 : micro_;
      }
      n = std::min(n, B - i);
      const char* src = static_cast<const char*>(pcm) + (size_t)i * cb;
      if (!src_pinned) {   // callee copies (process.go:280-291): stage through our pinned buffer
        memcpy(static_cast<char*>(S.h_in) + (size_t)i * cb, src, (size_t)n * cb);
        src = static_cast<const char*>(S.h_in) + (size_t)i * cb;
      }
      BNB_CUDA(cudaMemcpyAsync(static_cast<char*>(S.d_in) + (size_t)i * cb, src, (size_t)n * cb, cudaMemcpyHostToDevice, copy_));
      BNB_CUDA(cudaEventRecord(S.ev_h2d[mi], copy_));
      Lane& L = lanes_[mi % lanes];
      BNB_CUDA(cudaStreamWaitEvent(L.stream, S.ev_h2d[mi], 0));
      run_front(static_cast<const char*>(S.d_in) + (size_t)i * cb, fmt, n, i, L, L.stream);
      i += n;
    }
    for (int l = 0; l < lanes; ++l) { BNB_CUDA(cudaEventRecord(lanes_[l].done, lanes_[l].stream)); BNB_CUDA(cudaStreamWaitEvent(compute_, lanes_[l].done, 0)); }
    run_back(ws_mid_, B, S.d_logits, S.d_emb, compute_);
    if (k > 0) { ProfScope ps(this, C_TOPK, compute_); launch_sigmoid_topk(S.d_logits, B, n_species_, sensitivity, k, S.d_idx, S.d_conf, compute_, lc_); }
  }
  // results -> pinned staging (asynchronous); wait_host hands them to the caller's buffers
  if (k > 0 && idx != nullptr) {                                      // idx == nullptr: the top-k stays on the device (detect_host)
    BNB_CUDA(cudaMemcpyAsync(S.h_out + S.off_idx, S.d_idx, (size_t)B * k * 4, cudaMemcpyDeviceToHost, compute_));
    BNB_CUDA(cudaMemcpyAsync(S.h_out + S.off_conf, S.d_conf, (size_t)B * k * 4, cudaMemcpyDeviceToHost, compute_));
  }
  if (logits) BNB_CUDA(cudaMemcpyAsync(S.h_out, S.d_logits, (size_t)B * n_species_ * 4, cudaMemcpyDeviceToHost, compute_));
  if (emb) BNB_CUDA(cudaMemcpyAsync(S.h_out + S.off_emb, S.d_emb, (size_t)B * emb_dim_ * 4, cudaMemcpyDeviceToHost, compute_));
  BNB_CUDA(cudaEventRecord(S.done, compute_));
  S.busy = true; S.ticket = ++tickets_;
  next_slot_ ^= 1;
  return (int)(S.ticket & 0x7fffffff);
}

void Engine::wait_host(int ticket) {
  BNB_CUDA(cudaSetDevice(device_));
  Slot* S = nullptr;
  for (Slot& c : slots_) if (c.busy && (int)(c.ticket & 0x7fffffff) == ticket) S = &c;
  if (!S) throw std::invalid_argument("unknown or already completed ticket");
  BNB_CUDA(cudaEventSynchronize(S->done));
  S->busy = false;
  BNB_CUDA(cudaEventElapsedTime(&last_ms_, S->start, S->done));
  const size_t B = (size_t)S->B, k = (size_t)S->k;
  if (S->k > 0 && S->u_idx) { memcpy(S->u_idx, S->h_out + S->off_idx, B * k * 4); memcpy(S->u_conf, S->h_out + S->off_conf, B * k * 4); }
  if (S->u_logits) memcpy(S->u_logits, S->h_out, B * n_species_ * 4);
  if (S->u_emb) memcpy(S->u_emb, S->h_out + S->off_emb, B * emb_dim_ * 4);
}

int Engine::detect_host(const void* pcm, int fmt, int B, float sensitivity, float threshold, int k, int max_det, int32_t* det_chunk,
                        int32_t* det_idx, float* det_conf, int32_t* counts) {
  BNB_CUDA(cudaSetDevice(device_));
  ensure_host_staging();                          // sets topk_cap_ (first host call of this handle)
  if (k <= 0 || k > topk_cap_) throw std::invalid_argument("k must be in 1..64");
  const size_t cap = (size_t)max_batch_ * topk_cap_;
  if (!d_det_) {                                   // [3][cap] (chunk, idx, conf bits) + counts[max_batch] + length
    const size_t bytes = (3 * cap + (size_t)max_batch_ + 1) * sizeof(int32_t);
    BNB_CUDA(cudaMalloc(&d_det_, bytes));
    BNB_CUDA(cudaMallocHost(&h_det_, bytes));
    det_cap_ = cap;
  }
  const int list_cap = (int)std::min<size_t>(cap, (size_t)std::max(0, max_det));
  const int si = next_slot_;
  const int ticket = submit_host(pcm, fmt, B, sensitivity, k, nullptr, nullptr, nullptr, nullptr);    // chain + sigmoid/top-k, results stay on the device
  Slot& S = slots_[si];
  int32_t* d_chunk = d_det_; int32_t* d_idx = d_det_ + cap; float* d_conf = reinterpret_cast<float*>(d_det_ + 2 * cap);
  int32_t* d_counts = d_det_ + 3 * cap; int32_t* d_n = d_counts + max_batch_;
  launch_compact_detections(S.d_idx, S.d_conf, B, k, threshold, list_cap, d_chunk, d_idx, d_conf, d_counts, d_n, compute_, lc_);
  int32_t* h_counts = h_det_ + 3 * cap; int32_t* h_n = h_counts + max_batch_;
  BNB_CUDA(cudaMemcpyAsync(h_counts, d_counts, ((size_t)max_batch_ + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, compute_));
  BNB_CUDA(cudaStreamSynchronize(compute_));
  wait_host(ticket);
  const int found = *h_n, n = std::min(found, list_cap);
  if (n > 0) {                                      // only the detections cross PCIe: 12 B each
    BNB_CUDA(cudaMemcpyAsync(h_det_, d_chunk, (size_t)n * 4, cudaMemcpyDeviceToHost, compute_));
    BNB_CUDA(cudaMemcpyAsync(h_det_ + cap, d_idx, (size_t)n * 4, cudaMemcpyDeviceToHost, compute_));
    BNB_CUDA(cudaMemcpyAsync(h_det_ + 2 * cap, d_conf, (size_t)n * 4, cudaMemcpyDeviceToHost, compute_));
    BNB_CUDA(cudaStreamSynchronize(compute_));
    memcpy(det_chunk, h_det_, (size_t)n * 4); memcpy(det_idx, h_det_ + cap, (size_t)n * 4); memcpy(det_conf, h_det_ + 2 * cap, (size_t)n * 4);
  }
  if (counts) memcpy(counts, h_counts, (size_t)B * 4);
  return found;
}

// ------------------------------------------------------------------------------------------------
Engine::ProfScope::ProfScope(Engine* eng, int cat, cudaStream_t st) : e(eng), s(st), idx(-1) {
  if (!e->profiling_) return;
  if (e->prof_used_ + 2 > e->prof_ev_.size()) {
    for (int i = 0; i < 2; ++i) { cudaEvent_t ev; BNB_CUDA(cudaEventCreate(&ev)); e->prof_ev_.push_back(ev); }
  }
  idx = (int)e->prof_used_; e->prof_used_ += 2; e->prof_cat_.push_back(cat);
  cudaEventRecord(e->prof_ev_[idx], s);
}
Engine::ProfScope::~ProfScope() { if (idx >= 0) cudaEventRecord(e->prof_ev_[idx + 1], s); }

void Engine::profile_begin() { profiling_ = true; prof_used_ = 0; prof_cat_.clear(); }

int Engine::profile_end(float* ms, long long* launches, int cap) {
  BNB_CUDA(cudaSetDevice(device_));
  BNB_CUDA(cudaDeviceSynchronize());
  profiling_ = false;
  for (int i = 0; i < cap; ++i) { ms[i] = 0.f; launches[i] = 0; }
  prof_last_ms_.assign(prof_cat_.size(), 0.f);
  for (size_t i = 0; i < prof_cat_.size(); ++i) {
    float t = 0.f;
    BNB_CUDA(cudaEventElapsedTime(&t, prof_ev_[2 * i], prof_ev_[2 * i + 1]));
    prof_last_ms_[i] = t;
    const int c = prof_cat_[i];
    if (c < cap) { ms[c] += t; launches[c]++; }
  }
  return C_COUNT;
}

int Engine::profile_launches(float* ms, int* cat, int cap) const {
  const int n = (int)std::min<size_t>(prof_last_ms_.size(), (size_t)cap);
  for (int i = 0; i < n; ++i) { ms[i] = prof_last_ms_[i]; cat[i] = prof_cat_[i]; }
  return n;
}

long long Engine::read_tensor(int tensor, float* out, size_t cap) {
  BNB_CUDA(cudaSetDevice(device_));
  auto it = views_.find(tensor);
  if (it == views_.end()) return BNB_ERR_INVALID_ARGUMENT;
  const size_t n = it->second.per_chunk * (size_t)it->second.chunks;
  if (n > cap) return BNB_ERR_INVALID_ARGUMENT;
  BNB_CUDA(cudaDeviceSynchronize());
  if (it->second.kind == 2 || it->second.kind == 3) {      // pre-tiled hi/lo image: gather + join on the host (test hook)
    const TensorView& v = it->second;
    const long long pixels = (long long)(n / (size_t)v.ch);
    const int c8n = (v.ch + 7) / 8;
    const size_t bytes = v.kind == 2 ? RowTiles::make(v.ch).bytes(pixels) : v.patch.bytes(v.chunks);
    std::vector<uint8_t> img(bytes);
    BNB_CUDA(cudaMemcpy(img.data(), v.img, bytes, cudaMemcpyDeviceToHost));
    const RowTiles rt = RowTiles::make(v.ch);
    for (long long m = 0; m < pixels; ++m)
      for (int c8 = 0; c8 < c8n; ++c8) {
        size_t off_hi, lo_delta;
        if (v.kind == 2) { off_hi = rt.piece(m, c8); lo_delta = 16384; }
        else {
          int b, h, w, ty, tx, pr, pc, st, chunk;
          v.patch.split_pixel((uint32_t)m, &b, &h, &w);
          v.patch.stage_of(c8, &st, &chunk);
          bool found = false;
          v.patch.for_each_tile(h, w, [&](int y, int x, int r, int c) { if (!found) { ty = y; tx = x; pr = r; pc = c; found = true; } });
          off_hi = v.patch.tile_base(b, ty, tx) + v.patch.in_tile(pr, pc, st, chunk); lo_delta = v.patch.st_plane[st];
        }
        const __half* hp = reinterpret_cast<const __half*>(img.data() + off_hi);
        const __half* lp = reinterpret_cast<const __half*>(img.data() + off_hi + lo_delta);
        for (int e = 0; e < 8 && 8 * c8 + e < v.ch; ++e) out[(size_t)m * v.ch + 8 * c8 + e] = __half2float(hp[e]) + __half2float(lp[e]);
      }
    return (long long)n;
  }
  if (it->second.h != nullptr) {            // fp16 hi/lo planes: join and drop the pitch padding on the host
    const TensorView& v = it->second;
    const size_t pixels = n / (size_t)v.ch, elems = pixels * (size_t)v.pitch;
    std::vector<__half> h(elems), l(elems);
    BNB_CUDA(cudaMemcpy(h.data(), v.h, elems * sizeof(__half), cudaMemcpyDeviceToHost));
    BNB_CUDA(cudaMemcpy(l.data(), v.l, elems * sizeof(__half), cudaMemcpyDeviceToHost));
    for (size_t p = 0; p < pixels; ++p)
      for (int c = 0; c < v.ch; ++c) out[p * v.ch + c] = __half2float(h[p * v.pitch + c]) + __half2float(l[p * v.pitch + c]);
    return (long long)n;
  }
  BNB_CUDA(cudaMemcpy(out, it->second.ptr, n * sizeof(float), cudaMemcpyDeviceToHost));
  return (long long)n;
}

}  // namespace bnb
