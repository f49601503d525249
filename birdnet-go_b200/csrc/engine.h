// engine.h — device-side state of one classifier handle (weights, workspaces, streams) and the
// kernel chain that replaces `interpreter.Invoke()`
// (/root/reference/internal/inference/tflite/classifier.go:107).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <tuple>
#include <memory>
#include <string>
#include <vector>

#include "../../include/birdnet_b200.h"
#include "kernels.h"
#include "net_plan.h"
#include "mbconv2.h"
#include "pw2.h"
#include "pw_tc.h"

namespace bnb {

struct DevConv { const float* w = nullptr; const float* b = nullptr; PwTcLayer tc; const uint8_t* tc_img = nullptr; };

struct DevBlock {
  BlockPlan g;          // geometry + tensor ids (host weight pointers are dead after upload)
  DevConv expand, dw, se1, se2, proj;
  // F16X3 path (mbconv2.cu): tile plan, (unit, stage) weight image, expand bias padded to whole units, and the PatchTiles
  // layout of the block INPUT (what the producer of this block's input must write)
  Mb2Plan mb2; const uint8_t* mb2_img = nullptr; const float* mb2_bias = nullptr; PatchTiles in_patch;
};

// fp16 hi / lo planes of one activation tensor (x ~= hi + lo); `pitch` = channels per pixel in memory (multiple of 8)
struct Planes { __half* h = nullptr; __half* l = nullptr; };
inline int plane_pitch(int c) { return (c + 7) / 8 * 8; }

struct TensorView {
  const float* ptr = nullptr; size_t per_chunk = 0; int chunks = 0;
  const __half* h = nullptr; const __half* l = nullptr; int pitch = 0, ch = 0;     // plain plane tensors: per_chunk counts pixels * ch
  const uint8_t* img = nullptr; int kind = 0;                                      // kind 2: RowTiles image (ch = K), 3: PatchTiles image
  PatchTiles patch;
};

class Engine {
 public:
  Engine(const void* tflite, size_t len, const bnb_options& opts);
  ~Engine();
  void init(const void* tflite, size_t len, const bnb_options& opts);
  void release() noexcept;
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  int n_species() const { return n_species_; }
  int n_samples() const { return n_samples_; }
  int emb_dim() const { return emb_dim_; }
  int max_batch() const { return max_batch_; }
  int device() const { return device_; }
  const char* device_name() const { return device_name_.c_str(); }
  const char* precision_name() const { return precision_name_.c_str(); }
  long long launches() const { return lc_.n; }
  float last_device_ms() const { return last_ms_; }

  // device-resident inputs/outputs, enqueued on `s` (nullptr -> own compute stream), no sync
  void predict_device(const void* d_pcm, int fmt, int B, float* d_logits, float* d_emb, cudaStream_t s);
  void analyze_device(const void* d_pcm, int fmt, int B, float sensitivity, int k, int32_t* d_idx, float* d_conf,
                      float* d_logits_or_null, cudaStream_t s);
  // host buffers: pinned staging, H2D/compute overlap per micro-batch, D2H, synchronous
  void predict_host(const void* pcm, int fmt, int B, float* logits, float* emb);
  void analyze_host(const void* pcm, int fmt, int B, float sensitivity, int k, int32_t* idx, float* conf, float* logits);
  // asynchronous form: returns a ticket once the work is enqueued; the outputs are valid after wait_host(ticket).  At most two
  // tickets may be outstanding.  k == 0 -> no top-k; logits / emb may be null.
  int submit_host(const void* pcm, int fmt, int B, float sensitivity, int k, int32_t* idx, float* conf, float* logits, float* emb);
  void wait_host(int ticket);
  // N1: the same path, but sigma >= threshold and the compaction of the per-chunk top-k into ONE detection list run on the
  // device; only counts[B], the list length and the list itself cross PCIe (returns the number of detections found, which may
  // exceed max_det: the list is then truncated, counts[] is not)
  int detect_host(const void* pcm, int fmt, int B, float sensitivity, float threshold, int k, int max_det, int32_t* det_chunk,
                  int32_t* det_idx, float* det_conf, int32_t* counts);

  void keep_intermediates(bool on) { keep_ = on; }
  // per-category device timing (CUDA events around every launch); see bnb_profile_* in the C ABI
  enum Cat : int { C_MINMAX = 0, C_FRONTEND, C_STEM_MIX, C_PW_EXPAND, C_DW, C_SE, C_PW_PROJECT, C_POST_CONV, C_ROW_MEAN, C_FC, C_TOPK, C_COUNT };
  void profile_begin();
  int profile_end(float* ms, long long* launches, int cap);
  int profile_launches(float* ms, int* cat, int cap) const;   // per launch, in issue order, of the last profile_end
  long long read_tensor(int tensor, float* out, size_t cap);

 private:
  void pw(const PwArgs& a, const DevConv& c, int cat, cudaStream_t s);
  static constexpr int kMaxDwParts = 32;
  struct Work { float *x0 = nullptr, *x1 = nullptr, *e = nullptr, *d = nullptr, *g = nullptr, *sep = nullptr; size_t cap_n = 0; };
  // F16X3 path: x0/x1 = block input / output as PatchTiles images, d = depthwise output as a RowTiles image (bytes per chunk)
  struct Work2 { uint8_t *x0 = nullptr, *x1 = nullptr, *d = nullptr; float* g = nullptr; float* sep = nullptr; size_t x_bytes = 0, d_bytes = 0; };
  static constexpr int kMaxLanes = 4;
  struct Lane { Work w; Work2 w2; float* partial = nullptr; float* fe = nullptr; cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; };
  float* run_blocks(int lo, int hi, float* cur, int n, Work& w, cudaStream_t s);
  // ---- F16X3 path on fp16 hi/lo planes (mbconv2.cu + pw2.cu) ----
  // blocks [lo, hi): input `cur` (PatchTiles image of block lo).  The last block writes `final_img` (the PatchTiles image of
  // block hi, when hi < #blocks) or, for the last block of the network, the plain planes `final_plain`.
  void run_blocks2(int lo, int hi, const uint8_t* cur, int n, Work2& w, cudaStream_t s, uint8_t* final_img, Planes final_plain);
  void run_back2(int n, float* d_logits, float* d_emb, cudaStream_t s);
  uint8_t* scratch_img(int tensor_id, uint8_t* normal, size_t bytes_per_chunk, int n);
  void record_rows(int tensor_id, const uint8_t* img, int rows_per_chunk, int K, int n) {
    if (tensor_id >= 0) { TensorView v; v.per_chunk = (size_t)rows_per_chunk * K; v.chunks = n; v.img = img; v.kind = 2; v.ch = K; views_[tensor_id] = v; }
  }
  void record_patch(int tensor_id, const uint8_t* img, const PatchTiles& t, int n) {
    if (tensor_id >= 0) { TensorView v; v.per_chunk = (size_t)t.H * t.W * t.C; v.chunks = n; v.img = img; v.kind = 3; v.ch = t.C; v.patch = t; views_[tensor_id] = v; }
  }
  void record_planes(int tensor_id, Planes p, int pixels, int ch, int n) {
    if (tensor_id >= 0) { TensorView v; v.per_chunk = (size_t)pixels * ch; v.chunks = n; v.h = p.h; v.l = p.l; v.pitch = plane_pitch(ch); v.ch = ch; views_[tensor_id] = v; }
  }
  Planes alloc_planes(size_t elems);
  uint8_t* alloc_img(size_t bytes);
  bool v2_ = false;        // F16X3 precision: every block runs mbconv2 + pw2 on pre-tiled fp16 hi/lo images
  Work2 work2_back_; uint8_t* mid2_ = nullptr; size_t mid2_bytes_ = 0; uint8_t *im2col2_ = nullptr, *emb2_ = nullptr; Planes last2_;
  void run_front(const void* d_pcm, int fmt, int n, int chunk0, Lane& L, cudaStream_t s);   // chunk0: slot of the first chunk in the split-point buffer
  void run_back(const float* mid, int n, float* d_logits, float* d_emb, cudaStream_t s);
  float* scratch(int tensor_id, float* normal, size_t per_chunk, int n);
  void record(int tensor_id, const float* p, size_t per_chunk, int n) { if (tensor_id >= 0) { TensorView v; v.ptr = p; v.per_chunk = per_chunk; v.chunks = n; views_[tensor_id] = v; } }
  void upload_weights(const NetPlan& P);
  void alloc_workspace();
  void ensure_host_staging();
  const float* up(const float* host, size_t n);  // copy into the weight arena, returns device ptr
  template <class T> const T* up_t(const T* host, size_t n);

  struct ProfScope {
    Engine* e; cudaStream_t s; int idx;
    ProfScope(Engine* eng, int cat, cudaStream_t st);
    ~ProfScope();
  };
  bool profiling_ = false;
  std::vector<cudaEvent_t> prof_ev_;          // pairs (start, stop)
  std::vector<int> prof_cat_;
  std::vector<float> prof_last_ms_;
  size_t prof_used_ = 0;

  int device_ = 0, n_species_ = 0, n_samples_ = 0, emb_dim_ = 0, max_batch_ = 256, micro_ = 32, precision_ = BNB_PRECISION_F32;
  std::string device_name_, precision_name_;
  LaunchCounter lc_;
  float last_ms_ = 0.f;
  bool keep_ = false;
  bool fused_force_ = false;
  bool fused_ = true;      // fused expand+depthwise tcgen05 kernel (BNB_FUSED=0 falls back to the two-kernel chain)

  // weights
  std::vector<void*> allocs_;                 // every cudaMalloc owned by the handle
  FrontendDev fe_{};
  StemMixDev stem_{};
  int stem_tensor_ = -1, fe_tensor_ = -1, mix_tensor_ = -1;
  std::vector<DevBlock> blocks_;
  PostPlan post_g_{}; DevConv post_conv_; const float* post_mul_ = nullptr; const float* post_add_ = nullptr;
  DevConv fc_; int logits_tensor_ = -1;

  // workspaces (capacity: micro_ chunks)
  float *ws_mid_ = nullptr, *ws_pc_ = nullptr, *ws_emb_ = nullptr, *ws_im2col_ = nullptr;
  Lane lanes_[kMaxLanes]; // front phase: micro-batches round-robin over `n_lanes_` streams so kernel ramps/tails overlap
  int n_lanes_ = 2;
  cudaEvent_t ev_in_ = nullptr;
  Work work_back_;        // back phase (whole batch)
  int split_ = 0;         // first block of the back phase
  size_t mid_sz_ = 0;     // floats per chunk of the split-point tensor
  std::map<int, std::pair<float*, size_t>> keep_bufs_;   // tensor id -> (device buffer, capacity in floats)
  std::map<int, std::pair<uint8_t*, size_t>> keep_imgs_;  // tensor id -> (image buffer, capacity in bytes)
  std::map<int, TensorView> views_;

  // full-batch device buffers for the host path
  // host-path submission slots (two batches in flight: copy + front of i+1 overlap the back phase of i)
  struct Slot {
    void* d_in = nullptr; float* d_logits = nullptr; float* d_emb = nullptr; int32_t* d_idx = nullptr; float* d_conf = nullptr;
    void* h_in = nullptr; uint8_t* h_out = nullptr; size_t h_in_bytes = 0, off_emb = 0, off_idx = 0, off_conf = 0;
    std::vector<cudaEvent_t> ev_h2d; cudaEvent_t start = nullptr, done = nullptr;
    float* mid = nullptr; uint8_t* mid2 = nullptr;          // split-point buffer of this slot (fp32 chain / plane images)
    bool busy = false; long long ticket = 0; int B = 0, k = 0;
    int32_t* u_idx = nullptr; float *u_conf = nullptr, *u_logits = nullptr, *u_emb = nullptr;   // the caller's output buffers
  };
  Slot slots_[2]; int next_slot_ = 0; long long tickets_ = 0;
  void small_batch_chain(Slot& S, int fmt, int B, float sensitivity, int k, cudaStream_t s);
  // CUDA graphs of the single-micro-batch chain, keyed by (slot, format, batch, k, sensitivity bits)
  struct GraphKey { int slot, fmt, B, k; uint32_t sens; bool operator<(const GraphKey& o) const { return std::tie(slot, fmt, B, k, sens) < std::tie(o.slot, o.fmt, o.B, o.k, o.sens); } };
  struct GraphEntry { cudaGraphExec_t exec = nullptr; long long kernels = 0; };
  std::map<GraphKey, GraphEntry> graphs_;
  bool use_graphs_ = false;
  cudaEvent_t ev_small_ = nullptr;                       // end of the last single-stream (small batch) chain: lane streams order after it
  float* ws_mid_base_ = nullptr; uint8_t* mid2_base_ = nullptr;
  int topk_cap_ = 0;
  // detection compaction buffers (device + pinned host), allocated on first use: [max_batch * topk_cap] triples, counts, length
  int32_t* d_det_ = nullptr; int32_t* h_det_ = nullptr; size_t det_cap_ = 0;
  cudaStream_t compute_ = nullptr, copy_ = nullptr;

};

}  // namespace bnb
