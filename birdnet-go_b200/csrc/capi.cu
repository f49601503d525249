// capi.cu — the extern "C" surface declared in include/birdnet_b200.h.
//
// Error convention follows the reference's native backends: status int + thread-local message
// (/root/reference/internal/inference/openvino/backend_openvino.go:462-470), never abort.
#include <mutex>
#include <new>
#include <set>

#include "engine.h"
#include "mbconv2.h"
#include "pw2.h"
#include "mbconv_tc.h"

using namespace bnb;

struct bnb_classifier {
  Engine* eng = nullptr;
  bool closed = false;
};

namespace {

thread_local std::string g_last_error;
std::mutex g_init_mu;
bool g_inited = false;
int g_devices = 0;

int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

template <class F>
int guarded(F&& f) {
  try {
    f();
    return BNB_OK;
  } catch (const unsupported_model& e) { return fail(BNB_ERR_UNSUPPORTED_MODEL, e.what());
  } catch (const cuda_error& e) {
    cudaGetLastError();
    return fail(e.code == cudaErrorMemoryAllocation ? BNB_ERR_OUT_OF_MEMORY : BNB_ERR_CUDA, e.what());
  } catch (const std::invalid_argument& e) { return fail(BNB_ERR_INVALID_ARGUMENT, e.what());
  } catch (const std::bad_alloc&) { return fail(BNB_ERR_OUT_OF_MEMORY, "host allocation failed");
  } catch (const std::exception& e) { return fail(BNB_ERR_INTERNAL, e.what());
  } catch (...) { return fail(BNB_ERR_INTERNAL, "unknown exception"); }
}

// Registry of live handles: a destroyed handle is recognised WITHOUT dereferencing it (bnb_classifier_destroy frees the
// object), so use-after-Close reports BNB_ERR_CLOSED like the reference's ErrSessionClosed (onnx.go:89-91) instead of
// reading freed memory.  (A later create may reuse the address; such a handle is then simply live again.)
std::mutex g_live_mu;
std::set<const bnb_classifier*> g_live;

int check_handle(const bnb_classifier* h) {
  if (!h) return fail(BNB_ERR_INVALID_ARGUMENT, "classifier handle is NULL");
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    if (!g_live.count(h)) return fail(BNB_ERR_CLOSED, "classifier is closed");
  }
  if (h->closed || !h->eng) return fail(BNB_ERR_CLOSED, "classifier is closed");
  return BNB_OK;
}

// every entry point leaves the calling thread's current CUDA device as it found it
struct DeviceRestore {
  int prev = -1;
  DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
  ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_batch(const bnb_classifier* h, const void* pcm, int format, int B) {
  if (int rc = check_handle(h)) return rc;
  if (!pcm) return fail(BNB_ERR_INVALID_ARGUMENT, "pcm pointer is NULL");
  if (format != BNB_PCM_F32 && format != BNB_PCM_S16) return fail(BNB_ERR_INVALID_ARGUMENT, "unknown pcm format");
  if (B < 0) return fail(BNB_ERR_INVALID_ARGUMENT, "negative batch size");
  return BNB_OK;
}

}  // namespace

namespace bnb {
int capi_fail(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace bnb

extern "C" {

int bnb_abi_version(void) { return BNB_ABI_VERSION; }

const char* bnb_last_error(void) { return g_last_error.c_str(); }

int bnb_init(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_inited) return BNB_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(BNB_ERR_NO_DEVICE, std::string("b200: no CUDA device: ") + cudaGetErrorString(e)); }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p{};
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  if (ok == 0) return fail(BNB_ERR_NO_DEVICE, "b200: no compute-capability 10.x (Blackwell) device present");
  g_devices = ok; g_inited = true;
  return BNB_OK;
}

int bnb_device_count(void) {
  int rc = bnb_init();
  return rc == BNB_OK ? g_devices : rc;
}

int bnb_classifier_create(const void* tflite, size_t tflite_len, const bnb_options* opts, bnb_classifier** out) {
  if (!out) return fail(BNB_ERR_INVALID_ARGUMENT, "out pointer is NULL");
  *out = nullptr;
  if (!tflite || tflite_len == 0) return fail(BNB_ERR_INVALID_ARGUMENT, "cannot create model from data (0 bytes)");
  bnb_options o{};
  o.device = -1;
  if (opts) {
    if (opts->struct_size != 0 && opts->struct_size > sizeof(bnb_options)) return fail(BNB_ERR_INVALID_ARGUMENT, "bnb_options.struct_size is larger than this library knows");
    memcpy(&o, opts, opts->struct_size ? opts->struct_size : sizeof(bnb_options));
  }
  // model structure problems are reported before any device is touched
  int rc = guarded([&] { TfModel m = parse_tflite(tflite, tflite_len); (void)build_plan(m); });
  if (rc == BNB_ERR_INTERNAL) return fail(BNB_ERR_UNSUPPORTED_MODEL, g_last_error);
  if (rc != BNB_OK) return rc;
  if ((rc = bnb_init()) != BNB_OK) return rc;
  bnb_classifier* h = new (std::nothrow) bnb_classifier();
  if (!h) return fail(BNB_ERR_OUT_OF_MEMORY, "host allocation failed");
  {
    DeviceRestore dr;
    rc = guarded([&] { h->eng = new Engine(tflite, tflite_len, o); });
  }
  if (rc != BNB_OK) { delete h; return rc; }
  { std::lock_guard<std::mutex> lk(g_live_mu); g_live.insert(h); }
  *out = h;
  return BNB_OK;
}

void bnb_classifier_destroy(bnb_classifier* h) {
  if (!h) return;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    if (!g_live.erase(h)) return;                 // already destroyed: Close() is idempotent (tflite/classifier.go:129-134)
  }
  DeviceRestore dr;
  try { delete h->eng; } catch (...) {}
  h->eng = nullptr; h->closed = true;
  delete h;
}

int bnb_num_species(const bnb_classifier* h) { if (int rc = check_handle(h)) return rc; return h->eng->n_species(); }
int bnb_num_samples(const bnb_classifier* h) { if (int rc = check_handle(h)) return rc; return h->eng->n_samples(); }
int bnb_embedding_dim(const bnb_classifier* h) { if (int rc = check_handle(h)) return rc; return h->eng->emb_dim(); }
int bnb_max_batch(const bnb_classifier* h) { if (int rc = check_handle(h)) return rc; return h->eng->max_batch(); }
const char* bnb_runtime_device(const bnb_classifier* h) { return check_handle(h) ? "" : h->eng->device_name(); }
const char* bnb_runtime_precision(const bnb_classifier* h) { return check_handle(h) ? "" : h->eng->precision_name(); }

int bnb_predict(bnb_classifier* h, const float* samples, size_t n_samples, float* logits) {
  return bnb_predict_with_embeddings(h, samples, n_samples, logits, nullptr);
}

int bnb_predict_with_embeddings(bnb_classifier* h, const float* samples, size_t n_samples, float* logits, float* embeddings) {
  if (int rc = check_handle(h)) return rc;
  if (!samples || !logits) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL samples/logits pointer");
  if (n_samples != (size_t)h->eng->n_samples())
    return fail(BNB_ERR_INVALID_ARGUMENT, "input size mismatch: expected " + std::to_string(h->eng->n_samples()) + " samples, got " + std::to_string(n_samples));
  DeviceRestore dr;
  return guarded([&] { h->eng->predict_host(samples, BNB_PCM_F32, 1, logits, embeddings); });
}

int bnb_predict_batch(bnb_classifier* h, const void* pcm, int format, int B, float* logits, float* embeddings) {
  if (int rc = check_batch(h, pcm, format, B)) return rc;
  if (!logits) return fail(BNB_ERR_INVALID_ARGUMENT, "logits pointer is NULL");
  if (B > h->eng->max_batch()) return fail(BNB_ERR_INVALID_ARGUMENT, "batch exceeds max_batch");
  if (B == 0) return BNB_OK;
  DeviceRestore dr;
  return guarded([&] { h->eng->predict_host(pcm, format, B, logits, embeddings); });
}

int bnb_analyze_batch(bnb_classifier* h, const void* pcm, int format, int B, float sensitivity, int k, int32_t* idx, float* conf,
                      float* logits_or_null) {
  if (int rc = check_batch(h, pcm, format, B)) return rc;
  if (!idx || !conf || k <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "idx/conf NULL or k <= 0");
  if (B > h->eng->max_batch()) return fail(BNB_ERR_INVALID_ARGUMENT, "batch exceeds max_batch");
  if (B == 0) return BNB_OK;
  DeviceRestore dr;
  return guarded([&] { h->eng->analyze_host(pcm, format, B, sensitivity, k, idx, conf, logits_or_null); });
}

int bnb_analyze_batch_submit(bnb_classifier* h, const void* pcm, int format, int B, float sensitivity, int k, int32_t* idx, float* conf,
                             float* logits_or_null, int32_t* ticket) {
  if (int rc = check_batch(h, pcm, format, B)) return rc;
  if (!idx || !conf || k <= 0 || !ticket) return fail(BNB_ERR_INVALID_ARGUMENT, "idx/conf/ticket NULL or k <= 0");
  if (B <= 0 || B > h->eng->max_batch()) return fail(BNB_ERR_INVALID_ARGUMENT, "batch must be in 1..max_batch");
  DeviceRestore dr;
  return guarded([&] { *ticket = h->eng->submit_host(pcm, format, B, sensitivity, k, idx, conf, logits_or_null, nullptr); });
}

int bnb_analyze_batch_detections(bnb_classifier* h, const void* pcm, int format, int B, float sensitivity, float threshold, int k,
                                 int max_det, int32_t* det_chunk, int32_t* det_idx, float* det_conf, int32_t* counts_or_null, int32_t* n_det) {
  if (int rc = check_batch(h, pcm, format, B)) return rc;
  if (!n_det || k <= 0 || max_det < 0 || (max_det > 0 && (!det_chunk || !det_idx || !det_conf))) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL output or k <= 0");
  if (B > h->eng->max_batch()) return fail(BNB_ERR_INVALID_ARGUMENT, "batch exceeds max_batch");
  *n_det = 0;
  if (B == 0) return BNB_OK;
  DeviceRestore dr;
  return guarded([&] { *n_det = h->eng->detect_host(pcm, format, B, sensitivity, threshold, k, max_det, det_chunk, det_idx, det_conf, counts_or_null); });
}

int bnb_wait(bnb_classifier* h, int32_t ticket) {
  if (int rc = check_handle(h)) return rc;
  DeviceRestore dr;
  return guarded([&] { h->eng->wait_host(ticket); });
}

int bnb_predict_batch_device(bnb_classifier* h, const void* d_pcm, int format, int B, float* d_logits, float* d_embeddings, void* stream) {
  if (int rc = check_batch(h, d_pcm, format, B)) return rc;
  if (!d_logits) return fail(BNB_ERR_INVALID_ARGUMENT, "logits pointer is NULL");
  if (B == 0) return BNB_OK;
  DeviceRestore dr;
  return guarded([&] { h->eng->predict_device(d_pcm, format, B, d_logits, d_embeddings, static_cast<cudaStream_t>(stream)); });
}

int bnb_analyze_batch_device(bnb_classifier* h, const void* d_pcm, int format, int B, float sensitivity, int k, int32_t* d_idx,
                             float* d_conf, float* d_logits_or_null, void* stream) {
  if (int rc = check_batch(h, d_pcm, format, B)) return rc;
  if (!d_idx || !d_conf || k <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "idx/conf NULL or k <= 0");
  if (!d_logits_or_null && B > h->eng->max_batch()) return fail(BNB_ERR_INVALID_ARGUMENT, "batch exceeds max_batch (pass a logits buffer for larger batches)");
  if (B == 0) return BNB_OK;
  DeviceRestore dr;
  return guarded([&] { h->eng->analyze_device(d_pcm, format, B, sensitivity, k, d_idx, d_conf, d_logits_or_null, static_cast<cudaStream_t>(stream)); });
}

int64_t bnb_kernel_launches(const bnb_classifier* h) { if (int rc = check_handle(h)) return rc; return h->eng->launches(); }
float bnb_last_device_ms(const bnb_classifier* h) { return check_handle(h) ? -1.f : h->eng->last_device_ms(); }

int bnb_profile_begin(bnb_classifier* h) {
  if (int rc = check_handle(h)) return rc;
  h->eng->profile_begin();
  return BNB_OK;
}

int bnb_profile_end(bnb_classifier* h, float* ms, int64_t* launches, int cap) {
  if (int rc = check_handle(h)) return rc;
  if (!ms || !launches || cap <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL ms/launches or cap <= 0");
  int n = 0;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  int rc = guarded([&] { n = h->eng->profile_end(ms, reinterpret_cast<long long*>(launches), cap); });
  return rc != BNB_OK ? rc : n;
}

int bnb_profile_launches(bnb_classifier* h, float* ms, int32_t* cat, int cap) {
  if (int rc = check_handle(h)) return rc;
  if (!ms || !cat || cap <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL ms/cat or cap <= 0");
  return h->eng->profile_launches(ms, cat, cap);
}

int bnb_debug_pw_tiling(int M, int N, int K, int* bn, int* stages, int64_t* smem_bytes) {
  if (!bn || !stages || !smem_bytes || M <= 0 || N <= 0 || K <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "bad tiling query");
  PwTcLayer L; L.N = N; L.K = K; L.n_pad = (N + 15) / 16 * 16; L.k_pad = (K + 15) / 16 * 16; L.k_stages = (K + 63) / 64;
  size_t sm = 0;
  pw_tc_tiling(L, M, bn, stages, &sm);
  *smem_bytes = (int64_t)sm;
  return BNB_OK;
}

int bnb_debug_mbconv_geometry(int H, int W, int Ho, int Wo, int stride, int Cin, int C, int B, int max_tiles, int* out10, int64_t* smem_bytes) {
  if (!out10 || !smem_bytes) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL out");
  const MbGeom g = mbconv_geometry(H, W, Ho, Wo, stride, Cin, C, B, max_tiles);
  const int v[10] = {g.th, g.tw, g.ph, g.pw, g.tiles_h, g.tiles_w, g.k_stages, g.box_c, g.a_slots, g.b_slots};
  for (int i = 0; i < 10; ++i) out10[i] = v[i];
  *smem_bytes = (int64_t)g.smem_bytes;
  return BNB_OK;
}

// ---- N2: bat pipeline helpers (no classifier handle: they work on raw PCM / embeddings) ------------------------------------
namespace {
struct DevMem {
  void* p = nullptr;
  explicit DevMem(size_t n) { BNB_CUDA(cudaMalloc(&p, n ? n : 1)); }
  ~DevMem() { if (p) cudaFree(p); }
  DevMem(const DevMem&) = delete; DevMem& operator=(const DevMem&) = delete;
};
int pick_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return -1; }
  return device < 0 ? 0 : (device < n ? device : -1);
}
}  // namespace

int bnb_ultrasonic_cv_batch(int device, const void* pcm, int format, int B, int n_samples, int sample_rate, int fft_size, int hop_size,
                            int frequency_split_hz, double* cv, int32_t* ok) {
  if (!pcm || !cv || !ok || B < 0 || n_samples < 0 || (format != BNB_PCM_F32 && format != BNB_PCM_S16)) return fail(BNB_ERR_INVALID_ARGUMENT, "bad ultrasonic filter arguments");
  const int frames = ultrasonic_frames(n_samples, sample_rate, fft_size, hop_size, frequency_split_hz);
  if (frames == 0 || fft_size > 8192) {                       // the reference's (0, false): not an error (filter.go:21-39)
    for (int b = 0; b < B; ++b) { cv[b] = 0.0; ok[b] = 0; }
    return frames == 0 ? BNB_OK : fail(BNB_ERR_INVALID_ARGUMENT, "FFT sizes above 8192 are not supported by the GPU filter");
  }
  if (B == 0) return BNB_OK;
  const int dev = pick_device(device);
  if (dev < 0) return fail(BNB_ERR_NO_DEVICE, "no usable CUDA device");
  DeviceRestore dr;
  return guarded([&] {
    BNB_CUDA(cudaSetDevice(dev));
    const size_t in_bytes = (size_t)B * n_samples * (format == BNB_PCM_S16 ? 2 : 4);
    DevMem d_in(in_bytes), ws(ultrasonic_workspace_bytes(B, frames, fft_size));
    BNB_CUDA(cudaMemcpy(d_in.p, pcm, in_bytes, cudaMemcpyHostToDevice));
    double* d_cv = reinterpret_cast<double*>(static_cast<char*>(ws.p) + ultrasonic_workspace_bytes(B, frames, fft_size) - (size_t)B * sizeof(double));
    launch_ultrasonic_cv(d_in.p, format, B, n_samples, sample_rate, fft_size, hop_size, frequency_split_hz, ws.p, d_cv, nullptr);
    BNB_CUDA(cudaMemcpy(cv, d_cv, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) ok[b] = 1;
  });
}

int bnb_dense_head_batch(int device, const float* embeddings, int B, int n_in, const float* weights, const float* bias, int n_out, float* scores) {
  if (!embeddings || !weights || !bias || !scores || B < 0 || n_in <= 0 || n_out <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "bad dense head arguments");
  if (B == 0) return BNB_OK;
  const int dev = pick_device(device);
  if (dev < 0) return fail(BNB_ERR_NO_DEVICE, "no usable CUDA device");
  DeviceRestore dr;
  return guarded([&] {
    BNB_CUDA(cudaSetDevice(dev));
    DevMem d_e((size_t)B * n_in * 4), d_w((size_t)n_out * n_in * 4), d_b((size_t)n_out * 4), d_o((size_t)B * n_out * 4);
    BNB_CUDA(cudaMemcpy(d_e.p, embeddings, (size_t)B * n_in * 4, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(d_w.p, weights, (size_t)n_out * n_in * 4, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(d_b.p, bias, (size_t)n_out * 4, cudaMemcpyHostToDevice));
    launch_dense_head(static_cast<float*>(d_e.p), static_cast<float*>(d_w.p), static_cast<float*>(d_b.p), B, n_in, n_out, static_cast<float*>(d_o.p), nullptr);
    BNB_CUDA(cudaMemcpy(scores, d_o.p, (size_t)B * n_out * 4, cudaMemcpyDeviceToHost));
  });
}

int bnb_debug_mb2_plan(int H, int W, int Ho, int Wo, int stride, int Cin, int C, int* out12, int64_t* smem_bytes) {
  if (!out12 || !smem_bytes) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL out");
  const Mb2Plan P = mb2_plan(H, W, Ho, Wo, stride, Cin, C);
  const int v[12] = {P.ok ? 1 : 0, P.TH, P.TW, P.PH, P.PW, P.n_mma, P.n_units, P.k_stages, P.a_resident, P.a_slots, P.b_slots, (int)P.b_pair_bytes};
  for (int i = 0; i < 12; ++i) out12[i] = v[i];
  *smem_bytes = (int64_t)P.smem_bytes;
  return BNB_OK;
}

int bnb_debug_pw2_tiling(int M, int N, int K, int gated, int* out4, int64_t* smem_bytes) {
  if (!out4 || !smem_bytes || M <= 0 || N <= 0 || K <= 0) return fail(BNB_ERR_INVALID_ARGUMENT, "bad tiling query");
  PwTcLayer L; L.N = N; L.K = K; L.n_pad = (N + 15) / 16 * 16; L.k_pad = (K + 15) / 16 * 16; L.k_stages = (K + 63) / 64;
  int bn = 0, stages = 0, b_res = 0; size_t sm = 0;
  pw2_tiling(L, M, gated != 0, &bn, &stages, &sm, &b_res);
  out4[0] = bn; out4[1] = stages; out4[2] = b_res; out4[3] = (L.n_pad + bn - 1) / bn;
  *smem_bytes = (int64_t)sm;
  return BNB_OK;
}

int bnb_describe_model(const void* tflite, size_t tflite_len, char* json, size_t cap) {
  if (!tflite || !json) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL pointer");
  std::string s;
  int rc = guarded([&] { TfModel m = parse_tflite(tflite, tflite_len); s = describe_plan(build_plan(m)); });
  if (rc == BNB_ERR_INTERNAL) return fail(BNB_ERR_UNSUPPORTED_MODEL, g_last_error);
  if (rc != BNB_OK) return rc;
  if (s.size() + 1 > cap) return fail(BNB_ERR_INVALID_ARGUMENT, "json buffer too small");
  memcpy(json, s.c_str(), s.size() + 1);
  return (int)s.size();
}

int64_t bnb_debug_read_tensor(bnb_classifier* h, int tensor, float* out, size_t cap) {
  if (int rc = check_handle(h)) return rc;
  if (!out) return fail(BNB_ERR_INVALID_ARGUMENT, "NULL out pointer");
  long long n = 0;
  int rc = guarded([&] { n = h->eng->read_tensor(tensor, out, cap); });
  if (rc != BNB_OK) return rc;
  if (n < 0) return fail((int)n, "tensor not materialised by the last call (or buffer too small)");
  return n;
}

int bnb_debug_keep_intermediates(bnb_classifier* h, int on) {
  if (int rc = check_handle(h)) return rc;
  h->eng->keep_intermediates(on != 0);
  return BNB_OK;
}

}  // extern "C"
