// range_filter.cu — the BirdNET v2.4 range-filter ("meta") model on the GPU: species occurrence scores for a
// (latitude, longitude, week) triple, single and batched.
//
// Replaces inference.RangeFilter / BatchRangeFilter (/root/reference/internal/inference/backend.go:55-76), implemented in
// the reference by a TFLite interpreter (/root/reference/internal/inference/tflite/rangefilter.go:22-94) and used in
// batches by the heat-map builder (/root/reference/internal/classifier/orchestrator.go:1846-1882).
//
// The graph of BirdNET_GLOBAL_6K_V2.4_MData_Model_V2_FP16.tflite (56 ops, fp16 constants behind DEQUANTIZE) is matched
// structurally:   f_i = amp * sin(((x_i +- c0_i) * c1_i) * w + phase[0..47])   for lat, lon, week;
//                 the week features are multiplied by (week > lo ? on : off) * (week < hi ? on : off);
//                 concat(144) -> FC 256 ReLU -> FC 512 ReLU -> FC 1024 ReLU -> FC 6522 -> sigmoid.
// Arithmetic is fp32 with the fp16 weights widened exactly, like TFLite's DEQUANTIZE + float kernels.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/birdnet_b200.h"
#include "common.cuh"
#include "net_plan.h"
#include "tflite_model.h"

namespace bnb {
int capi_fail(int code, const std::string& msg);   // capi.cu

namespace {

constexpr int OP_DEQUANTIZE = 6, OP_SIN = 66, OP_GREATER = 61, OP_LESS = 58, OP_SELECT_V2 = 123;
constexpr int kEnc = 48;

struct RfPlan {
  float add[3], mul1[3], wmul, amp, phase[kEnc];
  float gt_thr, lt_thr, gt_on, gt_off, lt_on, lt_off;       // week mask = (w > gt_thr ? gt_on : gt_off) * (w < lt_thr ? lt_on : lt_off)
  struct Fc { std::vector<float> w, b; int n_out = 0, n_in = 0, relu = 0; } fc[4];
  int n_species = 0;
};

struct RfMatcher {
  const TfModel& m;
  std::vector<int> producer;
  explicit RfMatcher(const TfModel& mm) : m(mm), producer(mm.tensors.size(), -1) {
    for (size_t i = 0; i < m.ops.size(); ++i) for (int o : m.ops[i].out) producer[o] = (int)i;
  }
  [[noreturn]] void fail(const std::string& s) const { throw unsupported_model("range filter: " + s); }
  const TfOp* prod(int t) const { return (t >= 0 && producer[t] >= 0) ? &m.ops[producer[t]] : nullptr; }
  const TfOp& need(int t, int code, const char* what) const {
    const TfOp* p = prod(t);
    if (!p || p->code != code) fail(std::string(what) + ": unexpected producer");
    return *p;
  }
  int skip(int t) const {       // EXPAND_DIMS / RESHAPE are shape plumbing
    for (;;) { const TfOp* p = prod(t); if (p && (p->code == OP_EXPAND_DIMS || p->code == OP_RESHAPE)) t = p->in[0]; else return t; }
  }
  // constant (optionally behind DEQUANTIZE) as floats; empty when `t` is not a constant
  std::vector<float> cval(int t) const {
    if (t < 0) return {};
    const TfOp* p = prod(t);
    if (p && p->code == OP_DEQUANTIZE) t = p->in[0];
    else if (p) return {};
    const TfTensor& c = m.tensors[t];
    std::vector<float> v;
    if (c.type == TT_F32 && c.has_bytes(4)) { v.assign(c.f32(), c.f32() + c.numel()); }
    else if (c.type == TT_F16 && c.has_bytes(2)) {
      v.resize(c.numel());
      const __half* h = reinterpret_cast<const __half*>(c.data);
      for (size_t i = 0; i < v.size(); ++i) v[i] = __half2float(h[i]);
    }
    return v;
  }
  float cscalar(int t, const char* what) const { auto v = cval(t); if (v.size() != 1) fail(std::string(what) + ": expected a scalar constant"); return v[0]; }
  // binary op with one constant operand -> (other operand, constant values)
  int split(const TfOp& op, std::vector<float>* c, const char* what) const {
    auto a = cval(op.in[0]), b = cval(op.in[1]);
    if (a.empty() == b.empty()) fail(std::string(what) + ": expected exactly one constant operand");
    *c = a.empty() ? b : a;
    return a.empty() ? op.in[0] : op.in[1];
  }
  int slice_index(int t, const char* what) const {      // STRIDED_SLICE(input, begin=[0, i], ...) -> i
    const TfOp& ss = need(t, OP_STRIDED_SLICE, what);
    if (ss.in[0] != m.inputs[0]) fail(std::string(what) + ": slice of something else than the input");
    const TfTensor& b = m.tensors[ss.in[1]];
    if (b.type != TT_I32 || !b.has_bytes(4) || b.numel() != 2) fail(std::string(what) + ": slice begin");
    return b.i32()[1];
  }
};

RfPlan build_range_plan(const TfModel& m) {
  RfMatcher M(m);
  RfPlan P{};
  if (m.inputs.size() != 1 || m.outputs.size() != 1) M.fail("expected one input and one output");
  int t = m.outputs[0];
  t = M.need(t, OP_LOGISTIC, "output activation").in[0];
  for (int l = 3; l >= 0; --l) {
    const TfOp& fc = M.need(t, OP_FULLY_CONNECTED, "dense layer");
    RfPlan::Fc& F = P.fc[l];
    F.w = M.cval(fc.in[1]);
    const TfOp* dq = M.prod(fc.in[1]);
    const TfTensor& wt = m.tensors[dq && dq->code == OP_DEQUANTIZE ? dq->in[0] : fc.in[1]];
    if (wt.shape.size() != 2 || F.w.size() != (size_t)wt.shape[0] * wt.shape[1]) M.fail("dense weights");
    F.n_out = wt.shape[0]; F.n_in = wt.shape[1]; F.relu = fc.act == 1 ? 1 : 0;
    if (fc.act != 0 && fc.act != 1) M.fail("dense activation");
    F.b = fc.in.size() > 2 ? M.cval(fc.in[2]) : std::vector<float>();
    if (F.b.empty()) F.b.assign(F.n_out, 0.f);
    if ((int)F.b.size() != F.n_out) M.fail("dense bias");
    t = fc.in[0];
  }
  P.n_species = P.fc[3].n_out;
  for (int l = 1; l < 4; ++l) if (P.fc[l].n_in != P.fc[l - 1].n_out) M.fail("dense chain widths");
  const TfOp& cat = M.need(t, OP_CONCATENATION, "feature concat");
  if (cat.in.size() != 3 || P.fc[0].n_in != 3 * kEnc) M.fail("expected three 48-wide encodings");
  bool seen[3] = {false, false, false};
  bool have_common = false;
  for (int k = 0; k < 3; ++k) {
    const TfOp& top = M.need(cat.in[k], OP_MUL, "encoding scale");
    int enc_t = -1, mask_t = -1;
    std::vector<float> c;
    if (!M.cval(top.in[0]).empty() || !M.cval(top.in[1]).empty()) enc_t = cat.in[k];          // amp * sin(..)
    else {                                                                                       // (amp * sin(..)) * mask
      const int a = M.skip(top.in[0]), b = M.skip(top.in[1]);
      auto is_sin = [&](int x) { const TfOp* q = M.prod(M.skip(x)); return q && q->code == OP_SIN; };
      const TfOp* pa = M.prod(a);
      const bool a_is_enc = pa && pa->code == OP_MUL && (is_sin(pa->in[0]) || is_sin(pa->in[1]));
      enc_t = a_is_enc ? a : b; mask_t = a_is_enc ? b : a;
    }
    const TfOp& ampmul = M.need(enc_t, OP_MUL, "encoding amplitude");
    const int sin_t = M.split(ampmul, &c, "encoding amplitude");
    if (c.size() != 1) M.fail("amplitude must be a scalar");
    const float amp = c[0];
    const TfOp& sn = M.need(sin_t, OP_SIN, "sin");
    const TfOp& ph = M.need(sn.in[0], OP_ADD, "phase add");
    std::vector<float> phase;
    const int arg_t = M.split(ph, &phase, "phase add");
    if ((int)phase.size() != kEnc) M.fail("phase vector length");
    const TfOp& wm = M.need(arg_t, OP_MUL, "frequency scale");
    const int u_t = M.skip(M.split(wm, &c, "frequency scale"));
    if (c.size() != 1) M.fail("frequency scale must be a scalar");
    const float wmul = c[0];
    const TfOp& m1 = M.need(u_t, OP_MUL, "normalisation scale");
    const int s_t = M.split(m1, &c, "normalisation scale");
    if (c.size() != 1) M.fail("normalisation scale must be a scalar");
    const float mul1 = c[0];
    const TfOp* sh = M.prod(s_t);
    if (!sh || (sh->code != OP_ADD && sh->code != OP_SUB)) M.fail("normalisation shift");
    const int x_t = M.split(*sh, &c, "normalisation shift");
    if (c.size() != 1) M.fail("normalisation shift must be a scalar");
    if (sh->code == OP_SUB && M.cval(sh->in[1]).empty()) M.fail("constant - x is not supported");
    const float add = sh->code == OP_ADD ? c[0] : -c[0];
    const int idx = M.slice_index(x_t, "input slice");
    if (idx < 0 || idx > 2 || seen[idx] || idx != k) M.fail("feature order");
    seen[idx] = true;
    P.add[idx] = add; P.mul1[idx] = mul1;
    if (!have_common) { P.wmul = wmul; P.amp = amp; memcpy(P.phase, phase.data(), sizeof(P.phase)); have_common = true; }
    else if (P.wmul != wmul || P.amp != amp || memcmp(P.phase, phase.data(), sizeof(P.phase)) != 0) M.fail("the three encodings do not share their constants");
    if (mask_t >= 0) {
      if (idx != 2) M.fail("only the week feature may be masked");
      const TfOp& mm = M.need(mask_t, OP_MUL, "week mask");
      for (int j = 0; j < 2; ++j) {
        const TfOp& sel = M.need(mm.in[j], OP_SELECT_V2, "week mask select");
        const TfOp* cmp = M.prod(sel.in[0]);
        if (!cmp || (cmp->code != OP_GREATER && cmp->code != OP_LESS) || M.slice_index(cmp->in[0], "week mask input") != 2) M.fail("week mask compare");
        const float thr = M.cscalar(cmp->in[1], "week mask threshold"), on = M.cscalar(sel.in[1], "mask on"), off = M.cscalar(sel.in[2], "mask off");
        if (cmp->code == OP_GREATER) { P.gt_thr = thr; P.gt_on = on; P.gt_off = off; } else { P.lt_thr = thr; P.lt_on = on; P.lt_off = off; }
      }
    } else if (idx == 2) { P.gt_thr = -INFINITY; P.lt_thr = INFINITY; P.gt_on = P.lt_on = P.gt_off = P.lt_off = 1.f; }
  }
  return P;
}

// ---- kernels ---------------------------------------------------------------------------------------------------------
struct RfEnc { float add[3], mul1[3], wmul, amp, gt_thr, lt_thr, gt_on, gt_off, lt_on, lt_off; };

__global__ void rf_encode_kernel(const float* __restrict__ in, const float* __restrict__ phase, RfEnc e, float* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 3 * kEnc) return;
  const int b = i / (3 * kEnc), r = i - b * 3 * kEnc, f = r / kEnc, j = r - f * kEnc;
  const float x = in[b * 3 + f];
  float v = e.amp * sinf(((x + e.add[f]) * e.mul1[f]) * e.wmul + phase[j]);
  if (f == 2) v *= (x > e.gt_thr ? e.gt_on : e.gt_off) * (x < e.lt_thr ? e.lt_on : e.lt_off);
  out[i] = v;
}

// y[b][n] = act(sum_k x[b][k] * w[n][k] + bias[n]); one warp per output neuron and 8 batch rows
constexpr int kRfRows = 8;
__global__ void __launch_bounds__(256)
rf_dense_kernel(const float* __restrict__ x, const __half* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                int B, int N, int K, int act) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = warp, b0 = blockIdx.y * kRfRows;
  if (n >= N) return;
  float acc[kRfRows];
#pragma unroll
  for (int r = 0; r < kRfRows; ++r) acc[r] = 0.f;
  const __half* wr = w + (size_t)n * K;
  for (int k = 2 * lane; k < K; k += 64) {
    const float2 wv = __half22float2(*reinterpret_cast<const __half2*>(wr + k));
#pragma unroll
    for (int r = 0; r < kRfRows; ++r) {
      if (b0 + r < B) {
        const float2 xv = __ldg(reinterpret_cast<const float2*>(x + (size_t)(b0 + r) * K + k));
        acc[r] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, acc[r]));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kRfRows; ++r) {
    const float v = warp_sum(acc[r]) + bias[n];
    if (lane == 0 && b0 + r < B) y[(size_t)(b0 + r) * N + n] = act == 1 ? fmaxf(v, 0.f) : (act == 2 ? 1.0f / (1.0f + expf(-v)) : v);
  }
}

}  // namespace
}  // namespace bnb

using namespace bnb;

struct bnb_range_filter {
  int device = 0, n_species = 0, cap = 0;
  RfEnc enc{};
  float* d_phase = nullptr;
  __half* d_w[4] = {nullptr, nullptr, nullptr, nullptr};
  float* d_b[4] = {nullptr, nullptr, nullptr, nullptr};
  int n_out[4] = {0, 0, 0, 0}, n_in[4] = {0, 0, 0, 0};
  float *d_in = nullptr, *d_a = nullptr, *d_b2 = nullptr, *d_out = nullptr;
  cudaStream_t stream = nullptr;
  ~bnb_range_filter() {
    cudaSetDevice(device);
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    cudaFree(d_phase); cudaFree(d_in); cudaFree(d_a); cudaFree(d_b2); cudaFree(d_out);
    for (int l = 0; l < 4; ++l) { cudaFree(d_w[l]); cudaFree(d_b[l]); }
  }
  void reserve(int B) {
    if (B <= cap) return;
    cudaFree(d_in); cudaFree(d_a); cudaFree(d_b2); cudaFree(d_out);
    d_in = d_a = d_b2 = d_out = nullptr; cap = 0;
    int widest = 3 * kEnc;
    for (int l = 0; l < 3; ++l) widest = std::max(widest, n_out[l]);
    BNB_CUDA(cudaMalloc(&d_in, (size_t)B * 3 * sizeof(float)));
    BNB_CUDA(cudaMalloc(&d_a, (size_t)B * widest * sizeof(float)));
    BNB_CUDA(cudaMalloc(&d_b2, (size_t)B * widest * sizeof(float)));
    BNB_CUDA(cudaMalloc(&d_out, (size_t)B * n_species * sizeof(float)));
    cap = B;
  }
  void run(const float* inputs, int B, float* scores) {
    BNB_CUDA(cudaSetDevice(device));
    reserve(B);
    BNB_CUDA(cudaMemcpyAsync(d_in, inputs, (size_t)B * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
    rf_encode_kernel<<<ceil_div(B * 3 * kEnc, 256), 256, 0, stream>>>(d_in, d_phase, enc, d_a, B);
    float* cur = d_a;
    for (int l = 0; l < 4; ++l) {
      float* dst = l == 3 ? d_out : (cur == d_a ? d_b2 : d_a);
      dim3 grid(ceil_div(n_out[l] * 32, 256), ceil_div(B, kRfRows));
      rf_dense_kernel<<<grid, 256, 0, stream>>>(cur, d_w[l], d_b[l], dst, B, n_out[l], n_in[l], l == 3 ? 2 : 1);
      cur = dst;
    }
    BNB_CUDA(cudaGetLastError());
    BNB_CUDA(cudaMemcpyAsync(scores, d_out, (size_t)B * n_species * sizeof(float), cudaMemcpyDeviceToHost, stream));
    BNB_CUDA(cudaStreamSynchronize(stream));
  }
};

namespace {
std::mutex g_rf_mu;
template <class F>
int rf_guarded(F&& f) {
  try { f(); return BNB_OK; }
  catch (const unsupported_model& e) { return capi_fail(BNB_ERR_UNSUPPORTED_MODEL, e.what()); }
  catch (const cuda_error& e) { cudaGetLastError(); return capi_fail(e.code == cudaErrorMemoryAllocation ? BNB_ERR_OUT_OF_MEMORY : BNB_ERR_CUDA, e.what()); }
  catch (const std::bad_alloc&) { return capi_fail(BNB_ERR_OUT_OF_MEMORY, "host allocation failed"); }
  catch (const std::exception& e) { return capi_fail(BNB_ERR_UNSUPPORTED_MODEL, e.what()); }
}
struct DevRestore { int prev = -1; DevRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } } ~DevRestore() { if (prev >= 0) cudaSetDevice(prev); } };
}  // namespace

extern "C" {

int bnb_range_filter_create(const void* tflite, size_t len, int device, bnb_range_filter** out) {
  if (!out) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "out pointer is NULL");
  *out = nullptr;
  if (!tflite || len == 0) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "cannot create TFLite range filter model from data (0 bytes)");
  RfPlan P;
  int rc = rf_guarded([&] { TfModel m = parse_tflite(tflite, len); P = build_range_plan(m); });
  if (rc != BNB_OK) return rc;
  if ((rc = bnb_init()) != BNB_OK) return rc;
  DevRestore dr;
  bnb_range_filter* h = new (std::nothrow) bnb_range_filter();
  if (!h) return capi_fail(BNB_ERR_OUT_OF_MEMORY, "host allocation failed");
  rc = rf_guarded([&] {
    int dev = device;
    if (dev < 0) BNB_CUDA(cudaGetDevice(&dev));
    BNB_CUDA(cudaSetDevice(dev));
    h->device = dev; h->n_species = P.n_species;
    for (int f = 0; f < 3; ++f) { h->enc.add[f] = P.add[f]; h->enc.mul1[f] = P.mul1[f]; }
    h->enc.wmul = P.wmul; h->enc.amp = P.amp; h->enc.gt_thr = P.gt_thr; h->enc.lt_thr = P.lt_thr;
    h->enc.gt_on = P.gt_on; h->enc.gt_off = P.gt_off; h->enc.lt_on = P.lt_on; h->enc.lt_off = P.lt_off;
    BNB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    BNB_CUDA(cudaMalloc(&h->d_phase, sizeof(P.phase)));
    BNB_CUDA(cudaMemcpy(h->d_phase, P.phase, sizeof(P.phase), cudaMemcpyHostToDevice));
    for (int l = 0; l < 4; ++l) {
      const RfPlan::Fc& F = P.fc[l];
      if (F.n_in % 2) throw unsupported_model("range filter: odd layer width");
      h->n_out[l] = F.n_out; h->n_in[l] = F.n_in;
      std::vector<__half> wh(F.w.size());
      for (size_t i = 0; i < wh.size(); ++i) wh[i] = __float2half_rn(F.w[i]);          // exact: the file stores fp16
      BNB_CUDA(cudaMalloc(&h->d_w[l], wh.size() * sizeof(__half)));
      BNB_CUDA(cudaMemcpy(h->d_w[l], wh.data(), wh.size() * sizeof(__half), cudaMemcpyHostToDevice));
      BNB_CUDA(cudaMalloc(&h->d_b[l], F.b.size() * sizeof(float)));
      BNB_CUDA(cudaMemcpy(h->d_b[l], F.b.data(), F.b.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    h->reserve(64);
  });
  if (rc != BNB_OK) { delete h; return rc; }
  *out = h;
  return BNB_OK;
}

void bnb_range_filter_destroy(bnb_range_filter* h) { if (h) { DevRestore dr; delete h; } }

int bnb_range_filter_num_species(const bnb_range_filter* h) { return h ? h->n_species : capi_fail(BNB_ERR_INVALID_ARGUMENT, "range filter handle is NULL"); }

int bnb_range_filter_predict_batch(bnb_range_filter* h, const float* inputs, int batch_size, float* scores) {
  if (!h) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "range filter handle is NULL");
  if (!inputs || !scores || batch_size < 0) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "NULL inputs/scores or negative batch");
  if (batch_size == 0) return BNB_OK;
  DevRestore dr;
  return rf_guarded([&] { h->run(inputs, batch_size, scores); });
}

int bnb_range_filter_predict(bnb_range_filter* h, float latitude, float longitude, float week, float* scores) {
  const float in[3] = {latitude, longitude, week};
  return bnb_range_filter_predict_batch(h, in, 1, scores);
}

}  // extern "C"
