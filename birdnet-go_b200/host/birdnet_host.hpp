// birdnet_host.hpp — C++ mirror of the reference's host-side surface for the hot path, on top of the C ABI.
//
// The reference is Go; there is no Go toolchain in the build image, so the host side above
// include/birdnet_b200.h is mirrored in C++ with the same names, argument meaning and error behaviour:
//
//   inference::Classifier / EmbeddingExtractor   /root/reference/internal/inference/backend.go:8-29
//   B200Classifier (a 4th backend)               shaped like tflite/classifier.go:29-134
//   BirdNET::Predict                             /root/reference/internal/classifier/analyze.go:25-110
//   customSigmoid / getTopKResults               analyze.go:113-115, 197-208, 220-253
//   convert16BitToFloat32                        /root/reference/internal/analysis/process.go:479-497
//   AnalysisBuffer (overwrite ring + overlap)    /root/reference/internal/audiocore/buffer/analysis.go:30-251
//   Results / ResultsQueue message               /root/reference/internal/classifier/queue.go:10-28
//   bufferOverrunTracker / recordBufferOverrun   /root/reference/internal/analysis/process.go:44-215, 351-370
//   AnalyzeFileBatched (config-2 driver)         doc/wiki/file-analysis.md:1-13 semantics through the batched int16 entry point
//
// Header-only; link with -lbirdnet_b200.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/birdnet_b200.h"

namespace birdnet {

struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};
// mirror of ErrOpenVINOUnavailable (openvino/openvino.go:31): the caller falls back to TFLite
struct ErrB200Unavailable : Error { using Error::Error; };

namespace inference {

// backend.go:8-19 — NOT thread-safe; callers synchronize.
class Classifier {
 public:
  virtual ~Classifier() = default;
  virtual std::vector<float> Predict(const std::vector<float>& samples) = 0;  // raw logits, label order
  virtual int NumSpecies() const = 0;
  virtual void Close() = 0;
};
// backend.go:23-29
class EmbeddingExtractor : public Classifier {
 public:
  virtual void PredictWithEmbeddings(const std::vector<float>& samples, std::vector<float>* logits, std::vector<float>* emb) = 0;
};

struct B200Options { int Device = -1; int MaxBatch = 256; int MicroBatch = 0; int Precision = BNB_PRECISION_DEFAULT; };

class B200Classifier final : public EmbeddingExtractor {
 public:
  // NewB200Classifier(modelData, opts): modelData = bytes of the embedded .tflite (tflite/classifier.go:38-40)
  B200Classifier(const std::vector<uint8_t>& modelData, const B200Options& o = B200Options()) {
    if (modelData.empty()) throw Error(BNB_ERR_INVALID_ARGUMENT, "cannot create model from data (0 bytes)");
    bnb_options opts{};
    opts.struct_size = sizeof(opts); opts.device = o.Device; opts.max_batch = o.MaxBatch; opts.micro_batch = o.MicroBatch; opts.precision = o.Precision;
    int rc = bnb_classifier_create(modelData.data(), modelData.size(), &opts, &h_);
    if (rc == BNB_ERR_NO_DEVICE) throw ErrB200Unavailable(rc, std::string("b200: unavailable: ") + bnb_last_error());
    if (rc != BNB_OK) throw Error(rc, std::string("b200: classifier_create failed: status=") + std::to_string(rc) + ": " + bnb_last_error());
    numSpecies_ = bnb_num_species(h_); numSamples_ = bnb_num_samples(h_); embDim_ = bnb_embedding_dim(h_);
  }
  ~B200Classifier() override { Close(); }

  std::vector<float> Predict(const std::vector<float>& samples) override {
    check_open();
    if ((int)samples.size() != numSamples_)   // same text as tflite/classifier.go:102-104
      throw Error(BNB_ERR_INVALID_ARGUMENT, "input size mismatch: expected " + std::to_string(numSamples_) + " samples, got " + std::to_string(samples.size()));
    std::vector<float> out((size_t)numSpecies_);   // freshly allocated, tflite/classifier.go:115-118
    check(bnb_predict(h_, samples.data(), samples.size(), out.data()), "predict");
    return out;
  }
  void PredictWithEmbeddings(const std::vector<float>& samples, std::vector<float>* logits, std::vector<float>* emb) override {
    check_open();
    if ((int)samples.size() != numSamples_)
      throw Error(BNB_ERR_INVALID_ARGUMENT, "input size mismatch: expected " + std::to_string(numSamples_) + " samples, got " + std::to_string(samples.size()));
    logits->assign((size_t)numSpecies_, 0.f); emb->assign((size_t)embDim_, 0.f);
    check(bnb_predict_with_embeddings(h_, samples.data(), samples.size(), logits->data(), emb->data()), "predict_with_embeddings");
  }
  // batched surface (the only precedent is onnx PredictBatch, onnx/classifier.go:372-430)
  void PredictBatch(const void* pcm, int format, int B, std::vector<float>* logits) {
    check_open();
    logits->assign((size_t)B * numSpecies_, 0.f);
    check(bnb_predict_batch(h_, pcm, format, B, logits->data(), nullptr), "predict_batch");
  }
  void AnalyzeBatch(const void* pcm, int format, int B, float sensitivity, int k, std::vector<int32_t>* idx, std::vector<float>* conf) {
    check_open();
    idx->assign((size_t)B * k, 0); conf->assign((size_t)B * k, 0.f);
    check(bnb_analyze_batch(h_, pcm, format, B, sensitivity, k, idx->data(), conf->data(), nullptr), "analyze_batch");
  }
  // N1: threshold + compaction on the device; chunk / species index / confidence triples in chunk-then-confidence order
  int AnalyzeBatchDetections(const void* pcm, int format, int B, float sensitivity, float threshold, int k, std::vector<int32_t>* chunk,
                             std::vector<int32_t>* idx, std::vector<float>* conf) {
    check_open();
    const size_t cap = (size_t)B * k;
    chunk->assign(cap, 0); idx->assign(cap, 0); conf->assign(cap, 0.f);
    int32_t n = 0;
    check(bnb_analyze_batch_detections(h_, pcm, format, B, sensitivity, threshold, k, (int)cap, chunk->data(), idx->data(), conf->data(), nullptr, &n), "analyze_batch_detections");
    chunk->resize((size_t)n); idx->resize((size_t)n); conf->resize((size_t)n);
    return n;
  }
  int NumSpecies() const override { return numSpecies_; }
  int NumSamples() const { return numSamples_; }
  int EmbeddingDim() const { return embDim_; }
  std::string Device() const { return h_ ? bnb_runtime_device(h_) : ""; }
  std::string Precision() const { return h_ ? bnb_runtime_precision(h_) : ""; }
  float LastInvokeMs() const { return h_ ? bnb_last_device_ms(h_) : -1.f; }
  void Close() override { if (h_) { bnb_classifier_destroy(h_); h_ = nullptr; } }   // idempotent (tflite/classifier.go:129-134)

 private:
  void check_open() const { if (!h_) throw Error(BNB_ERR_CLOSED, "classifier is closed"); }
  void check(int rc, const char* op) const {
    if (rc != BNB_OK) throw Error(rc, std::string("b200: ") + op + " failed: status=" + std::to_string(rc) + ": " + bnb_last_error());
  }
  bnb_classifier* h_ = nullptr;
  int numSpecies_ = 0, numSamples_ = 0, embDim_ = 0;
};

}  // namespace inference

// datastore.Results as used on this path
struct Result { std::string Species; float Confidence = 0.f; };

// analyze.go:113-115
inline double customSigmoid(double x, double sensitivity) { return 1.0 / (1.0 + std::exp(-sensitivity * x)); }
// analyze.go:186-194
inline std::vector<float> applySigmoidToPredictions(const std::vector<float>& predictions, double sensitivity) {
  std::vector<float> c(predictions.size());
  for (size_t i = 0; i < predictions.size(); ++i) c[i] = (float)customSigmoid((double)predictions[i], sensitivity);
  return c;
}
// analyze.go:125-137
inline std::vector<Result> pairLabelsAndConfidence(const std::vector<std::string>& labels, const std::vector<float>& preds) {
  if (labels.size() != preds.size())
    throw std::invalid_argument("mismatched labels and predictions lengths: " + std::to_string(labels.size()) + " vs " + std::to_string(preds.size()));
  std::vector<Result> r(labels.size());
  for (size_t i = 0; i < labels.size(); ++i) { r[i].Species = labels[i]; r[i].Confidence = preds[i]; }
  return r;
}
// analyze.go:220-253: k best by descending confidence, returned as a fresh copy (never aliases the scratch buffer)
inline std::vector<Result> getTopKResults(std::vector<Result> results, int k) {
  if (results.empty() || k <= 0) return {};
  const size_t n = std::min((size_t)k, results.size());
  auto gt = [](const Result& a, const Result& b) { return a.Confidence > b.Confidence; };
  std::partial_sort(results.begin(), results.begin() + n, results.end(), gt);
  return std::vector<Result>(results.begin(), results.begin() + n);
}
// process.go:479-497
inline std::vector<float> convert16BitToFloat32(const uint8_t* sample, size_t nbytes) {
  const size_t length = nbytes / 2;
  std::vector<float> out(length);
  const float divisor = 32768.0f;
  for (size_t i = 0; i < length; ++i) {
    const int16_t s = (int16_t)((uint16_t)sample[i * 2] | ((uint16_t)sample[i * 2 + 1] << 8));
    out[i] = (float)s / divisor;
  }
  return out;
}

constexpr int defaultTopKResults = 10;   // tracing.go:59

// BirdNET model instance: lock, backend call, sensitivity-sigmoid, label pairing, top-10 (analyze.go:25-110)
class BirdNET {
 public:
  BirdNET(std::unique_ptr<inference::Classifier> backend, std::vector<std::string> labels, double sensitivity = 1.0)
      : classifier_(std::move(backend)), labels_(std::move(labels)), sensitivity_(sensitivity) {
    if ((int)labels_.size() != classifier_->NumSpecies())   // validateModelAndLabels, birdnet.go:1248-1282
      throw std::invalid_argument("mismatched labels and predictions lengths: " + std::to_string(labels_.size()) + " vs " + std::to_string(classifier_->NumSpecies()));
  }
  std::vector<Result> Predict(const std::vector<std::vector<float>>& sample) {
    if (sample.empty() || sample[0].empty()) throw std::invalid_argument("empty audio sample");
    std::lock_guard<std::mutex> lk(mu_);                      // bn.mu held for the whole native call (birdnet.go:111-119)
    if (!classifier_) throw std::runtime_error("classifier backend is not initialized");
    const std::vector<float> predictions = classifier_->Predict(sample[0]);   // only sample[0] is used (analyze.go:60)
    const std::vector<float> confidence = applySigmoidToPredictions(predictions, sensitivity_);
    return getTopKResults(pairLabelsAndConfidence(labels_, confidence), defaultTopKResults);
  }
  void Delete() { std::lock_guard<std::mutex> lk(mu_); if (classifier_) { classifier_->Close(); classifier_.reset(); } }
  const std::vector<std::string>& Labels() const { return labels_; }

 private:
  std::unique_ptr<inference::Classifier> classifier_;
  std::vector<std::string> labels_;
  double sensitivity_;
  std::mutex mu_;
};

// ---- inference counters + the orchestrator's predict surface ------------------------------------------------
// inferencestats.Counters / CounterMap (internal/classifier/inferencestats/counters.go:29-251): invoke count, total and max
// microseconds (the interval max is reset by Snapshot, the lifetime max is not), errors, and a ring of the last 1024 durations
// for a nearest-rank percentile (idx = ceil(p n) - 1, clamped).
constexpr int latencyWindowSize = 1024;
constexpr double healthLatencyPercentile = 0.95;
struct CounterSnapshot { long long InvokeCount = 0, InvokeTotalUs = 0, InvokeMaxUs = 0, InvokeErrors = 0; };
struct CounterPeek { long long InvokeCount = 0, InvokeTotalUs = 0, InvokeMaxUsLifetime = 0, RecentP95Us = 0, InvokeErrors = 0, BatchWindows = 0; };
class Counters {
 public:
  void RecordInvoke(long long durationUs, long long windows = 1) {
    std::lock_guard<std::mutex> lk(mu_);
    ++count_; total_ += durationUs; max_ = std::max(max_, durationUs); maxLife_ = std::max(maxLife_, durationUs); windows_ += windows;
    ring_[pos_] = durationUs; pos_ = (pos_ + 1) % latencyWindowSize; if (len_ < latencyWindowSize) ++len_;
  }
  void RecordError() { std::lock_guard<std::mutex> lk(mu_); ++errors_; }
  long long RecentPercentileUs(double p) const {
    std::vector<long long> v;
    { std::lock_guard<std::mutex> lk(mu_); if (len_ == 0) return 0; v.assign(ring_, ring_ + len_); }
    std::sort(v.begin(), v.end());
    const int n = (int)v.size();
    int idx = (int)std::ceil(p * n) - 1;
    idx = std::min(std::max(idx, 0), n - 1);
    return v[(size_t)idx];
  }
  CounterSnapshot Snapshot() { std::lock_guard<std::mutex> lk(mu_); CounterSnapshot s{count_, total_, max_, errors_}; max_ = 0; return s; }
  CounterPeek Peek() const {
    CounterPeek p;
    { std::lock_guard<std::mutex> lk(mu_); p.InvokeCount = count_; p.InvokeTotalUs = total_; p.InvokeMaxUsLifetime = maxLife_; p.InvokeErrors = errors_; p.BatchWindows = windows_; }
    p.RecentP95Us = RecentPercentileUs(healthLatencyPercentile);
    return p;
  }
 private:
  mutable std::mutex mu_;
  long long count_ = 0, total_ = 0, max_ = 0, maxLife_ = 0, errors_ = 0, windows_ = 0;
  long long ring_[latencyWindowSize] = {0};
  int pos_ = 0, len_ = 0;
};
inline std::string SanitizeModelID(const std::string& id) {            // counters.go:123-130
  std::string o = id;
  for (char& c : o) if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_')) c = '_';
  return o;
}
inline std::string MetricKey(const std::string& id) { return "inference." + SanitizeModelID(id) + ".avg_ms"; }
class CounterMap {
 public:
  void RecordInvoke(const std::string& id, long long us, long long windows = 1) { get(id).RecordInvoke(us, windows); }
  void RecordError(const std::string& id) { get(id).RecordError(); }
  std::map<std::string, CounterSnapshot> SnapshotAll() { std::map<std::string, CounterSnapshot> r; std::lock_guard<std::mutex> lk(mu_); for (auto& kv : m_) r[kv.first] = kv.second->Snapshot(); return r; }
  std::map<std::string, CounterPeek> PeekAll() { std::map<std::string, CounterPeek> r; std::lock_guard<std::mutex> lk(mu_); for (auto& kv : m_) r[kv.first] = kv.second->Peek(); return r; }
  void Delete(const std::string& id) { std::lock_guard<std::mutex> lk(mu_); m_.erase(id); }
 private:
  Counters& get(const std::string& id) { std::lock_guard<std::mutex> lk(mu_); auto& p = m_[id]; if (!p) p.reset(new Counters()); return *p; }
  std::mutex mu_;
  std::map<std::string, std::unique_ptr<Counters>> m_;
};

// Orchestrator.PredictModel (internal/classifier/orchestrator.go:507-572): a lock on the models map to fetch the entry, then the
// global inferenceMu (one model runs at a time), then the entry's own mutex (instance lifecycle), counters on the way out.
// PredictModelBatch is the additive batched form: one pass through the same locks for a whole batch of windows.
class ModelInstance {
 public:
  virtual ~ModelInstance() = default;
  virtual std::vector<Result> Predict(const std::vector<std::vector<float>>& sample) = 0;
  virtual std::vector<std::vector<Result>> PredictBatch(const std::vector<std::vector<float>>& windows) {
    std::vector<std::vector<Result>> out;
    for (const auto& w : windows) out.push_back(Predict({w}));
    return out;
  }
  virtual void Close() {}
};
class Orchestrator {
 public:
  void Register(const std::string& id, std::shared_ptr<ModelInstance> inst) {
    std::lock_guard<std::mutex> lk(mu_);
    auto e = std::make_shared<Entry>(); e->instance = std::move(inst); models_[id] = e;
    if (primary_.empty()) primary_ = id;
  }
  void CloseModel(const std::string& id) {
    std::shared_ptr<Entry> e = find(id);
    if (!e) return;
    std::shared_ptr<ModelInstance> inst;
    { std::lock_guard<std::mutex> lk(e->mu); inst.swap(e->instance); }
    if (inst) inst->Close();
  }
  void DeleteModel(const std::string& id) { CloseModel(id); { std::lock_guard<std::mutex> lk(mu_); models_.erase(id); } counters.Delete(id); }
  std::vector<Result> Predict(const std::vector<std::vector<float>>& sample) { std::string id; { std::lock_guard<std::mutex> lk(mu_); id = primary_; } return PredictModel(id, sample); }
  std::vector<Result> PredictModel(const std::string& id, const std::vector<std::vector<float>>& sample) {
    std::vector<Result> out;
    run(id, 1, [&](ModelInstance& m) { out = m.Predict(sample); });
    return out;
  }
  std::vector<std::vector<Result>> PredictModelBatch(const std::string& id, const std::vector<std::vector<float>>& windows) {
    std::vector<std::vector<Result>> out;
    run(id, (long long)windows.size(), [&](ModelInstance& m) { out = m.PredictBatch(windows); });
    return out;
  }
  CounterMap counters;

 private:
  struct Entry { std::mutex mu; std::shared_ptr<ModelInstance> instance; };
  std::shared_ptr<Entry> find(const std::string& id) { std::lock_guard<std::mutex> lk(mu_); auto it = models_.find(id); return it == models_.end() ? nullptr : it->second; }
  template <class F>
  void run(const std::string& id, long long windows, F&& call) {
    std::shared_ptr<Entry> e = find(id);                               // map lock released before the inference / model locks
    if (!e) throw std::invalid_argument("unknown model: " + id);
    std::lock_guard<std::mutex> inf(inferenceMu_);
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->instance) throw std::runtime_error("model " + id + " has been closed");
    const auto t0 = std::chrono::steady_clock::now();
    try { call(*e->instance); }
    catch (...) { counters.RecordError(id); throw; }
    counters.RecordInvoke(id, std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count(), windows);
  }
  std::mutex mu_, inferenceMu_;
  std::map<std::string, std::shared_ptr<Entry>> models_;
  std::string primary_;
};

// Tumbling-window overrun accounting per (source, model): process.go:44-215.  `now` is a monotonic clock in seconds supplied by
// the caller (the reference uses time.Now()); a report is produced when a window of bufferOverrunReportCooldown expires with at
// least bufferOverrunMinCount overruns (the reference sends it to Sentry).
constexpr double bufferOverrunReportCooldown = 3600.0;   // process.go:36
constexpr long long bufferOverrunMinCount = 10;          // process.go:39
struct OverrunReport { std::string source, modelID; long long overrunCount = 0; double maxElapsed = 0, bufferLength = 0, window = 0; };
class BufferOverrunTracker {
 public:
  BufferOverrunTracker(std::string source, std::string modelID) : source_(std::move(source)), model_(std::move(modelID)) {}
  // recordBufferOverrun (process.go:177-215); returns true and fills *rep when the expired window is reported
  bool Record(double elapsed, double bufferLen, double now, OverrunReport* rep = nullptr) {
    std::lock_guard<std::mutex> lk(mu_);
    bool reported = false;
    if (!started_) { windowStart_ = now; started_ = true; }
    if (now - windowStart_ >= bufferOverrunReportCooldown) {
      if (count_ >= bufferOverrunMinCount) {
        if (rep) { rep->source = source_; rep->modelID = model_; rep->overrunCount = count_; rep->maxElapsed = maxElapsed_; rep->bufferLength = bufferLength_; rep->window = now - windowStart_; }
        reported = true;
      }
      count_ = 0; maxElapsed_ = 0; windowStart_ = now;
    }
    ++count_;
    if (elapsed > maxElapsed_) { maxElapsed_ = elapsed; bufferLength_ = bufferLen; }
    return reported;
  }
  long long Count() const { return count_; }
  double MaxElapsed() const { return maxElapsed_; }
 private:
  std::string source_, model_;
  long long count_ = 0; bool started_ = false; double windowStart_ = 0, maxElapsed_ = 0, bufferLength_ = 0;
  std::mutex mu_;
};
// getOverrunTracker (process.go:66-76): one tracker per "source:modelID"
class OverrunTrackers {
 public:
  BufferOverrunTracker& Get(const std::string& source, const std::string& modelID) {
    std::lock_guard<std::mutex> lk(mu_);
    auto& p = map_[source + ":" + modelID];
    if (!p) p.reset(new BufferOverrunTracker(source, modelID));
    return *p;
  }
  size_t Size() const { return map_.size(); }
 private:
  std::map<std::string, std::unique_ptr<BufferOverrunTracker>> map_;
  std::mutex mu_;
};
// the check of ProcessData (process.go:351-370): elapsed > BufferInterval -> record an overrun for that source
inline bool checkProcessingOverrun(OverrunTrackers& trackers, const std::string& source, const std::string& modelID, double elapsed,
                                   double bufferInterval, double now, OverrunReport* rep = nullptr) {
  if (elapsed <= bufferInterval) return false;
  trackers.Get(source, modelID).Record(elapsed, bufferInterval, now, rep);
  return true;
}

// One detection row of the offline file analysis
struct Detection { double Begin = 0, End = 0; std::string Species; float Confidence = 0.f; };

// Batched offline file analysis (BASELINE config 2): slide the 3 s window with `overlapSeconds` of overlap over mono int16 PCM, run
// ALL windows through bnb_analyze_batch(BNB_PCM_S16) in batches of <= maxBatch (device-side /32768, sigmoid(sensitivity * logit),
// top-10) and keep the results >= threshold, window order, descending confidence inside a window.  A trailing partial window of
// >= half the window length is zero-padded.
inline std::vector<Detection> AnalyzeFileBatched(inference::B200Classifier& clf, const std::vector<std::string>& labels, const int16_t* pcm,
                                                 size_t nSamples, double overlapSeconds, double sensitivity, double threshold, int maxBatch = 256,
                                                 int sampleRate = 48000) {
  const size_t n = (size_t)clf.NumSamples();
  const long long step = (long long)n - (long long)std::llround(overlapSeconds * sampleRate);
  if (step <= 0) throw std::invalid_argument("overlap must be shorter than the analysis window");
  std::vector<size_t> starts;
  for (size_t s0 = 0; s0 + n <= nSamples; s0 += (size_t)step) starts.push_back(s0);
  const size_t tail = starts.empty() ? 0 : starts.back() + (size_t)step;
  if (nSamples > tail && nSamples - tail >= n / 2) starts.push_back(tail);
  std::vector<Detection> out;
  std::vector<int16_t> win;
  std::vector<int32_t> chunk, idx; std::vector<float> conf;
  for (size_t i = 0; i < starts.size(); i += (size_t)maxBatch) {
    const int B = (int)std::min((size_t)maxBatch, starts.size() - i);
    win.assign((size_t)B * n, 0);
    for (int j = 0; j < B; ++j) {
      const size_t s0 = starts[i + j], len = std::min(n, nSamples - s0);
      std::memcpy(&win[(size_t)j * n], pcm + s0, len * sizeof(int16_t));
    }
    // sigmoid, top-10, threshold and compaction all on the device: only the detections come back (window, then confidence order)
    const int found = clf.AnalyzeBatchDetections(win.data(), BNB_PCM_S16, B, (float)sensitivity, (float)threshold, defaultTopKResults, &chunk, &idx, &conf);
    for (int r = 0; r < found; ++r) {
      Detection d; d.Begin = (double)starts[i + (size_t)chunk[(size_t)r]] / sampleRate; d.End = d.Begin + (double)n / sampleRate;
      d.Species = labels[(size_t)idx[(size_t)r]]; d.Confidence = conf[(size_t)r];
      out.push_back(d);
    }
  }
  return out;
}

// Overwrite-mode byte ring + overlap prefix: consecutive reads share `overlapSize` bytes (analysis.go:30-251).
class AnalysisBuffer {
 public:
  AnalysisBuffer(int capacity, int overlapSize, int readSize, const std::string& sourceID)
      : ring_((size_t)std::max(capacity, 1)), overlap_(overlapSize), read_(readSize) {
    if (capacity <= 0) throw std::invalid_argument("invalid analysis buffer capacity: " + std::to_string(capacity) + ", must be greater than 0");
    if (overlapSize < 0) throw std::invalid_argument("invalid overlap size: " + std::to_string(overlapSize) + ", must be >= 0");
    if (readSize <= 0) throw std::invalid_argument("invalid read size: " + std::to_string(readSize) + ", must be greater than 0");
    if (readSize < overlapSize) throw std::invalid_argument("read size " + std::to_string(readSize) + " must be >= overlap size " + std::to_string(overlapSize));
    if (capacity < readSize) throw std::invalid_argument("capacity " + std::to_string(capacity) + " must be >= read size " + std::to_string(readSize));
    if (sourceID.empty()) throw std::invalid_argument("source ID must not be empty");
  }
  // Write: when the ring is full the oldest bytes are overwritten (analysis.go:152-174)
  void Write(const uint8_t* data, size_t n) {
    std::lock_guard<std::mutex> lk(mu_);
    if (n > ring_.size() - len_) ++overwrites_;
    if (n >= ring_.size()) { data += n - ring_.size(); n = ring_.size(); head_ = 0; len_ = 0; }
    for (size_t i = 0; i < n; ++i) {
      ring_[(head_ + len_) % ring_.size()] = data[i];
      if (len_ < ring_.size()) ++len_; else head_ = (head_ + 1) % ring_.size();
    }
  }
  // Read: empty vector = "try again later"; else [overlap prefix | readSize fresh bytes] (analysis.go:187-251)
  std::vector<uint8_t> Read() {
    std::lock_guard<std::mutex> lk(mu_);
    if ((int)len_ < read_) return {};
    std::vector<uint8_t> window((size_t)overlap_ + read_, 0);
    if (overlap_ > 0 && (int)prev_.size() == overlap_) std::memcpy(window.data(), prev_.data(), (size_t)overlap_);
    for (int i = 0; i < read_; ++i) { window[(size_t)overlap_ + i] = ring_[head_]; head_ = (head_ + 1) % ring_.size(); }
    len_ -= (size_t)read_;
    if (overlap_ > 0) prev_.assign(window.end() - overlap_, window.end());
    return window;
  }
  int64_t OverwriteCount() const { return overwrites_; }

 private:
  std::vector<uint8_t> ring_;
  size_t head_ = 0, len_ = 0;
  std::vector<uint8_t> prev_;
  int overlap_, read_;
  int64_t overwrites_ = 0;
  std::mutex mu_;
};

}  // namespace birdnet
