// host_test.cc — exercises the C++ host mirror (birdnet_host.hpp) the way the reference's Go tests exercise the
// corresponding Go code:
//   AnalysisBuffer overlap/content/overwrite    <- internal/audiocore/buffer/analysis_test.go:23-300
//   sigmoid / pair labels / top-k / alias-safety <- internal/classifier/analyze_test.go:14-822
//   16-bit PCM conversion                        <- internal/analysis/process_alloc_test.go:38
// and, with a GPU, the drop-in path end to end:
//   host_test analyze <model.tflite> <labels.txt> <wav> [sensitivity] [threshold]
//     -> one line per 3 s chunk whose top-1 confidence >= threshold (the doc/wiki/file-analysis.md table)
//
//   g++ -std=c++17 -O2 host_test.cc -I../../include -L../lib -lbirdnet_b200 -Wl,-rpath,'$ORIGIN/../lib' -o host_test
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>

#include "birdnet_host.hpp"

using namespace birdnet;

static int g_fail = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } } while (0)

static void test_analysis_buffer() {
  // TestAnalysisBuffer_OverlapRead: second window starts with the tail of the first
  AnalysisBuffer ab(64, 4, 8, "src");
  CHECK(ab.Read().empty());                                 // not enough data yet -> "try again later"
  std::vector<uint8_t> d(16);
  for (int i = 0; i < 16; ++i) d[i] = (uint8_t)(i + 1);
  ab.Write(d.data(), d.size());
  auto w1 = ab.Read();
  CHECK(w1.size() == 12);
  CHECK(w1[0] == 0 && w1[3] == 0 && w1[4] == 1 && w1[11] == 8);   // first window: zero overlap prefix
  auto w2 = ab.Read();
  CHECK(w2.size() == 12 && w2[0] == 5 && w2[3] == 8 && w2[4] == 9 && w2[11] == 16);
  CHECK(ab.Read().empty());
  // BirdNET geometry: 288000-byte windows, 50 % overlap (model.go:41-46): 144000 prefix + 144000 fresh
  AnalysisBuffer big(3 * 288000, 144000, 144000, "mic");
  std::vector<uint8_t> pcm(288000);
  for (size_t i = 0; i < pcm.size(); ++i) pcm[i] = (uint8_t)(i * 7);
  big.Write(pcm.data(), pcm.size());
  auto a = big.Read(), b = big.Read();
  CHECK(a.size() == 288000 && b.size() == 288000);
  CHECK(std::memcmp(b.data(), a.data() + 144000, 144000) == 0);  // content parity across consecutive reads
  // overwrite mode: writing more than the capacity keeps the newest bytes
  AnalysisBuffer small(8, 0, 4, "s");
  std::vector<uint8_t> x(12);
  for (int i = 0; i < 12; ++i) x[i] = (uint8_t)i;
  small.Write(x.data(), 6); small.Write(x.data() + 6, 6);
  CHECK(small.OverwriteCount() >= 1);
  auto r = small.Read();
  CHECK(r.size() == 4 && r[0] == 4 && r[3] == 7);
  // constructor validation (analysis.go:55-111)
  int thrown = 0;
  try { AnalysisBuffer bad(10, 8, 4, "s"); } catch (const std::invalid_argument&) { ++thrown; }
  try { AnalysisBuffer bad(2, 0, 4, "s"); } catch (const std::invalid_argument&) { ++thrown; }
  try { AnalysisBuffer bad(10, 0, 4, ""); } catch (const std::invalid_argument&) { ++thrown; }
  CHECK(thrown == 3);
}

static void test_postprocessing() {
  CHECK(std::fabs(customSigmoid(0.0, 1.0) - 0.5) < 1e-12);
  CHECK(std::fabs(customSigmoid(1.4766, 1.5) - 0.9016) < 1e-4);       // doc table row 1
  auto c = applySigmoidToPredictions({0.f, 10.f, -10.f}, 1.0);
  CHECK(c.size() == 3 && std::fabs(c[0] - 0.5f) < 1e-7 && c[1] > 0.9999f && c[2] < 1e-4f);
  std::vector<std::string> labels = {"a", "b", "c", "d"};
  auto paired = pairLabelsAndConfidence(labels, {0.1f, 0.9f, 0.5f, 0.3f});
  auto top = getTopKResults(paired, 2);
  CHECK(top.size() == 2 && top[0].Species == "b" && top[1].Species == "c");
  CHECK(paired[0].Species == "a");                                     // input not aliased / mutated
  CHECK(getTopKResults(paired, 10).size() == 4 && getTopKResults(paired, 0).empty() && getTopKResults({}, 3).empty());
  bool threw = false;
  try { pairLabelsAndConfidence(labels, {0.1f}); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  const uint8_t raw[] = {0x00, 0x00, 0x01, 0x00, 0xff, 0xff, 0xff, 0x7f, 0x00, 0x80};
  auto f = convert16BitToFloat32(raw, sizeof(raw));
  CHECK(f.size() == 5 && f[0] == 0.f && f[1] == 1.f / 32768.f && f[2] == -1.f / 32768.f && f[3] == 32767.f / 32768.f && f[4] == -1.f);
}

// inferencestats counters + Orchestrator.PredictModel / PredictModelBatch (counters_test.go cases, orchestrator.go:507-572)
struct FakeModel : ModelInstance {
  bool fail = false, closed = false; int calls = 0, batchCalls = 0;
  std::vector<Result> Predict(const std::vector<std::vector<float>>& sample) override {
    ++calls; if (fail) throw std::runtime_error("backend failed");
    Result r; r.Species = "species"; r.Confidence = (float)sample[0].size(); return {r};
  }
  std::vector<std::vector<Result>> PredictBatch(const std::vector<std::vector<float>>& w) override { ++batchCalls; return ModelInstance::PredictBatch(w); }
  void Close() override { closed = true; }
};
static void test_counters_and_orchestrator() {
  Counters c;
  c.RecordInvoke(100); c.RecordInvoke(300); c.RecordInvoke(200); c.RecordError();
  CounterSnapshot s = c.Snapshot();
  CHECK(s.InvokeCount == 3 && s.InvokeTotalUs == 600 && s.InvokeMaxUs == 300 && s.InvokeErrors == 1);
  CHECK(c.Snapshot().InvokeMaxUs == 0 && c.Peek().InvokeMaxUsLifetime == 300);           // interval max reset, lifetime max kept
  struct { std::vector<long long> v; double p; long long want; } cases[] = {{{}, 0.95, 0}, {{42}, 0.95, 42}, {{10, 20, 30, 40, 50, 60, 70}, 0.95, 70},
                                                                            {{1, 2, 3, 4}, 1.0, 4}, {{10, 20, 30, 40, 50}, 0.5, 30}, {{5, 6, 7}, 0.0, 5}};
  for (auto& t : cases) { Counters k; for (long long x : t.v) k.RecordInvoke(x); CHECK(k.RecentPercentileUs(t.p) == t.want); }
  Counters ring;
  for (int i = 0; i < 1024; ++i) ring.RecordInvoke(9000);
  for (int i = 0; i < 1024; ++i) ring.RecordInvoke(10);
  CHECK(ring.RecentPercentileUs(0.95) == 10);                                              // old samples evicted from the 1024 ring
  CHECK(SanitizeModelID("BirdNET_V2.4-fp32") == "BirdNET_V2_4_fp32" && MetricKey("a.b") == "inference.a_b.avg_ms");

  Orchestrator o;
  auto good = std::make_shared<FakeModel>(); auto bad = std::make_shared<FakeModel>(); bad->fail = true;
  o.Register("birdnet", good); o.Register("bad", bad);
  CHECK(o.Predict({std::vector<float>(5, 0.f)})[0].Confidence == 5.f);                     // Predict = PredictModel(primary)
  bool threw = false;
  try { o.PredictModel("nope", {{0.f}}); } catch (const std::invalid_argument& e) { threw = std::string(e.what()) == "unknown model: nope"; }
  CHECK(threw);
  threw = false;
  try { o.PredictModel("bad", {{0.f}}); } catch (const std::runtime_error&) { threw = true; }
  CHECK(threw);
  auto out = o.PredictModelBatch("birdnet", std::vector<std::vector<float>>(7, std::vector<float>(3, 0.f)));
  CHECK(out.size() == 7 && good->batchCalls == 1);
  auto peek = o.counters.PeekAll();
  CHECK(peek["birdnet"].InvokeCount == 2 && peek["birdnet"].BatchWindows == 8);            // one invoke for the whole batch
  CHECK(peek["bad"].InvokeErrors == 1 && peek["bad"].InvokeCount == 0);
  o.CloseModel("birdnet");
  CHECK(good->closed);
  threw = false;
  try { o.PredictModel("birdnet", {{0.f}}); } catch (const std::runtime_error& e) { threw = std::string(e.what()) == "model birdnet has been closed"; }
  CHECK(threw);
  o.DeleteModel("birdnet");
  CHECK(o.counters.PeekAll().count("birdnet") == 0);
}

static void test_overrun_tracker() {
  OverrunTrackers tr;
  OverrunReport rep;
  CHECK(!checkProcessingOverrun(tr, "rtsp_1", "BirdNET_V2.4", 0.4, 1.5, 0.0));          // faster than the buffer interval: nothing recorded
  CHECK(tr.Size() == 0);
  double now = 1.0;
  for (int i = 0; i < 12; ++i, now += 1.5) CHECK(checkProcessingOverrun(tr, "rtsp_1", "BirdNET_V2.4", 2.0 + i, 1.5, now, &rep));
  BufferOverrunTracker& t = tr.Get("rtsp_1", "BirdNET_V2.4");
  CHECK(t.Count() == 12 && t.MaxElapsed() == 13.0 && tr.Size() == 1);
  CHECK(t.Record(3.0, 1.5, 1.0 + bufferOverrunReportCooldown + 1.0, &rep));               // window expired with >= 10 overruns: one report
  CHECK(rep.overrunCount == 12 && rep.maxElapsed == 13.0 && rep.bufferLength == 1.5 && rep.source == "rtsp_1");
  CHECK(t.Count() == 1);
  BufferOverrunTracker few("a", "m");
  for (int i = 0; i < 3; ++i) few.Record(2.0, 1.5, (double)i);
  CHECK(!few.Record(2.0, 1.5, bufferOverrunReportCooldown + 10.0, &rep) && few.Count() == 1);   // below the minimum count: reset silently
}

static std::vector<uint8_t> slurp(const char* p) {
  std::ifstream f(p, std::ios::binary);
  if (!f) throw std::runtime_error(std::string("cannot open ") + p);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), {});
}

static int analyze(int argc, char** argv) {
  const double sensitivity = argc > 5 ? atof(argv[5]) : 1.5;
  const double threshold = argc > 6 ? atof(argv[6]) : 0.1;
  auto model = slurp(argv[2]);
  std::vector<std::string> labels;
  { std::ifstream f(argv[3]); std::string ln; while (std::getline(f, ln)) if (!ln.empty()) labels.push_back(ln); }
  auto wav = slurp(argv[4]);
  // minimal RIFF walk: 16-bit mono PCM
  size_t p = 12, data_off = 0, data_len = 0; int bits = 0;
  while (p + 8 <= wav.size()) {
    uint32_t sz; std::memcpy(&sz, &wav[p + 4], 4);
    if (!std::memcmp(&wav[p], "fmt ", 4)) { uint16_t b; std::memcpy(&b, &wav[p + 8 + 14], 2); bits = b; }
    if (!std::memcmp(&wav[p], "data", 4)) { data_off = p + 8; data_len = sz; }
    p += 8 + sz + (sz & 1);
  }
  if (bits != 16 || !data_off) throw std::runtime_error("expected 16-bit PCM wav");
  std::unique_ptr<inference::Classifier> backend;
  try { backend.reset(new inference::B200Classifier(model)); }
  catch (const ErrB200Unavailable& e) { std::printf("UNAVAILABLE %s\n", e.what()); return 3; }   // caller would fall back to TFLite
  {
    // the batched offline driver (config 2): every window of the file through ONE bnb_analyze_batch(int16) call
    inference::B200Classifier& b200 = static_cast<inference::B200Classifier&>(*backend);
    const auto det = AnalyzeFileBatched(b200, labels, reinterpret_cast<const int16_t*>(&wav[data_off]), data_len / 2, 0.0, sensitivity, threshold, 64);
    double last = -1.0;
    for (const Detection& d : det) if (d.Begin != last) { std::printf("batched\t%.1f\t%s\t%.4f\n", d.Begin, d.Species.c_str(), d.Confidence); last = d.Begin; }
  }
  BirdNET bn(std::move(backend), labels, sensitivity);
  const size_t win = 144000 * 2;
  for (size_t off = 0, i = 0; off + win <= data_len; off += win, ++i) {
    auto chunk = convert16BitToFloat32(&wav[data_off + off], win);
    auto res = bn.Predict({chunk});
    if (!res.empty() && res[0].Confidence >= threshold)
      std::printf("%.1f\t%s\t%.4f\n", 3.0 * (double)i, res[0].Species.c_str(), res[0].Confidence);
  }
  // error behaviour of the drop-in boundary
  bool threw = false;
  try { bn.Predict({std::vector<float>(1000, 0.f)}); } catch (const Error& e) { threw = std::string(e.what()).find("input size mismatch") != std::string::npos; }
  std::printf("size-mismatch-error %s\n", threw ? "ok" : "MISSING");
  bn.Delete(); bn.Delete();
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc > 4 && std::string(argv[1]) == "analyze") return analyze(argc, argv);
    test_analysis_buffer();
    test_postprocessing();
    test_overrun_tracker();
    test_counters_and_orchestrator();
    std::printf(g_fail ? "FAILED %d\n" : "OK\n", g_fail);
    return g_fail ? 1 : 0;
  } catch (const std::exception& e) { std::printf("EXCEPTION %s\n", e.what()); return 2; }
}
