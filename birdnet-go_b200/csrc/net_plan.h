// net_plan.h — turn the decoded .tflite graph into the fused layer plan the CUDA engine executes.
//
// The reference hands the whole graph to an interpreter
// (/root/reference/internal/inference/tflite/classifier.go:69-107); here the graph is matched
// *structurally*, walking producers backwards from the logits, into:
//
//   frontend (min/max normalise -> 2x {frame, Hann, real-DFT (real part), mel, square, pow,
//   flip} -> concat -> BN affine)  ->  stem conv  ->  avg/max pool + concat + 1x1  ->
//   N x MBConv {1x1 expand+SiLU, [pad] 3x3 depthwise+SiLU, [SE], 1x1 project, [+residual]}  ->
//   affine+ReLU -> KxK VALID conv+ReLU -> global mean (embedding) -> FC (logits).
//
// Nothing is keyed on tensor numbers, so a retrained BirdNET-v2.4-shaped file loads too; any
// other topology is rejected with BNB_ERR_UNSUPPORTED_MODEL at create time so the Go caller can
// fall back to TFLite (/root/reference/internal/classifier/birdnet.go:321-335).
#pragma once
#include <cmath>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "tflite_model.h"

namespace bnb {

struct ConvW {
  const float* w = nullptr;  // OHWI (conv) / 1HWC (depthwise) / [O,I] (fc), as stored in the file
  const float* b = nullptr;  // may be null
  int cout = 0, kh = 1, kw = 1, cin = 0;
  int w_tensor = -1, b_tensor = -1;
};

struct SpecPlan {
  int frame_len = 0, hop = 0, n_frames = 0, n_bins = 0, n_mel = 0;
  const float* window = nullptr;  // [frame_len]
  const float* mel = nullptr;     // [n_mel][n_bins]
  float pow_exp = 0.f;
};

struct FrontendPlan {
  int n_samples = 0;
  float eps = 0.f, center = 0.f, gain = 0.f;  // x = ((x-min)/(max-min+eps) - center) * gain
  SpecPlan spec[2];
  float bn_scale[2] = {0, 0}, bn_shift[2] = {0, 0};
  int out_tensor = -1;  // [B, n_mel, n_frames, 2]
};

struct StemPlan {
  ConvW conv; int stride_h = 1, stride_w = 1, pad_t = 0, pad_l = 0;
  int in_h = 0, in_w = 0, out_h = 0, out_w = 0; int out_tensor = -1;
};

struct PoolMixPlan {
  ConvW conv;            // 1x1 over concat(max, avg)
  bool max_first = true; // channel order of the concat
  int out_h = 0, out_w = 0; int pooled_tensor = -1, out_tensor = -1;
};

struct BlockPlan {
  int in_h = 0, in_w = 0, out_h = 0, out_w = 0;
  int cin = 0, cexp = 0, cse = 0, cout = 0, stride = 1;
  ConvW expand, dw, se1, se2, proj;
  bool has_se = false, residual = false;
  int in_tensor = -1, exp_tensor = -1, dw_tensor = -1, gate_tensor = -1, out_tensor = -1;
};

struct PostPlan {
  const float* mul = nullptr; const float* add = nullptr;  // per-channel affine then ReLU
  ConvW conv; int in_h = 0, in_w = 0, out_h = 0, out_w = 0;
  int affine_tensor = -1, conv_tensor = -1, emb_tensor = -1;
};

struct HeadPlan { ConvW fc; int out_tensor = -1; };

struct NetPlan {
  FrontendPlan fe; StemPlan stem; PoolMixPlan mix; std::vector<BlockPlan> blocks; PostPlan post; HeadPlan head;
  int n_species() const { return head.fc.cout; }
  int emb_dim() const { return head.fc.cin; }
};

class unsupported_model : public std::runtime_error {
 public: explicit unsupported_model(const std::string& s) : std::runtime_error("unsupported model: " + s) {}
};

namespace detail {

struct Matcher {
  const TfModel& m;
  std::vector<int> producer;  // tensor -> op index or -1
  explicit Matcher(const TfModel& mm) : m(mm), producer(mm.tensors.size(), -1) {
    for (size_t i = 0; i < m.ops.size(); ++i) for (int o : m.ops[i].out) producer[o] = (int)i;
  }
  [[noreturn]] void fail(const std::string& s) const { throw unsupported_model(s); }
  const TfOp& prod(int t, int want_code, const char* what) const {
    if (t < 0 || producer[t] < 0) fail(std::string(what) + ": tensor has no producer");
    const TfOp& op = m.ops[producer[t]];
    if (op.code != want_code) { std::ostringstream s; s << what << ": expected op " << want_code << " got " << op.code << " (op #" << producer[t] << ")"; fail(s.str()); }
    return op;
  }
  int prod_code(int t) const { return (t >= 0 && producer[t] >= 0) ? m.ops[producer[t]].code : -1; }
  bool is_const(int t) const { return t >= 0 && m.tensors[t].is_const(); }
  // follow pure re-shapes backwards
  int skip_reshape(int t) const {
    for (;;) {
      int c = prod_code(t);
      if (c == OP_RESHAPE || c == OP_EXPAND_DIMS || c == OP_SQUEEZE) t = m.ops[producer[t]].in[0]; else return t;
    }
  }
  const float* cf32(int t, size_t n, const char* what) const {
    if (!is_const(t) || m.tensors[t].type != TT_F32 || m.tensors[t].numel() != n || m.tensors[t].nbytes < n * 4) {
      std::ostringstream s; s << what << ": expected float32 constant of " << n << " elements (tensor " << t << ")"; fail(s.str());
    }
    return m.tensors[t].f32();
  }
  float cscalar(int t, const char* what) const { return *cf32(t, 1, what); }
  // binary op with exactly one constant operand -> (activation tensor, const tensor)
  void split_const(const TfOp& op, int* act, int* cst, const char* what) const {
    if (op.in.size() != 2) fail(std::string(what) + ": not binary");
    bool c0 = is_const(op.in[0]), c1 = is_const(op.in[1]);
    if (c0 == c1) fail(std::string(what) + ": expected exactly one constant operand");
    *act = c0 ? op.in[1] : op.in[0]; *cst = c0 ? op.in[0] : op.in[1];
  }
  // t = x * sigmoid(x)  -> returns x, else -1
  int match_silu(int t) const {
    if (prod_code(t) != OP_MUL) return -1;
    const TfOp& mul = m.ops[producer[t]];
    for (int k = 0; k < 2; ++k) {
      int x = mul.in[k], s = mul.in[1 - k];
      if (prod_code(s) == OP_LOGISTIC && m.ops[producer[s]].in[0] == x) return x;
    }
    return -1;
  }
  ConvW conv_weights(const TfOp& op, bool depthwise, const char* what) const {
    ConvW c;
    if (op.in.size() < 2 || !is_const(op.in[1])) fail(std::string(what) + ": weights not constant");
    const TfTensor& w = m.tensors[op.in[1]];
    if (w.type != TT_F32 || w.shape.size() != 4 || !w.has_bytes(4)) fail(std::string(what) + ": weights must be a 4-D float32 constant of matching size");
    c.w = w.f32(); c.w_tensor = op.in[1];
    c.kh = w.shape[1]; c.kw = w.shape[2];
    if (depthwise) { if (w.shape[0] != 1) fail(std::string(what) + ": depth multiplier"); c.cin = c.cout = w.shape[3]; }
    else { c.cout = w.shape[0]; c.cin = w.shape[3]; }
    if (op.in.size() > 2 && op.in[2] >= 0) { c.b = cf32(op.in[2], (size_t)c.cout, what); c.b_tensor = op.in[2]; }
    return c;
  }
};

inline void same_pad(int in, int k, int s, int* out, int* before) {
  *out = (in + s - 1) / s;
  int total = (*out - 1) * s + k - in; if (total < 0) total = 0;
  *before = total / 2;
}

}  // namespace detail

inline NetPlan build_plan(const TfModel& m) {
  using namespace detail;
  Matcher M(m);
  NetPlan P;
  if (m.inputs.size() != 1 || m.outputs.size() != 1) M.fail("expected one input and one output");
  const TfTensor& tin = m.tensors[m.inputs[0]];
  if (tin.shape.size() != 2 || tin.type != TT_F32) M.fail("input must be [batch, samples] float32");
  P.fe.n_samples = tin.shape[1];

  // ---- head: logits <- FC(embedding) ----------------------------------------------------------
  {
    const TfOp& fc = M.prod(m.outputs[0], OP_FULLY_CONNECTED, "head");
    if (fc.act != 0) M.fail("head: fused activation");
    const TfTensor& w = m.tensors[fc.in[1]];
    if (fc.in.size() < 2 || !w.is_const() || w.shape.size() != 2 || w.type != TT_F32 || !w.has_bytes(4)) M.fail("head: weights");
    P.head.fc.w = w.f32(); P.head.fc.w_tensor = fc.in[1];
    P.head.fc.cout = w.shape[0]; P.head.fc.cin = w.shape[1];
    if (fc.in.size() > 2 && fc.in[2] >= 0) { P.head.fc.b = M.cf32(fc.in[2], (size_t)w.shape[0], "head bias"); P.head.fc.b_tensor = fc.in[2]; }
    P.head.out_tensor = m.outputs[0];
    P.post.emb_tensor = fc.in[0];
  }
  // ---- post: embedding <- MEAN <- conv KxK VALID ReLU <- ReLU(x*mul+add) ----------------------
  int t;
  {
    const TfOp& mean = M.prod(P.post.emb_tensor, OP_MEAN, "global pool");
    const TfOp& conv = M.prod(mean.in[0], OP_CONV_2D, "post conv");
    if (conv.padding != 1 || conv.act != 1 || conv.stride_h != 1 || conv.stride_w != 1) M.fail("post conv: expected VALID, stride 1, ReLU");
    P.post.conv = M.conv_weights(conv, false, "post conv"); P.post.conv_tensor = mean.in[0];
    const TfOp& add = M.prod(conv.in[0], OP_ADD, "post affine add");
    if (add.act != 1) M.fail("post affine: expected fused ReLU");
    int a, c; M.split_const(add, &a, &c, "post affine add");
    P.post.add = M.cf32(c, (size_t)P.post.conv.cin, "post affine add"); P.post.affine_tensor = conv.in[0];
    const TfOp& mul = M.prod(a, OP_MUL, "post affine mul");
    M.split_const(mul, &a, &c, "post affine mul");
    P.post.mul = M.cf32(c, (size_t)P.post.conv.cin, "post affine mul");
    t = a;
  }
  // ---- MBConv blocks, last to first -------------------------------------------------------------
  std::vector<BlockPlan> rev;
  for (;;) {
    BlockPlan b; b.out_tensor = t;
    int proj_out = t, res_in = -1;
    if (M.prod_code(t) == OP_ADD) {
      const TfOp& add = m.ops[M.producer[t]];
      if (add.act != 0 || M.is_const(add.in[0]) || M.is_const(add.in[1])) M.fail("residual add");
      int k = (M.prod_code(add.in[0]) == OP_CONV_2D) ? 0 : 1;   // the projection output; other side = block input
      // both could be conv outputs (previous block without residual): the projection is the one whose
      // input is NOT the other operand's consumer; disambiguate below by checking the expand input.
      proj_out = add.in[k]; res_in = add.in[1 - k]; b.residual = true;
    }
    if (M.prod_code(proj_out) != OP_CONV_2D) M.fail("block: projection conv not found");
    const TfOp* proj = &m.ops[M.producer[proj_out]];
    // Is this the pool-mix 1x1 (input = CONCATENATION) rather than a block projection?
    if (!b.residual && M.prod_code(proj->in[0]) == OP_CONCATENATION) break;
    auto try_parse = [&](const TfOp* pj, int resid) -> bool {
      b.proj = M.conv_weights(*pj, false, "project");
      if (b.proj.kh != 1 || b.proj.kw != 1 || pj->act != 0) return false;
      int a = pj->in[0];
      // optional SE: a = x * sigmoid(se2(silu(se1(mean(x)))))
      int x = M.match_silu(a);
      if (x < 0) {
        if (M.prod_code(a) != OP_MUL) return false;
        const TfOp& sc = m.ops[M.producer[a]];
        int gate = -1, xx = -1;
        for (int k = 0; k < 2; ++k) if (M.prod_code(sc.in[k]) == OP_LOGISTIC) { gate = sc.in[k]; xx = sc.in[1 - k]; }
        if (gate < 0) return false;
        const TfOp& lg = m.ops[M.producer[gate]];
        const TfOp& c2 = M.prod(lg.in[0], OP_CONV_2D, "SE expand");
        b.se2 = M.conv_weights(c2, false, "SE expand");
        int s1 = M.match_silu(c2.in[0]);
        if (s1 < 0) M.fail("SE: expected SiLU after reduce conv");
        const TfOp& c1 = M.prod(s1, OP_CONV_2D, "SE reduce");
        b.se1 = M.conv_weights(c1, false, "SE reduce");
        int mean_out = M.skip_reshape(c1.in[0]);
        const TfOp& mean = M.prod(mean_out, OP_MEAN, "SE mean");
        if (mean.in[0] != xx) M.fail("SE: mean input differs from scaled tensor");
        if (b.se1.kh != 1 || b.se1.kw != 1 || b.se2.kh != 1 || b.se2.kw != 1 || b.se1.cout != b.se2.cin || b.se2.cout != b.se1.cin) M.fail("SE: shapes");
        b.has_se = true; b.cse = b.se1.cout; b.gate_tensor = gate;
        a = xx; x = M.match_silu(a);
        if (x < 0) M.fail("block: expected SiLU after depthwise conv");
      }
      b.dw_tensor = a;
      const TfOp& dw = M.prod(x, OP_DEPTHWISE_CONV_2D, "depthwise");
      b.dw = M.conv_weights(dw, true, "depthwise");
      if (b.dw.kh != 3 || b.dw.kw != 3 || dw.act != 0 || dw.stride_h != dw.stride_w) M.fail("depthwise: expected 3x3, square stride, no fused act");
      b.stride = dw.stride_h;
      int e = dw.in[0];
      bool explicit_pad = false;
      if (M.prod_code(e) == OP_PAD) {
        const TfOp& pad = m.ops[M.producer[e]];
        const TfTensor& pc = m.tensors[pad.in[1]];
        if (!pc.is_const() || pc.numel() != 8 || pc.type != TT_I32 || !pc.has_bytes(4)) M.fail("pad: constant");
        const int32_t* p = pc.i32();
        if (!(p[0] == 0 && p[1] == 0 && p[2] == 1 && p[3] == 1 && p[4] == 1 && p[5] == 1 && p[6] == 0 && p[7] == 0)) M.fail("pad: expected 1 pixel on H and W");
        explicit_pad = true; e = pad.in[0];
      }
      // supported depthwise geometries: (SAME, stride 1) and (explicit pad 1 + VALID, stride 2)
      if (explicit_pad ? !(dw.padding == 1 && b.stride == 2) : !(dw.padding == 0 && b.stride == 1)) M.fail("depthwise: unsupported padding/stride combination");
      b.exp_tensor = e;
      int c = M.match_silu(e);
      if (c < 0) M.fail("block: expected SiLU after expand conv");
      const TfOp& ex = M.prod(c, OP_CONV_2D, "expand");
      b.expand = M.conv_weights(ex, false, "expand");
      if (b.expand.kh != 1 || b.expand.kw != 1 || ex.act != 0) M.fail("expand: expected 1x1 without fused act");
      b.in_tensor = ex.in[0];
      b.cin = b.expand.cin; b.cexp = b.expand.cout; b.cout = b.proj.cout;
      if (b.dw.cout != b.cexp || b.proj.cin != b.cexp || (b.has_se && b.se1.cin != b.cexp)) M.fail("block: channel mismatch");
      if (resid >= 0 && (resid != b.in_tensor || b.cin != b.cout || b.stride != 1)) return false;
      return true;
    };
    bool ok = try_parse(proj, res_in);
    if (!ok && b.residual) {  // operands the other way round
      std::swap(proj_out, res_in);
      if (M.prod_code(proj_out) == OP_CONV_2D) { proj = &m.ops[M.producer[proj_out]]; b = BlockPlan(); b.out_tensor = t; b.residual = true; ok = try_parse(proj, res_in); }
    }
    if (!ok) M.fail("block: structure not recognised");
    rev.push_back(b);
    t = b.in_tensor;
    if (rev.size() > 64) M.fail("too many blocks");
  }
  if (rev.empty()) M.fail("no MBConv blocks found");
  P.blocks.assign(rev.rbegin(), rev.rend());
  // ---- pool mix: t <- conv1x1(concat(max_pool(s), avg_pool(s))) --------------------------------
  int stem_out;
  {
    const TfOp& conv = M.prod(t, OP_CONV_2D, "pool-mix conv");
    P.mix.conv = M.conv_weights(conv, false, "pool-mix conv");
    if (P.mix.conv.kh != 1 || P.mix.conv.kw != 1 || conv.act != 0) M.fail("pool-mix conv: expected 1x1 without act");
    P.mix.out_tensor = t; P.mix.pooled_tensor = conv.in[0];
    const TfOp& cat = M.prod(conv.in[0], OP_CONCATENATION, "pool concat");
    if (cat.in.size() != 2 || cat.axis != 3) M.fail("pool concat: expected 2 inputs on channel axis");
    int c0 = M.prod_code(cat.in[0]), c1 = M.prod_code(cat.in[1]);
    if (!((c0 == OP_MAX_POOL_2D && c1 == OP_AVERAGE_POOL_2D) || (c0 == OP_AVERAGE_POOL_2D && c1 == OP_MAX_POOL_2D))) M.fail("pool concat: expected max+avg pool");
    P.mix.max_first = (c0 == OP_MAX_POOL_2D);
    const TfOp& p0 = m.ops[M.producer[cat.in[0]]]; const TfOp& p1 = m.ops[M.producer[cat.in[1]]];
    for (const TfOp* p : {&p0, &p1})
      if (p->filter_h != 1 || p->filter_w != 2 || p->stride_h != 1 || p->stride_w != 2 || p->act != 0) M.fail("pool: expected 1x2 stride (1,2)");
    if (p0.in[0] != p1.in[0]) M.fail("pool: inputs differ");
    stem_out = p0.in[0];
  }
  // ---- stem conv ------------------------------------------------------------------------------------
  int fe_out;
  {
    const TfOp& conv = M.prod(stem_out, OP_CONV_2D, "stem conv");
    P.stem.conv = M.conv_weights(conv, false, "stem conv");
    if (conv.padding != 0 || conv.act != 1) M.fail("stem conv: expected SAME + ReLU");
    P.stem.stride_h = conv.stride_h; P.stem.stride_w = conv.stride_w; P.stem.out_tensor = stem_out;
    fe_out = conv.in[0];
  }
  // ---- frontend -------------------------------------------------------------------------------------
  {
    FrontendPlan& F = P.fe; F.out_tensor = fe_out;
    const TfOp& add = M.prod(fe_out, OP_ADD, "frontend BN shift");
    int a, c; M.split_const(add, &a, &c, "frontend BN shift");
    const float* sh = M.cf32(c, 2, "frontend BN shift");
    const TfOp& mul = M.prod(a, OP_MUL, "frontend BN scale");
    M.split_const(mul, &a, &c, "frontend BN scale");
    const float* sc = M.cf32(c, 2, "frontend BN scale");
    if (add.act != 0 || mul.act != 0) M.fail("frontend BN: fused act");
    for (int i = 0; i < 2; ++i) { F.bn_scale[i] = sc[i]; F.bn_shift[i] = sh[i]; }
    const TfOp& cat = M.prod(a, OP_CONCATENATION, "spectrogram concat");
    if (cat.in.size() != 2 || cat.axis != 3) M.fail("spectrogram concat");
    int norm_tensor = -1;
    for (int s = 0; s < 2; ++s) {
      SpecPlan& S = F.spec[s];
      int x = M.skip_reshape(cat.in[s]);
      const TfOp& tr = M.prod(x, OP_TRANSPOSE, "spec transpose");
      const TfOp& rv = M.prod(tr.in[0], OP_REVERSE_V2, "spec reverse");
      const TfOp& pw = M.prod(rv.in[0], OP_POW, "spec pow");
      S.pow_exp = M.cscalar(pw.in[1], "spec pow exponent");
      const TfOp& sq = M.prod(pw.in[0], OP_MUL, "spec square");
      if (sq.in[0] != sq.in[1]) M.fail("spec square: expected x*x");
      int y = M.skip_reshape(sq.in[0]);
      const TfOp& fc = M.prod(y, OP_FULLY_CONNECTED, "mel projection");
      const TfTensor& mw = m.tensors[fc.in[1]];
      if (!mw.is_const() || mw.shape.size() != 2 || mw.type != TT_F32 || !mw.has_bytes(4) || (fc.in.size() > 2 && fc.in[2] >= 0)) M.fail("mel projection: expected bias-free constant matrix");
      S.n_mel = mw.shape[0]; S.n_bins = mw.shape[1]; S.mel = mw.f32();
      int z = M.skip_reshape(fc.in[0]);
      const TfOp& cast = M.prod(z, OP_CAST, "complex->real cast");
      int r = M.skip_reshape(cast.in[0]);
      const TfOp& fft = M.prod(r, OP_RFFT2D, "rfft");
      const TfTensor& fl = m.tensors[fft.in[1]];
      if (!fl.is_const() || fl.numel() != 2 || fl.type != TT_I32 || !fl.has_bytes(4) || fl.i32()[0] != 1) M.fail("rfft: fft_length");
      S.frame_len = fl.i32()[1];
      if (S.n_bins != S.frame_len / 2 + 1) M.fail("rfft: bin count vs mel matrix");
      int w = M.skip_reshape(fft.in[0]);
      const TfOp& win = M.prod(w, OP_MUL, "window");
      int fr, wc; M.split_const(win, &fr, &wc, "window");
      S.window = M.cf32(wc, (size_t)S.frame_len, "window");
      int g = M.skip_reshape(fr);
      const TfOp& gat = M.prod(g, OP_GATHER, "framing gather");
      int src = gat.in[0];
      const TfTensor& srct = m.tensors[src];
      int group = srct.shape.empty() ? 0 : srct.shape.back();
      const TfOp& idx_add = M.prod(gat.in[1], OP_ADD, "frame index");
      int ia, ic; M.split_const(idx_add, &ia, &ic, "frame index");
      const TfTensor& offs = m.tensors[ic];
      if (offs.type != TT_I32 || !offs.has_bytes(4) || group <= 0 || (int)offs.numel() * group != S.frame_len) M.fail("framing: offsets vs frame length");
      for (size_t i = 0; i < offs.numel(); ++i) if (offs.i32()[i] != (int)i) M.fail("framing: offsets not contiguous");
      int im = M.skip_reshape(ia);
      const TfOp& idx_mul = M.prod(im, OP_MUL, "frame stride");
      int ra, rc; M.split_const(idx_mul, &ra, &rc, "frame stride");
      const TfTensor& st = m.tensors[rc];
      if (st.type != TT_I32 || st.numel() != 1 || !st.has_bytes(4)) M.fail("frame stride const");
      S.hop = st.i32()[0] * group;
      if (S.hop <= 0 || F.n_samples < S.frame_len) M.fail("framing geometry");
      S.n_frames = (F.n_samples - S.frame_len) / S.hop + 1;
      // source of the gather: reshape(strided_slice(normalised input))
      int nsrc = M.skip_reshape(src);
      const TfOp& ss = M.prod(nsrc, OP_STRIDED_SLICE, "framing slice");
      if (norm_tensor < 0) norm_tensor = ss.in[0]; else if (norm_tensor != ss.in[0]) M.fail("spectrograms read different inputs");
    }
    if (F.spec[0].n_frames != F.spec[1].n_frames || F.spec[0].n_mel != F.spec[1].n_mel) M.fail("spectrogram shapes differ");
    // normalisation: ((x - min) / (max(x - min) + eps) - center) * gain
    const TfOp& g = M.prod(norm_tensor, OP_MUL, "norm gain");
    int a1, c1; M.split_const(g, &a1, &c1, "norm gain"); F.gain = M.cscalar(c1, "norm gain");
    const TfOp& sb = M.prod(a1, OP_SUB, "norm center");
    if (!M.is_const(sb.in[1])) M.fail("norm center"); F.center = M.cscalar(sb.in[1], "norm center");
    const TfOp& dv = M.prod(sb.in[0], OP_DIV, "norm div");
    const TfOp& den = M.prod(dv.in[1], OP_ADD, "norm eps");
    int a2, c2; M.split_const(den, &a2, &c2, "norm eps"); F.eps = M.cscalar(c2, "norm eps");
    const TfOp& mx = M.prod(a2, OP_REDUCE_MAX, "norm max");
    const TfOp& sub0 = M.prod(dv.in[0], OP_SUB, "norm min-subtract");
    if (mx.in[0] != dv.in[0] || sub0.in[0] != m.inputs[0]) M.fail("norm structure");
    const TfOp& mn = M.prod(sub0.in[1], OP_REDUCE_MIN, "norm min");
    if (mn.in[0] != m.inputs[0]) M.fail("norm min input");
  }
  // ---- geometry -----------------------------------------------------------------------------------
  {
    StemPlan& S = P.stem;
    S.in_h = P.fe.spec[0].n_mel; S.in_w = P.fe.spec[0].n_frames;
    if (S.conv.cin != 2) M.fail("stem: expected 2 input channels");
    detail::same_pad(S.in_h, S.conv.kh, S.stride_h, &S.out_h, &S.pad_t);
    detail::same_pad(S.in_w, S.conv.kw, S.stride_w, &S.out_w, &S.pad_l);
    P.mix.out_h = S.out_h; P.mix.out_w = S.out_w / 2;
    if (S.out_w % 2 || P.mix.conv.cin != 2 * S.conv.cout) M.fail("pool-mix geometry");
    int h = P.mix.out_h, w = P.mix.out_w, c = P.mix.conv.cout;
    for (BlockPlan& b : P.blocks) {
      if (b.cin != c) M.fail("block input channels");
      b.in_h = h; b.in_w = w;
      if (b.stride == 2) { b.out_h = (h + 2 - 3) / 2 + 1; b.out_w = (w + 2 - 3) / 2 + 1; } else { b.out_h = h; b.out_w = w; }
      h = b.out_h; w = b.out_w; c = b.cout;
    }
    PostPlan& Q = P.post;
    if (Q.conv.cin != c) M.fail("post conv channels");
    Q.in_h = h; Q.in_w = w; Q.out_h = h - Q.conv.kh + 1; Q.out_w = w - Q.conv.kw + 1;
    if (Q.out_h < 1 || Q.out_w < 1) M.fail("post conv geometry");
    if (P.head.fc.cin != Q.conv.cout) M.fail("head input dim");
  }
  return P;
}

inline std::string describe_plan(const NetPlan& P) {
  std::ostringstream o;
  o << "{\"n_samples\":" << P.fe.n_samples << ",\"n_species\":" << P.n_species() << ",\"embedding_dim\":" << P.emb_dim()
    << ",\"norm\":{\"eps\":" << P.fe.eps << ",\"center\":" << P.fe.center << ",\"gain\":" << P.fe.gain << "}"
    << ",\"frontend_out_tensor\":" << P.fe.out_tensor << ",\"spec\":[";
  for (int s = 0; s < 2; ++s) {
    const SpecPlan& S = P.fe.spec[s];
    o << (s ? "," : "") << "{\"frame_len\":" << S.frame_len << ",\"hop\":" << S.hop << ",\"n_frames\":" << S.n_frames
      << ",\"n_bins\":" << S.n_bins << ",\"n_mel\":" << S.n_mel << ",\"pow\":" << S.pow_exp
      << ",\"bn_scale\":" << P.fe.bn_scale[s] << ",\"bn_shift\":" << P.fe.bn_shift[s] << "}";
  }
  o << "],\"stem\":{\"k\":[" << P.stem.conv.kh << "," << P.stem.conv.kw << "],\"stride\":[" << P.stem.stride_h << "," << P.stem.stride_w
    << "],\"pad\":[" << P.stem.pad_t << "," << P.stem.pad_l << "],\"cout\":" << P.stem.conv.cout << ",\"out\":[" << P.stem.out_h << "," << P.stem.out_w
    << "],\"out_tensor\":" << P.stem.out_tensor << "}"
    << ",\"mix\":{\"cin\":" << P.mix.conv.cin << ",\"cout\":" << P.mix.conv.cout << ",\"max_first\":" << (P.mix.max_first ? "true" : "false")
    << ",\"out\":[" << P.mix.out_h << "," << P.mix.out_w << "],\"out_tensor\":" << P.mix.out_tensor << "},\"blocks\":[";
  for (size_t i = 0; i < P.blocks.size(); ++i) {
    const BlockPlan& b = P.blocks[i];
    o << (i ? "," : "") << "{\"in\":[" << b.in_h << "," << b.in_w << "," << b.cin << "],\"cexp\":" << b.cexp << ",\"stride\":" << b.stride
      << ",\"se\":" << (b.has_se ? b.cse : 0) << ",\"out\":[" << b.out_h << "," << b.out_w << "," << b.cout << "],\"residual\":" << (b.residual ? "true" : "false")
      << ",\"tensors\":{\"in\":" << b.in_tensor << ",\"exp\":" << b.exp_tensor << ",\"dw\":" << b.dw_tensor << ",\"gate\":" << b.gate_tensor << ",\"out\":" << b.out_tensor << "}"
      << ",\"weights\":{\"expand\":" << b.expand.w_tensor << ",\"dw\":" << b.dw.w_tensor << ",\"proj\":" << b.proj.w_tensor << ",\"proj_bias\":" << b.proj.b_tensor << "}}";
  }
  o << "],\"post\":{\"k\":[" << P.post.conv.kh << "," << P.post.conv.kw << "],\"cin\":" << P.post.conv.cin << ",\"cout\":" << P.post.conv.cout
    << ",\"in\":[" << P.post.in_h << "," << P.post.in_w << "],\"out\":[" << P.post.out_h << "," << P.post.out_w << "],\"emb_tensor\":" << P.post.emb_tensor
    << "},\"head\":{\"w_tensor\":" << P.head.fc.w_tensor << ",\"out_tensor\":" << P.head.out_tensor << "}}";
  return o.str();
}

}  // namespace bnb
