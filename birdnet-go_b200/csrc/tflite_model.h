// tflite_model.h — dependency-free reader for the TFLite flatbuffer the Go side hands us.
//
// birdnet-go embeds BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite and passes its bytes to the backend
// constructor (/root/reference/internal/classifier/birdnet.go:338-351,
// /root/reference/internal/inference/tflite/classifier.go:38-40).  The B200 backend keeps that
// contract: bnb_classifier_create() receives the same bytes, so the library must decode the
// flatbuffer itself.  Field slots follow the public TFLite schema v3 (schema.fbs).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace bnb {

enum TfBuiltin : int {
  OP_ADD = 0, OP_AVERAGE_POOL_2D = 1, OP_CONCATENATION = 2, OP_CONV_2D = 3, OP_DEPTHWISE_CONV_2D = 4,
  OP_FULLY_CONNECTED = 9, OP_LOGISTIC = 14, OP_MAX_POOL_2D = 17, OP_MUL = 18, OP_RESHAPE = 22,
  OP_PAD = 34, OP_GATHER = 36, OP_TRANSPOSE = 39, OP_MEAN = 40, OP_SUB = 41, OP_DIV = 42,
  OP_SQUEEZE = 43, OP_STRIDED_SLICE = 45, OP_CAST = 53, OP_MAXIMUM = 55, OP_EXPAND_DIMS = 70,
  OP_SHAPE = 77, OP_POW = 78, OP_REDUCE_PROD = 81, OP_REDUCE_MAX = 82, OP_PACK = 83,
  OP_REDUCE_MIN = 89, OP_FLOOR_DIV = 90, OP_RANGE = 96, OP_SPLIT_V = 102, OP_REVERSE_V2 = 105,
  OP_RFFT2D = 131
};

enum TfType : int { TT_F32 = 0, TT_F16 = 1, TT_I32 = 2, TT_I64 = 4, TT_C64 = 8 };

struct TfTensor {
  std::vector<int32_t> shape;
  int type = 0;
  const uint8_t* data = nullptr;  // points into the caller's buffer; null for activations
  size_t nbytes = 0;
  std::string name;
  bool is_const() const { return data != nullptr && nbytes > 0; }
  size_t numel() const { size_t n = 1; for (int d : shape) n *= (size_t)d; return n; }
  // constant with at least numel * elem_bytes bytes behind it (negative / dynamic dims make numel huge -> false)
  bool has_bytes(size_t elem_bytes) const { const size_t n = numel(); return is_const() && n <= nbytes / elem_bytes; }
  const float* f32() const { return reinterpret_cast<const float*>(data); }
  const int32_t* i32() const { return reinterpret_cast<const int32_t*>(data); }
};

struct TfOp {
  int code = -1;
  std::vector<int32_t> in, out;
  // decoded builtin options (only the fields this backend needs)
  int padding = 0;  // 0 SAME, 1 VALID
  int stride_w = 1, stride_h = 1, filter_w = 1, filter_h = 1;
  int act = 0;      // 0 none, 1 RELU
  int axis = 0;
  bool keep_dims = false;
};

struct TfModel {
  std::vector<TfTensor> tensors;
  std::vector<TfOp> ops;
  std::vector<int32_t> inputs, outputs;
};

class FlatReader {
 public:
  FlatReader(const uint8_t* b, size_t n) : b_(b), n_(n) {}
  template <class T> T rd(size_t p) const {
    if (p > n_ || sizeof(T) > n_ - p) throw std::runtime_error("tflite: read past end of buffer");   // no wrap-around for huge p
    T v; std::memcpy(&v, b_ + p, sizeof(T)); return v;
  }
  size_t field(size_t table, int slot) const {
    int32_t so = rd<int32_t>(table);
    const int64_t vt64 = (int64_t)table - so;
    if (vt64 < 0 || (uint64_t)vt64 > n_) throw std::runtime_error("tflite: vtable offset out of range");
    size_t vt = (size_t)vt64;
    uint16_t vlen = rd<uint16_t>(vt);
    size_t off_pos = 4 + 2 * (size_t)slot;
    if (off_pos + 2 > vlen) return 0;
    uint16_t off = rd<uint16_t>(vt + off_pos);
    return off ? table + off : 0;
  }
  size_t indirect(size_t p) const { return p + rd<uint32_t>(p); }
  size_t table_field(size_t table, int slot) const { size_t p = field(table, slot); return p ? indirect(p) : 0; }
  // returns element start, sets len
  size_t vec(size_t table, int slot, uint32_t* len) const {
    size_t p = field(table, slot);
    if (!p) { *len = 0; return 0; }
    size_t v = indirect(p);
    *len = rd<uint32_t>(v);
    if ((uint64_t)*len > n_) throw std::runtime_error("tflite: vector length exceeds the buffer");
    return v + 4;
  }
  std::vector<size_t> table_vec(size_t table, int slot) const {
    uint32_t n; size_t s = vec(table, slot, &n);
    std::vector<size_t> r(n);
    for (uint32_t i = 0; i < n; ++i) r[i] = indirect(s + 4 * (size_t)i);
    return r;
  }
  std::vector<int32_t> i32_vec(size_t table, int slot) const {
    uint32_t n; size_t s = vec(table, slot, &n);
    std::vector<int32_t> r(n);
    for (uint32_t i = 0; i < n; ++i) r[i] = rd<int32_t>(s + 4 * (size_t)i);
    return r;
  }
  std::string str(size_t table, int slot) const {
    uint32_t n; size_t s = vec(table, slot, &n);
    if (!s) return std::string();
    if (s + n > n_) throw std::runtime_error("tflite: string past end");
    return std::string(reinterpret_cast<const char*>(b_ + s), n);
  }
  template <class T> T scalar(size_t table, int slot, T dflt) const {
    size_t p = field(table, slot);
    return p ? rd<T>(p) : dflt;
  }
  const uint8_t* base() const { return b_; }
  size_t size() const { return n_; }
 private:
  const uint8_t* b_;
  size_t n_;
};

// Decode subgraph 0.  Throws std::runtime_error on malformed input.
inline TfModel parse_tflite(const void* data, size_t len) {
  const uint8_t* b = static_cast<const uint8_t*>(data);
  if (!b || len < 16 || std::memcmp(b + 4, "TFL3", 4) != 0) throw std::runtime_error("tflite: not a TFL3 flatbuffer");
  FlatReader fb(b, len);
  size_t model = fb.rd<uint32_t>(0);
  std::vector<int> opcodes;
  for (size_t oc : fb.table_vec(model, 1)) {
    int dep = fb.scalar<int8_t>(oc, 0, 0);
    int neu = fb.scalar<int32_t>(oc, 3, 0);
    opcodes.push_back(dep > neu ? dep : neu);
  }
  struct Buf { size_t start; uint32_t len; };
  std::vector<Buf> bufs;
  for (size_t bt : fb.table_vec(model, 4)) {
    uint32_t n; size_t s = fb.vec(bt, 0, &n);
    if (s && s + n > len) throw std::runtime_error("tflite: buffer past end");
    bufs.push_back({s, n});
  }
  auto subgraphs = fb.table_vec(model, 2);
  if (subgraphs.empty()) throw std::runtime_error("tflite: no subgraph");
  size_t sg = subgraphs[0];
  TfModel m;
  for (size_t tt : fb.table_vec(sg, 0)) {
    TfTensor t;
    t.shape = fb.i32_vec(tt, 0);
    t.type = fb.scalar<int8_t>(tt, 1, 0);
    uint32_t bi = fb.scalar<uint32_t>(tt, 2, 0);
    t.name = fb.str(tt, 3);
    if (bi < bufs.size() && bufs[bi].len) { t.data = b + bufs[bi].start; t.nbytes = bufs[bi].len; }
    m.tensors.push_back(std::move(t));
  }
  for (size_t ot : fb.table_vec(sg, 3)) {
    TfOp op;
    uint32_t oi = fb.scalar<uint32_t>(ot, 0, 0);
    if (oi >= opcodes.size()) throw std::runtime_error("tflite: bad opcode index");
    op.code = opcodes[oi];
    op.in = fb.i32_vec(ot, 1);
    op.out = fb.i32_vec(ot, 2);
    size_t o = fb.table_field(ot, 4);
    if (o) {
      switch (op.code) {
        case OP_CONV_2D:
          op.padding = fb.scalar<int8_t>(o, 0, 0); op.stride_w = fb.scalar<int32_t>(o, 1, 0);
          op.stride_h = fb.scalar<int32_t>(o, 2, 0); op.act = fb.scalar<int8_t>(o, 3, 0); break;
        case OP_DEPTHWISE_CONV_2D:
          op.padding = fb.scalar<int8_t>(o, 0, 0); op.stride_w = fb.scalar<int32_t>(o, 1, 0);
          op.stride_h = fb.scalar<int32_t>(o, 2, 0); op.act = fb.scalar<int8_t>(o, 4, 0); break;
        case OP_AVERAGE_POOL_2D: case OP_MAX_POOL_2D:
          op.padding = fb.scalar<int8_t>(o, 0, 0); op.stride_w = fb.scalar<int32_t>(o, 1, 0);
          op.stride_h = fb.scalar<int32_t>(o, 2, 0); op.filter_w = fb.scalar<int32_t>(o, 3, 0);
          op.filter_h = fb.scalar<int32_t>(o, 4, 0); op.act = fb.scalar<int8_t>(o, 5, 0); break;
        case OP_FULLY_CONNECTED: op.act = fb.scalar<int8_t>(o, 0, 0); break;
        case OP_ADD: case OP_MUL: case OP_SUB: case OP_DIV: op.act = fb.scalar<int8_t>(o, 0, 0); break;
        case OP_CONCATENATION: op.axis = fb.scalar<int32_t>(o, 0, 0); op.act = fb.scalar<int8_t>(o, 1, 0); break;
        case OP_MEAN: case OP_REDUCE_MAX: case OP_REDUCE_MIN: op.keep_dims = fb.scalar<int8_t>(o, 0, 0) != 0; break;
        case OP_GATHER: op.axis = fb.scalar<int32_t>(o, 0, 0); break;
        default: break;
      }
    }
    for (int32_t i : op.in) if (i < -1 || i >= (int32_t)m.tensors.size()) throw std::runtime_error("tflite: bad tensor index");   // -1 = optional input absent
    // minimum arity of the ops the plan matcher dereferences (a malformed file must fail here, not read past op.in)
    {
      size_t need = 1;
      switch (op.code) {
        case OP_CONV_2D: case OP_DEPTHWISE_CONV_2D: case OP_FULLY_CONNECTED: case OP_ADD: case OP_MUL: case OP_SUB: case OP_DIV:
        case OP_PAD: case OP_GATHER: case OP_POW: case OP_RFFT2D: case OP_RESHAPE: case OP_TRANSPOSE: case OP_MEAN: case OP_REVERSE_V2:
        case OP_MAXIMUM: case OP_REDUCE_MAX: case OP_REDUCE_MIN: case OP_EXPAND_DIMS: need = 2; break;
        case OP_STRIDED_SLICE: need = 4; break;
        default: break;
      }
      if (op.code == OP_RESHAPE) need = 1;      // the shape operand of RESHAPE is optional in the schema
      if (op.in.size() < need) throw std::runtime_error("tflite: operator has too few inputs");
      for (size_t k = 0; k < need; ++k) if (op.in[k] < 0) throw std::runtime_error("tflite: required operator input is absent");
      if (op.out.empty()) throw std::runtime_error("tflite: operator has no output");
    }
    for (int32_t i : op.out) if (i < 0 || i >= (int32_t)m.tensors.size()) throw std::runtime_error("tflite: bad tensor index");
    m.ops.push_back(std::move(op));
  }
  m.inputs = fb.i32_vec(sg, 1);
  m.outputs = fb.i32_vec(sg, 2);
  for (int32_t i : m.inputs) if (i < 0 || i >= (int32_t)m.tensors.size()) throw std::runtime_error("tflite: bad graph input index");
  for (int32_t i : m.outputs) if (i < 0 || i >= (int32_t)m.tensors.size()) throw std::runtime_error("tflite: bad graph output index");
  return m;
}

}  // namespace bnb
