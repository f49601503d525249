"""ORACLE (test infrastructure, never shipped): minimal TFLite flatbuffer reader.

The reference runs the hot path inside a third-party dependency that is not under
/root/reference: github.com/tphakala/go-tflite v0.2.2-0.20260514101223-29408e53fff7
(go.mod:39) -> libtensorflowlite_c 2.17.1 + XNNPACK (Taskfile.yml:6).  What *is* under
/root/reference is the model file itself
(internal/classifier/data/BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite, embedded by
internal/classifier/models_embedded.go:14-15), so the `.tflite` graph is the spec the
oracle restates.  This module decodes that file with nothing but `struct`/numpy (the
`flatbuffers` Python module is not installed); field slots follow the public TFLite
schema v3 (schema.fbs), see SURVEY.md Appendix B.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may import this.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# builtin operator codes that occur in the BirdNET v2.4 graphs (schema.fbs BuiltinOperator)
BUILTIN = {
    0: "ADD", 1: "AVERAGE_POOL_2D", 2: "CONCATENATION", 3: "CONV_2D", 4: "DEPTHWISE_CONV_2D",
    6: "DEQUANTIZE", 9: "FULLY_CONNECTED", 14: "LOGISTIC", 17: "MAX_POOL_2D", 18: "MUL",
    19: "RELU", 22: "RESHAPE", 34: "PAD", 36: "GATHER", 39: "TRANSPOSE", 40: "MEAN", 41: "SUB",
    42: "DIV", 43: "SQUEEZE", 45: "STRIDED_SLICE", 53: "CAST", 55: "MAXIMUM", 58: "LESS",
    61: "GREATER", 66: "SIN", 70: "EXPAND_DIMS", 77: "SHAPE", 78: "POW", 81: "REDUCE_PROD",
    82: "REDUCE_MAX", 83: "PACK", 89: "REDUCE_MIN", 90: "FLOOR_DIV", 96: "RANGE",
    102: "SPLIT_V", 105: "REVERSE_V2", 123: "SELECT_V2", 131: "RFFT2D",
}

TENSOR_DTYPE = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8, 4: np.int64,
                6: np.bool_, 7: np.int16, 8: np.complex64, 9: np.int8}


class _FB:
    """Just enough flatbuffer decoding: tables, vectors, strings, scalars."""

    def __init__(self, buf: bytes):
        self.b = buf

    def u8(self, p):  return self.b[p]
    def i8(self, p):  return struct.unpack_from("<b", self.b, p)[0]
    def u16(self, p): return struct.unpack_from("<H", self.b, p)[0]
    def i32(self, p): return struct.unpack_from("<i", self.b, p)[0]
    def u32(self, p): return struct.unpack_from("<I", self.b, p)[0]

    def root(self):
        return self.u32(0)

    def field(self, table, slot):
        """Absolute position of the field, or 0 if absent (default value)."""
        vt = table - self.i32(table)
        vt_len = self.u16(vt)
        off_pos = 4 + 2 * slot
        if off_pos >= vt_len:
            return 0
        off = self.u16(vt + off_pos)
        return table + off if off else 0

    def indirect(self, p):
        return p + self.u32(p)

    def table_field(self, table, slot):
        p = self.field(table, slot)
        return self.indirect(p) if p else 0

    def vector(self, table, slot):
        """(start_of_elements, length) or (0, 0)."""
        p = self.field(table, slot)
        if not p:
            return 0, 0
        v = self.indirect(p)
        return v + 4, self.u32(v)

    def string(self, table, slot):
        s, n = self.vector(table, slot)
        return self.b[s:s + n].decode("utf-8", "replace") if s else ""

    def table_vector(self, table, slot):
        s, n = self.vector(table, slot)
        return [self.indirect(s + 4 * i) for i in range(n)]

    def i32_vector(self, table, slot):
        s, n = self.vector(table, slot)
        return list(struct.unpack_from("<%di" % n, self.b, s)) if n else []

    def scalar(self, table, slot, fmt, default=0):
        p = self.field(table, slot)
        return struct.unpack_from(fmt, self.b, p)[0] if p else default


@dataclass
class Tensor:
    index: int
    name: str
    shape: list
    dtype: type
    buffer: int
    shape_signature: list
    data: np.ndarray | None = None   # constant payload, None for activations


@dataclass
class Op:
    index: int
    kind: str
    inputs: list
    outputs: list
    opts: dict = field(default_factory=dict)


@dataclass
class Graph:
    tensors: list
    ops: list
    inputs: list
    outputs: list
    description: str = ""


def _decode_options(fb: _FB, kind: str, t: int) -> dict:
    """Builtin option tables (schema.fbs); `t` = table position or 0 (all defaults)."""
    if not t:
        t_field = lambda slot, fmt, d=0: d       # noqa: E731
    else:
        t_field = lambda slot, fmt, d=0: fb.scalar(t, slot, fmt, d)   # noqa: E731
    if kind == "CONV_2D":
        return dict(padding=t_field(0, "<b"), stride_w=t_field(1, "<i"), stride_h=t_field(2, "<i"),
                    act=t_field(3, "<b"), dil_w=t_field(4, "<i", 1), dil_h=t_field(5, "<i", 1))
    if kind == "DEPTHWISE_CONV_2D":
        return dict(padding=t_field(0, "<b"), stride_w=t_field(1, "<i"), stride_h=t_field(2, "<i"),
                    depth_multiplier=t_field(3, "<i"), act=t_field(4, "<b"),
                    dil_w=t_field(5, "<i", 1), dil_h=t_field(6, "<i", 1))
    if kind in ("AVERAGE_POOL_2D", "MAX_POOL_2D"):
        return dict(padding=t_field(0, "<b"), stride_w=t_field(1, "<i"), stride_h=t_field(2, "<i"),
                    filter_w=t_field(3, "<i"), filter_h=t_field(4, "<i"), act=t_field(5, "<b"))
    if kind == "FULLY_CONNECTED":
        return dict(act=t_field(0, "<b"), weights_format=t_field(1, "<b"),
                    keep_num_dims=bool(t_field(2, "<b")))
    if kind in ("ADD", "MUL", "SUB", "DIV"):
        return dict(act=t_field(0, "<b"))
    if kind == "CONCATENATION":
        return dict(axis=t_field(0, "<i"), act=t_field(1, "<b"))
    if kind in ("MEAN", "REDUCE_MAX", "REDUCE_MIN", "REDUCE_PROD"):
        return dict(keep_dims=bool(t_field(0, "<b")))
    if kind == "GATHER":
        return dict(axis=t_field(0, "<i"), batch_dims=t_field(1, "<i"))
    if kind == "STRIDED_SLICE":
        return dict(begin_mask=t_field(0, "<i"), end_mask=t_field(1, "<i"),
                    ellipsis_mask=t_field(2, "<i"), new_axis_mask=t_field(3, "<i"),
                    shrink_axis_mask=t_field(4, "<i"))
    if kind == "PACK":
        return dict(values_count=t_field(0, "<i"), axis=t_field(1, "<i"))
    if kind == "SQUEEZE":
        return dict(squeeze_dims=fb.i32_vector(t, 0) if t else [])
    if kind == "CAST":
        return dict(in_type=t_field(0, "<b"), out_type=t_field(1, "<b"))
    if kind == "SHAPE":
        return dict(out_type=t_field(0, "<b", 2))
    if kind == "RESHAPE":
        return dict(new_shape=fb.i32_vector(t, 0) if t else [])
    return {}


def read_tflite(data: bytes) -> Graph:
    """Decode subgraph 0 of a TFLite model into plain Python/numpy objects."""
    if data[4:8] != b"TFL3":
        raise ValueError("not a TFL3 flatbuffer")
    fb = _FB(data)
    model = fb.root()
    # operator codes: builtin code = max(deprecated i8 slot 0, i32 slot 3)
    opcodes = []
    for oc in fb.table_vector(model, 1):
        dep = fb.scalar(oc, 0, "<b")
        new = fb.scalar(oc, 3, "<i")
        opcodes.append(max(dep, new))
    buffers = []
    for bt in fb.table_vector(model, 4):
        s, n = fb.vector(bt, 0)
        buffers.append((s, n))
    sg = fb.table_vector(model, 2)[0]
    tensors = []
    for i, tt in enumerate(fb.table_vector(sg, 0)):
        shape = fb.i32_vector(tt, 0)
        ttype = fb.scalar(tt, 1, "<b")
        buf = fb.scalar(tt, 2, "<I")
        name = fb.string(tt, 3)
        sig = fb.i32_vector(tt, 7)
        dtype = TENSOR_DTYPE[ttype]
        t = Tensor(i, name, shape, dtype, buf, sig)
        s, n = buffers[buf] if buf < len(buffers) else (0, 0)
        if n:
            arr = np.frombuffer(data, dtype=dtype, count=n // np.dtype(dtype).itemsize, offset=s)
            t.data = arr.reshape(shape) if shape else arr.reshape(())
        elif buf != 0 and shape and 0 in shape:
            t.data = np.zeros(shape, dtype)      # zero-length constant (e.g. the scalar RESHAPE target `[]`)
        tensors.append(t)
    ops = []
    for i, ot in enumerate(fb.table_vector(sg, 3)):
        code = opcodes[fb.scalar(ot, 0, "<I")]
        kind = BUILTIN.get(code, "OP_%d" % code)
        opt_t = fb.table_field(ot, 4)
        ops.append(Op(i, kind, fb.i32_vector(ot, 1), fb.i32_vector(ot, 2), _decode_options(fb, kind, opt_t)))
    return Graph(tensors, ops, fb.i32_vector(sg, 1), fb.i32_vector(sg, 2), fb.string(model, 3))


def load(path: str) -> Graph:
    with open(path, "rb") as f:
        return read_tflite(f.read())


if __name__ == "__main__":
    import collections
    import sys
    g = load(sys.argv[1])
    print("tensors", len(g.tensors), "ops", len(g.ops), "in", g.inputs, "out", g.outputs, g.description)
    print(collections.Counter(o.kind for o in g.ops).most_common())
    if len(sys.argv) > 2:
        for o in g.ops:
            ins = ["%d%s" % (i, "c" if i >= 0 and g.tensors[i].data is not None else "") for i in o.inputs]
            print(o.index, o.kind, ins, "->", o.outputs, [g.tensors[x].shape for x in o.outputs], o.opts)
