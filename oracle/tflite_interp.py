"""ORACLE (test infrastructure, never shipped): op-by-op CPU interpretation of a `.tflite` graph.

This is the CPU restatement of what `tfliteClassifier.Predict`
(/root/reference/internal/inference/tflite/classifier.go:95-119: copy in -> Invoke -> copy out)
computes.  The arithmetic itself lives in libtensorflowlite_c 2.17.1 (+XNNPACK), which is NOT
under /root/reference (go.mod:39, Taskfile.yml:6); the restatement therefore executes the op
list of the reference's own model file
(/root/reference/internal/classifier/data/BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite) one TFLite
builtin at a time, following the published TFLite op semantics (NHWC, OHWI conv weights,
TF SAME padding = extra pad at bottom/right, RFFT2D -> complex, CAST complex->float keeps the
real part).  Nothing about the network structure is assumed: shape-plumbing ops
(SHAPE/PACK/STRIDED_SLICE/RANGE/...) are executed too, so any batch size works.

PINNED against the reference's own published detections (doc/wiki/file-analysis.md:20-44) in
tests/test_oracle_golden.py.

Runs in float64 (truth) or float32 on torch-CPU.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline/reference arm may import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from tflite_reader import Graph, load, read_tflite  # noqa: F401


def _same_pad(i, k, s, d=1):
    ke = (k - 1) * d + 1
    o = -(-i // s)
    total = max((o - 1) * s + ke - i, 0)
    return total // 2, total - total // 2


def _act(x, code):
    if code == 0:
        return x
    if code == 1:
        return torch.relu(x)
    if code == 3:
        return torch.clamp(x, 0, 6)
    raise NotImplementedError("fused activation %d" % code)


class Interpreter:
    """Executes subgraph 0.  `dtype` is the float type every float32 tensor is computed in."""

    def __init__(self, graph: Graph, dtype=torch.float64):
        self.g = graph
        self.dtype = dtype
        self.cdtype = torch.complex128 if dtype == torch.float64 else torch.complex64
        self.const = {}
        for t in graph.tensors:
            if t.data is not None:
                a = torch.from_numpy(np.array(t.data))
                if a.dtype in (torch.float32, torch.float16):
                    a = a.to(dtype)
                self.const[t.index] = a

    # ------------------------------------------------------------------ helpers
    def _ints(self, v):
        return [int(x) for x in v.reshape(-1).tolist()]

    # ------------------------------------------------------------------ run
    @torch.no_grad()
    def run(self, x: torch.Tensor, fetch=()):
        """x: [B, n_samples].  Returns {tensor_index: tensor} for graph outputs + `fetch`."""
        g = self.g
        v = dict(self.const)
        v[g.inputs[0]] = x.to(self.dtype)
        want = set(fetch) | set(g.outputs)
        # last use of each activation so memory can be dropped early
        last = {}
        for op in g.ops:
            for i in op.inputs:
                if i >= 0:
                    last[i] = op.index
        out = {}
        for op in g.ops:
            res = self._exec(op, [v[i] if i >= 0 else None for i in op.inputs])
            if not isinstance(res, (list, tuple)):
                res = [res]
            for o, r in zip(op.outputs, res):
                v[o] = r
                if o in want:
                    out[o] = r
            for i in op.inputs:
                if i >= 0 and i not in self.const and last.get(i) == op.index and i not in want:
                    v.pop(i, None)
        return out

    # ------------------------------------------------------------------ ops
    def _exec(self, op, a):
        k, o = op.kind, op.opts
        if k == "ADD":
            return _act(a[0] + a[1], o["act"])
        if k == "SUB":
            return _act(a[0] - a[1], o["act"])
        if k == "MUL":
            return _act(a[0] * a[1], o["act"])
        if k == "DIV":
            return _act(a[0] / a[1], o["act"])
        if k == "FLOOR_DIV":
            return torch.div(a[0], a[1], rounding_mode="floor")
        if k == "MAXIMUM":
            return torch.maximum(a[0], a[1])
        if k == "POW":
            return torch.pow(a[0], a[1])
        if k == "LOGISTIC":
            return torch.sigmoid(a[0])
        # ---- ops of the range-filter meta-model (BirdNET_GLOBAL_6K_V2.4_MData_Model_V2_FP16.tflite) ----
        if k == "DEQUANTIZE":                                   # fp16 constants -> float (exact)
            return a[0].to(self.dtype)
        if k == "SIN":
            return torch.sin(a[0])
        if k == "GREATER":
            return a[0] > a[1]
        if k == "LESS":
            return a[0] < a[1]
        if k == "SELECT_V2":
            return torch.where(a[0], a[1], a[2])
        if k in ("REDUCE_MIN", "REDUCE_MAX", "MEAN", "REDUCE_PROD"):
            axes = tuple(self._ints(a[1]))
            x = a[0]
            if k == "REDUCE_MIN":
                return torch.amin(x, dim=axes, keepdim=o["keep_dims"])
            if k == "REDUCE_MAX":
                return torch.amax(x, dim=axes, keepdim=o["keep_dims"])
            if k == "MEAN":
                return torch.mean(x, dim=axes, keepdim=o["keep_dims"])
            r = x
            for ax in sorted((d % x.dim() for d in axes), reverse=True):
                r = torch.prod(r, dim=ax, keepdim=o["keep_dims"])
            return r
        if k == "SHAPE":
            return torch.tensor(list(a[0].shape), dtype=torch.int32)
        if k == "SPLIT_V":
            sizes = self._ints(a[1])
            axis = self._ints(a[2])[0]
            return list(torch.split(a[0], sizes, dim=axis))
        if k == "RESHAPE":
            shape = self._ints(a[1]) if len(a) > 1 and a[1] is not None else o["new_shape"]
            return a[0].reshape(shape)
        if k == "PACK":
            return torch.stack([t.reshape(()) if t.dim() == 0 else t for t in a], dim=o["axis"])
        if k == "CONCATENATION":
            return _act(torch.cat(a, dim=o["axis"]), o["act"])
        if k == "STRIDED_SLICE":
            if any(o[m] for m in ("ellipsis_mask", "new_axis_mask")):
                raise NotImplementedError("STRIDED_SLICE ellipsis/new_axis masks")
            b, e, s = self._ints(a[1]), self._ints(a[2]), self._ints(a[3])
            idx = []
            for d, (bi, ei, si) in enumerate(zip(b, e, s)):
                if o["shrink_axis_mask"] >> d & 1:
                    idx.append(bi)                       # take one element, drop the axis
                else:
                    idx.append(slice(None if o["begin_mask"] >> d & 1 else bi,
                                     None if o["end_mask"] >> d & 1 else ei, si))
            return a[0][tuple(idx)]
        if k == "RANGE":
            s, e, d = (self._ints(t)[0] for t in a)
            return torch.arange(s, e, d, dtype=torch.int32)
        if k == "GATHER":
            if o.get("batch_dims", 0):
                raise NotImplementedError("GATHER batch_dims")
            idx = a[1].long()
            axis = o["axis"] % a[0].dim()
            r = torch.index_select(a[0], axis, idx.reshape(-1))
            return r.reshape(list(a[0].shape[:axis]) + list(idx.shape) + list(a[0].shape[axis + 1:]))
        if k == "EXPAND_DIMS":
            return a[0].unsqueeze(self._ints(a[1])[0])
        if k == "SQUEEZE":
            r = a[0]
            for d in sorted((d % a[0].dim() for d in o["squeeze_dims"]), reverse=True):
                r = r.squeeze(d)
            return r
        if k == "RFFT2D":
            h, w = self._ints(a[1])
            return torch.fft.rfft2(a[0].to(self.dtype), s=(h, w))
        if k == "CAST":
            x = a[0]
            # complex64 -> float32 in the BirdNET graph: TFLite's cast keeps the real part
            return x.real.to(self.dtype) if x.is_complex() else x
        if k == "REVERSE_V2":
            return torch.flip(a[0], dims=self._ints(a[1]))
        if k == "TRANSPOSE":
            return a[0].permute(self._ints(a[1]))
        if k == "PAD":
            p = a[1].reshape(-1, 2).tolist()
            flat = []
            for lo, hi in reversed(p):
                flat += [int(lo), int(hi)]
            return F.pad(a[0], flat)
        if k == "FULLY_CONNECTED":
            x = a[0]
            if not o["keep_num_dims"]:
                x = x.reshape(-1, a[1].shape[1])
            y = x @ a[1].t()
            if a[2] is not None:
                y = y + a[2]
            return _act(y, o["act"])
        if k in ("CONV_2D", "DEPTHWISE_CONV_2D"):
            x = a[0].permute(0, 3, 1, 2)                       # NHWC -> NCHW
            if k == "CONV_2D":
                w = a[1].permute(0, 3, 1, 2)                   # OHWI -> OIHW
                groups = 1
            else:
                w = a[1].permute(3, 0, 1, 2)                   # 1HWC -> C1HW
                groups = x.shape[1]
                if o["depth_multiplier"] != 1:
                    raise NotImplementedError
            kh, kw = w.shape[2], w.shape[3]
            sh, sw, dh, dw = o["stride_h"], o["stride_w"], o["dil_h"], o["dil_w"]
            if o["padding"] == 0:                              # SAME
                pt, pb = _same_pad(x.shape[2], kh, sh, dh)
                pl, pr = _same_pad(x.shape[3], kw, sw, dw)
                x = F.pad(x, (pl, pr, pt, pb))
            y = F.conv2d(x, w, a[2], stride=(sh, sw), dilation=(dh, dw), groups=groups)
            return _act(y.permute(0, 2, 3, 1), o["act"])
        if k in ("AVERAGE_POOL_2D", "MAX_POOL_2D"):
            x = a[0].permute(0, 3, 1, 2)
            kh, kw, sh, sw = o["filter_h"], o["filter_w"], o["stride_h"], o["stride_w"]
            if o["padding"] == 0:
                pt, pb = _same_pad(x.shape[2], kh, sh)
                pl, pr = _same_pad(x.shape[3], kw, sw)
                if pt or pb or pl or pr:
                    raise NotImplementedError("padded pooling")
            y = (F.avg_pool2d if k == "AVERAGE_POOL_2D" else F.max_pool2d)(x, (kh, kw), (sh, sw))
            return _act(y.permute(0, 2, 3, 1), o["act"])
        raise NotImplementedError(k)
