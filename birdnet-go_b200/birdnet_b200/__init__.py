"""ctypes binding of libbirdnet_b200.so + the host-side mirror of the reference's classifier surface.

`B200Classifier` mirrors `inference.Classifier` / `EmbeddingExtractor`
(/root/reference/internal/inference/backend.go:8-29): `predict(samples) -> raw logits`,
`num_species()`, `close()`.  `BirdNET` mirrors `BirdNET.Predict`
(/root/reference/internal/classifier/analyze.go:25-110): sensitivity-sigmoid, label pairing,
top-10.  Everything numeric happens in the CUDA library; this module never computes the model
on the CPU and raises if the library cannot be loaded (no fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libbirdnet_b200.so")
REPO = os.path.dirname(os.path.dirname(_HERE))
DEFAULT_MODEL = os.path.join(REPO, "assets", "BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite")
DEFAULT_LABELS = os.path.join(REPO, "assets", "BirdNET_GLOBAL_6K_V2.4_Labels_en_us.txt")

PCM_F32, PCM_S16 = 0, 1
PRECISION_DEFAULT, PRECISION_F32, PRECISION_F16X3 = 0, 1, 2
OK = 0
ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_UNSUPPORTED_MODEL, ERR_CUDA, ERR_CLOSED, ERR_OOM, ERR_INTERNAL = -1, -2, -3, -4, -5, -6, -7
DEFAULT_TOP_K = 10   # internal/classifier/tracing.go:59


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("max_batch", C.c_int32), ("micro_batch", C.c_int32),
                ("precision", C.c_int32), ("use_graphs", C.c_int32), ("lanes", C.c_int32), ("reserved", C.c_int32 * 8)]


class B200Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("b200: status=%d: %s" % (status, msg))
        self.status = status


class B200Unavailable(B200Error):
    """Mirror of ErrOpenVINOUnavailable (openvino/openvino.go:31): caller should fall back."""


_lib = None

# name -> (restype, argtypes): every symbol include/birdnet_b200.h declares
SYMBOLS = {
    "bnb_init": (C.c_int, []),
    "bnb_device_count": (C.c_int, []),
    "bnb_abi_version": (C.c_int, []),
    "bnb_last_error": (C.c_char_p, []),
    "bnb_classifier_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(Options), C.POINTER(C.c_void_p)]),
    "bnb_classifier_destroy": (None, [C.c_void_p]),
    "bnb_num_species": (C.c_int, [C.c_void_p]),
    "bnb_num_samples": (C.c_int, [C.c_void_p]),
    "bnb_embedding_dim": (C.c_int, [C.c_void_p]),
    "bnb_max_batch": (C.c_int, [C.c_void_p]),
    "bnb_runtime_device": (C.c_char_p, [C.c_void_p]),
    "bnb_runtime_precision": (C.c_char_p, [C.c_void_p]),
    "bnb_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "bnb_predict_with_embeddings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "bnb_predict_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bnb_analyze_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bnb_analyze_batch_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "bnb_wait": (C.c_int, [C.c_void_p, C.c_int32]),
    "bnb_predict_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bnb_analyze_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bnb_kernel_launches": (C.c_int64, [C.c_void_p]),
    "bnb_last_device_ms": (C.c_float, [C.c_void_p]),
    "bnb_profile_begin": (C.c_int, [C.c_void_p]),
    "bnb_profile_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "bnb_profile_launches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "bnb_debug_pw_tiling": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "bnb_analyze_batch_detections": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "bnb_ultrasonic_cv_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bnb_dense_head_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bnb_debug_mb2_plan": (C.c_int, [C.c_int] * 7 + [C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "bnb_debug_pw2_tiling": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "bnb_debug_mbconv_geometry": (C.c_int, [C.c_int] * 9 + [C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "bnb_describe_model": (C.c_int, [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "bnb_debug_read_tensor": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "bnb_debug_keep_intermediates": (C.c_int, [C.c_void_p, C.c_int]),
    "bnb_range_filter_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "bnb_range_filter_destroy": (None, [C.c_void_p]),
    "bnb_range_filter_num_species": (C.c_int, [C.c_void_p]),
    "bnb_range_filter_predict": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "bnb_range_filter_predict_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bnb_debug_tmem_probe": (C.c_int, [C.c_void_p]),
    "bnb_debug_mbconv2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bnb_debug_pw2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def load_library(path=LIB_PATH):
    """dlopen the in-tree library (built by birdnet_b200.build).  Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise B200Unavailable(ERR_NO_DEVICE, "libbirdnet_b200.so not built (%s); run `python -m birdnet_b200.build`" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)           # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error():
    return load_library().bnb_last_error().decode("utf-8", "replace")


def _check(rc):
    if rc < 0:
        msg = last_error()
        raise (B200Unavailable if rc == ERR_NO_DEVICE else B200Error)(rc, msg)
    return rc


def pw_tiling(M, N, K):
    """(bn, stages, smem_bytes) the tcgen05 GEMM launcher picks for an [M,K] x [N,K]^T layer."""
    lib = load_library()
    bn, st, sm = C.c_int(), C.c_int(), C.c_int64()
    _check(lib.bnb_debug_pw_tiling(M, N, K, C.byref(bn), C.byref(st), C.byref(sm)))
    return bn.value, st.value, sm.value


def ultrasonic_cv_batch(pcm, sample_rate, fft_size=8192, hop_size=4096, frequency_split_hz=20000, device=-1):
    """N2: ultrasonic.ComputeUSFrameCV for a batch of chunks on the GPU.  pcm: [B, n] int16 (scaled by 1/32768) or float32.
    Returns (cv float64 [B], ok bool [B]); ok is False exactly where the reference returns (0, false)."""
    lib = load_library()
    pcm = np.ascontiguousarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[None]
    if pcm.dtype not in (np.int16, np.float32):
        raise ValueError("pcm must be int16 or float32")
    B, n = pcm.shape
    cv = np.zeros(B, np.float64); ok = np.zeros(B, np.int32)
    _check(lib.bnb_ultrasonic_cv_batch(device, _ptr(pcm), PCM_S16 if pcm.dtype == np.int16 else PCM_F32, B, n, int(sample_rate), int(fft_size),
                                       int(hop_size), int(frequency_split_hz), _ptr(cv), _ptr(ok)))
    return cv, ok.astype(bool)


def dense_head_batch(embeddings, weights, bias, device=-1):
    """N2: custom classification head on embeddings: sigmoid(embeddings @ weights.T + bias), float32, on the GPU."""
    lib = load_library()
    e = np.ascontiguousarray(embeddings, np.float32); w = np.ascontiguousarray(weights, np.float32); b = np.ascontiguousarray(bias, np.float32)
    if e.ndim != 2 or w.ndim != 2 or e.shape[1] != w.shape[1] or b.shape != (w.shape[0],):
        raise ValueError("shapes: embeddings [B, n_in], weights [n_out, n_in], bias [n_out]")
    out = np.empty((e.shape[0], w.shape[0]), np.float32)
    _check(lib.bnb_dense_head_batch(device, _ptr(e), e.shape[0], e.shape[1], _ptr(w), _ptr(b), w.shape[0], _ptr(out)))
    return out


def mb2_plan(H, W, Ho, Wo, stride, Cin, C_exp):
    """Plan of the F16X3 fused expand+depthwise kernel (mbconv2.cu) for one block: tile, MMA width, weight residency, smem."""
    lib = load_library()
    out = (C.c_int * 12)()
    sm = C.c_int64()
    _check(lib.bnb_debug_mb2_plan(H, W, Ho, Wo, stride, Cin, C_exp, out, C.byref(sm)))
    keys = ("ok", "th", "tw", "ph", "pw", "n_mma", "units", "k_stages", "resident", "a_slots", "pair_slots", "pair_bytes")
    d = dict(zip(keys, list(out)))
    d["smem_bytes"] = sm.value
    return d


def pw2_tiling(M, N, K, gated=False):
    """(bn, stages, weights_resident, n_tiles, smem_bytes) of the F16X3 project GEMM (pw2.cu)."""
    lib = load_library()
    out = (C.c_int * 4)()
    sm = C.c_int64()
    _check(lib.bnb_debug_pw2_tiling(M, N, K, 1 if gated else 0, out, C.byref(sm)))
    return out[0], out[1], out[2], out[3], sm.value


def mbconv_geometry(H, W, Ho, Wo, stride, Cin, C_exp=0, B=0, max_tiles=0):
    """Tile geometry + smem bytes of the fused expand+depthwise kernel for one block."""
    lib = load_library()
    out = (C.c_int * 10)()
    sm = C.c_int64()
    _check(lib.bnb_debug_mbconv_geometry(H, W, Ho, Wo, stride, Cin, C_exp, B, max_tiles, out, C.byref(sm)))
    keys = ("th", "tw", "ph", "pw", "tiles_h", "tiles_w", "k_stages", "box_c", "a_slots", "b_slots")
    d = dict(zip(keys, list(out)))
    d["smem_bytes"] = sm.value
    return d


def describe_model(model_bytes: bytes) -> str:
    lib = load_library()
    buf = C.create_string_buffer(1 << 16)
    n = _check(lib.bnb_describe_model(model_bytes, len(model_bytes), buf, len(buf)))
    return buf.raw[:n].decode()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _Ticket:
    def __init__(self, clf, ticket, pcm, idx, conf, logits):
        self._clf, self._t, self._pcm, self._idx, self._conf, self._logits = clf, ticket, pcm, idx, conf, logits

    def wait(self):
        _check(self._clf._lib.bnb_wait(self._clf._h, self._t))
        return (self._idx, self._conf, self._logits) if self._logits is not None else (self._idx, self._conf)


class B200Classifier:
    """inference.Classifier + EmbeddingExtractor backed by libbirdnet_b200 (NOT thread-safe: backend.go:7)."""

    def __init__(self, model_data: bytes | None = None, device=-1, max_batch=256, micro_batch=0,
                 precision=PRECISION_DEFAULT, use_graphs=0, lanes=0):
        lib = load_library()
        if model_data is None:
            with open(DEFAULT_MODEL, "rb") as f:
                model_data = f.read()
        o = Options()
        o.struct_size = C.sizeof(Options)
        o.device, o.max_batch, o.micro_batch, o.precision, o.use_graphs = device, max_batch, micro_batch, precision, use_graphs
        o.lanes = lanes            # concurrent front-phase streams (0 = library default)
        h = C.c_void_p()
        _check(lib.bnb_classifier_create(model_data, len(model_data), C.byref(o), C.byref(h)))
        self._lib, self._h = lib, h
        self.n_species = lib.bnb_num_species(h)
        self.n_samples = lib.bnb_num_samples(h)
        self.emb_dim = lib.bnb_embedding_dim(h)
        self.max_batch = lib.bnb_max_batch(h)

    # --- inference.Classifier ---------------------------------------------------------------------
    def predict(self, samples):
        """Raw logits (pre-activation) for exactly `n_samples` float32 samples."""
        if self._h is None:
            raise B200Error(ERR_CLOSED, "classifier is closed")
        x = np.ascontiguousarray(samples, np.float32).reshape(-1)
        out = np.empty(self.n_species, np.float32)
        _check(self._lib.bnb_predict(self._h, _ptr(x), x.size, _ptr(out)))
        return out

    def predict_with_embeddings(self, samples):
        if self._h is None:
            raise B200Error(ERR_CLOSED, "classifier is closed")
        x = np.ascontiguousarray(samples, np.float32).reshape(-1)
        out = np.empty(self.n_species, np.float32)
        emb = np.empty(self.emb_dim, np.float32)
        _check(self._lib.bnb_predict_with_embeddings(self._h, _ptr(x), x.size, _ptr(out), _ptr(emb)))
        return out, emb

    def num_species(self):
        return self.n_species

    def close(self):
        if self._h is not None:
            self._lib.bnb_classifier_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- batched surface ---------------------------------------------------------------------------
    def _fmt(self, pcm):
        if pcm.dtype == np.int16:
            return PCM_S16
        if pcm.dtype == np.float32:
            return PCM_F32
        raise TypeError("pcm must be float32 or int16")

    def _check_batch(self, pcm, k=None):
        """Shared argument validation of the batch entry points: the C side copies B*n_samples samples out of the buffer,
        so a wrong-width or 1-D array must be rejected here with the reference's "input size mismatch" error
        (tflite/classifier.go:102-104)."""
        if self._h is None:
            raise B200Error(ERR_CLOSED, "classifier is closed")
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim != 2 or pcm.shape[1] != self.n_samples:
            raise B200Error(ERR_INVALID_ARGUMENT, "input size mismatch: expected [B,%d], got %s" % (self.n_samples, pcm.shape))
        if k is not None and not (0 < int(k) <= 64):
            raise B200Error(ERR_INVALID_ARGUMENT, "k must be in 1..64, got %r" % (k,))
        return pcm

    def predict_batch(self, pcm, with_embeddings=False, out=None):
        """pcm [B,144000] float32|int16 (host) -> logits [B,6522] (and embeddings [B,1024])."""
        pcm = self._check_batch(pcm)
        B = pcm.shape[0]
        logits = out if out is not None else np.empty((B, self.n_species), np.float32)
        emb = np.empty((B, self.emb_dim), np.float32) if with_embeddings else None
        _check(self._lib.bnb_predict_batch(self._h, _ptr(pcm), self._fmt(pcm), B, _ptr(logits), _ptr(emb) if emb is not None else None))
        return (logits, emb) if with_embeddings else logits

    def analyze_batch(self, pcm, sensitivity=1.0, k=DEFAULT_TOP_K, want_logits=False):
        pcm = self._check_batch(pcm, k)
        B = pcm.shape[0]
        idx = np.empty((B, k), np.int32)
        conf = np.empty((B, k), np.float32)
        logits = np.empty((B, self.n_species), np.float32) if want_logits else None
        _check(self._lib.bnb_analyze_batch(self._h, _ptr(pcm), self._fmt(pcm), B, float(sensitivity), k, _ptr(idx), _ptr(conf),
                                           _ptr(logits) if logits is not None else None))
        return (idx, conf, logits) if want_logits else (idx, conf)

    def analyze_batch_submit(self, pcm, sensitivity=1.0, k=DEFAULT_TOP_K, want_logits=False):
        """Asynchronous analyze_batch: returns a ticket object; call .wait() for (idx, conf[, logits]).  At most two tickets may be
        outstanding.  `pcm` must stay alive and unchanged until wait() (it is referenced by the ticket)."""
        pcm = self._check_batch(pcm, k)
        B = pcm.shape[0]
        idx = np.empty((B, k), np.int32)
        conf = np.empty((B, k), np.float32)
        logits = np.empty((B, self.n_species), np.float32) if want_logits else None
        t = C.c_int32()
        _check(self._lib.bnb_analyze_batch_submit(self._h, _ptr(pcm), self._fmt(pcm), B, float(sensitivity), k, _ptr(idx), _ptr(conf),
                                                  _ptr(logits) if logits is not None else None, C.byref(t)))
        return _Ticket(self, t.value, pcm, idx, conf, logits)

    def analyze_batch_detections(self, pcm, sensitivity=1.0, threshold=0.1, k=DEFAULT_TOP_K, max_det=None):
        """N1: sigmoid, threshold and compaction on the device.  Returns (chunk, species_idx, conf, counts): the detections with
        conf >= threshold among each chunk's top-k, ordered by chunk then by descending confidence; counts[b] per chunk."""
        pcm = self._check_batch(pcm, k)
        B = pcm.shape[0]
        cap = B * k if max_det is None else int(max_det)
        chunk = np.empty(max(cap, 1), np.int32); idx = np.empty(max(cap, 1), np.int32); conf = np.empty(max(cap, 1), np.float32)
        counts = np.zeros(max(B, 1), np.int32)
        n = C.c_int32()
        _check(self._lib.bnb_analyze_batch_detections(self._h, _ptr(pcm), self._fmt(pcm), B, float(sensitivity), float(threshold), k, cap,
                                                      _ptr(chunk), _ptr(idx), _ptr(conf), _ptr(counts), C.byref(n)))
        m = min(n.value, cap)
        return chunk[:m], idx[:m], conf[:m], counts[:B]

    # raw-pointer variants (device memory owned by the caller, e.g. torch tensors' data_ptr())
    def predict_batch_device(self, d_pcm, fmt, B, d_logits, d_emb=0, stream=0):
        _check(self._lib.bnb_predict_batch_device(self._h, d_pcm, fmt, B, d_logits, d_emb or None, stream or None))

    def analyze_batch_device(self, d_pcm, fmt, B, sensitivity, k, d_idx, d_conf, d_logits=0, stream=0):
        _check(self._lib.bnb_analyze_batch_device(self._h, d_pcm, fmt, B, float(sensitivity), k, d_idx, d_conf, d_logits or None, stream or None))

    # --- introspection -------------------------------------------------------------------------------
    def kernel_launches(self):
        return int(self._lib.bnb_kernel_launches(self._h))

    def last_device_ms(self):
        return float(self._lib.bnb_last_device_ms(self._h))

    def runtime_info(self):
        return (self._lib.bnb_runtime_device(self._h).decode(), "B200-native", self._lib.bnb_runtime_precision(self._h).decode())

    PROFILE_CATEGORIES = ("minmax", "frontend", "stem_mix", "pw_expand", "depthwise", "se_gate", "pw_project", "post_conv",
                          "row_mean", "fc_head", "topk")

    def profile_begin(self):
        _check(self._lib.bnb_profile_begin(self._h))

    def profile_end(self):
        """-> {category: (ms, launches)} accumulated since profile_begin()."""
        n = len(self.PROFILE_CATEGORIES)
        ms = np.zeros(n, np.float32)
        cnt = np.zeros(n, np.int64)
        _check(self._lib.bnb_profile_end(self._h, _ptr(ms), _ptr(cnt), n))
        return {c: (float(ms[i]), int(cnt[i])) for i, c in enumerate(self.PROFILE_CATEGORIES)}

    def profile_launches(self, cap=1 << 16):
        """[(category, ms)] per launch of the region closed by the last profile_end(), in issue order."""
        ms = np.zeros(cap, np.float32)
        cat = np.zeros(cap, np.int32)
        n = _check(self._lib.bnb_profile_launches(self._h, _ptr(ms), _ptr(cat), cap))
        return [(self.PROFILE_CATEGORIES[int(cat[i])], float(ms[i])) for i in range(n)]

    def keep_intermediates(self, on=True):
        _check(self._lib.bnb_debug_keep_intermediates(self._h, int(on)))

    def read_tensor(self, tensor, max_elems=1 << 26):
        buf = np.empty(max_elems, np.float32)
        n = _check(self._lib.bnb_debug_read_tensor(self._h, tensor, _ptr(buf), buf.size))
        return buf[:n].copy()


# --- kernel-level test hooks (tests/test_gpu_kernels.py) ------------------------------------------------------------
def debug_tmem_probe():
    out = np.zeros(4, np.int32)
    _check(load_library().bnb_debug_tmem_probe(_ptr(out)))
    return out


def debug_mbconv2(x, w_exp, b_exp, w_dw, b_dw, stride, flags=0):
    """x [B,H,W,Cin], w_exp [C,Cin], w_dw [9,C] -> (d [B,Ho,Wo,C], se_sum [B,C], info dict) through mbconv2_kernel."""
    x = np.ascontiguousarray(x, np.float32); w_exp = np.ascontiguousarray(w_exp, np.float32)
    b_exp = np.ascontiguousarray(b_exp, np.float32); w_dw = np.ascontiguousarray(w_dw, np.float32); b_dw = np.ascontiguousarray(b_dw, np.float32)
    B, H, W, Cin = x.shape
    Cc = w_exp.shape[0]
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    d = np.zeros((B, Ho, Wo, Cc), np.float32)
    se = np.zeros((B, Cc), np.float32)
    info = np.zeros(10, np.int32)
    _check(load_library().bnb_debug_mbconv2(_ptr(x), B, H, W, Cin, _ptr(w_exp), _ptr(b_exp), _ptr(w_dw), _ptr(b_dw), Cc, stride, flags,
                                            _ptr(d), _ptr(se), _ptr(info)))
    keys = ("TH", "TW", "PH", "PW", "n_mma", "k_stages", "a_resident", "a_slots", "b_slots", "smem")
    return d, se, dict(zip(keys, info.tolist()))


def debug_pw2(A, Wt, bias, gate=None, rows_per_chunk=0, residual=None, act=0, out_mode=1, geom=None):
    """out [M,N] = act(A' W^T + bias) (+ residual) through pw2_kernel; A [M,K], Wt [N,K].  out_mode 0 fp32, 1 plain planes,
    2 patch image of a consuming block with input map geom = (H, W, stride, C_exp); a residual needs geom (H, W, ..) too."""
    A = np.ascontiguousarray(A, np.float32); Wt = np.ascontiguousarray(Wt, np.float32); bias = np.ascontiguousarray(bias, np.float32)
    M, K = A.shape
    N = Wt.shape[0]
    out = np.zeros((M, N), np.float32)
    info = np.zeros(4, np.int32)
    g = np.ascontiguousarray(gate, np.float32) if gate is not None else None
    r = np.ascontiguousarray(residual, np.float32) if residual is not None else None
    gm = np.ascontiguousarray(list(geom) + [0] * (4 - len(geom)), np.int32) if geom is not None else None
    _check(load_library().bnb_debug_pw2(_ptr(A), M, K, _ptr(Wt), _ptr(bias), N, _ptr(g) if g is not None else None, rows_per_chunk,
                                        _ptr(r) if r is not None else None, act, int(out_mode), _ptr(gm) if gm is not None else None,
                                        _ptr(out), _ptr(info)))
    return out, dict(zip(("bn", "stages", "b_res", "smem"), info.tolist()))


DEFAULT_RANGE_MODEL = os.path.join(REPO, "assets", "BirdNET_GLOBAL_6K_V2.4_MData_Model_V2_FP16.tflite")


class B200RangeFilter:
    """inference.RangeFilter + BatchRangeFilter (backend.go:55-76) on the GPU: species occurrence scores for
    (latitude, longitude, week).  NOT thread-safe, like the reference's."""

    def __init__(self, model_data: bytes | None = None, device=-1):
        lib = load_library()
        if model_data is None:
            with open(DEFAULT_RANGE_MODEL, "rb") as f:
                model_data = f.read()
        h = C.c_void_p()
        _check(lib.bnb_range_filter_create(model_data, len(model_data), device, C.byref(h)))
        self._lib, self._h = lib, h
        self.n_species = lib.bnb_range_filter_num_species(h)

    def num_species(self):
        return self.n_species

    def predict(self, latitude, longitude, week):
        if self._h is None:
            raise B200Error(ERR_CLOSED, "range filter is closed")
        out = np.empty(self.n_species, np.float32)
        _check(self._lib.bnb_range_filter_predict(self._h, float(latitude), float(longitude), float(week), _ptr(out)))
        return out

    def predict_batch(self, inputs, batch_size=None):
        """inputs: flat [batch*3] or [batch, 3] float32 (lat, lon, week) -> [batch, n_species] scores (PredictBatch, backend.go:70-76)."""
        if self._h is None:
            raise B200Error(ERR_CLOSED, "range filter is closed")
        x = np.ascontiguousarray(inputs, np.float32).reshape(-1)
        n = x.size // 3 if batch_size is None else int(batch_size)
        if x.size != n * 3:
            raise B200Error(ERR_INVALID_ARGUMENT, "input size mismatch: expected %d values, got %d" % (n * 3, x.size))
        out = np.empty((n, self.n_species), np.float32)
        _check(self._lib.bnb_range_filter_predict_batch(self._h, _ptr(x), n, _ptr(out)))
        return out

    def close(self):
        if self._h is not None:
            self._lib.bnb_range_filter_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Result:
    __slots__ = ("species", "confidence")

    def __init__(self, species, confidence):
        self.species, self.confidence = species, confidence

    def __repr__(self):
        return "Result(%r, %.4f)" % (self.species, self.confidence)


class BirdNET:
    """Mirror of classifier.BirdNET's Predict surface (analyze.go:25-110) on the B200 backend."""

    def __init__(self, labels=None, sensitivity=1.0, **kw):
        if labels is None:
            with open(DEFAULT_LABELS, encoding="utf-8") as f:
                labels = [ln.rstrip("\n") for ln in f if ln.strip()]
        self.classifier = B200Classifier(**kw)
        if len(labels) != self.classifier.num_species():     # validateModelAndLabels, birdnet.go:1248-1282
            n = self.classifier.num_species()
            self.classifier.close()
            raise ValueError("mismatched labels and predictions lengths: %d vs %d" % (len(labels), n))
        self.labels, self.sensitivity = labels, sensitivity

    def predict(self, sample):
        """sample: [[float32 x 144000]] like Predict(ctx, [][]float32); uses sample[0] (analyze.go:60)."""
        if self.classifier is None:
            raise RuntimeError("classifier backend is not initialized")
        if len(sample) == 0 or len(sample[0]) == 0:
            raise ValueError("empty audio sample")
        x = np.ascontiguousarray(sample[0], np.float32)[None]
        if x.shape[1] != self.classifier.n_samples:
            raise B200Error(ERR_INVALID_ARGUMENT, "input size mismatch: expected %d samples, got %d" % (self.classifier.n_samples, x.shape[1]))
        idx, conf = self.classifier.analyze_batch(x, self.sensitivity, DEFAULT_TOP_K)
        return [Result(self.labels[i], float(c)) for i, c in zip(idx[0], conf[0])]

    def predict_batch(self, chunks, k=DEFAULT_TOP_K):
        idx, conf = self.classifier.analyze_batch(chunks, self.sensitivity, k)
        return [[Result(self.labels[i], float(c)) for i, c in zip(ri, rc)] for ri, rc in zip(idx, conf)]

    def analyze_file(self, pcm_int16, overlap_s=0.0, threshold=0.1, batch=None, sample_rate=48000):
        """Batched offline file analysis (BASELINE config 2; doc/wiki/file-analysis.md:1-13): slide a 3 s window with `overlap_s`
        seconds of overlap over mono int16 PCM, run ALL windows through the batched int16 detection entry point (device-side /32768,
        sigmoid(sensitivity * logit), top-10, threshold and compaction: bnb_analyze_batch_detections) which keeps, per window,
        the results at or above `threshold`
        (the reference's file mode keeps what passes --threshold).  Returns [(begin_s, end_s, species, confidence)], window order,
        descending confidence inside a window.  The trailing partial window is zero-padded when it holds >= 1.5 s."""
        pcm = np.ascontiguousarray(pcm_int16, np.int16).reshape(-1)
        n = self.classifier.n_samples
        step = n - int(round(overlap_s * sample_rate))
        if step <= 0:
            raise ValueError("overlap must be shorter than the 3 s window")
        starts = list(range(0, max(len(pcm) - n, 0) + 1, step)) if len(pcm) >= n else []
        tail = starts[-1] + step if starts else 0
        if len(pcm) - tail >= n // 2:                         # last partial window, padded with silence
            starts.append(tail)
        B = batch or self.classifier.max_batch
        out = []
        for i in range(0, len(starts), B):
            win = np.zeros((len(starts[i:i + B]), n), np.int16)
            for j, s0 in enumerate(starts[i:i + B]):
                seg = pcm[s0:s0 + n]
                win[j, :len(seg)] = seg
            # threshold + compaction on the device (N1): only the detections come back, already in window / confidence order
            ch, sp, cf, _ = self.classifier.analyze_batch_detections(win, self.sensitivity, threshold, DEFAULT_TOP_K)
            for j, ii, cc in zip(ch, sp, cf):
                s0 = starts[i + int(j)]
                out.append((s0 / sample_rate, (s0 + n) / sample_rate, self.labels[ii], float(cc)))
        return out

    def close(self):
        if self.classifier is not None:
            self.classifier.close()
            self.classifier = None
