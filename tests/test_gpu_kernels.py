"""GPU: each tensor-core kernel of the F16X3 path on its own, against a float64 numpy restatement of the same layer
(SURVEY.md App. C shapes: every MBConv block geometry of BirdNET v2.4, every GEMM shape).

These are unit tests of the kernels behind bnb_predict* (the end-to-end parity tests are in test_gpu_parity.py); the
tolerance is relative to the layer's output range: the 3-term fp16 hi/lo split carries ~21 bits."""
import numpy as np
import pytest

import birdnet_b200 as bb

pytestmark = pytest.mark.gpu
REL_TOL = 3e-5


def _silu(x):
    return x / (1.0 + np.exp(-x))


def _mbconv_ref(x, w_exp, b_exp, w_dw, b_dw, stride):
    x = x.astype(np.float64)
    e = _silu(x @ w_exp.astype(np.float64).T + b_exp.astype(np.float64))          # [B,H,W,C]
    B, H, W, C = e.shape
    ep = np.zeros((B, H + 2, W + 2, C))
    ep[:, 1:-1, 1:-1] = e
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    acc = np.zeros((B, Ho, Wo, C)) + b_dw.astype(np.float64)
    for kh in range(3):
        for kw in range(3):
            acc += ep[:, kh:kh + (Ho - 1) * stride + 1:stride, kw:kw + (Wo - 1) * stride + 1:stride] * w_dw[kh * 3 + kw].astype(np.float64)
    d = _silu(acc)
    return d, d.sum(axis=(1, 2))


def _describe(d, ref):
    """where the error lives: per 32-channel block, per output row, per output column (max abs error / range)"""
    e = np.abs(d - ref) / np.abs(ref).max()
    C = e.shape[-1]
    blocks = [float(e[..., c:c + 32].max()) for c in range(0, C, 32)]
    return "by32ch=%s rows=%s cols=%s chunks=%s" % (np.array2string(np.array(blocks), precision=1), np.array2string(e.max(axis=(0, 2, 3)), precision=1),
                                                  np.array2string(e.max(axis=(0, 1, 3)), precision=1), np.array2string(e.max(axis=(1, 2, 3))[:8], precision=1))


def test_tmem_unaligned_column_loads(lib_path):
    """tcgen05.ld.32x32b at arbitrary column offsets (the depthwise phase walks patch rows of 18 / 17 / 10 columns)."""
    bad18, bad17, bad10, threads = bb.debug_tmem_probe()
    assert threads == 128 and (bad18, bad17, bad10) == (0, 0, 0)


# (name, H, W, Cin, C, stride, B): the eight distinct block geometries of the model; B chosen so that tiles > 148 CTAs
BLOCKS = [
    ("b1", 48, 128, 24, 72, 2, 4), ("b2", 24, 64, 36, 72, 1, 10), ("b4", 24, 64, 36, 288, 2, 13), ("b5", 12, 32, 72, 288, 1, 40),
    ("b8", 12, 32, 72, 864, 2, 40), ("b9", 6, 16, 108, 864, 1, 150), ("b13", 6, 16, 108, 1536, 2, 150), ("b14", 3, 8, 192, 1536, 1, 150),
]


@pytest.mark.parametrize("name,H,W,Cin,C,stride,B", BLOCKS, ids=[b[0] for b in BLOCKS])
@pytest.mark.parametrize("flags", [0, 1], ids=["sw-small", "sw128-only"])
def test_mbconv2_matches_numpy(lib_path, name, H, W, Cin, C, stride, B, flags):
    rng = np.random.default_rng(hash(name) % 1000)
    x = (rng.standard_normal((B, H, W, Cin)) * 3).astype(np.float32)
    x[0, 0, 0, :] = 40.0                                                   # large activations as seen in the real network
    w_exp = (rng.standard_normal((C, Cin)) / np.sqrt(Cin)).astype(np.float32)
    b_exp = rng.standard_normal(C).astype(np.float32)
    w_dw = (rng.standard_normal((9, C)) / 3).astype(np.float32)
    b_dw = rng.standard_normal(C).astype(np.float32)
    d, se, info = bb.debug_mbconv2(x, w_exp, b_exp, w_dw, b_dw, stride, flags)
    ref, ref_se = _mbconv_ref(x, w_exp, b_exp, w_dw, b_dw, stride)
    err = np.abs(d - ref).max() / np.abs(ref).max()
    err_se = np.abs(se - ref_se).max() / np.abs(ref_se).max()
    print(name, info, "rel err %.2e se %.2e" % (err, err_se))
    if err > REL_TOL:
        print(_describe(d, ref))
    assert err <= REL_TOL and err_se <= REL_TOL


def test_mbconv2_single_chunk_and_borders(lib_path):
    """B = 1 (fewer tiles than CTAs) and an input that is non-zero only on the image border (padding / mask paths)."""
    rng = np.random.default_rng(7)
    for (H, W, Cin, C, stride) in [(12, 32, 72, 288, 1), (24, 64, 36, 288, 2)]:
        x = np.zeros((1, H, W, Cin), np.float32)
        x[:, 0] = rng.standard_normal((W, Cin)); x[:, -1] = rng.standard_normal((W, Cin))
        x[:, :, 0] = rng.standard_normal((H, Cin)); x[:, :, -1] = rng.standard_normal((H, Cin))
        w_exp = (rng.standard_normal((C, Cin)) / np.sqrt(Cin)).astype(np.float32)
        b_exp = rng.standard_normal(C).astype(np.float32)
        w_dw = (rng.standard_normal((9, C)) / 3).astype(np.float32)
        b_dw = rng.standard_normal(C).astype(np.float32)
        d, se, _ = bb.debug_mbconv2(x, w_exp, b_exp, w_dw, b_dw, stride)
        ref, ref_se = _mbconv_ref(x, w_exp, b_exp, w_dw, b_dw, stride)
        assert np.abs(d - ref).max() <= REL_TOL * np.abs(ref).max()
        assert np.abs(se - ref_se).max() <= REL_TOL * np.abs(ref_se).max()


# (name, B, map H x W of the GEMM rows, K, N, gated, residual, act, out_mode, consumer (stride, C_exp) or None)
GEMMS = [
    ("b1->b2", 33, 24, 64, 72, 36, False, False, 0, 2, (1, 72)), ("b2->b3", 2, 24, 64, 72, 36, False, True, 0, 2, (1, 72)),
    ("b3->b4", 3, 24, 64, 72, 36, False, True, 0, 2, (2, 288)), ("b4->b5", 5, 12, 32, 288, 72, True, False, 0, 2, (1, 288)),
    ("b5->b6", 4, 12, 32, 288, 72, True, True, 0, 2, (1, 288)), ("b7->b8", 4, 12, 32, 288, 72, True, True, 0, 2, (2, 864)),
    ("b9->b10", 9, 6, 16, 864, 108, True, True, 0, 2, (1, 864)), ("b12->b13", 9, 6, 16, 864, 108, True, True, 0, 2, (2, 1536)),
    ("b14->b15", 11, 3, 8, 1536, 192, True, True, 0, 2, (1, 1536)), ("b16->post", 11, 3, 8, 1536, 192, True, True, 0, 1, None),
    ("post", 37, 1, 6, 1728, 1024, False, False, 1, 0, None), ("fc5", 5, 1, 1, 1024, 6522, False, False, 0, 0, None),
    ("fc256", 256, 1, 1, 1024, 6522, False, False, 0, 0, None), ("silu32", 700, 1, 1, 24, 72, False, False, 2, 0, None),
]


@pytest.mark.parametrize("name,B,H,W,K,N,gated,res,act,mode,cons", GEMMS, ids=[g[0] for g in GEMMS])
def test_pw2_matches_numpy(lib_path, name, B, H, W, K, N, gated, res, act, mode, cons):
    rng = np.random.default_rng(B + K + N)
    M, rpc = B * H * W, H * W
    A = (rng.standard_normal((M, K)) * 2).astype(np.float32)
    Wt = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    gate = rng.uniform(0.0, 1.0, (B, K)).astype(np.float32) if gated else None
    resid = (rng.standard_normal((M, N)) * 5).astype(np.float32) if res else None
    geom = (H, W) + (cons if cons else (1, 128)) if (res or mode == 2) else None
    out, info = bb.debug_pw2(A, Wt, bias, gate, rpc if gated else 0, resid, act, mode, geom)
    A64 = A.astype(np.float64)
    if gated:
        A64 = A64 * np.repeat(gate.astype(np.float64), rpc, axis=0)
    ref = A64 @ Wt.astype(np.float64).T + bias
    if act == 1:
        ref = np.maximum(ref, 0)
    elif act == 2:
        ref = _silu(ref)
    if res:
        ref = ref + resid
    err = np.abs(out - ref).max() / np.abs(ref).max()
    print(name, (M, K, N), info, "rel err %.2e" % err)
    assert err <= REL_TOL
