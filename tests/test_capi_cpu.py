"""CPU: the C-ABI library builds, loads, exports every symbol include/birdnet_b200.h declares, parses the
model into the expected layer plan, and fails closed without a GPU (no compute calls here)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import birdnet_b200 as bb

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(REPO, "include", "birdnet_b200.h")).read()
    return sorted(set(re.findall(r"BNB_API [^;(]*?\b(bnb_[a-z_0-9]+)\(", hdr)))


def test_library_exports_every_declared_symbol(lib_path):
    declared = _declared_symbols()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (bnb_[a-z_0-9]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(bb.SYMBOLS) == set(declared)          # the ctypes binding covers the whole header
    lib = bb.load_library()
    assert lib.bnb_abi_version() == 1


def test_sass_is_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sass_uses_the_blackwell_paths(lib_path):
    """The shipped library's SASS (not just the arch string): the F16X3 kernels issue tcgen05 MMAs (UTCHMMA), read
    TMEM (LDTM), fetch operands with bulk copies (UBLKCP), wait for the previous grid (ACQBULK = griddepcontrol.wait), use the
    256-bit stores, and the MMA / copy issue is NOT wrapped in waterfall loops (BRA.U.ANY) any more; the frontend and stem use the
    packed fp32 pipe.  profiles/r02_sass_tc.txt is the committed listing of the same counts (tools/sass_counts.py)."""
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    per, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per[cur] = []
        elif cur is not None:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                per[cur].append(m.group(1))
    def kernel(sub):
        ks = [k for k in per if sub in k]
        assert ks, sub
        return [op for k in ks for op in per[k]]
    mb2, pw2 = kernel("mbconv2_kernel"), kernel("pw2_kernel")
    for ops in (mb2, pw2):
        assert sum(o.startswith("UTCHMMA") for o in ops) >= 12 and any(o.startswith("LDTM") for o in ops)
        assert any(o.startswith("UBLKCP") for o in ops) and any(o.startswith("ACQBULK") for o in ops)
        assert sum(o.startswith("BRA.U.ANY") for o in ops) <= 1
    assert any(".256" in o and o.startswith("STG") for o in pw2)
    assert sum(o.startswith("FFMA2") for o in kernel("stem_mix_kernel")) > 1000
    assert any(o.startswith("FADD2") for o in kernel("frontend_kernel"))


def test_plan_extraction_matches_the_graph(lib_path):
    d = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    assert d["n_samples"] == 144000 and d["n_species"] == 6522 and d["embedding_dim"] == 1024
    assert [s["frame_len"] for s in d["spec"]] == [2048, 1024] and [s["hop"] for s in d["spec"]] == [278, 280]
    assert all(s["n_frames"] == 511 and s["n_mel"] == 96 for s in d["spec"])
    assert abs(d["spec"][0]["pow"] - 0.22952409) < 1e-6 and abs(d["spec"][1]["pow"] - 0.1905273) < 1e-6
    assert d["frontend_out_tensor"] == 265 and d["stem"]["out_tensor"] == 266 and d["mix"]["out_tensor"] == 270
    assert d["stem"]["pad"] == [1, 3] and d["stem"]["out"] == [48, 256]
    b = d["blocks"]
    assert len(b) == 16
    assert [x["cexp"] for x in b] == [72] * 3 + [288] * 4 + [864] * 5 + [1536] * 4
    assert [x["stride"] for x in b] == [2, 1, 1, 2, 1, 1, 1, 2, 1, 1, 1, 1, 2, 1, 1, 1]
    assert [x["se"] for x in b] == [0] * 3 + [18] * 4 + [27] * 5 + [48] * 4
    assert [x["residual"] for x in b] == [False, True, True, False, True, True, True, False, True, True, True, True, False, True, True, True]
    assert b[0]["in"] == [48, 128, 24] and b[-1]["out"] == [3, 8, 192]
    assert b[-1]["tensors"]["out"] == 541 and b[3]["tensors"]["gate"] == 311
    assert d["post"]["out"] == [1, 6] and d["post"]["emb_tensor"] == 545 and d["head"]["out_tensor"] == 546
    # MAC count of the plan == SURVEY §8(d): 328,305,984 per chunk
    mac = 48 * 256 * 24 * 64 + 48 * 128 * 24 * 48
    for x in b:
        hin, win, cin = x["in"]; ho, wo, co = x["out"]
        mac += hin * win * cin * x["cexp"] + ho * wo * x["cexp"] * 9 + ho * wo * x["cexp"] * co + (2 * x["cexp"] * x["se"])
    mac += 6 * 1728 * 1024 + 1024 * 6522
    assert mac == 328305984


def test_rejects_garbage_and_foreign_models(lib_path):
    with pytest.raises(bb.B200Error) as e:
        bb.describe_model(b"\x1c\x00\x00\x00TFL3" + b"\x00" * 64)
    assert e.value.status == bb.ERR_UNSUPPORTED_MODEL
    with pytest.raises(bb.B200Error):
        bb.describe_model(b"not a flatbuffer at all")
    # the range-filter model in the reference is a valid TFLite file but not this topology; emulate by truncation
    data = open(bb.DEFAULT_MODEL, "rb").read()
    with pytest.raises(bb.B200Error):
        bb.describe_model(data[: len(data) // 2])


def test_create_fails_closed_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bb.B200Unavailable) as e:
        bb.B200Classifier()
    assert e.value.status == bb.ERR_NO_DEVICE          # caller falls back to TFLite (birdnet.go:321-335)
    lib = bb.load_library()
    assert lib.bnb_num_species(None) == bb.ERR_INVALID_ARGUMENT
    lib.bnb_classifier_destroy(None)                   # Close() on nil is a no-op
    assert "NULL" in bb.last_error() or bb.last_error()


def test_every_gemm_layer_fits_the_tensor_core_kernel(lib_path):
    """Host-side tiling of the tcgen05 GEMM for every pointwise / post / head layer at micro-batch, full-batch and
    batch-1 row counts: >= 2 pipeline stages, N tile <= 256 TMEM columns, <= 227 KB shared memory, 32-column store
    boxes never straddle two N tiles."""
    d = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    shapes = []
    for b in d["blocks"]:
        hin, win, cin = b["in"]; ho, wo, co = b["out"]
        shapes += [(hin * win, b["cexp"], cin), (ho * wo, co, b["cexp"])]
    shapes += [(6, 1024, 1728), (1, 6522, 1024)]
    for rows, N, K in shapes:
        for chunks in (1, 16, 32, 64, 256, 1024):
            bn, stages, smem = bb.pw_tiling(rows * chunks, N, K)
            n_pad = (N + 15) // 16 * 16
            assert 16 <= bn <= 256 and bn % 16 == 0 and stages >= 2 and smem <= 227 * 1024, (rows * chunks, N, K, bn, stages, smem)
            if bn < n_pad:
                assert bn % 32 == 0, (N, bn)


def test_fused_block_geometry_covers_every_block(lib_path):
    """Fused expand+depthwise kernel: for each of the 16 blocks the output tiles cover the image exactly once, the halo
    patch fits the 128 GEMM rows, shared memory fits, and SE blocks need <= 32 partial-sum slots per chunk."""
    d = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    for b in d["blocks"]:
        hin, win, cin = b["in"]; ho, wo, _ = b["out"]
        for B in (0, 1, 64, 256):                                            # the tile search is cost-model driven: any batch size
            g = bb.mbconv_geometry(hin, win, ho, wo, b["stride"], cin, b["cexp"], B, 32 if b["se"] else 0)
            assert g["ph"] == (g["th"] - 1) * b["stride"] + 3 and g["pw"] == (g["tw"] - 1) * b["stride"] + 3
            assert g["ph"] * g["pw"] <= 128
            # the shared-memory plan fits for every block the engine fuses by default (input map >= 256 pixels, K <= 128);
            # blocks with K > 128 fall back to the two-kernel chain (engine.cu checks the same limit)
            if hin * win >= 256 or cin <= 128:
                assert g["smem_bytes"] <= 227 * 1024, (b, g)
            assert g["tiles_h"] * g["th"] >= ho and (g["tiles_h"] - 1) * g["th"] < ho
            assert g["tiles_w"] * g["tw"] >= wo and (g["tiles_w"] - 1) * g["tw"] < wo
            assert g["k_stages"] == (cin + 63) // 64 and g["a_slots"] >= g["k_stages"]
            if b["se"]:
                assert g["tiles_h"] * g["tiles_w"] <= 32


def test_f16x3_plans_cover_every_block(lib_path):
    """Round-2 kernels (mbconv2.cu / pw2.cu): every block has a plan; a pair of halo patches fits one MMA (2 * n_mma <= 256
    TMEM columns, n_mma % 16 == 0), the accumulators fit the 512 TMEM columns at least twice, tiles cover the output exactly,
    shared memory <= 227 KB; the 288 -> 72 project GEMMs keep their weights resident with a 3-deep A ring (the r02 fix)."""
    d = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    for i, b in enumerate(d["blocks"]):
        hin, win, cin = b["in"]; ho, wo, co = b["out"]
        p = bb.mb2_plan(hin, win, ho, wo, b["stride"], cin, b["cexp"])
        assert p["ok"] == 1, (i, p)
        assert p["ph"] == (p["th"] - 1) * b["stride"] + 3 and p["pw"] == (p["tw"] - 1) * b["stride"] + 3
        assert p["ph"] * p["pw"] <= p["n_mma"] <= 128 and p["n_mma"] % 16 == 0
        assert ho % p["th"] == 0 and wo % p["tw"] == 0
        assert p["units"] == (b["cexp"] + 127) // 128 and p["k_stages"] >= 1
        assert p["pair_slots"] >= 1 and p["smem_bytes"] <= 227 * 1024, (i, p)
        assert p["pair_bytes"] % 1024 == 0                                   # swizzle atoms of the second tile stay aligned
        for chunks in (1, 64, 128, 256, 1024):
            bn, stages, b_res, n_tiles, smem = bb.pw2_tiling(ho * wo * chunks, co, b["cexp"], b["se"] > 0)
            assert 16 <= bn <= 256 and bn % 16 == 0 and stages >= 2 and smem <= 227 * 1024, (i, chunks, bn, stages, smem)
            assert n_tiles * bn >= co
        if b["cexp"] == 288 and b["stride"] == 1:
            bn, stages, b_res, n_tiles, smem = bb.pw2_tiling(ho * wo * 128, co, b["cexp"], True)
            assert b_res == 1 and stages >= 3, (bn, stages, b_res)
    for M, N, K in ((6 * 256, 1024, 1728), (256, 6522, 1024), (1, 6522, 1024)):
        bn, stages, b_res, n_tiles, smem = bb.pw2_tiling(M, N, K, False)
        assert stages >= 2 and smem <= 227 * 1024 and n_tiles * bn >= N


def test_go_backend_binds_only_declared_symbols(lib_path):
    """The Go side (go/internal/inference/b200, not compilable here: no Go toolchain) must reference only functions and constants
    that include/birdnet_b200.h declares and the library exports; the stub must offer the same method set as the cgo file."""
    import re
    root = os.path.join(REPO, "go", "internal", "inference", "b200")
    hdr = open(os.path.join(REPO, "include", "birdnet_b200.h")).read()
    cgo = open(os.path.join(root, "backend_b200.go")).read()
    stub = open(os.path.join(root, "stub_nob200.go")).read()
    used = set(re.findall(r"\bC\.(bnb_\w+|BNB_\w+)", cgo))
    assert len(used) >= 12
    lib = bb.load_library()
    for name in sorted(used):
        assert re.search(r"\b%s\b" % name, hdr), name
        if name.startswith("bnb_") and not name.endswith("_t") and name not in ("bnb_classifier", "bnb_options"):
            assert hasattr(lib, name), name
    assert "//go:build b200" in cgo and "//go:build !b200" in stub
    meth = lambda src: set(re.findall(r"^func \(c \*Classifier\) (\w+)\(", src, re.M)) | set(re.findall(r"^func (\w+)\(", src, re.M))
    assert meth(cgo) - {"sizeMismatch", "lastErr"} <= meth(stub) | {"analyze"}, sorted(meth(cgo) - meth(stub))
