// debug_api.cu — test hooks that run ONE tensor-core kernel on host data (tests/test_gpu_kernels.py compares them with
// numpy restatements of the same layer).  Not part of the product path; declared in include/birdnet_b200.h under
// "introspection / test hooks".
#include <cuda_fp16.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/birdnet_b200.h"
#include "common.cuh"
#include "layouts.h"
#include "mbconv2.h"
#include "pw2.h"
#include "tc_common.cuh"

using namespace bnb;

namespace bnb {
int capi_fail(int code, const std::string& msg);       // capi.cu
void tc_prepare_device(int device);
void tc_forget_devices();
}  // namespace bnb

namespace {

struct DevBuf {
  void* p = nullptr;
  explicit DevBuf(size_t bytes) { BNB_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); BNB_CUDA(cudaMemset(p, 0, bytes ? bytes : 16)); }
  ~DevBuf() { cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

void split1(float v, __half* h, __half* l) { *h = __float2half_rn(v); *l = __float2half_rn(v - __half2float(*h)); }

// [M][K] fp32 -> RowTiles image
std::vector<uint8_t> rows_encode(const float* x, long long M, int K) {
  const RowTiles t = RowTiles::make(K);
  std::vector<uint8_t> img(t.bytes(M), 0);
  for (long long m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      __half h, l; split1(x[m * K + k], &h, &l);
      const size_t off = t.piece(m, k >> 3) + (size_t)(k & 7) * 2;
      memcpy(img.data() + off, &h, 2); memcpy(img.data() + off + 16384, &l, 2);
    }
  return img;
}
void rows_decode(const uint8_t* img, long long M, int K, float* out) {
  const RowTiles t = RowTiles::make(K);
  for (long long m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      __half h, l;
      const size_t off = t.piece(m, k >> 3) + (size_t)(k & 7) * 2;
      memcpy(&h, img + off, 2); memcpy(&l, img + off + 16384, 2);
      out[m * K + k] = __half2float(h) + __half2float(l);
    }
}
// [B][H][W][C] fp32 -> PatchTiles image (every pixel into every tile that holds it)
std::vector<uint8_t> patch_encode(const float* x, int B, const PatchTiles& t) {
  std::vector<uint8_t> img(t.bytes(B), 0);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < t.H; ++h)
      for (int w = 0; w < t.W; ++w)
        t.for_each_tile(h, w, [&](int ty, int tx, int pr, int pc) {
          for (int c = 0; c < t.C; ++c) {
            int st, chunk; t.stage_of(c >> 3, &st, &chunk);
            __half hh, ll; split1(x[(((size_t)b * t.H + h) * t.W + w) * t.C + c], &hh, &ll);
            const size_t off = t.tile_base(b, ty, tx) + t.in_tile(pr, pc, st, chunk) + (size_t)(c & 7) * 2;
            memcpy(img.data() + off, &hh, 2); memcpy(img.data() + off + t.st_plane[st], &ll, 2);
          }
        });
  return img;
}
// PatchTiles image -> [B][H][W][C]; every copy of a pixel must agree (returns the number of disagreeing copies)
long long patch_decode(const uint8_t* img, int B, const PatchTiles& t, float* out) {
  long long bad = 0;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < t.H; ++h)
      for (int w = 0; w < t.W; ++w) {
        int copies = 0;
        t.for_each_tile(h, w, [&](int ty, int tx, int pr, int pc) {
          for (int c = 0; c < t.C; ++c) {
            int st, chunk; t.stage_of(c >> 3, &st, &chunk);
            __half hh, ll;
            const size_t off = t.tile_base(b, ty, tx) + t.in_tile(pr, pc, st, chunk) + (size_t)(c & 7) * 2;
            memcpy(&hh, img + off, 2); memcpy(&ll, img + off + t.st_plane[st], 2);
            const float v = __half2float(hh) + __half2float(ll);
            float& o = out[(((size_t)b * t.H + h) * t.W + w) * t.C + c];
            if (copies == 0) o = v; else if (o != v) ++bad;
          }
          ++copies;
        });
      }
  return bad;
}

template <class F>
int guarded_dbg(F&& f) {
  try { f(); return BNB_OK; }
  catch (const cuda_error& e) {
    // a trapped kernel leaves a sticky error: tear the context down so that the next test starts clean (test hook only)
    cudaGetLastError(); cudaDeviceReset(); tc_forget_devices();
    return capi_fail(BNB_ERR_CUDA, e.what());
  }
  catch (const std::exception& e) { return capi_fail(BNB_ERR_INTERNAL, e.what()); }
}

// ---- TMEM probe: tcgen05.st a known pattern, read it back with narrow loads at unaligned columns --------------------
__global__ void tmem_probe_kernel(int* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tb = slot + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c = 0; c < 64; ++c) tc::tmem_st1(tb + c, (uint32_t)(row * 1000 + c));
  tc::tmem_st_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  int bad18 = 0, bad17 = 0, bad10 = 0;
  { uint32_t r[18]; tc::tmem_ld_n<18>(tb + 5, r); tc::tmem_ld_wait(); for (int i = 0; i < 18; ++i) bad18 += r[i] != (uint32_t)(row * 1000 + 5 + i); }
  { uint32_t r[17]; tc::tmem_ld_n<17>(tb + 18, r); tc::tmem_ld_wait(); for (int i = 0; i < 17; ++i) bad17 += r[i] != (uint32_t)(row * 1000 + 18 + i); }
  { uint32_t r[10]; tc::tmem_ld_n<10>(tb + 3, r); tc::tmem_ld_wait(); for (int i = 0; i < 10; ++i) bad10 += r[i] != (uint32_t)(row * 1000 + 3 + i); }
  atomicAdd(out + 0, bad18); atomicAdd(out + 1, bad17); atomicAdd(out + 2, bad10); atomicAdd(out + 3, 1);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(64)); }
}

}  // namespace

extern "C" {

int bnb_debug_tmem_probe(int32_t* out4) {
  if (!out4) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "NULL out");
  return guarded_dbg([&] {
    DevBuf d(4 * sizeof(int));
    tmem_probe_kernel<<<1, 128>>>(d.as<int>());
    BNB_CUDA(cudaDeviceSynchronize());
    BNB_CUDA(cudaMemcpy(out4, d.p, 4 * sizeof(int), cudaMemcpyDeviceToHost));
  });
}

// x [B][H][W][Cin], w_exp [C][Cin], b_exp [C], w_dw [9][C], b_dw [C] -> d_out [B][Ho][Wo][C], se_sum [B][C] (sum over pixels)
// flags bit 0: forbid the 32/64-byte swizzle stage shapes.  info: {TH, TW, PH, PW, n_mma, k_stages, a_resident, a_slots, b_slots, smem}
int bnb_debug_mbconv2(const float* x, int B, int H, int W, int Cin, const float* w_exp, const float* b_exp, const float* w_dw,
                      const float* b_dw, int C, int stride, int flags, float* d_out, float* se_sum, int32_t* info10) {
  if (!x || !w_exp || !b_exp || !w_dw || !b_dw || !d_out) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "NULL pointer");
  return guarded_dbg([&] {
    int dev = 0; BNB_CUDA(cudaGetDevice(&dev)); tc_prepare_device(dev);
    const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
    const Mb2Plan P = mb2_plan(H, W, Ho, Wo, stride, Cin, C, !(flags & 1));
    if (info10) { const int v[10] = {P.TH, P.TW, P.PH, P.PW, P.n_mma, P.k_stages, P.a_resident, P.a_slots, P.b_slots, (int)P.smem_bytes}; memcpy(info10, v, sizeof(v)); }
    if (!P.ok) throw std::runtime_error("mbconv2: no plan for this layer shape");
    const PatchTiles pt = mb2_patch_layout(P, H, W);
    const std::vector<uint8_t> ximg = patch_encode(x, B, pt);
    std::vector<uint8_t> img;
    mb2_prepare_weights(P, w_exp, &img);
    std::vector<float> be((size_t)P.n_units * 128 + 64, 0.f);
    memcpy(be.data(), b_exp, (size_t)C * 4);
    const long long M = (long long)B * Ho * Wo;
    const RowTiles dt = RowTiles::make(C);
    const int tiles = P.tiles_h * P.tiles_w;
    DevBuf dx(ximg.size()), dimg(img.size()), dbe(be.size() * 4), dwd((size_t)9 * C * 4), dbd((size_t)C * 4), dd(dt.bytes(M)), dpart((size_t)B * tiles * C * 4);
    BNB_CUDA(cudaMemcpy(dx.p, ximg.data(), ximg.size(), cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dimg.p, img.data(), img.size(), cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dbe.p, be.data(), be.size() * 4, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dwd.p, w_dw, (size_t)9 * C * 4, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dbd.p, b_dw, (size_t)C * 4, cudaMemcpyHostToDevice));
    Mb2Launch L{};
    L.x_img = dx.as<uint8_t>(); L.Wimg = dimg.as<uint8_t>(); L.bias_e = dbe.as<float>();
    L.w_dw = dwd.as<float>(); L.bias_dw = dbd.as<float>(); L.d_img = dd.as<uint8_t>(); L.partial = dpart.as<float>();
    L.B = B; L.H = H; L.W = W; L.Ho = Ho; L.Wo = Wo;
    LaunchCounter lc;
    launch_mbconv2(P, L, nullptr, lc);
    BNB_CUDA(cudaDeviceSynchronize());
    std::vector<uint8_t> dh(dt.bytes(M));
    BNB_CUDA(cudaMemcpy(dh.data(), dd.p, dh.size(), cudaMemcpyDeviceToHost));
    rows_decode(dh.data(), M, C, d_out);
    if (se_sum) {
      std::vector<float> part((size_t)B * tiles * C);
      BNB_CUDA(cudaMemcpy(part.data(), dpart.p, part.size() * 4, cudaMemcpyDeviceToHost));
      for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
          float s = 0.f;
          for (int t = 0; t < tiles; ++t) s += part[((size_t)b * tiles + t) * C + c];
          se_sum[(size_t)b * C + c] = s;
        }
    }
  });
}

// A [M][K], W [N][K], bias [N], gate [M / rows_per_chunk][K] or NULL, residual [M][N] or NULL -> out [M][N].
// out_mode 0: fp32 epilogue; 1: plain hi/lo planes; 2: PatchTiles image of a consuming MBConv block with input map
// geom = {H, W, stride, C_exp} (M must be a multiple of H*W; the image is decoded on the host and every halo copy of a
// pixel must agree).  A residual needs geom too ({H, W} of the block, read as its stride-1 PatchTiles image).
int bnb_debug_pw2(const float* A, int M, int K, const float* W, const float* bias, int N, const float* gate, int rows_per_chunk,
                  const float* residual, int act, int out_mode, const int32_t* geom4, float* out, int32_t* info4) {
  if (!A || !W || !bias || !out) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "NULL pointer");
  if ((out_mode == 2 || residual) && !geom4) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "geometry required");
  return guarded_dbg([&] {
    int dev = 0; BNB_CUDA(cudaGetDevice(&dev)); tc_prepare_device(dev);
    std::vector<uint8_t> img;
    const PwTcLayer L = pw_tc_prepare(W, N, K, &img);
    const std::vector<uint8_t> aimg = rows_encode(A, M, K);
    std::vector<float> bz((size_t)L.n_pad + 64, 0.f);
    memcpy(bz.data(), bias, (size_t)N * 4);
    const int chunks = rows_per_chunk > 0 ? (M + rows_per_chunk - 1) / rows_per_chunk : 1;
    PatchTiles rp, op;
    int Bimg = 0;
    std::vector<uint8_t> rimg;
    if (residual) {       // the block input of a stride-1 block with N channels on an H x W map
      const Mb2Plan Pr = mb2_plan(geom4[0], geom4[1], geom4[0], geom4[1], 1, N, 128, true);
      if (!Pr.ok || M % (geom4[0] * geom4[1])) throw std::runtime_error("debug_pw2: bad residual geometry");
      rp = mb2_patch_layout(Pr, geom4[0], geom4[1]); Bimg = M / (geom4[0] * geom4[1]);
      rimg = patch_encode(residual, Bimg, rp);
    }
    if (out_mode == 2) {
      const int st = geom4[2], Ho = st == 1 ? geom4[0] : geom4[0] / 2, Wo = st == 1 ? geom4[1] : geom4[1] / 2;
      const Mb2Plan Po = mb2_plan(geom4[0], geom4[1], Ho, Wo, st, N, geom4[3], true);
      if (!Po.ok || M % (geom4[0] * geom4[1])) throw std::runtime_error("debug_pw2: bad output geometry");
      op = mb2_patch_layout(Po, geom4[0], geom4[1]); Bimg = M / (geom4[0] * geom4[1]);
    }
    const int o_pitch = (N + 7) / 8 * 8;
    const PatchTables Tr = residual ? patch_build_tables(rp) : PatchTables(), To = out_mode == 2 ? patch_build_tables(op) : PatchTables();
    DevBuf dtr(Tr.dst.size() * 4), dtrr(Tr.res.size() * 4), dto(To.dst.size() * 4), dtor(To.res.size() * 4);
    if (residual) {
      BNB_CUDA(cudaMemcpy(dtr.p, Tr.dst.data(), Tr.dst.size() * 4, cudaMemcpyHostToDevice)); BNB_CUDA(cudaMemcpy(dtrr.p, Tr.res.data(), Tr.res.size() * 4, cudaMemcpyHostToDevice));
      rp.dst_tbl = dtr.as<uint4>(); rp.res_tbl = dtrr.as<uint32_t>();
    }
    if (out_mode == 2) {
      BNB_CUDA(cudaMemcpy(dto.p, To.dst.data(), To.dst.size() * 4, cudaMemcpyHostToDevice)); BNB_CUDA(cudaMemcpy(dtor.p, To.res.data(), To.res.size() * 4, cudaMemcpyHostToDevice));
      op.dst_tbl = dto.as<uint4>(); op.res_tbl = dtor.as<uint32_t>();
    }
    DevBuf da(aimg.size()), dimg(img.size()), db(bz.size() * 4), dg(gate ? (size_t)chunks * K * 4 : 16), dr(rimg.size()),
        doh((size_t)M * o_pitch * 2), dol((size_t)M * o_pitch * 2), d32((size_t)M * N * 4), dop(out_mode == 2 ? op.bytes(Bimg) : 16);
    BNB_CUDA(cudaMemcpy(da.p, aimg.data(), aimg.size(), cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dimg.p, img.data(), img.size(), cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(db.p, bz.data(), bz.size() * 4, cudaMemcpyHostToDevice));
    if (gate) BNB_CUDA(cudaMemcpy(dg.p, gate, (size_t)chunks * K * 4, cudaMemcpyHostToDevice));
    if (residual) BNB_CUDA(cudaMemcpy(dr.p, rimg.data(), rimg.size(), cudaMemcpyHostToDevice));
    Pw2Launch p{};
    p.a_img = da.as<uint8_t>(); p.Wimg = dimg.as<uint8_t>(); p.bias = db.as<float>();
    p.gate = gate ? dg.as<float>() : nullptr;
    if (residual) { p.r_img = dr.as<uint8_t>(); p.r_patch = rp; }
    if (out_mode == 0) p.out32 = d32.as<float>();
    else if (out_mode == 1) { p.oh = doh.as<__half>(); p.ol = dol.as<__half>(); p.o_pitch = o_pitch; }
    else { p.o_img = dop.as<uint8_t>(); p.o_patch = op; }
    p.M = M; p.N = N; p.K = K; p.rows_per_chunk = rows_per_chunk; p.act = act;
    if (info4) { int bn, st, br; size_t sm; pw2_tiling(L, M, gate != nullptr, &bn, &st, &sm, &br); info4[0] = bn; info4[1] = st; info4[2] = br; info4[3] = (int)sm; }
    LaunchCounter lc;
    launch_pw2(L, p, nullptr, lc);
    BNB_CUDA(cudaDeviceSynchronize());
    if (out_mode == 1) {
      std::vector<__half> oh((size_t)M * o_pitch), ol((size_t)M * o_pitch);
      BNB_CUDA(cudaMemcpy(oh.data(), doh.p, oh.size() * 2, cudaMemcpyDeviceToHost));
      BNB_CUDA(cudaMemcpy(ol.data(), dol.p, ol.size() * 2, cudaMemcpyDeviceToHost));
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) out[(size_t)m * N + n] = __half2float(oh[(size_t)m * o_pitch + n]) + __half2float(ol[(size_t)m * o_pitch + n]);
    } else if (out_mode == 2) {
      std::vector<uint8_t> oi(op.bytes(Bimg));
      BNB_CUDA(cudaMemcpy(oi.data(), dop.p, oi.size(), cudaMemcpyDeviceToHost));
      const long long bad = patch_decode(oi.data(), Bimg, op, out);
      if (bad) throw std::runtime_error("debug_pw2: " + std::to_string(bad) + " halo copies of output pixels disagree");
    } else {
      BNB_CUDA(cudaMemcpy(out, d32.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    }
  });
}

}  // extern "C"
