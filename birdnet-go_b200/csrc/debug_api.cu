// debug_api.cu — test hooks that run ONE tensor-core kernel on host data (tests/test_gpu_kernels.py compares them with
// numpy restatements of the same layer).  Not part of the product path; declared in include/birdnet_b200.h under
// "introspection / test hooks".
#include <cuda_fp16.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/birdnet_b200.h"
#include "common.cuh"
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

void split_host(const float* x, size_t rows, int cols, int pitch, std::vector<__half>* hi, std::vector<__half>* lo) {
  hi->assign(rows * pitch, __float2half_rn(0.f)); lo->assign(rows * pitch, __float2half_rn(0.f));
  for (size_t r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float v = x[r * cols + c];
      const __half h = __float2half_rn(v);
      (*hi)[r * pitch + c] = h; (*lo)[r * pitch + c] = __float2half_rn(v - __half2float(h));
    }
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
    const int pitch = (Cin + 7) / 8 * 8;
    std::vector<__half> xh, xl;
    split_host(x, (size_t)B * H * W, Cin, pitch, &xh, &xl);
    std::vector<uint8_t> img;
    mb2_prepare_weights(P, w_exp, &img);
    std::vector<float> be((size_t)P.n_units * 128 + 64, 0.f);
    memcpy(be.data(), b_exp, (size_t)C * 4);
    const size_t n_out = (size_t)B * Ho * Wo * C;
    const int tiles = P.tiles_h * P.tiles_w;
    DevBuf dxh(xh.size() * 2), dxl(xl.size() * 2), dimg(img.size()), dbe(be.size() * 4), dwd((size_t)9 * C * 4), dbd((size_t)C * 4),
        ddh(n_out * 2), ddl(n_out * 2), dpart((size_t)B * tiles * C * 4);
    BNB_CUDA(cudaMemcpy(dxh.p, xh.data(), xh.size() * 2, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dxl.p, xl.data(), xl.size() * 2, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dimg.p, img.data(), img.size(), cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dbe.p, be.data(), be.size() * 4, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dwd.p, w_dw, (size_t)9 * C * 4, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dbd.p, b_dw, (size_t)C * 4, cudaMemcpyHostToDevice));
    Mb2Launch L{};
    L.xh = dxh.as<__half>(); L.xl = dxl.as<__half>(); L.x_pitch = pitch; L.Wimg = dimg.as<uint8_t>(); L.bias_e = dbe.as<float>();
    L.w_dw = dwd.as<float>(); L.bias_dw = dbd.as<float>(); L.dh = ddh.as<__half>(); L.dl = ddl.as<__half>(); L.partial = dpart.as<float>();
    L.B = B; L.H = H; L.W = W; L.Ho = Ho; L.Wo = Wo;
    LaunchCounter lc;
    launch_mbconv2(P, L, nullptr, lc);
    BNB_CUDA(cudaDeviceSynchronize());
    std::vector<__half> oh(n_out), ol(n_out);
    BNB_CUDA(cudaMemcpy(oh.data(), ddh.p, n_out * 2, cudaMemcpyDeviceToHost));
    BNB_CUDA(cudaMemcpy(ol.data(), ddl.p, n_out * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n_out; ++i) d_out[i] = __half2float(oh[i]) + __half2float(ol[i]);
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

// A [M][K], W [N][K], bias [N], gate [M / rows_per_chunk][K] or NULL, residual [M][N] or NULL -> out [M][N]
// planes_out = 1: the kernel writes hi/lo planes (joined on the host), else fp32
int bnb_debug_pw2(const float* A, int M, int K, const float* W, const float* bias, int N, const float* gate, int rows_per_chunk,
                  const float* residual, int act, int planes_out, float* out, int32_t* info4) {
  if (!A || !W || !bias || !out) return capi_fail(BNB_ERR_INVALID_ARGUMENT, "NULL pointer");
  return guarded_dbg([&] {
    int dev = 0; BNB_CUDA(cudaGetDevice(&dev)); tc_prepare_device(dev);
    std::vector<uint8_t> img;
    const PwTcLayer L = pw_tc_prepare(W, N, K, &img);
    const int a_pitch = (K + 7) / 8 * 8, o_pitch = (N + 7) / 8 * 8;
    std::vector<__half> ah, al, rh, rl;
    split_host(A, (size_t)M, K, a_pitch, &ah, &al);
    if (residual) split_host(residual, (size_t)M, N, o_pitch, &rh, &rl);
    std::vector<float> bz((size_t)L.n_pad + 64, 0.f);
    memcpy(bz.data(), bias, (size_t)N * 4);
    const int chunks = rows_per_chunk > 0 ? (M + rows_per_chunk - 1) / rows_per_chunk : 1;
    DevBuf dah(ah.size() * 2), dal(al.size() * 2), dimg(img.size()), db(bz.size() * 4), dg(gate ? (size_t)chunks * K * 4 : 16),
        drh(rh.size() * 2), drl(rl.size() * 2), doh((size_t)M * o_pitch * 2), dol((size_t)M * o_pitch * 2), d32((size_t)M * N * 4);
    BNB_CUDA(cudaMemcpy(dah.p, ah.data(), ah.size() * 2, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dal.p, al.data(), al.size() * 2, cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(dimg.p, img.data(), img.size(), cudaMemcpyHostToDevice));
    BNB_CUDA(cudaMemcpy(db.p, bz.data(), bz.size() * 4, cudaMemcpyHostToDevice));
    if (gate) BNB_CUDA(cudaMemcpy(dg.p, gate, (size_t)chunks * K * 4, cudaMemcpyHostToDevice));
    if (residual) { BNB_CUDA(cudaMemcpy(drh.p, rh.data(), rh.size() * 2, cudaMemcpyHostToDevice)); BNB_CUDA(cudaMemcpy(drl.p, rl.data(), rl.size() * 2, cudaMemcpyHostToDevice)); }
    Pw2Launch p{};
    p.ah = dah.as<__half>(); p.al = dal.as<__half>(); p.a_pitch = a_pitch; p.Wimg = dimg.as<uint8_t>(); p.bias = db.as<float>();
    p.gate = gate ? dg.as<float>() : nullptr;
    if (residual) { p.rh = drh.as<__half>(); p.rl = drl.as<__half>(); p.r_pitch = o_pitch; }
    if (planes_out) { p.oh = doh.as<__half>(); p.ol = dol.as<__half>(); p.o_pitch = o_pitch; } else p.out32 = d32.as<float>();
    p.M = M; p.N = N; p.K = K; p.rows_per_chunk = rows_per_chunk; p.act = act;
    if (info4) { int bn, st, br; size_t sm; pw2_tiling(L, M, gate != nullptr, &bn, &st, &sm, &br); info4[0] = bn; info4[1] = st; info4[2] = br; info4[3] = (int)sm; }
    LaunchCounter lc;
    launch_pw2(L, p, nullptr, lc);
    BNB_CUDA(cudaDeviceSynchronize());
    if (planes_out) {
      std::vector<__half> oh((size_t)M * o_pitch), ol((size_t)M * o_pitch);
      BNB_CUDA(cudaMemcpy(oh.data(), doh.p, oh.size() * 2, cudaMemcpyDeviceToHost));
      BNB_CUDA(cudaMemcpy(ol.data(), dol.p, ol.size() * 2, cudaMemcpyDeviceToHost));
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) out[(size_t)m * N + n] = __half2float(oh[(size_t)m * o_pitch + n]) + __half2float(ol[(size_t)m * o_pitch + n]);
    } else {
      BNB_CUDA(cudaMemcpy(out, d32.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    }
  });
}

}  // extern "C"
