/*
 * birdnet_b200.h — C ABI of libbirdnet_b200.so: the B200-native BirdNET v2.4 inference path.
 *
 * This is the drop-in boundary for birdnet-go's `inference.Classifier`
 * (/root/reference/internal/inference/backend.go:8-19).  A Go backend `internal/inference/b200`
 * binds these symbols through cgo exactly the way `internal/inference/openvino` binds
 * libopenvino_c (/root/reference/internal/inference/openvino/backend_openvino.go:16-413);
 * INTEGRATION.md shows that stub.  Plain pointers and sizes only — no torch / C++ types.
 *
 * Conventions (mirroring the reference's native-backend rules):
 *   - every function returns a bnb_status (0 = OK, negative = error) unless documented otherwise;
 *     the message for the calling thread's last failure is `bnb_last_error()` (thread-local, as
 *     `ovbind_get_last_err_msg`, backend_openvino.go:462-470);
 *   - the library never aborts or exits; configuration errors surface at create time so the
 *     caller can fall back to TFLite (birdnet.go:321-335, `ErrOpenVINOUnavailable` precedent);
 *   - a classifier handle is NOT thread-safe: the caller serializes Predict/Close
 *     (backend.go:7; birdnet.go:111-119 holds bn.mu across the native call and Close);
 *   - input pointers are only read during the call (callee copies — process.go:280-291 recycles
 *     the slice immediately); outputs are written into caller-provided buffers.
 */
#ifndef BIRDNET_B200_H_
#define BIRDNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNB_ABI_VERSION 1

#if defined(__GNUC__)
#define BNB_API __attribute__((visibility("default")))
#else
#define BNB_API
#endif

typedef enum bnb_status {
  BNB_OK = 0,
  BNB_ERR_INVALID_ARGUMENT = -1,  /* null pointer, bad size: tflite/classifier.go:102-104 "input size mismatch" */
  BNB_ERR_NO_DEVICE = -2,         /* no usable sm_100 GPU -> caller falls back (ErrB200Unavailable)          */
  BNB_ERR_UNSUPPORTED_MODEL = -3, /* flatbuffer is not a BirdNET-v2.4-shaped graph                           */
  BNB_ERR_CUDA = -4,              /* CUDA runtime/driver failure; message carries the cudaError string        */
  BNB_ERR_CLOSED = -5,            /* handle already destroyed (onnx.go:89-91 ErrSessionClosed analogue)       */
  BNB_ERR_OUT_OF_MEMORY = -6,
  BNB_ERR_INTERNAL = -7
} bnb_status;

typedef struct bnb_classifier bnb_classifier; /* opaque; owns device weights, workspaces, pinned staging, streams */

/* Arithmetic used for the dense 1x1 ("pointwise") convolutions. */
typedef enum bnb_precision {
  BNB_PRECISION_DEFAULT = 0, /* = BNB_PRECISION_F16X3 */
  BNB_PRECISION_F32 = 1,     /* CUDA-core fp32 FMA everywhere (truth path)                                      */
  BNB_PRECISION_F16X3 = 2    /* tcgen05 kind::f16, 3-term hi/lo split of both operands, fp32 accumulate in TMEM */
} bnb_precision;

/* Options, shaped like TFLiteClassifierOptions (tflite/classifier.go:17-27) / openvino Options
 * (openvino/openvino.go:49-62).  Zero-initialise for defaults. */
typedef struct bnb_options {
  uint32_t struct_size; /* = sizeof(bnb_options); lets the ABI grow                                  */
  int32_t device;       /* CUDA device ordinal; -1 = current device                                  */
  int32_t max_batch;    /* largest B accepted by the batch entry points; 0 -> 256                    */
  int32_t micro_batch;  /* chunks per kernel-chain launch (L2-resident tiling); 0 -> library default */
  int32_t precision;    /* bnb_precision                                                             */
  int32_t use_graphs;   /* 1 = replay the kernel chain from CUDA graphs; 0 = plain launches          */
  int32_t lanes;        /* concurrent front-phase streams (micro-batches in flight); 0 -> library default (2) */
  int32_t reserved[8];
} bnb_options;

/* Input sample formats for the batch entry points. */
typedef enum bnb_pcm_format {
  BNB_PCM_F32 = 0, /* float32 in [-1,1): what Predict receives (process.go:479-497 output)          */
  BNB_PCM_S16 = 1  /* int16 LE as it sits in the AnalysisBuffer; converted (/32768) on the device   */
} bnb_pcm_format;

/* ---- process-global -------------------------------------------------------------------------- */

/* Idempotent, retryable, mutex-guarded (InitONNXRuntime / InitOV semantics: onnx.go:297-318,
 * backend_openvino.go:477-506).  Verifies a compute-capability-10.x device is present. */
BNB_API int bnb_init(void);
/* Number of usable devices (>= 0) or a negative bnb_status. */
BNB_API int bnb_device_count(void);
BNB_API int bnb_abi_version(void);
/* Thread-local message of the calling thread's most recent failing call ("" if none). */
BNB_API const char* bnb_last_error(void);

/* ---- classifier lifetime ---------------------------------------------------------------------- */

/* Replaces tflite.NewTFLiteClassifier(modelData, opts) (tflite/classifier.go:38-92): `tflite` are
 * the bytes of BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite (embedded in the Go binary,
 * models_embedded.go:14-15); they are parsed and uploaded during the call and may be freed after. */
BNB_API int bnb_classifier_create(const void* tflite, size_t tflite_len, const bnb_options* opts,
                          bnb_classifier** out);
/* Close(): frees everything immediately; NULL is a no-op (tflite/classifier.go:129-134). */
BNB_API void bnb_classifier_destroy(bnb_classifier* h);

BNB_API int bnb_num_species(const bnb_classifier* h);   /* NumSpecies(): 6522 */
BNB_API int bnb_num_samples(const bnb_classifier* h);   /* 144000              */
BNB_API int bnb_embedding_dim(const bnb_classifier* h); /* 1024                */
BNB_API int bnb_max_batch(const bnb_classifier* h);
/* RuntimeInfo strings for setRuntimeInfo(device, backend, precision) (birdnet.go:1708). */
BNB_API const char* bnb_runtime_device(const bnb_classifier* h);
BNB_API const char* bnb_runtime_precision(const bnb_classifier* h);

/* ---- host-buffer entry points (what the Go backend calls) ------------------------------------ */

/* Classifier.Predict (backend.go:12; tflite/classifier.go:95-119): `n_samples` must equal
 * bnb_num_samples; writes bnb_num_species raw logits (pre-activation, label order). */
BNB_API int bnb_predict(bnb_classifier* h, const float* samples, size_t n_samples, float* logits);
/* EmbeddingExtractor.PredictWithEmbeddings (backend.go:23-29): also writes the 1024-d embedding. */
BNB_API int bnb_predict_with_embeddings(bnb_classifier* h, const float* samples, size_t n_samples,
                                float* logits, float* embeddings);
/* Batched form (the only batch API in the reference is onnx PredictBatch, onnx/classifier.go:372-430):
 * `pcm` holds B*144000 samples in `format`; logits [B,6522]; embeddings [B,1024] or NULL. */
BNB_API int bnb_predict_batch(bnb_classifier* h, const void* pcm, int format, int B, float* logits,
                      float* embeddings);
/* BirdNET.Predict post-processing on the device (analyze.go:82-99): conf = sigmoid(sensitivity*x),
 * top-k by descending confidence.  idx/conf are [B,k].  `logits_or_null` optionally receives the
 * raw logits too.  Replaces the 26 KB/chunk D2H by 8*k bytes. */
BNB_API int bnb_analyze_batch(bnb_classifier* h, const void* pcm, int format, int B, float sensitivity,
                      int k, int32_t* idx, float* conf, float* logits_or_null);

/* Asynchronous form of bnb_analyze_batch for ONE caller that wants the host-to-device copy of batch i+1 to overlap the
 * kernels of batch i (the handle stays single-caller: submit and wait come from the same serialized caller).
 * `submit` returns once the work is enqueued and writes a ticket; idx / conf (/ logits) are valid after bnb_wait(ticket).
 * At most TWO tickets may be outstanding (a third submit fails with BNB_ERR_INVALID_ARGUMENT).  Pageable `pcm` is copied
 * during the call (callee-copies, process.go:280-291); PINNED host memory is read by the copy engine asynchronously and
 * must stay unchanged until bnb_wait returns.  Batches of at most micro_batch chunks replay from a CUDA graph when
 * bnb_options.use_graphs is set. */
BNB_API int bnb_analyze_batch_submit(bnb_classifier* h, const void* pcm, int format, int B, float sensitivity, int k,
                                     int32_t* idx, float* conf, float* logits_or_null, int32_t* ticket);
BNB_API int bnb_wait(bnb_classifier* h, int32_t ticket);

/* sigma(sensitivity * x) >= threshold and the compaction of the per-chunk top-k run on the device; only the
 * detections cross PCIe.  The list is ordered by chunk, then by descending confidence (what the reference's loop over a chunk's
 * Results produces: internal/analysis/processor/processor.go:820-876 with analyze.go:82-99 before it).  det_* hold up to max_det
 * entries; *n_det is the number FOUND (> max_det means the list was truncated); counts_or_null[b] = detections of chunk b.
 * Same input contract as bnb_analyze_batch. */
BNB_API int bnb_analyze_batch_detections(bnb_classifier* h, const void* pcm, int format, int B, float sensitivity, float threshold, int k,
                                         int max_det, int32_t* det_chunk, int32_t* det_idx, float* det_conf, int32_t* counts_or_null,
                                         int32_t* n_det);

/* SURVEY 8(f) N2, bat pipeline.  bnb_ultrasonic_cv_batch = ultrasonic.ComputeUSFrameCV
 * (internal/audiocore/ultrasonic/filter.go:20-66) for B chunks of n_samples PCM samples at the SOURCE rate (int16 scaled by
 * 1/32768 like convert.BytesToFloat64PCM16, or float32): float64 STFT (fft_size <= 8192, power of two), energy above
 * frequency_split_hz per frame, cv[b] = std / mean of the frame powers; ok[b] = 0 (and cv[b] = 0) exactly where the reference
 * returns (0, false).  bnb_dense_head_batch = the custom classification head on embeddings
 * (internal/inference/onnx/custom_classifier.go:147-173): scores[b][c] = sigmoid(bias[c] + <embeddings[b], weights[c]>).
 * device < 0 selects device 0.  Both are stateless convenience entry points (they allocate and free their device buffers). */
BNB_API int bnb_ultrasonic_cv_batch(int device, const void* pcm, int format, int B, int n_samples, int sample_rate, int fft_size,
                                    int hop_size, int frequency_split_hz, double* cv, int32_t* ok);
BNB_API int bnb_dense_head_batch(int device, const float* embeddings, int B, int n_in, const float* weights, const float* bias,
                                 int n_out, float* scores);

/* ---- device-buffer entry points (batched offline driver, bench, multi-GPU harness) ----------- */

/* Same computation with inputs/outputs already resident in device memory of the handle's device;
 * enqueued on `stream` (a cudaStream_t, NULL = the handle's own stream) without synchronising. */
BNB_API int bnb_predict_batch_device(bnb_classifier* h, const void* d_pcm, int format, int B,
                             float* d_logits, float* d_embeddings, void* stream);
BNB_API int bnb_analyze_batch_device(bnb_classifier* h, const void* d_pcm, int format, int B,
                             float sensitivity, int k, int32_t* d_idx, float* d_conf,
                             float* d_logits_or_null, void* stream);

/* ---- range filter (the "meta" model: species occurrence by location and week) --------------------------------- */

typedef struct bnb_range_filter bnb_range_filter; /* opaque; owns device weights and a stream */

/* Replaces tflite.NewTFLiteRangeFilter(modelData, errorFunc) (tflite/rangefilter.go:22-58): `tflite` are the bytes of
 * BirdNET_GLOBAL_6K_V2.4_MData_Model_V2_FP16.tflite (embedded, models_embedded.go); `device` = CUDA ordinal, -1 = current.
 * A graph of another shape -> BNB_ERR_UNSUPPORTED_MODEL (the caller keeps its TFLite range filter). */
BNB_API int bnb_range_filter_create(const void* tflite, size_t tflite_len, int device, bnb_range_filter** out);
BNB_API void bnb_range_filter_destroy(bnb_range_filter* h);
BNB_API int bnb_range_filter_num_species(const bnb_range_filter* h); /* RangeFilter.NumSpecies() (backend.go:63) */
/* RangeFilter.Predict(latitude, longitude, week) (backend.go:58-61; tflite/rangefilter.go:64-94): scores[num_species]. */
BNB_API int bnb_range_filter_predict(bnb_range_filter* h, float latitude, float longitude, float week, float* scores);
/* BatchRangeFilter.PredictBatch(inputs, batchSize) (backend.go:70-76; used by the heat-map builder,
 * orchestrator.go:1846-1882): inputs = batch_size x [lat, lon, week]; scores = [batch_size, num_species] row-major. */
BNB_API int bnb_range_filter_predict_batch(bnb_range_filter* h, const float* inputs, int batch_size, float* scores);

/* ---- introspection / test hooks -------------------------------------------------------------- */

/* Number of kernels of THIS library launched by the handle since creation (bench `gpu_launches`). */
BNB_API int64_t bnb_kernel_launches(const bnb_classifier* h);
/* Device time of the most recent host-buffer call in ms (for RecordModelInvoke, analyze.go:73-79). */
BNB_API float bnb_last_device_ms(const bnb_classifier* h);
/* Per-category device timing: between begin and end every kernel launch of the handle is bracketed by
 * CUDA events on its stream.  `end` synchronises the device and writes, per category (order: minmax,
 * frontend, stem_mix, pw_expand, depthwise, se_gate, pw_project, post_conv, row_mean, fc_head, topk),
 * the summed milliseconds and the launch count; returns the number of categories (11). */
BNB_API int bnb_profile_begin(bnb_classifier* h);
BNB_API int bnb_profile_end(bnb_classifier* h, float* ms, int64_t* launches, int cap);
/* Per-launch device times (ms) and categories of the region closed by the last bnb_profile_end, in issue order;
 * returns the number of entries written (<= cap). */
BNB_API int bnb_profile_launches(bnb_classifier* h, float* ms, int32_t* cat, int cap);
/* Tiling the tensor-core GEMM would use for an [M,K]x[N,K]^T layer: N-tile width, pipeline stages, shared-memory
 * bytes (host logic only, no GPU needed; lets the CPU tests check every layer shape fits the 227 KB budget). */
BNB_API int bnb_debug_pw_tiling(int M, int N, int K, int* bn, int* stages, int64_t* smem_bytes);
/* Tile geometry of the fused expand+depthwise kernel for one block: out10 = {th, tw, ph, pw, tiles_h, tiles_w,
 * k_stages, box_c, a_slots, b_slots} (host logic only).  C = expanded channels and B = nominal launch size feed the tile
 * cost model (0 = unknown: fewest tiles); max_tiles caps the tiles per chunk (SE partial-sum slots, 0 = no cap). */
BNB_API int bnb_debug_mbconv_geometry(int H, int W, int Ho, int Wo, int stride, int Cin, int C, int B, int max_tiles, int* out10,
                                      int64_t* smem_bytes);
/* Plans of the F16X3 kernels (host logic only).  bnb_debug_mb2_plan: out12 = {ok, th, tw, ph, pw, n_mma (patch rows per tile in
 * the MMA), units of 128 expanded channels, k_stages, weights resident, weight slots, patch PAIR slots, bytes of a pair slot}.
 * bnb_debug_pw2_tiling: out4 = {N-tile width, A-ring stages, weights resident, n-tiles}. */
BNB_API int bnb_debug_mb2_plan(int H, int W, int Ho, int Wo, int stride, int Cin, int C, int* out12, int64_t* smem_bytes);
BNB_API int bnb_debug_pw2_tiling(int M, int N, int K, int gated, int* out4, int64_t* smem_bytes);
/* JSON description of the layer plan extracted from a .tflite (no GPU needed). Returns bytes
 * written (excluding NUL) or a negative status; `cap` too small -> BNB_ERR_INVALID_ARGUMENT. */
BNB_API int bnb_describe_model(const void* tflite, size_t tflite_len, char* json, size_t cap);
/* Copy an intermediate activation of the most recent batch call to the host.  `tensor` is the
 * tensor index in the .tflite subgraph (e.g. 265 = frontend output); layout NHWC float32;
 * returns the element count written (<= cap) or a negative status.  Test hook only. */
BNB_API int64_t bnb_debug_read_tensor(bnb_classifier* h, int tensor, float* out, size_t cap);
/* Keep every intermediate alive (no buffer reuse) so bnb_debug_read_tensor can see all of them. */
BNB_API int bnb_debug_keep_intermediates(bnb_classifier* h, int on);

/* Kernel-level test hooks (tests/test_gpu_kernels.py): run ONE tensor-core kernel of the F16X3 path on host data.
 * bnb_debug_tmem_probe: out4 = {mismatches of 18-, 17-, 10-column tcgen05.ld at unaligned columns, threads run}.
 * bnb_debug_mbconv2: fused 1x1 expand + SiLU + 3x3 depthwise + SiLU of one MBConv block; x [B,H,W,Cin], w_exp [C,Cin],
 *   w_dw [9,C] -> d_out [B,Ho,Wo,C] (+ se_sum [B,C] = per-channel sums over pixels, may be NULL); flags bit 0 forbids the
 *   32/64-byte swizzle stage shapes; info10 = {TH, TW, PH, PW, n_mma, k_stages, a_resident, a_slots, b_slots, smem}.
 * bnb_debug_pw2: out [M,N] = act(A' W^T + bias) (+ residual), A' = A (* gate[m / rows_per_chunk] when gate != NULL);
 *   out_mode 0 = fp32 epilogue, 1 = plain hi/lo planes, 2 = the pre-tiled patch image of a consuming MBConv block whose
 *   input map is geom4 = {H, W, stride, C_exp} (decoded on the host); a residual needs geom4 {H, W} too;
 *   info4 = {bn, stages, b_res, smem}. */
BNB_API int bnb_debug_tmem_probe(int32_t* out4);
BNB_API int bnb_debug_mbconv2(const float* x, int B, int H, int W, int Cin, const float* w_exp, const float* b_exp,
                              const float* w_dw, const float* b_dw, int C, int stride, int flags, float* d_out, float* se_sum,
                              int32_t* info10);
BNB_API int bnb_debug_pw2(const float* A, int M, int K, const float* W, const float* bias, int N, const float* gate,
                          int rows_per_chunk, const float* residual, int act, int out_mode, const int32_t* geom4, float* out,
                          int32_t* info4);

#ifdef __cplusplus
}
#endif
#endif /* BIRDNET_B200_H_ */
