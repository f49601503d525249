//go:build b200

package b200

/*
#cgo LDFLAGS: -lbirdnet_b200
#include <stdint.h>
#include <stdlib.h>
#include "birdnet_b200.h"
*/
import "C"

import (
	"runtime"
	"unsafe"

	"github.com/tphakala/birdnet-go/internal/errors"
)

// Supported reports whether the B200 backend is compiled into this build.
const Supported = true

// Classifier implements inference.Classifier and inference.EmbeddingExtractor
// (internal/inference/backend.go:8-29) on libbirdnet_b200.so, plus the batched surface the
// offline driver and the realtime coalescer use.  NOT goroutine-safe: callers serialize
// (backend.go:7; BirdNET.mu holds across the native call and Close, birdnet.go:111-119).
type Classifier struct {
	h          *C.bnb_classifier
	numSpecies int
	numSamples int
	embDim     int
	maxBatch   int
	device     string
	precision  string
}

// lastErr must run on the OS thread that made the failing call: bnb_last_error is thread-local
// (the rule of backend_openvino.go:478-480).
func lastErr(op string, rc C.int) error {
	return errors.Newf("b200: %s failed: status=%d: %s", op, int(rc), C.GoString(C.bnb_last_error())).
		Component("inference.b200").Category(errors.CategoryModelInit).Build()
}

// Init verifies the library loads and a compute-capability-10.x device is present.
// Idempotent and retryable (InitOV / InitONNXRuntime semantics).
func Init() error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := C.bnb_init(); rc != C.BNB_OK {
		return ErrB200Unavailable
	}
	return nil
}

// DeviceCount returns the number of usable B200-class devices.
func DeviceCount() (int, error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	n := C.bnb_device_count()
	if n < 0 {
		return 0, ErrB200Unavailable
	}
	return int(n), nil
}

// NewClassifier mirrors tflite.NewTFLiteClassifier(modelData, opts) (tflite/classifier.go:38-92):
// modelData is the embedded BirdNET v2.4 .tflite (models_embedded.go:14-15); the bytes are parsed
// and uploaded during the call.  Any failure is reported at construction so the caller can fall
// back to TFLite; ErrB200Unavailable means "no device / library".
func NewClassifier(modelData []byte, o Options) (*Classifier, error) {
	if len(modelData) == 0 {
		return nil, errors.Newf("cannot create model from data (0 bytes)").Component("inference.b200").Category(errors.CategoryModelInit).Build()
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := C.bnb_init(); rc != C.BNB_OK {
		return nil, ErrB200Unavailable
	}
	var opts C.bnb_options
	opts.struct_size = C.uint32_t(unsafe.Sizeof(opts))
	opts.device = C.int32_t(o.Device)
	opts.max_batch = C.int32_t(o.MaxBatch)
	opts.micro_batch = C.int32_t(o.MicroBatch)
	opts.precision = C.int32_t(o.Precision)
	opts.lanes = C.int32_t(o.Lanes)
	if o.UseGraphs {
		opts.use_graphs = 1
	}
	var h *C.bnb_classifier
	// modelData is only read during the call (the library copies what it keeps): allowed cgo pattern.
	rc := C.bnb_classifier_create(unsafe.Pointer(&modelData[0]), C.size_t(len(modelData)), &opts, &h)
	if rc == C.BNB_ERR_NO_DEVICE {
		return nil, ErrB200Unavailable
	}
	if rc != C.BNB_OK {
		return nil, lastErr("classifier_create", rc)
	}
	return &Classifier{
		h:          h,
		numSpecies: int(C.bnb_num_species(h)),
		numSamples: int(C.bnb_num_samples(h)),
		embDim:     int(C.bnb_embedding_dim(h)),
		maxBatch:   int(C.bnb_max_batch(h)),
		device:     C.GoString(C.bnb_runtime_device(h)),
		precision:  C.GoString(C.bnb_runtime_precision(h)),
	}, nil
}

func sizeMismatch(want, got int) error {
	// same message as tflite/classifier.go:102-104
	return errors.Newf("input size mismatch: expected %d samples, got %d", want, got).
		Component("inference.b200").Category(errors.CategoryValidation).Build()
}

// Predict implements inference.Classifier: exactly NumSamples float32 samples in, a freshly
// allocated slice of NumSpecies raw logits (pre-activation, label order) out.
func (c *Classifier) Predict(samples []float32) ([]float32, error) {
	if c == nil || c.h == nil {
		return nil, ErrClosed
	}
	if len(samples) != c.numSamples {
		return nil, sizeMismatch(c.numSamples, len(samples))
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	out := make([]float32, c.numSpecies)
	rc := C.bnb_predict(c.h, (*C.float)(unsafe.Pointer(&samples[0])), C.size_t(len(samples)), (*C.float)(unsafe.Pointer(&out[0])))
	if rc != C.BNB_OK {
		return nil, lastErr("predict", rc)
	}
	return out, nil
}

// PredictWithEmbeddings implements inference.EmbeddingExtractor: also returns the 1024-d
// GLOBAL_AVG_POOL embedding (the bat pipeline's input, bat_onnx.go:252).
func (c *Classifier) PredictWithEmbeddings(samples []float32) (logits, embeddings []float32, err error) {
	if c == nil || c.h == nil {
		return nil, nil, ErrClosed
	}
	if len(samples) != c.numSamples {
		return nil, nil, sizeMismatch(c.numSamples, len(samples))
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	logits = make([]float32, c.numSpecies)
	embeddings = make([]float32, c.embDim)
	rc := C.bnb_predict_with_embeddings(c.h, (*C.float)(unsafe.Pointer(&samples[0])), C.size_t(len(samples)),
		(*C.float)(unsafe.Pointer(&logits[0])), (*C.float)(unsafe.Pointer(&embeddings[0])))
	if rc != C.BNB_OK {
		return nil, nil, lastErr("predict_with_embeddings", rc)
	}
	return logits, embeddings, nil
}

// PredictBatch runs batchSize windows (flat [batchSize*NumSamples] float32) and returns flat
// [batchSize*NumSpecies] raw logits (the shape of onnx.Classifier.PredictBatch, onnx/classifier.go:372-430).
func (c *Classifier) PredictBatch(samples []float32, batchSize int) ([]float32, error) {
	if c == nil || c.h == nil {
		return nil, ErrClosed
	}
	if batchSize <= 0 || batchSize > c.maxBatch || len(samples) != batchSize*c.numSamples {
		return nil, sizeMismatch(batchSize*c.numSamples, len(samples))
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	out := make([]float32, batchSize*c.numSpecies)
	rc := C.bnb_predict_batch(c.h, unsafe.Pointer(&samples[0]), C.BNB_PCM_F32, C.int(batchSize), (*C.float)(unsafe.Pointer(&out[0])), nil)
	if rc != C.BNB_OK {
		return nil, lastErr("predict_batch", rc)
	}
	return out, nil
}

func (c *Classifier) analyze(pcm unsafe.Pointer, format C.int, batchSize int, sensitivity float32, k int) ([]TopK, error) {
	if k <= 0 || k > 64 {
		return nil, errors.Newf("top-k must be in 1..64, got %d", k).Component("inference.b200").Category(errors.CategoryValidation).Build()
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	idx := make([]int32, batchSize*k)
	conf := make([]float32, batchSize*k)
	rc := C.bnb_analyze_batch(c.h, pcm, format, C.int(batchSize), C.float(sensitivity), C.int(k),
		(*C.int32_t)(unsafe.Pointer(&idx[0])), (*C.float)(unsafe.Pointer(&conf[0])), nil)
	if rc != C.BNB_OK {
		return nil, lastErr("analyze_batch", rc)
	}
	rows := make([]TopK, batchSize)
	for i := range rows {
		rows[i] = TopK{Index: idx[i*k : (i+1)*k : (i+1)*k], Confidence: conf[i*k : (i+1)*k : (i+1)*k]}
	}
	return rows, nil
}

// AnalyzeBatch is BirdNET.Predict's post-processing done on the device for a whole batch
// (analyze.go:82-99): confidence = sigmoid(sensitivity*logit), top-k by descending confidence.
func (c *Classifier) AnalyzeBatch(samples []float32, batchSize int, sensitivity float32, k int) ([]TopK, error) {
	if c == nil || c.h == nil {
		return nil, ErrClosed
	}
	if batchSize <= 0 || batchSize > c.maxBatch || len(samples) != batchSize*c.numSamples {
		return nil, sizeMismatch(batchSize*c.numSamples, len(samples))
	}
	return c.analyze(unsafe.Pointer(&samples[0]), C.BNB_PCM_F32, batchSize, sensitivity, k)
}

// AnalyzeBatchInt16 is AnalyzeBatch fed with the int16 PCM the AnalysisBuffer holds
// (process.go:479-497's conversion runs on the device): half the host-to-device bytes.
func (c *Classifier) AnalyzeBatchInt16(pcm []int16, batchSize int, sensitivity float32, k int) ([]TopK, error) {
	if c == nil || c.h == nil {
		return nil, ErrClosed
	}
	if batchSize <= 0 || batchSize > c.maxBatch || len(pcm) != batchSize*c.numSamples {
		return nil, sizeMismatch(batchSize*c.numSamples, len(pcm))
	}
	return c.analyze(unsafe.Pointer(&pcm[0]), C.BNB_PCM_S16, batchSize, sensitivity, k)
}

// DetectBatchInt16 runs sigmoid(sensitivity*logit), the top-k, the confidence threshold and the
// compaction of the survivors on the device (SURVEY 8(f) N1); only the detections cross PCIe.
// The result holds one Detections per chunk (possibly empty), in descending confidence —
// what processor.go:820-876 keeps of BirdNET.Predict's ten Results.
func (c *Classifier) DetectBatchInt16(pcm []int16, batchSize int, sensitivity, threshold float32, k int) ([]Detections, error) {
	if c == nil || c.h == nil {
		return nil, ErrClosed
	}
	if batchSize <= 0 || batchSize > c.maxBatch || len(pcm) != batchSize*c.numSamples {
		return nil, sizeMismatch(batchSize*c.numSamples, len(pcm))
	}
	if k <= 0 {
		return nil, errors.Newf("k must be positive, got %d", k).Component("inference.b200").Category(errors.CategoryValidation).Build()
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	capacity := batchSize * k
	chunk := make([]int32, capacity)
	idx := make([]int32, capacity)
	conf := make([]float32, capacity)
	counts := make([]int32, batchSize)
	var found C.int32_t
	rc := C.bnb_analyze_batch_detections(c.h, unsafe.Pointer(&pcm[0]), C.BNB_PCM_S16, C.int(batchSize), C.float(sensitivity), C.float(threshold),
		C.int(k), C.int(capacity), (*C.int32_t)(unsafe.Pointer(&chunk[0])), (*C.int32_t)(unsafe.Pointer(&idx[0])),
		(*C.float)(unsafe.Pointer(&conf[0])), (*C.int32_t)(unsafe.Pointer(&counts[0])), &found)
	if rc != C.BNB_OK {
		return nil, lastErr("analyze_batch_detections", rc)
	}
	out := make([]Detections, batchSize)
	pos := 0
	for b := range out {
		n := int(counts[b])
		out[b] = Detections{Index: idx[pos : pos+n : pos+n], Confidence: conf[pos : pos+n : pos+n]}
		pos += n
	}
	return out, nil
}

// UltrasonicFrameCV is ultrasonic.ComputeUSFrameCV (internal/audiocore/ultrasonic/filter.go:20-66)
// for a batch of int16 chunks at the SOURCE rate, computed in float64 on the device.
// ok[b] is false exactly where the reference returns (0, false).
func UltrasonicFrameCV(device int, pcm []int16, batchSize, sampleRate, fftSize, hopSize, frequencySplitHz int) (cv []float64, ok []bool, err error) {
	if batchSize <= 0 || len(pcm)%batchSize != 0 {
		return nil, nil, errors.Newf("pcm length %d is not a multiple of the batch size %d", len(pcm), batchSize).
			Component("inference.b200").Category(errors.CategoryValidation).Build()
	}
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	cv = make([]float64, batchSize)
	flags := make([]int32, batchSize)
	rc := C.bnb_ultrasonic_cv_batch(C.int(device), unsafe.Pointer(&pcm[0]), C.BNB_PCM_S16, C.int(batchSize), C.int(len(pcm)/batchSize),
		C.int(sampleRate), C.int(fftSize), C.int(hopSize), C.int(frequencySplitHz), (*C.double)(unsafe.Pointer(&cv[0])),
		(*C.int32_t)(unsafe.Pointer(&flags[0])))
	if rc != C.BNB_OK {
		return nil, nil, lastErr("ultrasonic_cv_batch", rc)
	}
	ok = make([]bool, batchSize)
	for i, f := range flags {
		ok[i] = f != 0
	}
	return cv, ok, nil
}

// NumSpecies implements inference.Classifier (6522, read from the model).
func (c *Classifier) NumSpecies() int { return c.numSpecies }

// NumSamples is the window length the model expects (144000).
func (c *Classifier) NumSamples() int { return c.numSamples }

// EmbeddingDim is the embedding vector length (1024).
func (c *Classifier) EmbeddingDim() int { return c.embDim }

// MaxBatch is the largest batch the batch entry points accept.
func (c *Classifier) MaxBatch() int { return c.maxBatch }

// Device and Precision feed bn.setRuntimeInfo(device, backend, precision) (birdnet.go:1708).
func (c *Classifier) Device() string { return c.device }

// Precision reports the arithmetic of the dense layers, e.g. "FP16x3(tcgen05)+FP32".
func (c *Classifier) Precision() string { return c.precision }

// LastDeviceMillis is the device time of the most recent host-buffer call (RecordModelInvoke, analyze.go:73-79).
func (c *Classifier) LastDeviceMillis() float32 {
	if c == nil || c.h == nil {
		return -1
	}
	return float32(C.bnb_last_device_ms(c.h))
}

// Close implements inference.Classifier: frees device weights, workspaces, pinned staging and
// streams immediately; idempotent (tflite/classifier.go:129-134).
func (c *Classifier) Close() {
	if c != nil && c.h != nil {
		C.bnb_classifier_destroy(c.h)
		c.h = nil
	}
}
