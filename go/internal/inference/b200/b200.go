// Package b200 is the birdnet-go inference backend for NVIDIA B200 (sm_100a) GPUs: a thin cgo
// binding of libbirdnet_b200.so (include/birdnet_b200.h).  The real implementation is compiled
// only under the "b200" build tag; other builds get stubs that report ErrB200Unavailable so the
// classifier falls back to TFLite — the same arrangement as internal/inference/openvino
// (openvino.go:27-31, stub_noopenvino.go).
//
// Drop this directory into birdnet-go as internal/inference/b200 and apply
// go/patches/0001-classifier-select-b200-backend.patch.
package b200

import "github.com/tphakala/birdnet-go/internal/errors"

// ErrB200Unavailable is returned when the backend is not compiled in (no "b200" build tag), the
// shared library cannot initialise, or no compute-capability-10.x device is present.  Callers
// treat it as "fall back to TFLite" (errors.Is-comparable sentinel).
var ErrB200Unavailable = errors.NewStd("b200: backend unavailable")

// ErrClosed is returned by every method of a Classifier after Close.
var ErrClosed = errors.NewStd("b200: classifier is closed")

// Precision selects the arithmetic of the dense 1x1 convolutions (bnb_precision).
type Precision int

const (
	// PrecisionDefault is PrecisionF16x3.
	PrecisionDefault Precision = 0
	// PrecisionF32 runs every layer as fp32 FMA on CUDA cores (truth path).
	PrecisionF32 Precision = 1
	// PrecisionF16x3 runs the dense layers on tcgen05 tensor cores as a 3-term fp16 hi/lo split
	// with fp32 accumulation (sigmoid outputs within 1e-3 of the TFLite path).
	PrecisionF16x3 Precision = 2
)

// PCMFormat is the sample format of the batch entry points (bnb_pcm_format).
type PCMFormat int

const (
	// PCMFloat32 is float32 in [-1, 1): what Predict receives (process.go:479-497 output).
	PCMFloat32 PCMFormat = 0
	// PCMInt16 is int16 little endian as it sits in the AnalysisBuffer; /32768 happens on the device.
	PCMInt16 PCMFormat = 1
)

// Options configures a Classifier (shaped like tflite.TFLiteClassifierOptions /
// openvino.Options).  The zero value selects the library defaults.
type Options struct {
	// Device is the CUDA device ordinal; -1 selects the calling thread's current device.
	Device int
	// MaxBatch is the largest batch PredictBatch / AnalyzeBatch accept (0 -> 256).
	MaxBatch int
	// MicroBatch is the number of chunks per kernel-chain launch (0 -> library default).
	MicroBatch int
	// Lanes is the number of micro-batches in flight (0 -> library default).
	Lanes int
	// Precision selects the arithmetic of the dense layers.
	Precision Precision
	// UseGraphs replays small batches from CUDA graphs.
	UseGraphs bool
}

// TopK is one row of AnalyzeBatch: label indices and confidences in descending confidence.
type TopK struct {
	Index      []int32
	Confidence []float32
}

// Detections is the thresholded form of one chunk: every label whose confidence is >= the
// threshold, in descending confidence (at most the k passed to DetectBatchInt16).
type Detections struct {
	Index      []int32
	Confidence []float32
}
