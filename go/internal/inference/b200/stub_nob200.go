//go:build !b200

package b200

// This file is built when the "b200" tag is absent.  Every entry point reports
// ErrB200Unavailable so the classifier falls back to TFLite and no libbirdnet_b200 symbol is
// referenced (same arrangement as openvino/stub_noopenvino.go:1-27).

// Supported reports whether the B200 backend is compiled into this build.
const Supported = false

// Classifier is the stub form of the backend handle.
type Classifier struct{}

// Init always fails without the b200 build tag.
func Init() error { return ErrB200Unavailable }

// DeviceCount always fails without the b200 build tag.
func DeviceCount() (int, error) { return 0, ErrB200Unavailable }

// NewClassifier always fails without the b200 build tag.
func NewClassifier(_ []byte, _ Options) (*Classifier, error) { return nil, ErrB200Unavailable }

// Predict always fails without the b200 build tag.
func (c *Classifier) Predict(_ []float32) ([]float32, error) { return nil, ErrB200Unavailable }

// PredictWithEmbeddings always fails without the b200 build tag.
func (c *Classifier) PredictWithEmbeddings(_ []float32) (logits, embeddings []float32, err error) {
	return nil, nil, ErrB200Unavailable
}

// PredictBatch always fails without the b200 build tag.
func (c *Classifier) PredictBatch(_ []float32, _ int) ([]float32, error) {
	return nil, ErrB200Unavailable
}

// AnalyzeBatchInt16 always fails without the b200 build tag.
func (c *Classifier) AnalyzeBatchInt16(_ []int16, _ int, _ float32, _ int) ([]TopK, error) {
	return nil, ErrB200Unavailable
}

// AnalyzeBatch always fails without the b200 build tag.
func (c *Classifier) AnalyzeBatch(_ []float32, _ int, _ float32, _ int) ([]TopK, error) {
	return nil, ErrB200Unavailable
}

// DetectBatchInt16 always fails without the b200 build tag.
func (c *Classifier) DetectBatchInt16(_ []int16, _ int, _, _ float32, _ int) ([]Detections, error) {
	return nil, ErrB200Unavailable
}

// UltrasonicFrameCV always fails without the b200 build tag (callers keep using ultrasonic.ComputeUSFrameCV).
func UltrasonicFrameCV(_ int, _ []int16, _, _, _, _, _ int) ([]float64, []bool, error) {
	return nil, nil, ErrB200Unavailable
}

// NumSpecies reports 0 without the b200 build tag.
func (c *Classifier) NumSpecies() int { return 0 }

// NumSamples reports 0 without the b200 build tag.
func (c *Classifier) NumSamples() int { return 0 }

// EmbeddingDim reports 0 without the b200 build tag.
func (c *Classifier) EmbeddingDim() int { return 0 }

// MaxBatch reports 0 without the b200 build tag.
func (c *Classifier) MaxBatch() int { return 0 }

// Device reports "" without the b200 build tag.
func (c *Classifier) Device() string { return "" }

// Precision reports "" without the b200 build tag.
func (c *Classifier) Precision() string { return "" }

// LastDeviceMillis reports -1 without the b200 build tag.
func (c *Classifier) LastDeviceMillis() float32 { return -1 }

// Close is a no-op without the b200 build tag.
func (c *Classifier) Close() {}
