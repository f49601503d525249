// Package classifier — additive file for github.com/tphakala/birdnet-go/internal/classifier.
//
// PredictModelBatch is the batched form of Orchestrator.PredictModel (orchestrator.go:507-572): the same three-level
// locking protocol (read lock on the models map, then inferenceMu, then the entry's own mutex) and the same
// globalInferenceCounters bookkeeping, taken ONCE for a whole batch of analysis windows instead of once per window.
// With the reference's per-window PredictModel every source of a multi-stream deployment queues on inferenceMu
// (SURVEY §8(a) a6); a realtime coalescer or the offline file driver hands the windows that are ready to this method
// and gets one Results slice per window back.
//
// Models whose backend does not implement BatchPredictor are served by a loop over Predict under the same locks, so
// the method is safe to call for any registered model.
//
// Not compiled in this repository (no Go toolchain in the build image); kept in step with the reference's types by
// review. The Python mirror with the same semantics is birdnet-go_b200/birdnet_b200/orchestrator.py
// (tests/test_orchestrator.py).
package classifier

import (
	"context"
	"time"

	"github.com/tphakala/birdnet-go/internal/datastore"
	"github.com/tphakala/birdnet-go/internal/errors"
	"github.com/tphakala/birdnet-go/internal/logger"
)

// BatchPredictor is implemented by model instances that can analyse several windows in one native call
// (the B200 backend: b200.Classifier.AnalyzeBatchInt16 / DetectBatchInt16 behind BirdNET.PredictBatchInt16).
type BatchPredictor interface {
	// PredictBatchInt16 takes batchSize consecutive windows of int16 PCM (the bytes the AnalysisBuffer hands to
	// ProcessData, process.go:253-291) and returns, per window, what Predict returns for that window.
	PredictBatchInt16(ctx context.Context, pcm []int16, batchSize int) ([][]datastore.Results, error)
}

// PredictModelBatch runs inference for batchSize windows on the model identified by modelID.
func (o *Orchestrator) PredictModelBatch(ctx context.Context, modelID string, pcm []int16, batchSize int) ([][]datastore.Results, error) {
	log := GetLogger()

	o.mu.RLock()
	entry, ok := o.models[modelID]
	o.mu.RUnlock()

	if !ok {
		return nil, errors.Newf("unknown model: %s", modelID).
			Component("classifier.orchestrator").
			Category(errors.CategoryValidation).
			Context("model_id", modelID).
			Build()
	}
	if batchSize <= 0 || len(pcm)%batchSize != 0 {
		return nil, errors.Newf("batch of %d windows does not divide %d samples", batchSize, len(pcm)).
			Component("classifier.orchestrator").
			Category(errors.CategoryValidation).
			Context("model_id", modelID).
			Build()
	}

	o.inferenceMu.Lock()
	defer o.inferenceMu.Unlock()

	entry.mu.Lock()
	defer entry.mu.Unlock()
	if entry.instance == nil {
		return nil, errors.Newf("model %s has been closed", modelID).
			Component("classifier.orchestrator").
			Category(errors.CategoryValidation).
			Context("model_id", modelID).
			Build()
	}

	start := time.Now()
	var (
		results [][]datastore.Results
		err     error
	)
	if bp, isBatch := entry.instance.(BatchPredictor); isBatch {
		results, err = bp.PredictBatchInt16(ctx, pcm, batchSize)
	} else {
		// any other backend: the reference's per-window path, still under ONE acquisition of the locks
		window := len(pcm) / batchSize
		results = make([][]datastore.Results, 0, batchSize)
		buf := make([]float32, window)
		for b := 0; b < batchSize && err == nil; b++ {
			for i, v := range pcm[b*window : (b+1)*window] {
				buf[i] = float32(v) / 32768.0 // process.go:491-494
			}
			var r []datastore.Results
			r, err = entry.instance.Predict(ctx, [][]float32{buf})
			results = append(results, r)
		}
	}
	duration := time.Since(start)

	if err != nil {
		globalInferenceCounters.RecordError(modelID)
		log.Error("PredictModelBatch inference failed",
			logger.String("model_id", modelID),
			logger.Int("windows", batchSize),
			logger.Error(err),
			logger.Duration("duration", duration))
		return nil, err
	}
	globalInferenceCounters.RecordInvoke(modelID, duration.Microseconds())
	log.Debug("PredictModelBatch complete",
		logger.String("model_id", modelID),
		logger.Int("windows", batchSize),
		logger.Duration("duration", duration))
	return results, nil
}
