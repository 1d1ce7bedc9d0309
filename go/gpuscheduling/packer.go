package gpuscheduling

import (
	"fmt"
	"math"

	"inference.networking.x-k8s.io/llm-instance-gateway/pkg/ext-proc/backend"
)

// packedSnapshot is one refresh tick of PodMetricsProvider.AllPodMetrics() in the column layout
// include/lig.h describes.  The slice order of the provider IS the pod index: picks come back as
// indices into pods.
type packedSnapshot struct {
	epoch      uint64
	P, A       int
	kv         []float64
	q          []int32
	nActive    []uint16
	maxActive  []uint16
	bitmap     []uint32 // adapter-major, A x ceil(P/32)
	adapterIDs map[string]int32
	pods       []backend.Pod
}

// packSnapshot interns adapter names (order of first appearance) and narrows the Go-width
// metrics to the device record.  WaitingQueueSize must fit int32 (an out-of-range value is an
// error, never a silent wrap); MaxActiveModels saturates to [0, 65535], which cannot change
// len(ActiveModels) < MaxActiveModels because len(ActiveModels) <= 65534.
func packSnapshot(all []*backend.PodMetrics) (*packedSnapshot, error) {
	P := len(all)
	W := (P + 31) / 32
	s := &packedSnapshot{
		P: P, kv: make([]float64, P), q: make([]int32, P), nActive: make([]uint16, P),
		maxActive: make([]uint16, P), adapterIDs: map[string]int32{}, pods: make([]backend.Pod, P),
	}
	for i, pm := range all {
		s.pods[i] = pm.Pod
		s.kv[i] = pm.KVCacheUsagePercent
		if pm.WaitingQueueSize > math.MaxInt32 || pm.WaitingQueueSize < math.MinInt32 {
			return nil, fmt.Errorf("pod %v: WaitingQueueSize %d does not fit the device record", pm.Pod, pm.WaitingQueueSize)
		}
		s.q[i] = int32(pm.WaitingQueueSize)
		if len(pm.ActiveModels) > 65534 {
			return nil, fmt.Errorf("pod %v: %d active models exceed the device record", pm.Pod, len(pm.ActiveModels))
		}
		s.nActive[i] = uint16(len(pm.ActiveModels))
		switch m := pm.MaxActiveModels; {
		case m < 0:
			s.maxActive[i] = 0
		case m > 65535:
			s.maxActive[i] = 65535
		default:
			s.maxActive[i] = uint16(m)
		}
		for name := range pm.ActiveModels {
			if _, ok := s.adapterIDs[name]; !ok {
				s.adapterIDs[name] = int32(len(s.adapterIDs))
			}
		}
	}
	s.A = len(s.adapterIDs)
	s.bitmap = make([]uint32, s.A*W)
	for i, pm := range all {
		for name := range pm.ActiveModels {
			s.bitmap[int(s.adapterIDs[name])*W+i>>5] |= 1 << (uint(i) & 31)
		}
	}
	return s, nil
}

// adapterID resolves ResolvedTargetModel; a model in no pod's ActiveModels gets id A, which
// matches no bitmap row — the same outcome as the Go map miss in loRAAffinityPredicate.
func (s *packedSnapshot) adapterID(model string) int32 {
	if id, ok := s.adapterIDs[model]; ok {
		return id
	}
	return int32(s.A)
}
