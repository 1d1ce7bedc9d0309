package gpuscheduling

// UNVERIFIED (no Go toolchain in the build image).  Runs the reference's own golden vectors —
// the committed fixture tests/golden/go_filter_test_vectors.json, extracted from
// pkg/ext-proc/scheduling/filter_test.go and pkg/ext-proc/test/hermetic_test.go — through
// GPUScheduler.Schedule on a machine with a B200 and liblig.so built.  The same vectors are run
// through the C++ host runtime by tests/test_host_runtime.py, which IS executed on the GPU box.

import (
	"encoding/json"
	"os"
	"path/filepath"
	"testing"

	"google.golang.org/grpc/codes"
	"google.golang.org/grpc/status"

	"inference.networking.x-k8s.io/llm-instance-gateway/pkg/ext-proc/backend"
	"inference.networking.x-k8s.io/llm-instance-gateway/pkg/ext-proc/scheduling"
)

type fixturePod struct {
	Name      string   `json:"name"`
	Address   string   `json:"address"`
	Queue     int      `json:"waiting_queue_size"`
	KV        float64  `json:"kv_cache_usage_percent"`
	MaxActive int      `json:"max_active_models"`
	Active    []string `json:"active_models"`
}

type fixtureCase struct {
	Name   string `json:"name"`
	Filter struct {
		Name string `json:"name"`
	} `json:"filter"`
	Req *struct {
		Model    string `json:"model"`
		Resolved string `json:"resolved_target_model"`
		Critical bool   `json:"critical"`
	} `json:"req"`
	Input  []fixturePod `json:"input"`
	Output []fixturePod `json:"output"`
	Err    bool         `json:"err"`
}

type sliceProvider []*backend.PodMetrics

func (s sliceProvider) AllPodMetrics() []*backend.PodMetrics { return s }

func toPodMetrics(in []fixturePod) sliceProvider {
	out := make(sliceProvider, 0, len(in))
	for _, p := range in {
		active := map[string]int{}
		for _, a := range p.Active {
			active[a] = 1
		}
		out = append(out, &backend.PodMetrics{
			Pod: backend.Pod{Name: p.Name, Address: p.Address},
			Metrics: backend.Metrics{WaitingQueueSize: p.Queue, KVCacheUsagePercent: p.KV,
				MaxActiveModels: p.MaxActive, ActiveModels: active},
		})
	}
	return out
}

func TestGoldenVectorsThroughGPUScheduler(t *testing.T) {
	raw, err := os.ReadFile(filepath.Join("..", "..", "tests", "golden", "go_filter_test_vectors.json"))
	if err != nil {
		t.Fatal(err)
	}
	var fx struct {
		TestFilter []fixtureCase `json:"TestFilter"`
	}
	if err := json.Unmarshal(raw, &fx); err != nil {
		t.Fatal(err)
	}
	for _, c := range fx.TestFilter {
		if c.Filter.Name != "defaultFilter" {
			continue // the error-passthrough case is a property of the Go node engine, not of the path
		}
		t.Run(c.Name, func(t *testing.T) {
			opt := DefaultOptions()
			opt.MaxPods, opt.MaxAdapters, opt.MaxBatch = 64, 64, 256
			sched, err := NewGPUScheduler(toPodMetrics(c.Input), opt)
			if err != nil {
				t.Skipf("no GPU scheduler available: %v", err)
			}
			defer sched.Close()
			pod, err := sched.Schedule(&scheduling.LLMRequest{Model: c.Req.Model,
				ResolvedTargetModel: c.Req.Resolved, Critical: c.Req.Critical})
			if c.Err {
				if status.Code(err) != codes.ResourceExhausted {
					t.Fatalf("want ResourceExhausted (-> 429 at handlers/server.go:97-109), got %v", err)
				}
				return
			}
			if err != nil || len(c.Output) != 1 || pod.Name != c.Output[0].Name {
				t.Fatalf("got pod %v err %v, want %v", pod, err, c.Output)
			}
		})
	}
}
