package gpuscheduling

/*
#include "lig.h"
*/
import "C"

import (
	"fmt"
	"math/rand"
	"sync"
	"time"

	"google.golang.org/grpc/codes"
	"google.golang.org/grpc/status"
	klog "k8s.io/klog/v2"

	"inference.networking.x-k8s.io/llm-instance-gateway/pkg/ext-proc/backend"
	"inference.networking.x-k8s.io/llm-instance-gateway/pkg/ext-proc/scheduling"
)

// Options of the GPU scheduler.  Thresholds default to the reference's constants
// (pkg/ext-proc/scheduling/scheduler.go:15-24).
type Options struct {
	Device          int
	MaxPods         int
	MaxAdapters     int
	MaxBatch        int
	FlushSize       int           // flush when this many requests are pending ...
	BatchWindow     time.Duration // ... or when the oldest pending request is this old
	RefreshInterval time.Duration // snapshot re-pack period; use main.go's refreshMetricsInterval
	KVCacheThreshold       float64
	QueueThresholdCritical int64
	QueueingThresholdLoRA  int64
}

func DefaultOptions() Options {
	return Options{MaxPods: 4096, MaxAdapters: 1024, MaxBatch: 1 << 16, FlushSize: 4096,
		BatchWindow: 20 * time.Microsecond, RefreshInterval: 50 * time.Millisecond,
		KVCacheThreshold: 0.8, QueueThresholdCritical: 5, QueueingThresholdLoRA: 50}
}

type waiter struct {
	req  *scheduling.LLMRequest
	done chan result
}

type result struct {
	pod backend.Pod
	err error
}

// GPUScheduler implements handlers.Scheduler (pkg/ext-proc/handlers/server.go:37-39).  Swap it
// in at pkg/ext-proc/main.go:137:
//
//	sched, err := gpuscheduling.NewGPUScheduler(pp, gpuscheduling.DefaultOptions())
//	handlers.NewServer(pp, sched, *targetPodHeader, datastore)
type GPUScheduler struct {
	pmp  scheduling.PodMetricsProvider
	opt  Options
	c    *ctx
	in   *pinned
	out  *pinned
	seed uint64
	rng  *rand.Rand

	snapMu sync.RWMutex
	snap   *packedSnapshot
	epoch  uint64

	queue     chan *waiter
	stop      chan struct{}
	wg        sync.WaitGroup
	closeOnce sync.Once
}

// NewGPUScheduler mirrors scheduling.NewScheduler(pmp) (scheduler.go:93-99).
func NewGPUScheduler(pmp scheduling.PodMetricsProvider, opt Options) (*GPUScheduler, error) {
	c, err := newCtx(opt.Device, opt.MaxPods, opt.MaxAdapters, opt.MaxBatch)
	if err != nil {
		return nil, err // no CPU fallback: without the GPU the caller keeps the stock scheduler
	}
	if err := c.setThresholds(opt.KVCacheThreshold, opt.QueueThresholdCritical, opt.QueueingThresholdLoRA); err != nil {
		c.close()
		return nil, err
	}
	in, err := allocPinned(opt.MaxBatch * 16)
	if err != nil {
		c.close()
		return nil, err
	}
	out, err := allocPinned(opt.MaxBatch * 8)
	if err != nil {
		in.free()
		c.close()
		return nil, err
	}
	g := &GPUScheduler{pmp: pmp, opt: opt, c: c, in: in, out: out, seed: rand.Uint64(),
		rng: rand.New(rand.NewSource(time.Now().UnixNano())), queue: make(chan *waiter, opt.MaxBatch), stop: make(chan struct{})}
	if err := g.Refresh(); err != nil {
		g.Close()
		return nil, err
	}
	g.wg.Add(2)
	go g.batcher()
	go g.refresher()
	return g, nil
}

// Close stops the batcher and the refresher, answers every request that was still queued (or that
// races in while the scheduler shuts down) with codes.Unavailable, then frees the pinned buffers
// and the device context.  No Schedule call is left blocked.
func (g *GPUScheduler) Close() {
	g.closeOnce.Do(func() {
		close(g.stop)
		g.wg.Wait() // batcher and refresher are gone: nobody reads g.queue any more
		for {
			select {
			case w := <-g.queue:
				w.done <- result{err: status.Error(codes.Unavailable, "gpu scheduler is shut down")}
				continue
			default:
			}
			break
		}
		g.in.free()
		g.out.free()
		g.c.close()
	})
}

// Refresh re-packs the provider's current slice and uploads it as a new epoch.  One pack per
// refresh tick replaces the two AllPodMetrics() materialisations per request of
// scheduler.go:114-115.
func (g *GPUScheduler) Refresh() error {
	s, err := packSnapshot(g.pmp.AllPodMetrics())
	if err != nil {
		return err
	}
	g.snapMu.Lock()
	defer g.snapMu.Unlock()
	g.epoch++
	s.epoch = g.epoch
	if err := g.c.uploadSnapshot(s.epoch, s); err != nil {
		return err
	}
	g.snap = s
	return nil
}

func (g *GPUScheduler) refresher() {
	defer g.wg.Done()
	t := time.NewTicker(g.opt.RefreshInterval)
	defer t.Stop()
	for {
		select {
		case <-g.stop:
			return
		case <-t.C:
			if err := g.Refresh(); err != nil {
				klog.Errorf("gpu scheduler: snapshot refresh failed, keeping the previous one: %v", err)
			}
		}
	}
}

// Schedule finds the target pod based on metrics and the requested lora adapter
// (scheduler.go:113-122).  Goroutine-safe and blocking, like the reference's.
func (g *GPUScheduler) Schedule(req *scheduling.LLMRequest) (backend.Pod, error) {
	w := &waiter{req: req, done: make(chan result, 1)}
	select {
	case g.queue <- w:
	case <-g.stop:
		return backend.Pod{}, status.Error(codes.Unavailable, "gpu scheduler is shut down")
	}
	// A request that slipped into the queue while Close was draining it is answered by the
	// second arm: stop is closed for good, so no waiter can block forever.
	select {
	case r := <-w.done:
		return r.pod, r.err
	case <-g.stop:
		select {
		case r := <-w.done:
			return r.pod, r.err
		case <-time.After(g.opt.BatchWindow + 10*time.Millisecond):
			return backend.Pod{}, status.Error(codes.Unavailable, "gpu scheduler is shut down")
		}
	}
}

func (g *GPUScheduler) batcher() {
	defer g.wg.Done()
	batch := make([]*waiter, 0, g.opt.FlushSize)
	for {
		batch = batch[:0]
		select {
		case <-g.stop:
			return
		case w := <-g.queue:
			batch = append(batch, w)
		}
		deadline := time.NewTimer(g.opt.BatchWindow)
	fill:
		for len(batch) < g.opt.FlushSize && len(batch) < g.opt.MaxBatch {
			select {
			case w := <-g.queue:
				batch = append(batch, w)
			case <-deadline.C:
				break fill
			}
		}
		deadline.Stop()
		g.flush(batch)
	}
}

func (g *GPUScheduler) flush(batch []*waiter) {
	n := len(batch)
	var snap *packedSnapshot
	var rc C.int
	for attempt := 0; attempt < 4; attempt++ {
		g.snapMu.RLock()
		snap = g.snap
		g.snapMu.RUnlock()
		reqs := g.in.reqs(n)
		for i, w := range batch {
			reqs[i].adapter_id = C.int32_t(snap.adapterID(w.req.ResolvedTargetModel))
			reqs[i].flags = 0
			if w.req.Critical {
				reqs[i].flags = C.LIG_REQ_CRITICAL
			}
			reqs[i].rand_key = C.uint64_t(g.rng.Uint64())
		}
		rc = g.c.scheduleBatch(snap.epoch, g.seed, g.in, g.out, n)
		if rc != C.LIG_ERR_STALE_EPOCH {
			break
		}
	}
	if rc != 0 {
		err := status.Errorf(codes.Internal, "gpu scheduler: %v", ligError("lig_schedule_batch", rc))
		for _, w := range batch {
			w.done <- result{err: err}
		}
		return
	}
	picks := g.out.picks(n)
	for i, w := range batch {
		switch picks[i].status {
		case C.LIG_OK:
			w.done <- result{pod: snap.pods[int(picks[i].pod_idx)]}
		case C.LIG_DROP:
			// identical to the reference: the ResourceExhausted status error of scheduler.go:87
			// wrapped with %w like scheduler.go:117, so status.Code(err) is still ResourceExhausted
			// after request.go:74 wraps it again and server.go:97-109 answers 429.
			inner := status.Errorf(codes.ResourceExhausted, "dropping request due to limited backend resources")
			w.done <- result{err: fmt.Errorf("failed to apply filter, resulted %v pods, this should never happen: %w", 0, inner)}
		default:
			var nilErr error
			w.done <- result{err: fmt.Errorf("failed to apply filter, resulted %v pods, this should never happen: %w", 0, nilErr)}
		}
	}
}
