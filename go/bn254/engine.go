package bn254

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gopairingbasedcryptography_b200/lib -lbn254_b200 -Wl,-rpath,${SRCDIR}/../../gopairingbasedcryptography_b200/lib
#include <stdlib.h>
#include "bn254_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// ErrInvalidSizes is the error gnark returns from Pair / MillerLoop / PairingCheck (same text).
var ErrInvalidSizes = errors.New("invalid inputs sizes")

// device is one GPU: one engine context (its own two streams and pinned staging) and the mutex-free handle the C
// library locks internally.
type device struct {
	ordinal int
	ctx     *C.bn254_ctx
}

// pool holds one context per visible B200.  Batches are split into ceil(n/G) contiguous chunks, one per GPU, each
// driven by its own goroutine (locked to an OS thread for the duration of the call so the CUDA device selection
// the library makes stays with it); there is no inter-GPU traffic (SURVEY.md 8e).
type pool struct {
	devs []*device
}

var (
	thePool  *pool
	poolOnce sync.Once
	poolErr  error
)

func engines() (*pool, error) {
	poolOnce.Do(func() {
		n := int(C.bn254_device_count())
		if n == 0 {
			poolErr = errors.New("bn254: no CUDA device (this package has no CPU path)")
			return
		}
		p := &pool{}
		for i := 0; i < n; i++ {
			var ctx *C.bn254_ctx
			if rc := C.bn254_ctx_create(C.int(i), &ctx); rc != 0 {
				for _, d := range p.devs {
					C.bn254_ctx_destroy(d.ctx)
				}
				poolErr = fmt.Errorf("bn254: bn254_ctx_create(device %d) failed with code %d", i, int(rc))
				return
			}
			p.devs = append(p.devs, &device{ordinal: i, ctx: ctx})
		}
		thePool = p
	})
	return thePool, poolErr
}

// DeviceCount reports how many GPUs the batch entry points spread over.
func DeviceCount() int {
	p, err := engines()
	if err != nil {
		return 0
	}
	return len(p.devs)
}

// Close releases every context (tables created through this package must be closed first).
func Close() {
	if thePool != nil {
		for _, d := range thePool.devs {
			C.bn254_ctx_destroy(d.ctx)
		}
		thePool = nil
	}
}

func (d *device) check(rc C.int) error {
	switch rc {
	case C.BN254_OK:
		return nil
	case C.BN254_ERR_INVALID_SIZES:
		return ErrInvalidSizes
	default:
		return fmt.Errorf("bn254: device %d: %s (code %d)", d.ordinal, C.GoString(C.bn254_last_error(d.ctx)), int(rc))
	}
}

// shard calls f(dev, lo, hi) for contiguous chunks of [0, n): ceil(n/G) items per GPU, all GPUs in parallel.  Small
// batches (fewer than minPerDevice items per GPU) stay on fewer devices: a chunk below one wave of threads only adds
// launch latency.
func shard(n, minPerDevice int, f func(d *device, lo, hi int) error) error {
	p, err := engines()
	if err != nil {
		return err
	}
	g := len(p.devs)
	if minPerDevice > 0 && n/minPerDevice < g {
		g = n / minPerDevice
		if g < 1 {
			g = 1
		}
	}
	if g == 1 {
		return f(p.devs[0], 0, n)
	}
	per := (n + g - 1) / g
	errs := make([]error, g)
	var wg sync.WaitGroup
	for i := 0; i < g; i++ {
		lo, hi := i*per, (i+1)*per
		if hi > n {
			hi = n
		}
		if lo >= hi {
			continue
		}
		wg.Add(1)
		go func(i, lo, hi int) {
			defer wg.Done()
			runtime.LockOSThread()
			defer runtime.UnlockOSThread()
			errs[i] = f(p.devs[i], lo, hi)
		}(i, lo, hi)
	}
	wg.Wait()
	for _, e := range errs {
		if e != nil {
			return e
		}
	}
	return nil
}

// one runs f on device 0 (1-element calls of the gnark-named API).
func one(f func(d *device) error) error {
	p, err := engines()
	if err != nil {
		return err
	}
	return f(p.devs[0])
}

func ptr[T any](s []T) unsafe.Pointer {
	if len(s) == 0 {
		return nil
	}
	return unsafe.Pointer(&s[0])
}

// PinnedBytes allocates page-locked host memory (bn254_host_alloc): buffers carved from it are copied to / from the
// GPU directly, without the staging memcpy.  Free with FreePinned.
func PinnedBytes(n int) []byte {
	p := C.bn254_host_alloc(C.size_t(n))
	if p == nil {
		return nil
	}
	return unsafe.Slice((*byte)(p), n)
}

// FreePinned releases a PinnedBytes buffer.
func FreePinned(b []byte) {
	if len(b) > 0 {
		C.bn254_host_free(unsafe.Pointer(&b[0]))
	}
}
