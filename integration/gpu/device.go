package gpu

/*
#include "curdle_b200.h"
*/
import "C"

import (
	"unsafe"

	bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381"
	"github.com/consensys/gnark-crypto/ecc/bls12-381/fr"
)

// PointPool is an array of affine points that lives in HBM: the working vectors of an argument
// (G, G', T', U' of innerproductargument.go:100-172 / samemultiscalarargument.go:85-140) are
// uploaded once, folded in place on the device and used as MSM bases round after round.
type PointPool struct {
	c *Context
	p unsafe.Pointer
	n int
}

// NewPointPool allocates room for n points and uploads `init` to its front.
func (c *Context) NewPointPool(n int, init []bls12381.G1Affine) (*PointPool, error) {
	var d unsafe.Pointer
	if rc := C.cdl_dev_alloc(c.h, C.size_t(96*n), &d); rc != 0 {
		return nil, c.err(rc)
	}
	pp := &PointPool{c, d, n}
	if len(init) > 0 {
		if rc := C.cdl_dev_upload(c.h, d, unsafe.Pointer(&init[0]), C.size_t(96*len(init))); rc != 0 {
			pp.Free()
			return nil, c.err(rc)
		}
	}
	return pp, nil
}

// Free releases the device memory.
func (p *PointPool) Free() { C.cdl_dev_free(p.c.h, p.p) }

func (p *PointPool) at(i int) *C.cdl_g1_affine {
	return (*C.cdl_g1_affine)(unsafe.Add(p.p, 96*i))
}

// MultiExpBatch is (*G1Jac).MultiExp for k MSMs over bases that stay on the device: term t of MSM j
// is scalars[t] * pool[idx[t]] (bit 31 of idx[t] negates), t in offsets[j]:offsets[j+1].  Result j
// is stored into pool[outSlot[j]] (outSlot may be nil) and returned both affine and as the 48-byte
// compressed encoding transcript.AppendPoints hashes.
func (p *PointPool) MultiExpBatch(idx []uint32, scalars []fr.Element, offsets, outSlot []uint32) ([]bls12381.G1Affine, []byte, error) {
	k := len(offsets) - 1
	if k <= 0 {
		return nil, nil, nil
	}
	out := make([]bls12381.G1Affine, k)
	enc := make([]byte, 48*k)
	var ip, sp *C.uint32_t
	if len(idx) > 0 {
		ip = (*C.uint32_t)(unsafe.Pointer(&idx[0]))
	}
	if outSlot != nil {
		sp = (*C.uint32_t)(unsafe.Pointer(&outSlot[0]))
	}
	rc := C.cdl_g1_msm_batch_device(p.c.h, (*C.cdl_g1_affine)(p.p), ip, frPtr(scalars),
		(*C.uint32_t)(unsafe.Pointer(&offsets[0])), C.size_t(k), sp, affPtr(out), (*C.uint8_t)(unsafe.Pointer(&enc[0])))
	if rc != 0 {
		return nil, nil, p.c.err(rc)
	}
	return out, enc, nil
}

// Fold is pool[l+i] += x * pool[r+i] for i < n, on the device (the scalar is uploaded per call).
func (p *PointPool) Fold(l, r, n int, x *fr.Element) error {
	var dx unsafe.Pointer
	if rc := C.cdl_dev_alloc(p.c.h, 32, &dx); rc != 0 {
		return p.c.err(rc)
	}
	defer C.cdl_dev_free(p.c.h, dx)
	if rc := C.cdl_dev_upload(p.c.h, dx, unsafe.Pointer(x), 32); rc != 0 {
		return p.c.err(rc)
	}
	if rc := C.cdl_g1_fold_device(p.c.h, p.at(l), p.at(r), (*C.cdl_fr)(dx), C.size_t(n)); rc != 0 {
		return p.c.err(rc)
	}
	return nil
}

// Download copies pool[i:i+n] back to the host.
func (p *PointPool) Download(i, n int) ([]bls12381.G1Affine, error) {
	out := make([]bls12381.G1Affine, n)
	if n == 0 {
		return out, nil
	}
	if rc := C.cdl_dev_download(p.c.h, unsafe.Pointer(&out[0]), unsafe.Pointer(p.at(i)), C.size_t(96*n)); rc != 0 {
		return nil, p.c.err(rc)
	}
	return out, nil
}

// JoinComm joins the library's NCCL communicator (one process per GPU).  Rank 0 obtains the id with
// UniqueID and ships the 128 bytes to the other ranks over whatever transport the application has.
func (c *Context) JoinComm(id [128]byte, rank, world int) error {
	if rc := C.cdl_comm_init(c.h, (*C.uint8_t)(unsafe.Pointer(&id[0])), C.int32_t(rank), C.int32_t(world)); rc != 0 {
		return c.err(rc)
	}
	return nil
}

// UniqueID creates the communicator id on rank 0.
func UniqueID() (id [128]byte, ok bool) {
	return id, C.cdl_comm_unique_id((*C.uint8_t)(unsafe.Pointer(&id[0]))) == 0
}

// MultiExpSharded: every rank passes the same (replicated, device-resident) vectors and receives the
// same point.  Rank r sums the windows r, r+world, .. of the signed-digit decomposition; the only
// exchange is an NCCL all-gather of one 144-byte partial sum per rank inside the library.
func (c *Context) MultiExpSharded(points *PointPool, dScalars unsafe.Pointer, n int) (bls12381.G1Jac, error) {
	var out bls12381.G1Jac
	var d unsafe.Pointer
	if rc := C.cdl_dev_alloc(c.h, 144, &d); rc != 0 {
		return out, c.err(rc)
	}
	defer C.cdl_dev_free(c.h, d)
	if rc := C.cdl_g1_msm_sharded_device(c.h, (*C.cdl_g1_affine)(points.p), (*C.cdl_fr)(dScalars), C.size_t(n),
		(*C.cdl_g1_jac)(d), nil); rc != 0 {
		return out, c.err(rc)
	}
	if rc := C.cdl_dev_download(c.h, unsafe.Pointer(&out), d, 144); rc != 0 {
		return out, c.err(rc)
	}
	return out, nil
}
