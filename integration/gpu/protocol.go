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

// Rand is common.Rand (common/rand.go): the same SHAKE256 stream with the same rejection sampling,
// owned by the library so that the draw order a caller observes across CRS generation, proof
// generation and validation (whisk/whisk_test.go:36-50) is the reference's.
type Rand struct{ h *C.cdl_rand }

// NewRand is common.NewRand(seed).
func NewRand(seed uint64) (*Rand, error) {
	var h *C.cdl_rand
	if rc := C.cdl_rand_new(C.uint64_t(seed), &h); rc != 0 {
		return nil, errStatus(rc)
	}
	return &Rand{h}, nil
}

// Free releases the generator.
func (r *Rand) Free() { C.cdl_rand_free(r.h) }

// GetFrs is (*Rand).GetFrs.
func (r *Rand) GetFrs(n int) ([]fr.Element, error) {
	out := make([]fr.Element, n)
	if rc := C.cdl_rand_get_frs(r.h, C.size_t(n), frPtr(out)); rc != 0 {
		return nil, errStatus(rc)
	}
	return out, nil
}

type statusError int32

func (e statusError) Error() string { return "curdle_b200 status " + itoa(int(e)) }
func errStatus(rc C.int32_t) error  { return statusError(rc) }
func itoa(v int) string {
	if v == 0 {
		return "0"
	}
	neg := v < 0
	if neg {
		v = -v
	}
	var b [20]byte
	i := len(b)
	for v > 0 {
		i--
		b[i] = byte('0' + v%10)
		v /= 10
	}
	if neg {
		i--
		b[i] = '-'
	}
	return string(b[i:])
}

// CRS is curdleproof.CRS (crs.go:10-18) resident on the device.
type CRS struct {
	h   *C.cdl_crs
	Ell int
}

// GenerateCRS is curdleproof.GenerateCRS(size, rand) (crs.go:20-59).
func (c *Context) GenerateCRS(size int, r *Rand) (*CRS, error) {
	var h *C.cdl_crs
	if rc := C.cdl_crs_generate(c.h, C.size_t(size), r.h, &h); rc != 0 {
		return nil, c.err(rc)
	}
	return &CRS{h, size}, nil
}

// CRSFromPoints uploads an existing CRS: Gs[ell] | Hs[4] | H | Gt | Gu | Gsum | Hsum.
func (c *Context) CRSFromPoints(ell int, pts []bls12381.G1Affine) (*CRS, error) {
	var h *C.cdl_crs
	if rc := C.cdl_crs_from_points(c.h, C.size_t(ell), affPtr(pts), &h); rc != 0 {
		return nil, c.err(rc)
	}
	return &CRS{h, ell}, nil
}

// Free releases the CRS.
func (s *CRS) Free() { C.cdl_crs_free(s.h) }

// Prove is curdleproof.Prove (curdleproof.go:38-197) followed by Proof.Serialize (:358-387); the
// caller decodes the bytes with Proof.FromReader (or keeps them: they are what goes on the wire).
func (c *Context) Prove(crs *CRS, Rs, Ss, Ts, Us []bls12381.G1Affine, M *bls12381.G1Jac, perm []uint32,
	k *fr.Element, rsM []fr.Element, r *Rand) ([]byte, error) {
	buf := make([]byte, 48*(18+10*32)+32*7+40)
	var n C.size_t
	rc := C.cdl_prove(c.h, crs.h, affPtr(Rs), affPtr(Ss), affPtr(Ts), affPtr(Us),
		(*C.cdl_g1_jac)(unsafe.Pointer(M)), (*C.uint32_t)(unsafe.Pointer(&perm[0])),
		(*C.cdl_fr)(unsafe.Pointer(k)), frPtr(rsM), r.h, (*C.uint8_t)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)), &n)
	if rc != 0 {
		return nil, c.err(rc)
	}
	return buf[:n], nil
}

// Verify is Proof.FromReader + curdleproof.Verify (curdleproof.go:199-318): (bool, error) as in the
// reference — (false, nil) for an invalid proof, (false, err) for malformed input.
func (c *Context) Verify(crs *CRS, proof []byte, Rs, Ss, Ts, Us []bls12381.G1Affine, M *bls12381.G1Jac, r *Rand) (bool, error) {
	var ok C.int32_t
	rc := C.cdl_verify(c.h, crs.h, (*C.uint8_t)(unsafe.Pointer(&proof[0])), C.size_t(len(proof)),
		affPtr(Rs), affPtr(Ss), affPtr(Ts), affPtr(Us), (*C.cdl_g1_jac)(unsafe.Pointer(M)), r.h, &ok)
	if rc != 0 {
		return false, c.err(rc)
	}
	return ok == 1, nil
}

// WhiskTracker / WhiskShuffleProofBytes cross as the byte arrays they already are
// (whisk/types.go:21-24,74-77): 96 bytes per tracker, 4576 bytes per proof.
const (
	TrackerSize      = C.CDL_WHISK_TRACKER_SIZE
	ShuffleProofSize = C.CDL_WHISK_SHUFFLE_PROOF_SIZE
)

// GenerateWhiskShuffleProofs is whisk.GenerateWhiskShuffleProof (whisk/whisk.go:63-114) for B
// independent instances in lock step: pre holds B*ell trackers, rands one generator per instance.
func (c *Context) GenerateWhiskShuffleProofs(crs *CRS, pre []byte, rands []*Rand) (post, proofs []byte, status []int32, err error) {
	B := len(rands)
	post = make([]byte, len(pre))
	proofs = make([]byte, B*ShuffleProofSize)
	status = make([]int32, B)
	hs := make([]*C.cdl_rand, B)
	for i, r := range rands {
		hs[i] = r.h
	}
	rc := C.cdl_whisk_generate_shuffle_proof_batch(c.h, crs.h, C.size_t(B), (*C.uint8_t)(unsafe.Pointer(&pre[0])),
		(**C.cdl_rand)(unsafe.Pointer(&hs[0])), (*C.uint8_t)(unsafe.Pointer(&post[0])),
		(*C.uint8_t)(unsafe.Pointer(&proofs[0])), C.size_t(ShuffleProofSize), (*C.int32_t)(unsafe.Pointer(&status[0])))
	if rc != 0 {
		err = c.err(rc)
	}
	return
}

// IsValidWhiskShuffleProofs is whisk.IsValidWhiskShuffleProof (whisk/whisk.go:20-61) for B instances:
// ok[b] is the reference's bool, status[b] != 0 its error.
func (c *Context) IsValidWhiskShuffleProofs(crs *CRS, pre, post, proofs []byte, rands []*Rand) (ok, status []int32, err error) {
	B := len(rands)
	ok = make([]int32, B)
	status = make([]int32, B)
	hs := make([]*C.cdl_rand, B)
	for i, r := range rands {
		hs[i] = r.h
	}
	rc := C.cdl_whisk_is_valid_shuffle_proof_batch(c.h, crs.h, C.size_t(B), (*C.uint8_t)(unsafe.Pointer(&pre[0])),
		(*C.uint8_t)(unsafe.Pointer(&post[0])), (*C.uint8_t)(unsafe.Pointer(&proofs[0])), C.size_t(ShuffleProofSize),
		(**C.cdl_rand)(unsafe.Pointer(&hs[0])), (*C.int32_t)(unsafe.Pointer(&ok[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	if rc != 0 {
		err = c.err(rc)
	}
	return
}
