// goref — runs the UNMODIFIED reference (github.com/jsign/curdleproofs, resolved to the checkout
// named by CURDLE_REF, default /root/reference, through the `replace` line of go.mod) and writes
// its outputs in the schema of tests/golden/*.json, so that the CPU oracle and the GPU path can
// be pinned to bytes produced by the Go code itself.
//
// TEST INFRASTRUCTURE ONLY.  This image has no Go toolchain and no module cache (gnark-crypto
// v0.11.0, jsign/merlin and x/crypto are un-vendored dependencies), so the program is committed
// but has never been run here; `make -C oracle/goref` builds and runs it wherever `go` and the
// modules are available and drops the fixtures into tests/golden/ref_*.json, which
// tests/test_oracle_goref.py then compares with the oracle byte for byte.
//
// Flows (the reference's own tests):
//   ref_whisk_ell124.json   whisk/whisk_test.go:36-56 TestWhiskShuffleProof, one continuing Rand(0)
//   ref_prove_ell{60,124,252,508}.json   curdleproof_test.go:16-46 with setup :239-274, in two
//       variants: perm = Rand(42).GeneratePermutation (what tests/golden/prove_ell*.json use) and
//       perm = math/rand.New(NewSource(42)).Shuffle (the reference's own benchmark setup)
package main

import (
	"bytes"
	"crypto/sha256"
	"encoding/hex"
	"encoding/json"
	"fmt"
	mrand "math/rand"
	"os"
	"path/filepath"
	"unsafe"

	bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381"
	"github.com/consensys/gnark-crypto/ecc/bls12-381/fr"
	curdleproof "github.com/jsign/curdleproofs"
	"github.com/jsign/curdleproofs/common"
	"github.com/jsign/curdleproofs/whisk"
)

func must(err error) {
	if err != nil {
		panic(err)
	}
}

func encAff(pts ...bls12381.G1Affine) string {
	var b bytes.Buffer
	for i := range pts {
		e := pts[i].Bytes()
		b.Write(e[:])
	}
	return hex.EncodeToString(b.Bytes())
}

func jacToAff(p bls12381.G1Jac) bls12381.G1Affine {
	var a bls12381.G1Affine
	a.FromJacobian(&p)
	return a
}

func crsPoints(crs curdleproof.CRS) []bls12381.G1Affine {
	out := append([]bls12381.G1Affine{}, crs.Gs...)
	out = append(out, crs.Hs...)
	out = append(out, jacToAff(crs.H), jacToAff(crs.Gt), jacToAff(crs.Gu), crs.Gsum, crs.Hsum)
	return out
}

func frHex(x fr.Element) string {
	b := x.Bytes()
	return hex.EncodeToString(b[:])
}

// WhiskTracker is two [48]byte arrays with unexported names (whisk/types.go:74-77)
func trackerBytes(ts []whisk.WhiskTracker) string {
	var b bytes.Buffer
	for i := range ts {
		raw := (*[96]byte)(unsafe.Pointer(&ts[i]))
		b.Write(raw[:])
	}
	return hex.EncodeToString(b.Bytes())
}

func whiskFixture() map[string]any {
	rand, err := common.NewRand(0)
	must(err)
	crs, err := curdleproof.GenerateCRS(whisk.ELL, rand)
	must(err)
	_, _, g1Gen, _ := bls12381.Generators()
	pre := make([]whisk.WhiskTracker, whisk.ELL)
	for i := range pre { // whisk_test.go:116-124 generateShuffleTrackers: k then r per tracker
		k, err := rand.GetFr()
		must(err)
		r, err := rand.GetFr()
		must(err)
		var rG, krG bls12381.G1Affine
		rG.ScalarMultiplication(&g1Gen, common.FrToBigInt(&r))
		krG.ScalarMultiplication(&rG, common.FrToBigInt(&k))
		pre[i] = whisk.NewWhiskTracker(rG, krG)
	}
	post, proof, err := whisk.GenerateWhiskShuffleProof(crs, pre, rand)
	must(err)
	ok, err := whisk.IsValidWhiskShuffleProof(crs, pre, post, proof, rand)
	must(err)
	tail, err := rand.GetFr()
	must(err)
	used := len(bytes.TrimRight(proof[:], "\x00"))
	return map[string]any{
		"flow":                    "whisk/whisk_test.go:36-56 TestWhiskShuffleProof, Rand(0) continuing (Go reference run)",
		"ell":                     whisk.ELL,
		"crs":                     encAff(crsPoints(crs)...),
		"pre_trackers":            trackerBytes(pre),
		"post_trackers":           trackerBytes(post),
		"proof":                   hex.EncodeToString(proof[:]),
		"proof_used_bytes":        used,
		"valid":                   ok,
		"next_fr_after_roundtrip": frHex(tail),
	}
}

func proveFixture(ell int, goShuffle bool) map[string]any {
	rand, err := common.NewRand(0)
	must(err)
	crs, err := curdleproof.GenerateCRS(ell, rand)
	must(err)
	var perm []uint32
	if goShuffle { // curdleproof_test.go:254-260
		perm = make([]uint32, ell)
		for i := range perm {
			perm[i] = uint32(i)
		}
		srand := mrand.New(mrand.NewSource(42))
		srand.Shuffle(len(perm), func(i, j int) { perm[i], perm[j] = perm[j], perm[i] })
	} else {
		pr, err := common.NewRand(42)
		must(err)
		perm, err = pr.GeneratePermutation(ell)
		must(err)
	}
	k, err := rand.GetFr()
	must(err)
	Rs, err := rand.GetG1Affines(ell)
	must(err)
	Ss, err := rand.GetG1Affines(ell)
	must(err)
	Ts, Us, M, rsM, err := common.ShufflePermuteCommit(crs.Gs, crs.Hs, Rs, Ss, perm, k, rand)
	must(err)
	prand, err := common.NewRand(42)
	must(err)
	proof, err := curdleproof.Prove(crs, Rs, Ss, Ts, Us, M, perm, k, rsM, prand)
	must(err)
	var pb bytes.Buffer
	must(proof.Serialize(&pb))
	vrand, err := common.NewRand(43)
	must(err)
	ok, err := curdleproof.Verify(proof, crs, Rs, Ss, Ts, Us, M, vrand)
	must(err)
	crsSum := sha256.Sum256(mustHex(encAff(crsPoints(crs)...)))
	inst := encAff(Rs...) + encAff(Ss...) + encAff(Ts...) + encAff(Us...) + encAff(jacToAff(M))
	instSum := sha256.Sum256(mustHex(inst))
	rs := make([]string, len(rsM))
	for i := range rsM {
		rs[i] = frHex(rsM[i])
	}
	return map[string]any{
		"flow":            "curdleproof_test.go:16-46 (setup :239-274), Go reference run",
		"perm_source":     map[bool]string{true: "math/rand NewSource(42) Shuffle", false: "common.Rand(42).GeneratePermutation"}[goShuffle],
		"ell":             ell,
		"crs_sha256":      hex.EncodeToString(crsSum[:]),
		"instance_sha256": hex.EncodeToString(instSum[:]),
		"instance":        inst,
		"M":               encAff(jacToAff(M)),
		"rs_m":            rs,
		"k":               frHex(k),
		"perm":            perm,
		"proof":           hex.EncodeToString(pb.Bytes()),
		"valid":           ok,
	}
}

func mustHex(s string) []byte {
	b, err := hex.DecodeString(s)
	must(err)
	return b
}

func write(dir, name string, v map[string]any) {
	b, err := json.MarshalIndent(v, "", " ")
	must(err)
	must(os.WriteFile(filepath.Join(dir, name), b, 0o644))
	fmt.Println("wrote", filepath.Join(dir, name))
}

func main() {
	dir := "."
	if len(os.Args) > 1 {
		dir = os.Args[1]
	}
	write(dir, "ref_whisk_ell124.json", whiskFixture())
	for _, ell := range []int{60, 124, 252, 508} {
		write(dir, fmt.Sprintf("ref_prove_ell%d.json", ell), proveFixture(ell, false))
		write(dir, fmt.Sprintf("ref_prove_goshuffle_ell%d.json", ell), proveFixture(ell, true))
	}
}
