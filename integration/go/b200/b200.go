// Source a maintainer adds to gnark-whir (or to a gnark fork) to put libb200g16 behind gnark's groth16 API.
// NOT built or tested in this repository: the image has no Go toolchain (see INTEGRATION.md).  The same C-ABI,
// call for call, is exercised by gnark_whir_b200/lib.py + groth16.py in the test-suite.
//go:build b200

package b200

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -lb200g16
#include "b200g16.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"

	curve "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/pedersen"
	"github.com/consensys/gnark/backend"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	cs "github.com/consensys/gnark/constraint/bn254"
	"github.com/consensys/gnark/constraint/solver"
)

// ProvingKey = gnark's pk + lazily created device handles (like icicle's deviceInfo).
type ProvingKey struct {
	groth16_bn254.ProvingKey
	ctx       *C.b200g16_ctx
	dev       *C.b200g16_pk
	pedBasis  []*C.b200g16_bases // pk.CommitmentKeys[i].Basis
	pedSigma  []*C.b200g16_bases // pk.CommitmentKeys[i].BasisExpSigma
}

func check(st C.int) error {
	if st == 0 {
		return nil
	}
	return errors.New(C.GoString(C.b200g16_last_error()))
}

func u64(p unsafe.Pointer) *C.uint64_t { return (*C.uint64_t)(p) }

func boolsToBytes(b []bool) []byte { /* 1 byte per flag */ }

// setup uploads the key once; every slice is passed zero-copy (gnark-crypto's in-memory
// layout IS the ABI layout: fr/fp.Element = [4]uint64 Montgomery, G1Affine{X,Y}, G2Affine{X{A0,A1},Y{A0,A1}}).
func (pk *ProvingKey) setup(r1cs *cs.R1CS) error {
	if pk.dev != nil {
		return nil
	}
	if err := check(C.b200g16_init(0, &pk.ctx)); err != nil {
		return err // no CPU fallback: the caller decides what to do
	}
	nbWires := len(pk.InfinityA)
	// k_skip: wires that are NOT in the K MSM = public wires + BSB22 committed wires + commitment wires
	kSkip := make([]byte, nbWires)
	for i := 0; i < r1cs.GetNbPublicVariables(); i++ { kSkip[i] = 1 }
	for _, c := range r1cs.CommitmentInfo.(constraint.Groth16Commitments) {
		kSkip[c.CommitmentIndex] = 1
		for _, w := range c.PrivateCommitted { kSkip[w] = 1 }
	}
	infA, infB := boolsToBytes(pk.InfinityA), boolsToBytes(pk.InfinityB)
	d := C.b200g16_pk_desc{
		log2_domain: C.uint(bits.TrailingZeros64(pk.Domain.Cardinality)),
		n_wires:     C.size_t(nbWires),
		g1_a: u64(unsafe.Pointer(&pk.G1.A[0])), n_a: C.size_t(len(pk.G1.A)),
		g1_b: u64(unsafe.Pointer(&pk.G1.B[0])), n_b: C.size_t(len(pk.G1.B)),
		g1_k: u64(unsafe.Pointer(&pk.G1.K[0])), n_k: C.size_t(len(pk.G1.K)),
		g1_z: u64(unsafe.Pointer(&pk.G1.Z[0])), n_z: C.size_t(len(pk.G1.Z)),
		g2_b: u64(unsafe.Pointer(&pk.G2.B[0])),
		g1_alpha: u64(unsafe.Pointer(&pk.G1.Alpha)), g1_beta: u64(unsafe.Pointer(&pk.G1.Beta)),
		g1_delta: u64(unsafe.Pointer(&pk.G1.Delta)),
		g2_beta: u64(unsafe.Pointer(&pk.G2.Beta)), g2_delta: u64(unsafe.Pointer(&pk.G2.Delta)),
		infinity_a: (*C.uint8_t)(&infA[0]), infinity_b: (*C.uint8_t)(&infB[0]), k_skip: (*C.uint8_t)(&kSkip[0]),
		precompute: 1, // window tables over A, B1, K, Z, B2 (13 instead of 15-17 additions per point; ~13x the
		               // key's HBM footprint; identical proofs) — drop it when the key does not fit
	}
	// (cgo: the desc holds Go pointers -> pin them with runtime.Pinner for the duration of the call)
	var pin runtime.Pinner
	defer pin.Unpin()
	/* pin.Pin(&pk.G1.A[0]) ... for every pointer field */
	if err := check(C.b200g16_pk_upload(pk.ctx, &d, &pk.dev)); err != nil {
		return err
	}
	for i := range pk.CommitmentKeys {
		ck := &pk.CommitmentKeys[i]
		var b, s *C.b200g16_bases
		check(C.b200g16_bases_upload_g1(pk.ctx, u64(unsafe.Pointer(&ck.Basis[0])), C.size_t(len(ck.Basis)), &b))
		check(C.b200g16_bases_upload_g1(pk.ctx, u64(unsafe.Pointer(&ck.BasisExpSigma[0])), C.size_t(len(ck.BasisExpSigma)), &s))
		pk.pedBasis, pk.pedSigma = append(pk.pedBasis, b), append(pk.pedSigma, s)
	}
	return nil
}

func msmG1(ctx *C.b200g16_ctx, bases *C.b200g16_bases, scalars []fr.Element) (curve.G1Affine, error) {
	var out curve.G1Affine
	var p *C.uint64_t
	if len(scalars) > 0 { p = u64(unsafe.Pointer(&scalars[0])) }
	err := check(C.b200g16_msm_g1(ctx, bases, 0, p, C.size_t(len(scalars)), u64(unsafe.Pointer(&out))))
	return out, err
}

// Prove has gnark's signature: groth16.Prove(ccs, pk, fullWitness, opts...) at mt.go:496.
func Prove(r1cs *cs.R1CS, pk *ProvingKey, fullWitness witness.Witness, opts ...backend.ProverOption) (*groth16_bn254.Proof, error) {
	opt, err := backend.NewProverConfig(opts...)
	if err != nil { return nil, err }
	if err := pk.setup(r1cs); err != nil { return nil, err }

	commitmentInfo := r1cs.CommitmentInfo.(constraint.Groth16Commitments)
	proof := &groth16_bn254.Proof{Commitments: make([]curve.G1Affine, len(commitmentInfo))}
	privateCommittedValues := make([][]fr.Element, len(commitmentInfo))

	// BSB22 hint override: identical to gnark's, except Commit() is the GPU MSM on the resident basis
	solverOpts := opt.SolverOpts[:len(opt.SolverOpts):len(opt.SolverOpts)]
	for i := range commitmentInfo {
		i := i
		solverOpts = append(solverOpts, solver.OverrideHint(commitmentInfo[i].HintID,
			func(_ *big.Int, in []*big.Int, out []*big.Int) error {
				/* ... copy committed values exactly as gnark's prove.go does ... */
				proof.Commitments[i], err = msmG1(pk.ctx, pk.pedBasis[i], privateCommittedValues[i])
				/* ... hash_to_field("bsb22-commitment") over Marshal() || committed publics, unchanged ... */
				return err
			}))
	}
	_solution, err := r1cs.Solve(fullWitness, solverOpts...)           // Go, CPU (out of scope)
	if err != nil { return nil, err }
	solution := _solution.(*cs.R1CSSolution)
	wireValues := []fr.Element(solution.W)

	// Pedersen PoK (gnark: pedersen.BatchProve): one MSM per commitment on BasisExpSigma, folded with the
	// "G16-BSB22" challenge exactly as upstream
	/* proof.CommitmentPok = fold_i( msmG1(pk.ctx, pk.pedSigma[i], privateCommittedValues[i]) ) */

	// r, s: gnark samples them here; they are explicit in the ABI so CPU and GPU proofs can be compared bit for bit
	var _r, _s fr.Element
	_r.SetRandom(); _s.SetRandom()

	var out C.b200g16_proof
	err = check(C.b200g16_prove(pk.ctx, pk.dev,
		u64(unsafe.Pointer(&wireValues[0])), C.size_t(len(wireValues)),
		u64(unsafe.Pointer(&solution.A[0])), u64(unsafe.Pointer(&solution.B[0])), u64(unsafe.Pointer(&solution.C[0])),
		C.size_t(len(solution.A)),
		u64(unsafe.Pointer(&_r)), u64(unsafe.Pointer(&_s)), &out, nil))
	if err != nil { return nil, err }
	proof.Ar = *(*curve.G1Affine)(unsafe.Pointer(&out.ar))
	proof.Bs = *(*curve.G2Affine)(unsafe.Pointer(&out.bs))
	proof.Krs = *(*curve.G1Affine)(unsafe.Pointer(&out.krs))
	return proof, nil
}
