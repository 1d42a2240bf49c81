// Source a maintainer adds to gnark-whir (or to a gnark fork) to put libb200g16 behind gnark's groth16 API.
// NOT built or tested in this repository: the image has no Go toolchain (see INTEGRATION.md).  The same C-ABI,
// call for call, is exercised by gnark_whir_b200/lib.py + groth16.py in the test-suite.
//go:build b200

package b200

/*
#include "b200g16.h"
*/
import "C"

import (
	"errors"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/hash_to_field"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	"github.com/consensys/gnark/constraint"
)

// groth16.Verify(proof, vk, publicWitness) at mt.go:497
func Verify(ctx *C.b200g16_ctx, proof *groth16_bn254.Proof, vk *groth16_bn254.VerifyingKey, publicWitness fr.Vector) error {
	pub := append(fr.Vector{}, publicWitness...)
	var com, pok, pedG, pedGS *C.uint64_t
	if len(vk.PublicAndCommitmentCommitted) > 0 {           // one BSB22 commitment (what this circuit has)
		// challenge = hash_to_field("bsb22-commitment")(proof.Commitments[0].Marshal() || committed publics),
		// exactly as gnark v0.11.0 verify.go computes it
		h := hash_to_field.New([]byte(constraint.CommitmentDst))
		maxNb := len(vk.PublicAndCommitmentCommitted[0])
		buf := make([]byte, 0, 64+maxNb*fr.Bytes)
		buf = append(buf, proof.Commitments[0].Marshal()...)
		for _, w := range vk.PublicAndCommitmentCommitted[0] {
			b := pub[w-1].Bytes() // wire ids count the constant-one wire; publicWitness does not hold it
			buf = append(buf, b[:]...)
		}
		h.Write(buf)
		var challenge fr.Element
		challenge.SetBytes(h.Sum(nil)[:fr.Bytes])
		pub = append(pub, challenge)
		com, pok = u64(unsafe.Pointer(&proof.Commitments[0])), u64(unsafe.Pointer(&proof.CommitmentPok))
		pedG, pedGS = u64(unsafe.Pointer(&vk.CommitmentKeys[0].G)), u64(unsafe.Pointer(&vk.CommitmentKeys[0].GSigmaNeg))
	}
	d := C.b200g16_vk_desc{
		g1_alpha: u64(unsafe.Pointer(&vk.G1.Alpha)), g2_beta: u64(unsafe.Pointer(&vk.G2.Beta)),
		g2_gamma: u64(unsafe.Pointer(&vk.G2.Gamma)), g2_delta: u64(unsafe.Pointer(&vk.G2.Delta)),
		g1_k: u64(unsafe.Pointer(&vk.G1.K[0])), n_k: C.size_t(len(vk.G1.K)), ped_g: pedG, ped_g_sigma_neg: pedGS,
	}
	var ok C.int
	if err := call(func() C.int {
		return C.b200g16_verify(ctx, &d, u64(unsafe.Pointer(&proof.Ar)), u64(unsafe.Pointer(&proof.Bs)),
			u64(unsafe.Pointer(&proof.Krs)), com, pok, u64(unsafe.Pointer(&pub[0])), C.size_t(len(pub)), &ok)
	}); err != nil {
		return err                                           // malformed input ("invalid witness size", bad point)
	}
	if ok == 0 { return errors.New("pairing doesn't match") } // gnark's errPairingCheckFailed
	return nil
}
