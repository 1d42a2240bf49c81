// One Go process, several GPUs: gnark calls groth16.Prove once (mt.go:496); a group spreads that one call over the
// GPUs of the box through libb200g16's b200g16_group_* entry points (no NCCL, no helper processes).
// NOT built or tested in this repository (no Go toolchain); the same calls run in tests/test_gpu_group.py via ctypes.
//go:build b200

package b200

/*
#include "b200g16.h"
*/
import "C"

import (
	"math/bits"
	"runtime"
	"unsafe"

	curve "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	cs "github.com/consensys/gnark/constraint/bn254"
)

// GroupProvingKey = gnark's pk + its point-range shards resident on every GPU of the group.
type GroupProvingKey struct {
	groth16_bn254.ProvingKey
	grp *C.b200g16_group
	dev *C.b200g16_group_pk
}

// Upload cuts the key into one shard per device (devices = CUDA ordinals) and builds the window tables there.
func (pk *GroupProvingKey) Upload(devices []int, kSkip []byte) error {
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	if err := call(func() C.int { return C.b200g16_group_init(&devs[0], C.int(len(devs)), &pk.grp) }); err != nil {
		return err
	}
	infA, infB := boolsToBytes(pk.InfinityA), boolsToBytes(pk.InfinityB)
	var pin runtime.Pinner
	defer pin.Unpin()
	for _, p := range []any{&pk.G1.A[0], &pk.G1.B[0], &pk.G1.K[0], &pk.G1.Z[0], &pk.G2.B[0], &pk.G1.Alpha, &pk.G1.Beta,
		&pk.G1.Delta, &pk.G2.Beta, &pk.G2.Delta, &infA[0], &infB[0], &kSkip[0]} {
		pin.Pin(p)
	}
	d := C.b200g16_pk_desc{
		log2_domain: C.uint(bits.TrailingZeros64(pk.Domain.Cardinality)),
		n_wires:     C.size_t(len(pk.InfinityA)),
		g1_a:        u64(unsafe.Pointer(&pk.G1.A[0])), n_a: C.size_t(len(pk.G1.A)),
		g1_b:        u64(unsafe.Pointer(&pk.G1.B[0])), n_b: C.size_t(len(pk.G1.B)),
		g1_k:        u64(unsafe.Pointer(&pk.G1.K[0])), n_k: C.size_t(len(pk.G1.K)),
		g1_z:        u64(unsafe.Pointer(&pk.G1.Z[0])), n_z: C.size_t(len(pk.G1.Z)),
		g2_b:        u64(unsafe.Pointer(&pk.G2.B[0])),
		g1_alpha:    u64(unsafe.Pointer(&pk.G1.Alpha)), g1_beta: u64(unsafe.Pointer(&pk.G1.Beta)),
		g1_delta:    u64(unsafe.Pointer(&pk.G1.Delta)),
		g2_beta:     u64(unsafe.Pointer(&pk.G2.Beta)), g2_delta: u64(unsafe.Pointer(&pk.G2.Delta)),
		infinity_a:  (*C.uint8_t)(unsafe.Pointer(&infA[0])), infinity_b: (*C.uint8_t)(unsafe.Pointer(&infB[0])),
		k_skip:      (*C.uint8_t)(unsafe.Pointer(&kSkip[0])),
		precompute:  1,
	}
	return call(func() C.int { return C.b200g16_group_pk_upload(pk.grp, &d, &pk.dev) })
}

// ProveSolved runs everything after the constraint solver on all GPUs of the group: computeH split over the
// devices (2, 4 or 8: cross-GPU NTT levels over NVLink peer memory), the five MSMs on every shard, the partial
// points added on the host.  r, s as in b200.Prove.
func (pk *GroupProvingKey) ProveSolved(solution *cs.R1CSSolution, r, s *fr.Element) (*groth16_bn254.Proof, error) {
	wires := []fr.Element(solution.W)
	var out C.b200g16_proof
	err := call(func() C.int {
		return C.b200g16_group_prove(pk.grp, pk.dev,
			u64(unsafe.Pointer(&wires[0])), C.size_t(len(wires)),
			u64(unsafe.Pointer(&solution.A[0])), u64(unsafe.Pointer(&solution.B[0])), u64(unsafe.Pointer(&solution.C[0])),
			C.size_t(len(solution.A)), u64(unsafe.Pointer(r)), u64(unsafe.Pointer(s)), &out, nil)
	})
	if err != nil {
		return nil, err
	}
	proof := &groth16_bn254.Proof{}
	proof.Ar = *(*curve.G1Affine)(unsafe.Pointer(&out.ar))
	proof.Bs = *(*curve.G2Affine)(unsafe.Pointer(&out.bs))
	proof.Krs = *(*curve.G1Affine)(unsafe.Pointer(&out.krs))
	return proof, nil
}

func (pk *GroupProvingKey) Free() {
	if pk.dev != nil {
		C.b200g16_group_pk_free(pk.dev)
		pk.dev = nil
	}
	if pk.grp != nil {
		C.b200g16_group_destroy(pk.grp)
		pk.grp = nil
	}
}
