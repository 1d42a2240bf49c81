// abdump — the A/B kit: runs REAL gnark / gnark-crypto (the versions /root/reference/go.mod pins) on small seeded
// inputs and dumps every value that crosses libb200g16's C-ABI, so that `pytest tests/test_gnark_golden.py` turns
// "parity unpinned" into byte-for-byte parity on any box that has a Go toolchain:
//
//     cd integration/go/abdump && go mod tidy && go run . ../../../tests/golden/gnark
//     python -m pytest tests/test_gnark_golden.py            (CPU: pins the oracle)   and   -m gpu (pins the kernels)
//
// NOT built in this repository's image (no Go toolchain, no module cache, no network) — it is source for the
// maintainer's machine.  What is gnark's own code here: MultiExp (G1, G2), fft.Domain FFT / FFTInverse in every
// decimation / coset combination, the constraint solver (solution.W / A / B / C), Setup (pk, vk and their
// WriteRawTo / WriteTo bytes), Prove + Verify.  gnark draws the blinding scalars r, s inside Prove, so the proof with
// FIXED r, s is assembled below from gnark-crypto's MultiExp / fft results exactly as prove.go does (computeH and the
// Ar / Bs / Krs sums are restated in a few lines) and is then accepted by gnark's own groth16.Verify before it is
// written: the glue is pinned by the verifier, the arithmetic is gnark-crypto's.
//
// File format (all integers little-endian): a sequence of records  u32 name_len | name | u64 payload_len | payload.
// Field elements / points are dumped as they lie in memory (fr.Element / fp.Element = 4 x u64 Montgomery limbs,
// G1Affine = X, Y; G2Affine = X.A0, X.A1, Y.A0, Y.A1) — the layout the C-ABI takes.
package main

import (
	"bytes"
	"encoding/binary"
	"fmt"
	"math/big"
	"math/rand"
	"os"
	"path/filepath"

	"github.com/consensys/gnark-crypto/ecc"
	curve "github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/fft"
	"github.com/consensys/gnark/backend/groth16"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	cs "github.com/consensys/gnark/constraint/bn254"
	"github.com/consensys/gnark/frontend"
	"github.com/consensys/gnark/frontend/cs/r1cs"
)

type dump struct{ buf bytes.Buffer }

func (d *dump) rec(name string, payload []byte) {
	binary.Write(&d.buf, binary.LittleEndian, uint32(len(name)))
	d.buf.WriteString(name)
	binary.Write(&d.buf, binary.LittleEndian, uint64(len(payload)))
	d.buf.Write(payload)
}
func (d *dump) u64(name string, v uint64) {
	var b [8]byte
	binary.LittleEndian.PutUint64(b[:], v)
	d.rec(name, b[:])
}
func limbs(b *bytes.Buffer, e [4]uint64) {
	for _, l := range e {
		binary.Write(b, binary.LittleEndian, l)
	}
}
func frs(v []fr.Element) []byte {
	var b bytes.Buffer
	for i := range v {
		limbs(&b, v[i])
	}
	return b.Bytes()
}
func g1s(v []curve.G1Affine) []byte {
	var b bytes.Buffer
	for i := range v {
		limbs(&b, v[i].X)
		limbs(&b, v[i].Y)
	}
	return b.Bytes()
}
func g2s(v []curve.G2Affine) []byte {
	var b bytes.Buffer
	for i := range v {
		limbs(&b, v[i].X.A0)
		limbs(&b, v[i].X.A1)
		limbs(&b, v[i].Y.A0)
		limbs(&b, v[i].Y.A1)
	}
	return b.Bytes()
}
func bools(v []bool) []byte {
	out := make([]byte, len(v))
	for i, x := range v {
		if x {
			out[i] = 1
		}
	}
	return out
}
func (d *dump) save(dir, name string) {
	if err := os.WriteFile(filepath.Join(dir, name), d.buf.Bytes(), 0o644); err != nil {
		panic(err)
	}
	fmt.Println("wrote", name, d.buf.Len(), "bytes")
}

// seeded scalars: full-width, plus the zero / one / r-1 / byte-sized values the WHIR witness is made of
func scalars(rng *rand.Rand, n int) []fr.Element {
	out := make([]fr.Element, n)
	for i := range out {
		switch rng.Intn(6) {
		case 0:
			out[i].SetUint64(uint64(rng.Intn(2)))
		case 1:
			out[i].SetUint64(uint64(rng.Intn(256)))
		case 2:
			out[i].SetOne()
			out[i].Neg(&out[i]) // r - 1
		default:
			var b [32]byte
			rng.Read(b[:])
			out[i].SetBytes(b[:])
		}
	}
	return out
}

func dumpMSM(dir string, rng *rand.Rand) {
	for _, logn := range []int{6, 10, 14} {
		n := 1 << logn
		_, _, g1, g2 := curve.Generators()
		ks := scalars(rng, n)
		p1 := curve.BatchScalarMultiplicationG1(&g1, ks)
		p2 := curve.BatchScalarMultiplicationG2(&g2, ks)
		p1[3] = p1[2]                    // a repeated base and a point at infinity, as real keys have
		p1[5] = curve.G1Affine{}
		sc := scalars(rng, n)
		var r1 curve.G1Affine
		var r2 curve.G2Affine
		if _, err := r1.MultiExp(p1, sc, ecc.MultiExpConfig{}); err != nil {
			panic(err)
		}
		if _, err := r2.MultiExp(p2, sc, ecc.MultiExpConfig{}); err != nil {
			panic(err)
		}
		var d dump
		d.u64("log2n", uint64(logn))
		d.rec("g1_points", g1s(p1))
		d.rec("g2_points", g2s(p2))
		d.rec("scalars", frs(sc))
		d.rec("g1_result", g1s([]curve.G1Affine{r1}))
		d.rec("g2_result", g2s([]curve.G2Affine{r2}))
		d.save(dir, fmt.Sprintf("msm_%d.bin", logn))
	}
}

func dumpFFT(dir string, rng *rand.Rand) {
	for _, logn := range []int{3, 10} {
		n := 1 << logn
		domain := fft.NewDomain(uint64(n))
		in := scalars(rng, n)
		var d dump
		d.u64("log2n", uint64(logn))
		d.rec("input", frs(in))
		for _, inverse := range []bool{false, true} {
			for _, coset := range []bool{false, true} {
				for _, dec := range []fft.Decimation{fft.DIF, fft.DIT} {
					a := make([]fr.Element, n)
					copy(a, in)
					var opts []fft.Option
					if coset {
						opts = append(opts, fft.OnCoset())
					}
					if inverse {
						domain.FFTInverse(a, dec, opts...)
					} else {
						domain.FFT(a, dec, opts...)
					}
					decN := 0
					if dec == fft.DIT {
						decN = 1
					}
					d.rec(fmt.Sprintf("out_inv%d_coset%d_dec%d", b2i(inverse), b2i(coset), decN), frs(a))
				}
			}
		}
		d.save(dir, fmt.Sprintf("fft_%d.bin", logn))
	}
}

func b2i(b bool) int {
	if b {
		return 1
	}
	return 0
}

// ---- a small circuit: y == x^3 + x + 5 chained `Depth` times, one public input
type cubic struct {
	X frontend.Variable
	Y frontend.Variable `gnark:",public"`
}

const depth = 200

func (c *cubic) Define(api frontend.API) error {
	v := c.X
	for i := 0; i < depth; i++ {
		x3 := api.Mul(v, v, v)
		v = api.Add(x3, v, 5)
	}
	api.AssertIsEqual(c.Y, v)
	return nil
}

// prove.go computeH, restated on gnark-crypto's fft.Domain
func computeH(a, b, c []fr.Element, domain *fft.Domain) []fr.Element {
	n := int(domain.Cardinality)
	pad := func(v []fr.Element) []fr.Element { return append(v, make([]fr.Element, n-len(v))...) }
	a, b, c = pad(a), pad(b), pad(c)
	domain.FFTInverse(a, fft.DIF)
	domain.FFTInverse(b, fft.DIF)
	domain.FFTInverse(c, fft.DIF)
	domain.FFT(a, fft.DIT, fft.OnCoset())
	domain.FFT(b, fft.DIT, fft.OnCoset())
	domain.FFT(c, fft.DIT, fft.OnCoset())
	var den, one fr.Element
	one.SetOne()
	den.Exp(domain.FrMultiplicativeGen, big.NewInt(int64(domain.Cardinality)))
	den.Sub(&den, &one).Inverse(&den)
	for i := range a {
		a[i].Mul(&a[i], &b[i]).Sub(&a[i], &c[i]).Mul(&a[i], &den)
	}
	domain.FFTInverse(a, fft.DIF, fft.OnCoset())
	return a
}

func dumpProve(dir string, rng *rand.Rand) {
	var circuit cubic
	ccs, err := frontend.Compile(ecc.BN254.ScalarField(), r1cs.NewBuilder, &circuit)
	if err != nil {
		panic(err)
	}
	pkI, vkI, err := groth16.Setup(ccs)
	if err != nil {
		panic(err)
	}
	pk := pkI.(*groth16_bn254.ProvingKey)
	x := big.NewInt(3)
	y := new(big.Int).Set(x)
	mod := ecc.BN254.ScalarField()
	for i := 0; i < depth; i++ {
		t := new(big.Int).Exp(y, big.NewInt(3), mod)
		y.Add(t, y).Add(y, big.NewInt(5)).Mod(y, mod)
	}
	w, err := frontend.NewWitness(&cubic{X: x, Y: y}, ecc.BN254.ScalarField())
	if err != nil {
		panic(err)
	}
	pubW, _ := w.Public()
	solI, err := ccs.(*cs.R1CS).Solve(w)
	if err != nil {
		panic(err)
	}
	sol := solI.(*cs.R1CSSolution)
	wires := []fr.Element(sol.W)
	a, b, c := []fr.Element(sol.A), []fr.Element(sol.B), []fr.Element(sol.C)

	// gnark's own proof (random r, s) and its acceptance: cross-verification material
	proofI, err := groth16.Prove(ccs, pkI, w)
	if err != nil {
		panic(err)
	}
	if err := groth16.Verify(proofI, vkI, pubW); err != nil {
		panic(err)
	}

	// the same prove with FIXED r, s on gnark-crypto primitives (prove.go, steps after Solve)
	rs := scalars(rand.New(rand.NewSource(99)), 8)
	r, s := rs[6], rs[7]
	var wa, wb, wk []fr.Element
	nbPub := ccs.GetNbPublicVariables()
	for i := range wires {
		if !pk.InfinityA[i] {
			wa = append(wa, wires[i])
		}
		if !pk.InfinityB[i] {
			wb = append(wb, wires[i])
		}
		if i >= nbPub {
			wk = append(wk, wires[i]) // no commitment in this circuit: K = every private wire
		}
	}
	ac, bc, cc := append([]fr.Element{}, a...), append([]fr.Element{}, b...), append([]fr.Element{}, c...)
	h := computeH(ac, bc, cc, &pk.Domain)
	var mA, mB1, mK, mZ curve.G1Affine
	var mB2 curve.G2Affine
	cfg := ecc.MultiExpConfig{}
	mA.MultiExp(pk.G1.A, wa, cfg)
	mB1.MultiExp(pk.G1.B, wb, cfg)
	mK.MultiExp(pk.G1.K, wk, cfg)
	mZ.MultiExp(pk.G1.Z, h[:len(pk.G1.Z)], cfg)
	mB2.MultiExp(pk.G2.B, wb, cfg)
	var rB, sB, krB big.Int
	r.BigInt(&rB)
	s.BigInt(&sB)
	var kr fr.Element
	kr.Mul(&r, &s).Neg(&kr)
	kr.BigInt(&krB)
	var ar, bs1, krs, t curve.G1Affine
	var bs, t2 curve.G2Affine
	ar.Add(&mA, &pk.G1.Alpha)
	t.ScalarMultiplication(&pk.G1.Delta, &rB)
	ar.Add(&ar, &t)
	bs1.Add(&mB1, &pk.G1.Beta)
	t.ScalarMultiplication(&pk.G1.Delta, &sB)
	bs1.Add(&bs1, &t)
	krs.Add(&mK, &mZ)
	t.ScalarMultiplication(&pk.G1.Delta, &krB)
	krs.Add(&krs, &t)
	t.ScalarMultiplication(&ar, &sB)
	krs.Add(&krs, &t)
	t.ScalarMultiplication(&bs1, &rB)
	krs.Add(&krs, &t)
	bs.Add(&mB2, &pk.G2.Beta)
	t2.ScalarMultiplication(&pk.G2.Delta, &sB)
	bs.Add(&bs, &t2)
	fixed := &groth16_bn254.Proof{Ar: ar, Krs: krs, Bs: bs}
	if err := groth16.Verify(fixed, vkI, pubW); err != nil {
		panic(fmt.Errorf("the fixed-r,s proof assembled from gnark-crypto primitives does not verify: %w", err))
	}

	var d dump
	d.u64("log2_domain", uint64(bitLen(pk.Domain.Cardinality)-1))
	d.u64("nb_public", uint64(nbPub))
	d.u64("nb_wires", uint64(len(wires)))
	var pkRaw, pkComp, vkRaw, vkComp, prRaw, prComp, gnarkProof bytes.Buffer
	pk.WriteRawTo(&pkRaw)
	pk.WriteTo(&pkComp)
	vkI.WriteRawTo(&vkRaw)
	vkI.WriteTo(&vkComp)
	fixed.WriteRawTo(&prRaw)
	fixed.WriteTo(&prComp)
	proofI.WriteRawTo(&gnarkProof)
	d.rec("pk_raw", pkRaw.Bytes())
	d.rec("pk_compressed", pkComp.Bytes())
	d.rec("vk_raw", vkRaw.Bytes())
	d.rec("vk_compressed", vkComp.Bytes())
	d.rec("pk_g1_a", g1s(pk.G1.A))
	d.rec("pk_g1_b", g1s(pk.G1.B))
	d.rec("pk_g1_k", g1s(pk.G1.K))
	d.rec("pk_g1_z", g1s(pk.G1.Z))
	d.rec("pk_g2_b", g2s(pk.G2.B))
	d.rec("pk_g1_alpha_beta_delta", g1s([]curve.G1Affine{pk.G1.Alpha, pk.G1.Beta, pk.G1.Delta}))
	d.rec("pk_g2_beta_delta", g2s([]curve.G2Affine{pk.G2.Beta, pk.G2.Delta}))
	d.rec("infinity_a", bools(pk.InfinityA))
	d.rec("infinity_b", bools(pk.InfinityB))
	d.rec("wires", frs(wires))
	d.rec("a", frs(a))
	d.rec("b", frs(b))
	d.rec("c", frs(c))
	d.rec("r", frs([]fr.Element{r}))
	d.rec("s", frs([]fr.Element{s}))
	d.rec("h", frs(h))
	d.rec("msm_a", g1s([]curve.G1Affine{mA}))
	d.rec("msm_b1", g1s([]curve.G1Affine{mB1}))
	d.rec("msm_k", g1s([]curve.G1Affine{mK}))
	d.rec("msm_z", g1s([]curve.G1Affine{mZ}))
	d.rec("msm_b2", g2s([]curve.G2Affine{mB2}))
	d.rec("ar", g1s([]curve.G1Affine{ar}))
	d.rec("bs", g2s([]curve.G2Affine{bs}))
	d.rec("krs", g1s([]curve.G1Affine{krs}))
	d.rec("proof_raw", prRaw.Bytes())
	d.rec("proof_compressed", prComp.Bytes())
	d.rec("gnark_proof_raw", gnarkProof.Bytes())
	d.save(dir, "prove_cubic.bin")
}

func bitLen(v uint64) int {
	n := 0
	for v != 0 {
		n++
		v >>= 1
	}
	return n
}

func main() {
	dir := "."
	if len(os.Args) > 1 {
		dir = os.Args[1]
	}
	if err := os.MkdirAll(dir, 0o755); err != nil {
		panic(err)
	}
	rng := rand.New(rand.NewSource(20261018))
	dumpMSM(dir, rng)
	dumpFFT(dir, rng)
	dumpProve(dir, rng)
}
