"""CPU oracle for the B200 Groth16/BN254 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (gnark_whir_b200/) may import,
link or execute anything under oracle/.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, and there only as the checker
or as the reported CPU baseline.

PARITY UNPINNED by the reference: /root/reference ships no tests, no golden vectors,
and the arithmetic it calls lives in un-vendored Go modules (gnark v0.11.0,
gnark-crypto v0.14.1-0.20241217131346-b998989abdbe; go.mod:6-7) that cannot be built
here (no Go toolchain).  The oracle therefore restates the published algorithms and
is pinned by first-principles known answers instead (see tests/test_oracle_*.py):
  * Keccak-f[1600]: SHA3-256 / SHAKE128 built on it == hashlib; FIPS-202 zero-state KAT
  * Fp/Fr: python int arithmetic is the ground truth
  * G1/G2: generators on curve, [r]G = inf, homomorphism
  * MSM: known-discrete-log closed form
  * NTT: O(N^2) DFT definition; computeH: polynomial identity at a random point
  * Groth16: independent optimal-ate pairing check + closed-form Ar/Bs/Krs from the
    toxic waste
"""
