"""Developer probe: latency of verify / pairing_check and throughput of point decode (run under gpurun)."""
import sys, time, random
import numpy as np
sys.path.insert(0, ".")
from gnark_whir_b200 import lib, groth16 as g16
from oracle import bn254 as bn, groth16 as og
from oracle.bn254 import R
ctx = lib.Context(0)
rng = random.Random(1)
r1cs, w = og.synthetic_r1cs(64, 4, rng, with_commitment=True)
pk, vk = g16.Setup(ctx, r1cs)
def resolve(wit):
    L, Rr, O = r1cs.constraints[-1]
    wit[O[0][0]] = og.lc_eval(L, wit) * og.lc_eval(Rr, wit) % R
proof = g16.Prove(ctx, r1cs, pk, w, resolve=resolve)
pub = proof.debug["witness"][1:r1cs.nb_public]
g16.Verify(ctx, proof, vk, pub)
t0 = time.perf_counter()
for _ in range(3):
    g16.Verify(ctx, proof, vk, pub)
print(f"verify (with commitment: 4 + 2 Miller loops, 2 final exps): {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
P1, Q1 = bn.g1_to_array([bn.G1_GEN, bn.g1_neg(bn.G1_GEN)]), bn.g2_to_array([bn.G2_GEN, bn.G2_GEN])
ctx.pairing_check(P1, Q1)
t0 = time.perf_counter()
for _ in range(3):
    assert ctx.pairing_check(P1, Q1)
print(f"pairing_check (2 pairs): {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
rs = np.random.Generator(np.random.PCG64(5))
for group, logn in ((1, 22), (2, 20)):
    n = 1 << logn
    ks = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); ks[:, 3] &= np.uint64((1 << 60) - 1)
    gen = g16.g1_point(g16.G1_GEN) if group == 1 else g16.g2_point(g16.G2_GEN)
    pts = ctx.fixed_base_mul(gen, ks, group=group)
    t0 = time.perf_counter(); enc = ctx.encode_points(pts, group=group); t1 = time.perf_counter()
    dec, ok = ctx.decode_points(enc, group=group, subgroup_check=False); t2 = time.perf_counter()
    assert ok.all() and np.array_equal(dec, pts)
    print(f"G{group} 2^{logn}: encode {1e3*(t1-t0):.1f} ms, decode (sqrt per point, incl. PCIe) {1e3*(t2-t1):.1f} ms = {n/(t2-t1)/1e6:.1f} Mpts/s")
pk.free(); ctx.close()
