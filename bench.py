"""bench.py — BN254 G1 MSM throughput on B200 (BASELINE.json configs[1]), plus the other hot-path
numbers of the metric ("Groth16 prove ms; BN254 MSM Mpts/s, NTT GB/s") as extras.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # CPU arm: the C port of gnark's MultiExp on host cores

A "step" is one multi-scalar multiplication of 2^logn points (default 2^24, uniform scalars),
bases resident in HBM.  With N ranks the point range is sharded N ways (strong scaling), each
rank runs a local Pippenger, and one all_gather of N 64-byte partial points + N-1 host additions
produce the result on every rank.
  value  scalars already resident in HBM (b200g16_msm_g1_dev)
  e2e    scalars in pinned HOST memory through b200g16_msm_g1 (H2D inside the timed region)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
MAD_PER_MODMUL = 136          # 2*8*8 + 8 32x32 multiply-adds per Montgomery product
MODMUL_PER_MADD = 10          # XYZZ mixed add: 8M + 2S
# executed by k_accumulate per mixed add: 6 products, 2 dedicated squares (sqr_ptx: 108), and y3 = R(Q-X3) - Y*PPP
# as one fused two-term product (dot2_ptx: 200 instead of 2 x 136)
EXECUTED_MAD_PER_MADD = 6 * 136 + 2 * 108 + 200


def rand_fr(rs, n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)        # < 2^252 < r: every row is a valid Montgomery residue
    return a


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """All the host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank, which
    would silently run the CPU arm on one core — the thread count is passed to the C port explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_msm_sample(log_sample, steps, warmup):
    from oracle import cport
    nt = host_threads()
    rs = np.random.Generator(np.random.PCG64(SEED))
    n = 1 << log_sample
    k0, d = rand_fr(rs, 1), rand_fr(rs, 1)
    pts = cport.g1_progression(k0, d, n)
    sc = rand_fr(rs, n)
    for _ in range(warmup):
        cport.msm_g1(pts[: n >> 4], sc[: n >> 4], nt)
    times = []
    out = None
    for _ in range(steps):
        t0 = time.perf_counter()
        out = cport.msm_g1(pts, sc, nt)
        times.append(time.perf_counter() - t0)
    ok = bool(np.array_equal(out, cport.g1_gen_mul(cport.fr_dot_progression(sc, k0, d))))
    return n, times, nt, ok


def run_reference(args, rank):
    if rank != 0:
        return
    log_sample = min(args.logn, args.cpu_logn)
    n, times, cores, ok = cpu_msm_sample(log_sample, args.steps, min(args.warmup, 1))
    ms = 1e3 * sum(times) / len(times)
    val = n / (ms * 1e3)
    sample = (f"C port of gnark-crypto MultiExp (oracle/c/oracle.c; gnark itself cannot be built here: no Go "
              f"toolchain), G1 MSM of 2^{log_sample} points per step, uniform scalars, {cores} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": "BN254 G1 MSM throughput", "value": round(val, 4), "unit": "Mpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"G1 MSM, bounded sample 2^{log_sample} points of the 2^{args.logn}-point workload",
                   "result_checked": ok},
        "cpu_baseline": {"value": round(val, 4), "unit": "Mpts/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from gnark_whir_b200 import lib, sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = lib.Context(local_rank)
    peaks, peaks_kind = load_peaks()
    n = 1 << args.logn
    lo, hi = sharded.shard_range(n, rank, world)
    m = hi - lo
    rs = np.random.Generator(np.random.PCG64(SEED + 1000 * rank))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- inputs: bases = k_i * G generated on the GPU (known discrete logs), scalars uniform
    from gnark_whir_b200 import groth16 as g16
    from oracle import cport          # the checker for the parity gate below; never on the timed path
    gen = g16.g1_point(g16.G1_GEN)
    ks = rand_fr(rs, m)
    bases = ctx.fixed_base_mul(gen, ks, group=1, resident=True)
    table_ms = None
    if args.table:
        # window table over the resident bases, built once (like the pk upload): W rows of m points
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bases.precompute(0)
        table_ms = (time.perf_counter() - t0) * 1e3
    win_c, win_W = ctx.msm_plan(bases, m)
    sc_host_t = torch.empty((m, 4), dtype=torch.int64).pin_memory()
    sc_host = sc_host_t.numpy().view(np.uint64)
    sc_host[:] = rand_fr(rs, m)
    sc_dev = sc_host_t.to(dev)

    def step_resident():
        partial = ctx.msm(bases, sc_dev.data_ptr(), n=m)
        return sharded.exchange_and_combine(partial, 1, device=dev)

    def step_e2e():
        partial = ctx.msm(bases, sc_host)              # H2D of this step's scalars happens inside
        return sharded.exchange_and_combine(partial, 1, device=dev)

    # ---- parity gate before timing: shard result == closed form (C oracle: dot product, then [dot]G)
    partial = ctx.msm(bases, sc_dev.data_ptr(), n=m)
    expect = cport.g1_gen_mul(cport.fr_dot(ks, sc_host))
    if not np.array_equal(partial, expect):
        raise SystemExit("bench: GPU MSM result differs from the oracle's closed form — refusing to time it")
    checked = True

    # integer-pipe peak, measured live (burst): modmul/s with 4 independent chains, 8 CTAs/SM
    probe_rate, _ = ctx.modmul_probe(8, 4, 2000)
    tmad_peak = probe_rate * MAD_PER_MODMUL / 1e12

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        l0 = ctx.launch_count()
        phases = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
            phases.append(ctx.last_timings())
        e1.record()
        sync_all()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        return ms, phases, ctx.launch_count() - l0, res

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, phases, launches, res_a = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _, _, res_b = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    clocks = sampler.stop() if rank == 0 else None
    if not np.array_equal(res_a, res_b):
        raise SystemExit("bench: resident and end-to-end results differ")
    launches_total = int(sum_over_ranks(launches))

    # ---- dominant kernel: k_accumulate (phase index 2 of [digits, sort, accumulate, merge, reduce])
    acc_ms = max_over_ranks(statistics.mean(p[2] for p in phases if len(p) >= 5))
    msm_dev_ms = max_over_ranks(statistics.mean(sum(p) for p in phases if len(p) >= 5))
    algo_bytes = 96.0 * m                     # 64 B point + 32 B scalar per point
    # integer roofline of the dominant kernel: mixed adds = W digits per point, 10 products of 136 MADs each
    mads = float(win_W) * m * MODMUL_PER_MADD * MAD_PER_MODMUL
    ach = mads / (acc_ms * 1e-3) / 1e12
    traffic = args.ncu_traffic
    if traffic is None and args.logn == 24 and args.table and world == 1:
        traffic = 15.78e9        # profiles/r01c_ncu_k_accumulate_2p24_traffic.txt (ncu, this exact command)
    roofline = {"bound": "integer", "kernel": "k_accumulate<Fp>", "achieved": round(ach, 3), "peak": round(tmad_peak, 3),
                "unit": "TMAD/s", "frac": round(ach / tmad_peak, 4), "traffic": traffic,
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum); "
                                "algorithmic gather bytes per launch = adds_per_point * points * 68",
                "peak_source": "b200g16_modmul_probe measured in this run (burst; 136 IMAD.WIDE-class multiply-adds per "
                               "Montgomery product); MEASURED_PEAKS.json has no integer-pipe figure",
                "algorithmic_mads_per_launch": int(mads), "window_bits": win_c, "adds_per_point": win_W,
                "frac_executed": round(ach * EXECUTED_MAD_PER_MADD / (MODMUL_PER_MADD * MAD_PER_MODMUL) / tmad_peak, 4),
                "accumulate_ms": round(acc_ms, 4), "msm_device_ms": round(msm_dev_ms, 4),
                "note": "achieved counts SURVEY's algorithmic 1360 multiply-adds per mixed add; the kernel executes 1232 "
                        "(dedicated Montgomery square, fused two-term product), frac_executed is the pipe's own "
                        "utilisation.  "
                        "north_star: MSM is judged against the integer pipe (no dense contraction, 96 B/point of HBM "
                        "traffic against ~20k multiply-adds/point); the HBM view is roofline_hbm"}
    roofline_hbm = {"bound": "hbm", "kernel": "k_accumulate<Fp>", "achieved": round(algo_bytes / (acc_ms * 1e6), 2),
                    "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(algo_bytes / (acc_ms * 1e6) / peaks["hbm_gbs"], 5),
                    "peak_source": f"MEASURED_PEAKS.json ({peaks_kind})"}
    out = {
        "metric": "BN254 G1 MSM throughput", "value": round(n / (ms * 1e3), 3), "unit": "Mpts/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"G1 MSM 2^{args.logn} points (BASELINE.json configs[1]), uniform scalars, bases resident",
                   "window_table": ({"c": win_c, "rows": win_W, "build_ms_once": round(table_ms, 1),
                                     "bytes_per_gpu": int(win_W * m * 64)} if args.table else None),
                   "points": n, "points_per_gpu": m, "parallelism": f"point-range shards x{world} + all_gather of partial points",
                   "l2": "inputs (96 B/point, >= 190 MB per GPU) exceed the 126 MB L2; no flush needed",
                   "result_checked_vs_oracle": checked},
        "e2e": {"value": round(n / (ms_e2e * 1e3), 3), "unit": "Mpts/s", "ms_per_step": round(ms_e2e, 4),
                "h2d_bytes_per_step": int(32 * n), "d2h_bytes_per_step": int(world * ((1 if args.table else win_W) * 128 + 64))},
        "gpu_launches": launches_total,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
    }
    return ctx, out, dict(dev=dev, bases=bases, tmad_peak=tmad_peak, acc_ms=acc_ms, m=m, peaks=peaks)


def extras_single_gpu(ctx, st, args):
    """The other numbers of the metric, measured once at N=1 (not part of the timed MSM steps)."""
    import torch

    from gnark_whir_b200 import lib
    dev = st["dev"]
    ex = {}

    def t_ms(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def rnd(shape):
        a = torch.randint(0, 1 << 62, shape, dtype=torch.int64, device=dev)
        a[..., 3] &= (1 << 60) - 1
        return a
    L = args.ntt_logn
    N = 1 << L
    a, b, c = rnd((N, 4)), rnd((N, 4)), rnd((N, 4))
    ctx.ntt_dev(a.data_ptr(), L, coset=True, decimation=lib.DIF)            # warm tables
    ms = t_ms(lambda: ctx.ntt_dev(a.data_ptr(), L, decimation=lib.DIF))
    ex["ntt"] = {"log2n": L, "ms": round(ms, 4), "GBs_algorithmic_64N": round(64 * N / ms / 1e6, 1),
                 "frac_hbm": round(64 * N / ms / 1e6 / st["peaks"]["hbm_gbs"], 4),
                 "TMAD_s": round((N / 2) * L * MAD_PER_MODMUL / ms / 1e9, 3),
                 "frac_integer": round((N / 2) * L * MAD_PER_MODMUL / ms / 1e9 / st["tmad_peak"], 4)}
    ms = t_ms(lambda: ctx.compute_h_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), L))
    ex["compute_h"] = {"log2n": L, "ms": round(ms, 4), "GBs_algorithmic_576N": round(576 * N / ms / 1e6, 1)}
    del a, b, c
    nk = 1 << 22
    stt = torch.randint(0, 1 << 62, (nk, 25), dtype=torch.int64, device=dev)
    ms = t_ms(lambda: ctx.keccak_f_batch_dev(stt.data_ptr(), nk))
    ex["keccak_f_batch"] = {"states": nk, "ms": round(ms, 4), "Gperm_s": round(nk / ms / 1e6, 3),
                            "GBs_400B": round(400 * nk / ms / 1e6, 1)}
    del stt
    # synthetic prove: WHIR-verifier-shaped witness mix, N = 2^prove_logn constraints, wires = N
    Lp = args.prove_logn
    Np = 1 << Lp
    rs = np.random.Generator(np.random.PCG64(SEED + 7))
    from gnark_whir_b200 import groth16 as g16
    g1 = g16.g1_point(g16.G1_GEN)
    g2 = g16.g2_point(g16.G2_GEN)
    vecs = [ctx.fixed_base_mul(g1, rand_fr(rs, k), group=1, resident=True) for k in (Np, Np, Np - 1, Np - 1)]
    b2 = ctx.fixed_base_mul(g2, rand_fr(rs, Np), group=2, resident=True)
    if args.table:
        for v in vecs + [b2]:
            v.precompute(0)
    small = ctx.fixed_base_mul(g1, rand_fr(rs, 3), group=1)
    small2 = ctx.fixed_base_mul(g2, rand_fr(rs, 2), group=2)
    k_skip = np.zeros(Np, dtype=np.uint8)
    k_skip[0] = 1
    pk = ctx.pk_upload(Lp, Np, vecs[0], vecs[1], vecs[2], vecs[3], b2, small[0], small[1], small[2], small2[0],
                       small2[1], np.zeros(Np, np.uint8), np.zeros(Np, np.uint8), k_skip)
    u = torch.rand(Np, device=dev)
    wires = rnd((Np, 4))
    smallv = torch.randint(0, 256, (Np,), dtype=torch.int64, device=dev)
    # 40% zero/one, 30% bytes, 30% full width.  The small values must be small AFTER the library's
    # Montgomery->canonical conversion, so they are written as x*R mod r from a 256-entry table.
    mask01 = u < 0.4
    maskb = (u >= 0.4) & (u < 0.7)
    tbl = torch.from_numpy(g16.fr_array(list(range(256))).view(np.int64)).to(dev)
    wires[mask01] = tbl[(smallv[mask01] & 1)]
    wires[maskb] = tbl[smallv[maskb]]
    a, b = rnd((Np, 4)), rnd((Np, 4))
    # c = a*b pointwise is not needed for timing (any a,b,c give the same work); keep c random
    c = rnd((Np, 4))
    rr, ss = rand_fr(rs, 1)[0], rand_fr(rs, 1)[0]

    def run():
        aa, bb, cc = a.clone(), b.clone(), c.clone()
        ctx.prove_dev(pk, wires.data_ptr(), aa.data_ptr(), bb.data_ptr(), cc.data_ptr(), rr, ss)
    run()
    best, ph = 1e30, None
    for _ in range(3):
        aa, bb, cc = a.clone(), b.clone(), c.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.prove_dev(pk, wires.data_ptr(), aa.data_ptr(), bb.data_ptr(), cc.data_ptr(), rr, ss)
        dt = (time.perf_counter() - t0) * 1e3
        if dt < best:
            best, ph = dt, ctx.last_timings()
    ex["groth16_prove_synthetic"] = {
        "log2_constraints": Lp, "wires": Np, "witness_mix": "40% 0/1, 30% bytes, 30% uniform (SURVEY §8d config 1)",
        "ms": round(best, 3), "window_tables": bool(args.table),
        "phases_ms[h2d,gather,msmB2,msmB1,msmA,msmK,computeH,msmZ]": [round(x, 3) for x in (ph or [])],
        "note": "inputs resident in HBM; wall-clock around b200g16_prove_dev incl. host finish"}
    ctx.pk_free(pk)
    for v in vecs + [b2]:
        v.free()
    if not args.no_cpu_baseline:
        ex["cpu_port_same_box"] = cpu_extras()
    return ex


def cpu_extras():
    """The C port of the other kernels of the metric on this box's host cores (bounded samples; the
    checker, timed only as the reported CPU baseline)."""
    from oracle import cport
    rs = np.random.Generator(np.random.PCG64(SEED + 99))
    nt = host_threads()
    out = {"threads": nt, "kind": "port (oracle/c/oracle.c, OpenMP), not gnark"}

    def best_of(fn, reps=2):
        fn()
        return min(_timed(fn) for _ in range(reps))

    def _timed(fn):
        t0 = time.perf_counter()
        fn()
        return (time.perf_counter() - t0) * 1e3
    a = rand_fr(rs, 1 << 20)
    out["ntt_2^20_ms"] = round(best_of(lambda: cport.ntt(a, nthreads=nt)), 2)
    a, b, c = rand_fr(rs, 1 << 18), rand_fr(rs, 1 << 18), rand_fr(rs, 1 << 18)
    out["compute_h_2^18_ms"] = round(best_of(lambda: cport.compute_h(a, b, c, 18, nthreads=nt)), 2)
    st = rs.integers(0, 1 << 63, size=(1 << 18, 25), dtype=np.uint64)
    ms = best_of(lambda: cport.keccak_f_batch(st, nthreads=nt))
    out["keccak_f_2^18_states_ms"] = round(ms, 2)
    out["keccak_Mperm_s"] = round((1 << 18) / ms / 1e3, 2)
    return out


def run_prove_workload(args, rank, local_rank, world):
    """--workload prove (BASELINE.json configs[4]): synthetic WHIR-verifier-shaped Groth16 prove with
    N = 2^prove_logn constraints and as many wires, proving key sharded by point range over the
    ranks.  Every rank runs computeH (replicated) and the five MSMs on its shard; the five partial
    points are all-gathered over NCCL and the proof is finished on the host."""
    import torch
    import torch.distributed as dist

    from gnark_whir_b200 import groth16 as g16
    from gnark_whir_b200 import lib, sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = lib.Context(local_rank)
    L = args.prove_logn
    N = 1 << L
    rs = np.random.Generator(np.random.PCG64(SEED + 31 * rank))
    g1, g2 = g16.g1_point(g16.G1_GEN), g16.g2_point(g16.G2_GEN)
    lens = {"a": N, "b": N, "k": N - 1, "z": N - 1}
    # computeH for N > 1: spread over three ranks (default), the same overlapped with the other ranks' witness
    # MSMs and with weighted shards (--overlap-h; measured no better: NCCL's kernels do not get SMs under the
    # saturating accumulate kernel), or replicated (--replicated-h)
    distributed_h = world > 1 and not args.replicated_h
    overlap_h = distributed_h and args.overlap_h
    weights = sharded.prove_weights(world, args.h_share) if overlap_h else [1.0] * world
    spans = {k: (sharded.shard_range(v, rank, world) if k == "z" else sharded.shard_range_weighted(v, rank, weights))
             for k, v in lens.items()}
    vec = {k: ctx.fixed_base_mul(g1, rand_fr(rs, hi - lo), group=1, resident=True) for k, (lo, hi) in spans.items()}
    b2 = ctx.fixed_base_mul(g2, rand_fr(rs, spans["b"][1] - spans["b"][0]), group=2, resident=True)
    if args.table:
        for v in list(vec.values()) + [b2]:
            v.precompute(0)
    common = np.random.Generator(np.random.PCG64(SEED))          # identical on every rank
    small = ctx.fixed_base_mul(g1, rand_fr(common, 3), group=1)
    small2 = ctx.fixed_base_mul(g2, rand_fr(common, 2), group=2)
    k_skip = np.zeros(N, dtype=np.uint8)
    k_skip[0] = 1
    pk = ctx.pk_upload(L, N, vec["a"], vec["b"], vec["k"], vec["z"], b2, small[0], small[1], small[2], small2[0],
                       small2[1], np.zeros(N, np.uint8), np.zeros(N, np.uint8), k_skip, partial=world > 1,
                       offsets=(spans["a"][0], spans["b"][0], spans["k"][0], spans["z"][0]))
    gen = torch.Generator(device=dev).manual_seed(SEED)

    def rnd(shape):
        t = torch.randint(0, 1 << 62, shape, dtype=torch.int64, device=dev, generator=gen)
        t[..., 3] &= (1 << 60) - 1
        return t
    u = torch.rand(N, device=dev, generator=gen)
    wires = rnd((N, 4))
    smallv = torch.randint(0, 256, (N,), dtype=torch.int64, device=dev, generator=gen)
    tbl = torch.from_numpy(g16.fr_array(list(range(256))).view(np.int64)).to(dev)
    m01, mb = u < 0.4, (u >= 0.4) & (u < 0.7)
    wires[m01] = tbl[smallv[m01] & 1]
    wires[mb] = tbl[smallv[mb]]
    a, b, c = rnd((N, 4)), rnd((N, 4)), rnd((N, 4))
    rr, ss = rand_fr(common, 1)[0], rand_fr(common, 1)[0]

    def step():
        aa, bb, cc = a.clone(), b.clone(), c.clone()        # computeH works in place
        if overlap_h:       # a, b, c transformed on three ranks under the other ranks' witness MSMs, broadcast, finished everywhere
            torch.cuda.synchronize()                        # clones (torch stream) before the library's stream
            return sharded.prove_distributed(ctx, pk, wires, aa, bb, cc, L, rr, ss, device=dev)
        if distributed_h:   # a, b, c transformed on three ranks, broadcast, finished everywhere, then the five MSMs
            torch.cuda.synchronize()
            sharded.compute_h_distributed(ctx, aa, bb, cc, L)
            part = ctx.prove_h_dev(pk, wires.data_ptr(), aa.data_ptr(), rr, ss)
        else:
            part = ctx.prove_dev(pk, wires.data_ptr(), aa.data_ptr(), bb.data_ptr(), cc.data_ptr(), rr, ss)
        if world == 1:
            return part
        t = torch.from_numpy(sharded.pack_partials(part).view(np.int64).copy()).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        sums = sharded.sum_partials([o.cpu().numpy().view(np.uint64) for o in outs])
        return ctx.prove_finish(pk, sums["msm_a"], sums["msm_b1"], sums["msm_k"], sums["msm_z"], sums["msm_b2"], rr, ss)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    for _ in range(args.warmup):
        step()
    sync_all()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count() - l0
    ph = ctx.last_timings()
    e2e = None
    if world == 1:
        # end to end through b200g16_prove: witness and a, b, c in pinned HOST memory, H2D inside the timed region
        host = [t.cpu().pin_memory() for t in (wires, a, b, c)]
        hv = [t.numpy().view(np.uint64) for t in host]
        ctx.prove(pk, hv[0], hv[1], hv[2], hv[3], rr, ss)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res_h, _ = ctx.prove(pk, hv[0], hv[1], hv[2], hv[3], rr, ss)
        ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
        if not np.array_equal(res_h["krs"], res["krs"]):
            raise SystemExit("bench: host-path and device-path proofs differ")
        e2e = {"value": round(ms_e2e, 3), "unit": "ms", "h2d_bytes_per_step": int(4 * N * 32),
               "d2h_bytes_per_step": int(5 * 256 + 64)}
    if rank == 0:
        print(json.dumps({
            "metric": "WHIR-verifier-shaped Groth16 prove latency", "value": round(ms, 3), "unit": "ms",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"synthetic Groth16 prove, 2^{L} constraints, {N} wires, witness 40% 0/1 / 30% bytes / "
                                   f"30% uniform (SURVEY §8d config 1/5), inputs resident in HBM",
                       "window_tables": bool(args.table),
                       "parallelism": f"pk point-range shards x{world}; computeH " +
                                      ("replicated" if (world == 1 or args.replicated_h) else
                                       "spread over 3 ranks (a, b, c) + 3 NCCL broadcasts" +
                                       (f", under the other ranks' witness MSMs; witness-MSM shard weights "
                                        f"{[round(w, 2) for w in weights]}" if overlap_h else "")) +
                                      "; all_gather of 5 partial points"},
            "e2e": e2e,
            "gpu_launches": launches,
            "phases_ms_rank0[h2d,gather,msmB2,msmB1,msmA,msmK,computeH,msmZ]": [round(x, 3) for x in ph],
            "proof_krs_limb0": int(res["krs"][0])}))
    ctx.pk_free(pk)
    for v in list(vec.values()) + [b2]:
        v.free()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workload", default="msm", choices=["msm", "prove"],
                    help="msm = the headline (BASELINE configs[1]); prove = sharded synthetic Groth16 prove (configs[4])")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--logn", type=int, default=24, help="log2 of the MSM size (whole job)")
    ap.add_argument("--cpu-logn", type=int, default=20, help="log2 of the bounded CPU sample")
    ap.add_argument("--ntt-logn", type=int, default=24)
    ap.add_argument("--prove-logn", type=int, default=20)
    ap.add_argument("--no-table", dest="table", action="store_false",
                    help="run the MSM without the window table over the resident bases (b200g16_bases_precompute)")
    ap.add_argument("--overlap-h", action="store_true",
                    help="prove workload, N>1: sharded.prove_distributed (computeH stages overlapped with the witness MSMs)")
    ap.add_argument("--h-share", type=float, default=None,
                    help="prove workload, N>=3: relative witness-MSM shard size of the three ranks that also transform a, b, c")
    ap.add_argument("--replicated-h", action="store_true",
                    help="prove workload, N>1: every rank runs the whole computeH instead of spreading it over 3 ranks")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-traffic", type=float, default=None,
                    help="dram bytes per k_accumulate launch from profiles/ (ncu --set full), if captured")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if args.workload == "prove":
        run_prove_workload(args, rank, local_rank, world)
        return
    ctx, out, st = run_b200(args, rank, local_rank, world)
    if rank == 0:
        if world == 1 and not args.no_extras:
            try:
                out["extras"] = extras_single_gpu(ctx, st, args)
            except Exception as e:          # extras must never cost the headline line
                out["extras"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            n_s, times, cores, ok = cpu_msm_sample(min(args.logn, args.cpu_logn), 2, 1)
            v = n_s / (min(times) * 1e6)
            out["cpu_baseline"] = {"value": round(v, 4), "unit": "Mpts/s", "cores": cores, "kind": "port",
                                   "sample": f"C port of gnark-crypto MultiExp (oracle/c/oracle.c), G1 MSM of 2^{min(args.logn, args.cpu_logn)} "
                                             f"points, uniform scalars, {cores} OpenMP threads, best of 2; result_checked={ok}"}
        print(json.dumps(out))
    st["bases"].free()
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
