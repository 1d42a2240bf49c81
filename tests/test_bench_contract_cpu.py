"""bench.py's CPU arm (--impl reference) runs here without a GPU and prints the contract's JSON line; the
committed GPU lines under profiles/ carry every key the contract asks for."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports; the arm must not fall to one core
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-logn", "12"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS | {"impl", "cpu_baseline"} <= set(line)
    assert line["impl"] == "reference" and line["unit"] == "Mpts/s" and line["higher_is_better"] is True
    assert line["config"]["result_checked"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] >= 1
    try:
        assert cb["cores"] == len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    assert line["e2e"] == {"value": line["value"], "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_gpu_bench_lines_carry_every_contract_key():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_bench_default.json")))
    assert files, "no committed bench line under profiles/"
    d = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["gpu_launches"] > 0 and d["config"]["workload"] and d["config"]["result_checked_vs_oracle"] is True
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    # achieved counts the ALGORITHMIC 1360 multiply-adds per mixed add; the kernel executes fewer (dedicated square,
    # fused two-term product), so the fraction may exceed 1; frac_executed is the pipe utilisation
    assert 0 < d["roofline"]["frac"] <= 1.15 and 0 < d["roofline"].get("frac_executed", 0.5) <= 1.02
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if "r02" in os.path.basename(files[-1]):       # round-2 additions to the line
        assert d["cpu_baseline"]["sample"].count("2^24") and "e2e_pageable" in d and "peak_raw_imad_wide" in d["roofline"]
        proves = d["extras"]["groth16_prove"]
        assert [p["log2_constraints"] for p in proves] == [20, 24]
        for p in proves:
            assert p["result_checked_vs_oracle"] is True and p["e2e"]["h2d_bytes_per_step"] > 0 and p["cpu_baseline"]["kind"] == "port"


def test_reference_arm_for_the_prove_metric():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "prove",
                          "--prove-logn", "8", "--cpu-prove-logn", "8", "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False
    assert line["config"]["same_size_as_gpu_arm"] is True and line["cpu_baseline"]["kind"] == "port"
