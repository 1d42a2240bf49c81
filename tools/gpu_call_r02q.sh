# gpurun -- 'bash tools/gpu_call_r02q.sh r02q': MSM / prove suites after the CTA-local task histograms, phases, prove 2^20 line
TAG=${1:-r02q}
set -x
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_bench_sizes.py -m gpu -x -q > gpurun_out/${TAG}_pytest_msm_prove.log 2>&1); tail -3 gpurun_out/${TAG}_pytest_msm_prove.log
python tools/sweep.py --reduce-ab > gpurun_out/${TAG}_msm_phases.jsonl 2> gpurun_out/${TAG}_msm_phases.err; cat gpurun_out/${TAG}_msm_phases.jsonl
python bench.py --workload prove --prove-logn 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prove20.json 2> gpurun_out/${TAG}_prove20.err; tail -c 400 gpurun_out/${TAG}_prove20.json
