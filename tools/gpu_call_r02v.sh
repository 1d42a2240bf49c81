# gpurun -- 'bash tools/gpu_call_r02v.sh r02v': prove suites after the early host assembly (per-slot events), prove 2^20 line
TAG=${1:-r02v}
set -x
(time python -m pytest tests/test_gpu_prove.py tests/test_gpu_bench_sizes.py tests/test_gpu_group.py tests/test_gpu_verify.py -m gpu -x -q > gpurun_out/${TAG}_pytest_prove.log 2>&1); tail -3 gpurun_out/${TAG}_pytest_prove.log
python bench.py --workload prove --prove-logn 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prove20.json 2> gpurun_out/${TAG}_prove20.err; tail -c 500 gpurun_out/${TAG}_prove20.json
python bench.py --workload prove --prove-logn 18 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prove18.json 2> gpurun_out/${TAG}_prove18.err; tail -c 300 gpurun_out/${TAG}_prove18.json
