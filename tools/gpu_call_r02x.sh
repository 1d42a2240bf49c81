# gpurun -- 'bash tools/gpu_call_r02x.sh r02x': suites after the asynchronous PoK MSM, prove 2^20 / 2^18 lines
TAG=${1:-r02x}
set -x
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_bench_sizes.py tests/test_gpu_group.py tests/test_gpu_verify.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1); tail -5 gpurun_out/${TAG}_pytest.log
python bench.py --workload prove --prove-logn 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prove20.json 2> gpurun_out/${TAG}_prove20.err; tail -c 400 gpurun_out/${TAG}_prove20.json; tail -3 gpurun_out/${TAG}_prove20.err
python examples/prove_verify.py > gpurun_out/${TAG}_example.log 2>&1; tail -3 gpurun_out/${TAG}_example.log
