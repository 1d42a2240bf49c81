# gpurun -- 'bash tools/gpu_call_r02t.sh r02t': full GPU suite, config sweeps, default bench line (state after the task-order /
# reduce / merge / Merkle changes)
TAG=${1:-r02t}
set -x
(time python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1); tail -4 gpurun_out/${TAG}_pytest_gpu.log
python tools/sweep.py > gpurun_out/${TAG}_sweep_configs.jsonl 2> gpurun_out/${TAG}_sweep.err; grep -c . gpurun_out/${TAG}_sweep_configs.jsonl
(time python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err); tail -c 300 gpurun_out/${TAG}_bench_default.err
