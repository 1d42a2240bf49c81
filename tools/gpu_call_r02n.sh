# One gpurun call of round 2 (session 3): full GPU suite at HEAD, latency-kernel sweep of the Merkle recompute,
# table-window sweep on witness-shaped scalars, default bench line.  Usage: gpurun -- 'bash tools/gpu_call_r02n.sh r02n'
TAG=${1:-r02n}
set -x
(time python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1); tail -4 gpurun_out/${TAG}_pytest_gpu.log
python tools/sweep.py --merkle > gpurun_out/${TAG}_merkle_warp_sweep.jsonl 2> gpurun_out/${TAG}_merkle.err; grep -c . gpurun_out/${TAG}_merkle_warp_sweep.jsonl
python tools/sweep.py --windows-mix --logs=18,20,22 > gpurun_out/${TAG}_window_mix_sweep.jsonl 2> gpurun_out/${TAG}_window_mix.err; grep -c . gpurun_out/${TAG}_window_mix_sweep.jsonl
(time python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err); tail -c 300 gpurun_out/${TAG}_bench_default.err
