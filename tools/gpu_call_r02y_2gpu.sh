# gpurun --gpus 2 -- 'bash tools/gpu_call_r02y_2gpu.sh r02y': the driver's 2-GPU command (MSM line + sharded 2^24 prove in extras),
# group tests on two distinct devices
TAG=${1:-r02y}
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; tail -c 1500 gpurun_out/${TAG}_bench_2gpu.json; tail -5 gpurun_out/${TAG}_bench_2gpu.err
(time python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/${TAG}_pytest_group_2gpu.log 2>&1); tail -3 gpurun_out/${TAG}_pytest_group_2gpu.log
