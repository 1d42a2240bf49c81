# gpurun -- 'bash tools/gpu_call_r02o.sh r02o': MSM / Keccak suites on the new kernels, reduce A/B, Merkle latency sweep
TAG=${1:-r02o}
set -x
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_keccak.py tests/test_gpu_prove.py -m gpu -x -q > gpurun_out/${TAG}_pytest_msm_keccak_prove.log 2>&1); tail -3 gpurun_out/${TAG}_pytest_msm_keccak_prove.log
B200G16_REDUCE_INLINE=0 python tools/sweep.py --reduce-ab > gpurun_out/${TAG}_reduce_ab.jsonl 2> gpurun_out/${TAG}_reduce_ab.err
B200G16_REDUCE_INLINE=1 python tools/sweep.py --reduce-ab >> gpurun_out/${TAG}_reduce_ab.jsonl 2>> gpurun_out/${TAG}_reduce_ab.err
cat gpurun_out/${TAG}_reduce_ab.jsonl
python tools/sweep.py --merkle > gpurun_out/${TAG}_merkle_warp_sweep.jsonl 2> gpurun_out/${TAG}_merkle.err; grep -c . gpurun_out/${TAG}_merkle_warp_sweep.jsonl
