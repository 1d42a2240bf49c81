# gpurun -- 'bash tools/gpu_call_r02s.sh r02s': per-operation event trace of the sort phase (2^20 and 2^24, c = 20 table)
TAG=${1:-r02s}
set -x
B200G16_SORT_TRACE=1 python tools/sweep.py --reduce-ab --logs=20,24 > gpurun_out/${TAG}_sort_trace.jsonl 2> gpurun_out/${TAG}_sort_trace.err
grep "sort trace" gpurun_out/${TAG}_sort_trace.err | tail -4
