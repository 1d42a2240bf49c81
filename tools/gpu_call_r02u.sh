# gpurun -- 'bash tools/gpu_call_r02u.sh r02u': ranked sort (rank from the histogram atomic, scatter without atomics) against
# the two-atomic form: phases + per-operation trace at 2^20 / 2^22 / 2^24, then the MSM / prove suites on the ranked form
TAG=${1:-r02u}
set -x
for v in 0 1; do
  B200G16_RANKED_SORT=$v python tools/sweep.py --reduce-ab --logs=20,22,24 | sed "s/\"msm_reduce_ab\"/\"msm_ranked_sort_ab\", \"ranked_sort\": $v/" >> gpurun_out/${TAG}_ranked_sort_ab.jsonl 2>> gpurun_out/${TAG}_ab.err
  B200G16_RANKED_SORT=$v B200G16_SORT_TRACE=1 python tools/sweep.py --reduce-ab --logs=20,24 2>&1 >/dev/null | grep "sort trace" | sort | uniq -c | sort -rn | sed "s/^/ranked_sort=$v /" | awk 'NR<=40' >> gpurun_out/${TAG}_sort_trace.txt
done
cat gpurun_out/${TAG}_ranked_sort_ab.jsonl; tail -n 3 gpurun_out/${TAG}_sort_trace.txt
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_bench_sizes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/${TAG}_pytest_msm_prove.log 2>&1); tail -3 gpurun_out/${TAG}_pytest_msm_prove.log
