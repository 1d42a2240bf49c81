# gpurun -- 'bash tools/gpu_call_r02p.sh r02p': reduce A/B (levels / levels + tail expanded in place), launch list of a 2^20 prove
TAG=${1:-r02p}
set -x
for v in 1 2; do B200G16_REDUCE_INLINE=$v python tools/sweep.py --reduce-ab >> gpurun_out/${TAG}_reduce_ab.jsonl 2>> gpurun_out/${TAG}_reduce_ab.err; done
cat gpurun_out/${TAG}_reduce_ab.jsonl
CMD="python bench.py --workload prove --prove-logn 20 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_prove20_plain.json 2> gpurun_out/${TAG}_prove20_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_ncu_launches_prove_2p20.csv $CMD > gpurun_out/${TAG}_ncu_prove.log 2>&1
tail -c 600 gpurun_out/${TAG}_prove20_plain.json; wc -l gpurun_out/${TAG}_ncu_launches_prove_2p20.csv
