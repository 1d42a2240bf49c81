# gpurun -- 'bash tools/gpu_call_r02r.sh r02r': launch list of 2^20 G1 MSMs (c = 20 table, uniform scalars)
TAG=${1:-r02r}
set -x
CMD="python tools/sweep.py --reduce-ab --logs=20"
$CMD > gpurun_out/${TAG}_plain.jsonl 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_ncu_launches_msm_2p20.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_plain.jsonl; wc -l gpurun_out/${TAG}_ncu_launches_msm_2p20.csv
