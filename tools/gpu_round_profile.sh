# One gpurun call: config sweeps, window sweep, default bench line, ncu launch list + --set full captures (summarised on the
# box: the reports themselves are too large to bring back).  Usage: gpurun -- 'bash tools/gpu_round_profile.sh r02f'
TAG=${1:-r02}
set -x
python tools/sweep.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err; grep -c . gpurun_out/${TAG}_sweep.jsonl
python tools/sweep.py --windows > gpurun_out/${TAG}_window_sweep.jsonl 2>> gpurun_out/${TAG}_sweep.err; grep -c . gpurun_out/${TAG}_window_sweep.jsonl
(time python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err); tail -c 300 gpurun_out/${TAG}_bench_default.err
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench_2p24.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 0 -c 1 -f -o gpurun_out/${TAG}_accumulate_2p24 $CMD > gpurun_out/${TAG}_ncu_acc.log 2>&1
python tools/profile_families.py > gpurun_out/${TAG}_plain_families.log 2>&1 && ncu --set full --clock-control none -k regex:'k_accumulate|k_ntt_pass|k_reduce_first|k_reduce_level|k_scatter|k_digits|k_merkle|k_keccak_f|k_h_pointwise|k_merge' -c 60 -f -o /tmp/${TAG}_families python tools/profile_families.py > gpurun_out/${TAG}_ncu_fam.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_families.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt
python tools/ncu_summary.py gpurun_out/${TAG}_accumulate_2p24.ncu-rep > gpurun_out/${TAG}_ncu_k_accumulate_2p24_summary.txt
du -sh gpurun_out
