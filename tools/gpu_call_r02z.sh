# gpurun -- 'bash tools/gpu_call_r02z.sh r02z': artefacts of the final state — full GPU suite, smoke, default bench line (both arms),
# ncu launch list of the bench command, --set full capture of the dominant kernel (+ traffic artefact), kernel families
TAG=${1:-r02z}
set -x
(time python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1); tail -4 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
(time python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err); tail -c 300 gpurun_out/${TAG}_bench_default.err
(time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err); tail -c 400 gpurun_out/${TAG}_bench_reference.json
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench_2p24.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 0 -c 1 -f -o /tmp/${TAG}_accumulate_2p24 $CMD > gpurun_out/${TAG}_ncu_acc.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_accumulate_2p24.ncu-rep > gpurun_out/${TAG}_ncu_k_accumulate_2p24_summary.txt
python tools/ncu_traffic.py /tmp/${TAG}_accumulate_2p24.ncu-rep 'k_accumulate<Fp>' 24 1 > gpurun_out/${TAG}_ncu_traffic_k_accumulate_2p24.json; cat gpurun_out/${TAG}_ncu_traffic_k_accumulate_2p24.json
python tools/profile_families.py > gpurun_out/${TAG}_plain_families.log 2>&1 && ncu --set full --clock-control none -k regex:'k_accumulate|k_ntt_pass|k_reduce_first|k_reduce_level|k_reduce_tail|k_scatter|k_digits|k_merkle|k_keccak_f|k_h_pointwise|k_merge|k_task' -c 70 -f -o /tmp/${TAG}_families python tools/profile_families.py > gpurun_out/${TAG}_ncu_fam.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_families.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt
du -sh gpurun_out
