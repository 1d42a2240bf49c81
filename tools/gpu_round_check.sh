set -x
(time python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_group.py tests/test_gpu_bench_sizes.py -m gpu -x -q -k "not ntt_full and not compute_h_2p22" > gpurun_out/r02g_pytest.log 2>&1); tail -6 gpurun_out/r02g_pytest.log
python tools/sweep.py > gpurun_out/r02g_sweep.jsonl 2> gpurun_out/r02g_sweep.err; grep -c . gpurun_out/r02g_sweep.jsonl
(time python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err); tail -c 300 gpurun_out/r02g_bench_default.err
python bench.py --steps 5 --warmup 3 --no-table --no-extras > gpurun_out/r02g_bench_notable.json 2>> gpurun_out/r02g_bench_default.err
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/r02g_plain_bench.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 0 -c 1 -f -o /tmp/r02g_accumulate_2p24 $CMD > gpurun_out/r02g_ncu_acc.log 2>&1
python tools/ncu_summary.py /tmp/r02g_accumulate_2p24.ncu-rep > gpurun_out/r02g_ncu_k_accumulate_2p24_summary.txt
python tools/ncu_traffic.py /tmp/r02g_accumulate_2p24.ncu-rep 'k_accumulate<Fp>' 24 1 > gpurun_out/r02g_ncu_traffic_k_accumulate_2p24.json
cat gpurun_out/r02g_ncu_traffic_k_accumulate_2p24.json
