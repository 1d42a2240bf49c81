set -x
python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r01b_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r01b_bench_table.json 2> gpurun_out/r01b_bench_table.err
python bench.py --steps 5 --warmup 3 --no-table --no-extras --no-cpu-baseline > gpurun_out/r01b_bench_notable.json 2> gpurun_out/r01b_bench_notable.err
tail -5 gpurun_out/r01b_tests.log; cat gpurun_out/r01b_bench_table.json; tail -3 gpurun_out/r01b_bench_table.err
