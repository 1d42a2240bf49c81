# gpurun -- 'bash tools/gpu_call_r02w.sh r02w': CTA size of the G2 bucket-reduction levels (2^20, c = 20 table)
TAG=${1:-r02w}
set -x
for b in 128 64 32 96; do B200G16_G2_REDUCE_BLOCK=$b python tools/sweep.py --reduce-ab --g2 --logs=20 >> gpurun_out/${TAG}_g2_reduce_block.jsonl 2>> gpurun_out/${TAG}.err; done
cat gpurun_out/${TAG}_g2_reduce_block.jsonl
