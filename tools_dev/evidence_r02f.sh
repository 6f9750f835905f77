# S3 shape at E = 256 (tables 2 x 10.24 GB, HBM-resident): the wide-row window kernel against sgns_fast_kernel
set -x
mkdir -p gpurun_out
for k in window context; do
  timeout 70 python bench.py --emb 256 --steps 5 --warmup 3 --no-cpu-baseline --kernel $k > gpurun_out/r02f_bench_s3_e256_$k.json 2> gpurun_out/r02f_bench_s3_e256_$k.err
  python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_s3_e256_$k.json')); print('$k', d['value'], d['kernel_ms'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['survey_unit'], d['train_stats'], d['clocks'])"
done
