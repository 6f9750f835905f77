# round 2, last GPU call: full GPU test suite on the final tree, default bench line, wide-row window kernel (E = 256) against the kernel it
# replaces, bare row-mix ceiling of the S3 access pattern, smoke, one full ncu capture of the wide kernel
set -x
mkdir -p gpurun_out
timeout 330 python -m pytest tests -q -m gpu > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02e_pytest_gpu.log
timeout 150 python bench.py > gpurun_out/r02e_bench_default.json 2> gpurun_out/r02e_bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_default.json')); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
for k in window context; do
  timeout 90 python bench.py --workload s4 --emb 256 --kernel $k --steps 20 --warmup 3 > gpurun_out/r02e_bench_s4_e256_$k.json 2> gpurun_out/r02e_bench_s4_e256_$k.err
  python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_s4_e256_$k.json')); print('$k', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['train_stats'])"
done
timeout 60 ./tools_dev/micro/rowbw mix > gpurun_out/r02e_rowmix.txt 2>&1; cat gpurun_out/r02e_rowmix.txt
timeout 90 python bench.py --workload s4 --emb 256 --s4-power 0 --steps 20 --warmup 3 > gpurun_out/r02e_bench_s4_e256_uniform.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_s4_e256_uniform.json')); print('uniform', d['value'], d['roofline']['frac'])"
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02e_smoke.log 2>&1; tail -2 gpurun_out/r02e_smoke.log
timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sgns_winw -c 1 -f -o gpurun_out/r02e_s4_e256_winw python bench.py --workload s4 --emb 256 --steps 1 --warmup 3 > gpurun_out/r02e_ncu_full.log 2>&1; tail -2 gpurun_out/r02e_ncu_full.log
ls -la gpurun_out | tail -15
