set -x
python -m pytest tests -q -m gpu -x > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02c_pytest_gpu.log
python bench.py --impl reference > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err; tail -c 400 gpurun_out/r02c_bench_reference.json
python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_default.json')); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02c_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sgns_win_kernel -c 1 -f -o gpurun_out/r02c_s3_sgns python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_ncu_full.log 2>&1; tail -2 gpurun_out/r02c_ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:sgns_owned_pairs -c 1 -f -o gpurun_out/r02c_owned_pairs python tools_dev/negown_probe.py --iters 1 --modes all-pairs > gpurun_out/r02c_ncu_full2.log 2>&1; tail -2 gpurun_out/r02c_ncu_full2.log
ls -la gpurun_out/*.ncu-rep | tail -3
