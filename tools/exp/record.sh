set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.txt | tail -1 > gpurun_out/bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
python tools/bench_sweep.py --repeats 30 2>/dev/null | tail -1 > gpurun_out/sweep.json
python tools/exp/time_compact.py > gpurun_out/time_compact.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e']); print({k:v['value'] for k,v in d.get('e2e_variants',{}).items()}); print(d['cpu_baseline'])
for e in d.get('extra',[]): print(json.dumps(e)[:300])
"
cat gpurun_out/time_compact.txt; cut -c1-300 gpurun_out/bench_ref.json
if [ -n "$RECORD_NCU_FULL" ]; then
# ncu --set full digests of the kernels whose code changed late in the round (read with tools/ncu_summary.py on the CPU box)
ncu --set full --clock-control none --import-source on -k regex:"blur_masked_kernel" -s 6 -c 1 -o gpurun_out/prof_r2c_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-overlap > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blur_masked_kernel" -s 6 -c 1 -o gpurun_out/prof_r2c_cfg2h python bench.py --workload cfg2h --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-overlap > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"compact_taps" -s 3 -c 1 -o gpurun_out/prof_r2c_compact python tools/exp/time_compact.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
fi
