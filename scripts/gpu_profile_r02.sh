#!/bin/bash
# Round-2 evidence pass (1 GPU): parity tests, smoke, the default bench line + reference arm + Zipf line (none under a
# profiler), then ncu: launch list of the step, `--set full` of the probe and pull kernels, and the tier kernels (quad pull
# from a local world-1 tier, tier gather of a 125 M-row device table).  usage: bash scripts/gpu_profile_r02.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
timeout 300 python bench.py --value-only --zipf 1.05 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_zipf105.json 2> gpurun_out/bench_${TAG}_zipf105.err
timeout 300 python bench.py --value-only --local-tier --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_local_tier.json 2> gpurun_out/bench_${TAG}_local_tier.err
: > gpurun_out/sweep_${TAG}_quad_pcie.jsonl
for ctas in 37 74 148; do
  echo "{\"cfg\": \"quad pull over PCIe, $ctas CTAs\"}" >> gpurun_out/sweep_${TAG}_quad_pcie.jsonl
  timeout 200 python bench.py --value-only --debug-flags 8 --pull-ctas $ctas --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/sweep_${TAG}_quad_pcie.jsonl 2>> gpurun_out/sweep_${TAG}_quad_pcie.err
done
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline --core-arms-only --skip-extra > gpurun_out/ncu_bench.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"probe_gather_v8|pull_binned" -s 10 -c 4 -o gpurun_out/hot_${TAG} -f \
  python bench.py --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline --core-arms-only --skip-extra > gpurun_out/ncu_full.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/launches_${TAG}_tier.csv \
  python bench.py --value-only --local-tier --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline > gpurun_out/ncu_tier.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"pull_binned" -s 5 -c 1 -o gpurun_out/hot_${TAG}_tier -f \
  python bench.py --value-only --local-tier --steps 2 --warmup 3 --prefill 2 --no-cpu-baseline > gpurun_out/ncu_full_tier.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"tier_gather" -s 4 -c 1 -o gpurun_out/hot_${TAG}_gather -f \
  python scripts/c4_repro.py > gpurun_out/ncu_full_gather.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified_rows')}, 'e2e', d['e2e']['value'], 'host_out', (d.get('e2e_host_output') or {}).get('value'))
for k in ('c1','c5','c3'):
    print(k, {kk:vv for kk,vv in (d.get(k) or {}).items() if kk in ('value','ms_per_step','hit_rate_measured','error','us_per_request')})
print('miss dup', d['config'].get('miss_duplicates'))
PY
cut -c1-300 gpurun_out/bench_${TAG}_zipf105.json; cut -c1-300 gpurun_out/bench_${TAG}_local_tier.json
cut -c1-260 gpurun_out/sweep_${TAG}_quad_pcie.jsonl
ls -la gpurun_out/*.ncu-rep | tail -4
