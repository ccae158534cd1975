set -x
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_ensemble_quick.py tests/test_gpu_device_steady.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1gpu_fused.json 2> gpurun_out/r02_bench_1gpu_fused.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu_fused.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['roofline']); print([(k['kernel'][:30], round(k['ms'],3), round(k['frac'],3)) for k in d['hbm_kernels']['kernels']])
print(d['single_column']['steps_per_s'], d['single_column']['time_to_steady_state'].get('wall_s'))
PY
tail -3 gpurun_out/r02_bench_1gpu_fused.err
VK_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VK_FUSED=0', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
PYTHONHASHSEED=0 timeout 300 python oracle/dropin_in_reference.py --config HD189 --refdir oracle/_ref/HD189 --cuda --steady --out gpurun_out/r02_seam.jsonl 2>&1 | tail -1
PYTHONHASHSEED=0 timeout 300 python oracle/dropin_in_reference.py --config HD189 --refdir oracle/_ref/HD189 --cuda --steady --device-loop --out gpurun_out/r02_seam.jsonl 2>&1 | tail -1
