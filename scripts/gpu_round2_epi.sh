set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_lockstep.py -q -p no:cacheprovider -x 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_epi.json 2> gpurun_out/r02_bench_epi.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_epi.json').read().strip().splitlines()[-1]); print('EPI', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['column_groups'], d['gpu_launches']); print(d['roofline']['frac'], d['roofline']['kernel_ms'])"
tail -3 gpurun_out/r02_bench_epi.err
