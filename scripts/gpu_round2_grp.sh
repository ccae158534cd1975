set -x
timeout 600 python -m pytest tests/test_gpu_properties.py -q -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_groups.json 2> gpurun_out/r02_bench_groups.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_groups.json').read().strip().splitlines()[-1]); print('GROUPS', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['column_groups'], d['gpu_launches']); print([(k['kernel'][:30], round(k['ms'],3), round(k['frac'],3)) for k in d['hbm_kernels']['kernels']]); print(d['roofline'])"
tail -3 gpurun_out/r02_bench_groups.err
timeout 300 python bench.py --steps 10 --warmup 3 --columns 512 --no-cpu-baseline > gpurun_out/r02_bench_512.json 2> gpurun_out/r02_bench_512.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_512.json').read().strip().splitlines()[-1]); print('C512', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['column_groups'])"
