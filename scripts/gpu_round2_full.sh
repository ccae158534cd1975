set -x
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_gputests_full.log; tail -12 gpurun_out/r02_gputests_full.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1]); print('FINAL', d['value'], d['ms_per_step'], d['e2e']['value']); print([(k['kernel'][:40], round(k['ms'],3)) for k in d['hbm_kernels']['kernels']]); print({k: d[k] for k in d if k not in ('hbm_kernels','config','roofline')})"
tail -3 gpurun_out/r02_bench_final.err
