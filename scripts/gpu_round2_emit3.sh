set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "emitted_batch or batched" 2>&1 | tail -25
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_emitjac.json 2> gpurun_out/r02_bench_emitjac.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_emitjac.json').read().strip().splitlines()[-1]); print('EMITJAC', d['value'], d['ms_per_step'], d['e2e']['value']); print([(k['kernel'][:40], round(k['ms'],3)) for k in d['hbm_kernels']['kernels']]); print(d['roofline'])"
tail -3 gpurun_out/r02_bench_emitjac.err
