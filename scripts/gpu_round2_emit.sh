set -x
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_gputests_emit.log; tail -15 gpurun_out/r02_gputests_emit.log
timeout 300 python -m pytest tests/test_gpu_device_steady.py -q -s -k "condensing or jupiter" -p no:cacheprovider 2>&1 | grep -v "^Include\|^$" | cut -c1-700 | tail -12 > gpurun_out/r02_conden_loop.log; cat gpurun_out/r02_conden_loop.log
VK_EMIT=0 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_noemit.json 2> gpurun_out/r02_bench_noemit.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_noemit.json').read().strip().splitlines()[-1]); print('NOEMIT', d['value'], d['ms_per_step'], d['e2e']['value']); print([(k['kernel'][:40], round(k['ms'],3)) for k in d['hbm_kernels']['kernels']])"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_emit.json 2> gpurun_out/r02_bench_emit.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_emit.json').read().strip().splitlines()[-1]); print('EMIT', d['value'], d['ms_per_step'], d['e2e']['value']); print([(k['kernel'][:40], round(k['ms'],3)) for k in d['hbm_kernels']['kernels']]); print(d['roofline'])"
tail -3 gpurun_out/r02_bench_emit.err
