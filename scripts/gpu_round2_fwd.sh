set -x
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_gputests_fwd.log; tail -8 gpurun_out/r02_gputests_fwd.log
timeout 300 python -m pytest tests/test_gpu_device_steady.py -q -s -k "condensing or jupiter" -p no:cacheprovider 2>&1 | grep -v "^Include\|^$" | tail -12 > gpurun_out/r02_conden_loop.log; cat gpurun_out/r02_conden_loop.log
VK_FWD_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_nofwd.json 2> gpurun_out/r02_bench_nofwd.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_nofwd.json').read().strip().splitlines()[-1]); print('NOFWD', d['value'], d['ms_per_step'], d['e2e'])"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_fwd.json 2> gpurun_out/r02_bench_fwd.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_fwd.json').read().strip().splitlines()[-1]); print('FWD', d['value'], d['ms_per_step'], d['e2e']); print(json.dumps(d)[:3000])"
tail -3 gpurun_out/r02_bench_fwd.err
