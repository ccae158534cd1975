set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_gputests_final.log; tail -6 gpurun_out/r02_gputests_final.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err; tail -c 600 gpurun_out/r02_bench_ref_final.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_final2.json').read().strip().splitlines()[-1]); print('FINAL2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['column_groups'], d['gpu_launches']); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d.get('cpu_baseline')); print([(k['kernel'][:30], round(k['ms'],3), round(k['frac'],3)) for k in d['hbm_kernels']['kernels']])"
tail -3 gpurun_out/r02_bench_final2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
