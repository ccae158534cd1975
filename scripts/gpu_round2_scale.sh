for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1]); print('N$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['column_groups'], d['n_gpus'], d['final_gather_rows'], d['clocks'])"
done
