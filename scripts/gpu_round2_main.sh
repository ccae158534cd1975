set -x
cp -r oracle/_ref/HD189 /tmp/refA
(cd $GRAFT_REPO_ROOT; PYTHONHASHSEED=0 python oracle/dropin_in_reference.py --config HD189 --refdir /tmp/refA --reference --steady --out gpurun_out/r02_seam.jsonl > gpurun_out/r02_ref_run.log 2>&1) &
REFPID=$!
nproc; free -g | head -2
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_gputests_3.log; tail -8 gpurun_out/r02_gputests_3.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 1500 gpurun_out/r02_bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 3000 gpurun_out/r02_bench_1gpu.json; tail -5 gpurun_out/r02_bench_1gpu.err
PYTHONHASHSEED=0 timeout 300 python oracle/dropin_in_reference.py --config HD189 --refdir oracle/_ref/HD189 --cuda --steady --out gpurun_out/r02_seam.jsonl 2>&1 | tail -1
PYTHONHASHSEED=0 timeout 300 python oracle/dropin_in_reference.py --config HD189 --refdir oracle/_ref/HD189 --cuda --steady --device-loop --out gpurun_out/r02_seam.jsonl 2>&1 | tail -1
wait $REFPID
tail -2 gpurun_out/r02_ref_run.log
cat gpurun_out/r02_seam.jsonl
