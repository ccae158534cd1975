set -x
timeout 900 python -m pytest tests/test_gpu_device_steady.py tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_steady_state.py -q -s -p no:cacheprovider -k "against_reference_runs or emitted_batch or pipelined_host or replay_of_the_reference_time_grid" 2>&1 | grep -v "^Include\|^$" | cut -c1-400 | tail -40 > gpurun_out/r02_ens_reference.log; cat gpurun_out/r02_ens_reference.log
timeout 900 python scripts/ensemble_to_steady_state.py 4096 8000 gpurun_out/r02_ensemble4096_to_steady_state.json 2>&1 | tail -3
