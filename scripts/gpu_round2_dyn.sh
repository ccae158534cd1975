set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_steady.py -q -p no:cacheprovider -k "per_column_photolysis or emitted_batch or ensemble_runs_every or device_loop_matches" 2>&1 | tail -6
VK_EMIT=0 VK_EMIT_JAC=0 timeout 600 python scripts/ensemble_to_steady_state.py 256 2500 2>&1 | tail -1
timeout 600 python scripts/ensemble_to_steady_state.py 256 2500 gpurun_out/r02_ensemble256_to_steady_state.json 2>&1 | tail -1
