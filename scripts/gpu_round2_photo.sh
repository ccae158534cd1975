set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_steady.py tests/test_gpu_steady_state.py -q -p no:cacheprovider -k "photolysis or device_loop_matches or ensemble_runs_every or per_column_photolysis or hd189" 2>&1 | tail -6
python scripts/prof_photo.py 64 4 2>&1 | tail -1
python scripts/prof_photo.py 1 4 2>&1 | tail -1
timeout 600 python scripts/ensemble_to_steady_state.py 256 2500 gpurun_out/r02_ensemble256_to_steady_state.json 2>&1 | tail -1
