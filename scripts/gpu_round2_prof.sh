set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "emitted_batch" 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"negjac|chemdf|rhs_stencil|lhs_diag|layer_scal|factor_kernel|lu_solve" -c 11 -f -o gpurun_out/r02_step592_emit python scripts/prof_step.py 592 1 > gpurun_out/ncu_step_emit.log 2>&1
tail -3 gpurun_out/ncu_step_emit.log; ls -la gpurun_out/
