timeout 900 ncu --set full --clock-control none --import-source on -k regex:"negjac|lhs_diag" -c 2 -f -o gpurun_out/r02_negjac python scripts/prof_step.py 1184 1 > gpurun_out/ncu_negjac.log 2>&1
tail -2 gpurun_out/ncu_negjac.log
