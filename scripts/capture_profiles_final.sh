mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:interact_lists -s 8 -c 1 -f -o gpurun_out/r02_interact_lists_growth_1M python scripts/profile_step.py growth_1M 2 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_cubes -s 8 -c 1 -f -o gpurun_out/r02_sweep_cubes_relu_1M python scripts/profile_step.py relu_1M 2 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gabriel_lists -s 4 -c 1 -f -o gpurun_out/r02_gabriel_lists_1M python scripts/profile_step.py gabriel_1M 2 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
