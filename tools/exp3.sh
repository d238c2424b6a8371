ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 88 -c 8 -o gpurun_out/dgrad_prof -f python tools/profile_step.py --steps 2 > gpurun_out/ncu_dgrad.log 2>&1
tail -2 gpurun_out/ncu_dgrad.log
