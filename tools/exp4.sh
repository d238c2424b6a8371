ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 78 -c 6 -o gpurun_out/conv_wide -f python tools/profile_step.py --steps 2 > gpurun_out/ncu1.log 2>&1
tail -n 1 gpurun_out/ncu1.log
