ncu --set full --clock-control none --import-source on -k regex:wgrad_unpack -s 1 -c 1 -o gpurun_out/unpack_prof -f python tools/profile_step.py --steps 2 > gpurun_out/ncu_unpack.log 2>&1
tail -2 gpurun_out/ncu_unpack.log
