set -x
ncu --set full --clock-control none --import-source on -k regex:k_gym_run -s 1 -c 1 -o gpurun_out/gym_run python tools/n1n4_run.py gym > gpurun_out/ncu_gym.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_astar -s 1 -c 1 -o gpurun_out/astar python tools/n1n4_run.py astar > gpurun_out/ncu_astar.log 2>&1
tail -3 gpurun_out/ncu_gym.log gpurun_out/ncu_astar.log
ls -la gpurun_out/*.ncu-rep
