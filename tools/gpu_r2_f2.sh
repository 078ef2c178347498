#!/bin/bash
mkdir -p gpurun_out
AUVRRT_EDGES_BRUTE=1 timeout 600 ncu --set full --clock-control none -k regex:k_edges_dubins -s 1 -c 1 -o gpurun_out/f_dubins python tools/micro_run.py edges 4000000 > gpurun_out/f_ncu_dubins.log 2>&1
ls -la gpurun_out
