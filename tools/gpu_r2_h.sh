#!/bin/bash
# round 2, GPU call H (2 GPUs): full GPU tests, the whole bench line at N=1 (config 5, mode 3, python reference) and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/h_pytest.log
tail -8 gpurun_out/h_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err; echo "bench n1 rc $?"
tail -c 600 gpurun_out/h_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err; echo "bench n2 rc $?"
tail -c 600 gpurun_out/h_bench_n2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/h_ref.json 2> gpurun_out/h_ref.err; echo "ref rc $?"
nproc; ls -la gpurun_out
