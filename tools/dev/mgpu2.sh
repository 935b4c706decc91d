#!/bin/bash
# Development helper, run on a 2-GPU box: the multi-GPU test and the 2-GPU bench lines behind profiles/bench_<tag>_2gpu*.json
t=${1:-dev}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 300 python -m pytest tests/test_gpu_global_step.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${t}_2gpu_tests.log
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${t}_2gpu.json 2> gpurun_out/${t}_2gpu.err
timeout 300 $TR bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --global-step > gpurun_out/bench_${t}_2gpu_globalstep.json 2> gpurun_out/${t}_2gpu_gs.err
cat gpurun_out/${t}_2gpu_tests.log
