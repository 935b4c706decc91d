TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --config c4 > gpurun_out/bench_r2k_8gpu_c4.json 2> gpurun_out/r2k_8gpu_c4.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2k_8gpu.json 2> gpurun_out/r2k_8gpu.err
timeout 300 $TR bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --config c5 --batch 512 > gpurun_out/bench_r2k_8gpu_c5.json 2> gpurun_out/r2k_8gpu_c5.err
timeout 300 $TR bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --global-step > gpurun_out/bench_r2k_8gpu_globalstep.json 2> gpurun_out/r2k_8gpu_gs.err
tail -c 300 gpurun_out/bench_r2k_8gpu_c4.json
