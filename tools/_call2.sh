set -x
B=freeimpala_b200/_build/freeimpala_gpu
timeout 100 $B -p 2 -B 32 -M 32 -a 64 -T 200 -S 100 --game-steps 100 --agent-time 0 > gpurun_out/r2c_config3.json 2> gpurun_out/r2c_config3.err
timeout 100 $B -p 2 -B 32 -M 32 -a 64 -T 200 -S 100 --game-steps 100 --agent-time 0 --infer-every 1 > gpurun_out/r2c_config3_infer.json 2> gpurun_out/r2c_config3_infer.err
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/r2c_pytest_kernels.txt 2>&1
timeout 150 python tools/kernel_sweep.py > gpurun_out/r2c_sweep.json 2> gpurun_out/r2c_sweep.txt
nproc; lscpu | grep "Model name"
