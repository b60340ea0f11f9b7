set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt
timeout 150 python tools/kernel_sweep.py > gpurun_out/r2f_sweep.json 2> gpurun_out/r2f_sweep.txt
timeout 300 python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
FI_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-reference-workload --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1
nproc; lscpu | grep "Model name"
