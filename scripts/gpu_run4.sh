set -x
timeout 900 python -m pytest tests/test_selscan_v2_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r02_pytest4.log
timeout 600 python scripts/devbench.py --cfgs 8 --iters 10 > gpurun_out/r02_devbench4.log 2>&1
python scripts/prof_one.py --bwd --cfg 8 > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:selscan_bwd2_kernel -s 1 -c 1 -o gpurun_out/r02_bwd2_b python scripts/prof_one.py --bwd --cfg 8 > gpurun_out/ncu4.log 2>&1
cat gpurun_out/r02_pytest4.log gpurun_out/r02_devbench4.log
