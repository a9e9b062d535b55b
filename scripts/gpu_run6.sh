timeout 900 python -m pytest tests/test_selscan_v2_gpu.py tests/test_selscan_benchshape_gpu.py -q -x 2>&1 | tail -12
timeout 600 python scripts/devbench.py --cfgs 9,8 --iters 10 2>&1
python scripts/prof_one.py --cfg 8 > gpurun_out/plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:selscan_fwd2_kernel -s 1 -c 1 -o gpurun_out/r02_fwd2_a python scripts/prof_one.py --cfg 8 > gpurun_out/ncu6.log 2>&1
