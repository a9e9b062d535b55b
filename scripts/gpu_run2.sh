set -x
timeout 900 python -m pytest tests/test_selscan_v2_gpu.py tests/test_detect_gpu.py tests/test_mamba_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest2.log
timeout 600 python scripts/devbench.py --cfgs 9,8 --iters 10 > gpurun_out/r02_devbench2.log 2>&1
timeout 900 python -m pytest tests/test_selscan_benchshape_gpu.py -q 2>&1 | tail -30 > gpurun_out/r02_pytest2b.log
tail -5 gpurun_out/r02_pytest2.log gpurun_out/r02_pytest2b.log; cat gpurun_out/r02_devbench2.log
