cd /root/repo
timeout 2400 python -m pytest tests/test_selscan_v2_gpu.py tests/test_selscan_benchshape_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/devbench.py --cfgs 8 2>&1 | tail -2
timeout 300 python scripts/devbench.py --cfgs 8 --dtype bf16 2>&1 | tail -2
