timeout 900 python -m pytest tests/test_selscan_v2_gpu.py -q -x 2>&1 | tail -3
timeout 600 python scripts/devbench.py --cfgs 8 --iters 10 2>&1
