for ns in 0 200 400 800 1600; do echo "stagger $ns"; MMI_STAGGER_NS=$ns timeout 300 python scripts/devbench.py --cfgs 8 --iters 10 2>&1 | grep geomA; done
