timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest10.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
timeout 900 python scripts/sweep.py --out gpurun_out/r02_sweep.md > /dev/null 2> gpurun_out/r02_sweep.err
timeout 600 python scripts/detector_bench.py infer --arm ours --conf 0.001 --channels-last --steps 10 > gpurun_out/r02_infer_ours_fused_cl.json 2> gpurun_out/r02_infer_ours_fused_cl.err
cat gpurun_out/r02_pytest10.log gpurun_out/r02_bench_n1_a.json gpurun_out/r02_infer_ours_fused_cl.json; tail -3 gpurun_out/r02_bench_n1_a.err
