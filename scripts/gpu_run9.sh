timeout 300 python -m pytest tests/test_mamba_gpu.py -q -x -k "channels_last or inference_skips" 2>&1 | tail -3
timeout 600 python scripts/detector_bench.py train --arm ours --channels-last --steps 10 > gpurun_out/r02_train_ours_cl_n1.json 2> gpurun_out/r02_train_ours_cl_n1.err
timeout 600 python scripts/detector_bench.py train --arm pytorch --batch 4 --channels-last --steps 4 --warmup 2 > gpurun_out/r02_train_pytorch_cl_b4.json 2> gpurun_out/r02_train_pytorch_cl_b4.err
timeout 600 python scripts/detector_bench.py infer --arm ours --conf 0.001 --steps 10 > gpurun_out/r02_infer_ours_fused.json 2> gpurun_out/r02_infer_ours_fused.err
timeout 600 python scripts/detector_bench.py infer --arm ours --conf 0.001 --channels-last --steps 10 > gpurun_out/r02_infer_ours_fused_cl.json 2> gpurun_out/r02_infer_ours_fused_cl.err
timeout 600 python scripts/detector_bench.py infer --arm pytorch --dtype fp32 --batch 1 --conf 0.001 --steps 5 --warmup 2 > gpurun_out/r02_infer_pytorch_b1_fp32.json 2> gpurun_out/r02_infer_pytorch_b1_fp32.err
timeout 600 python scripts/detector_bench.py infer --arm ours --dtype fp32 --batch 1 --conf 0.001 --steps 10 > gpurun_out/r02_infer_ours_b1_fp32.json 2> gpurun_out/r02_infer_ours_b1_fp32.err
cat gpurun_out/r02_train_ours_cl_n1.json gpurun_out/r02_train_pytorch_cl_b4.json gpurun_out/r02_infer_ours_fused.json gpurun_out/r02_infer_ours_fused_cl.json gpurun_out/r02_infer_pytorch_b1_fp32.json gpurun_out/r02_infer_ours_b1_fp32.json
