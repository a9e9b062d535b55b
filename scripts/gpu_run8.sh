timeout 600 python scripts/detector_profile.py > gpurun_out/r02_detector_profile_train.txt 2>&1
timeout 600 python scripts/detector_profile.py --mode infer --size x --imgsz 1280 --batch 32 > gpurun_out/r02_detector_profile_infer.txt 2>&1
timeout 600 python scripts/detector_bench.py infer --arm ours --conf 0.001 --steps 10 > gpurun_out/r02_infer_ours_c001.json 2> gpurun_out/r02_infer_ours_c001.err
timeout 600 python scripts/detector_bench.py infer --arm pytorch --batch 1 --steps 5 --warmup 2 > gpurun_out/r02_infer_pytorch_b1.json 2> gpurun_out/r02_infer_pytorch_b1.err
timeout 600 python scripts/detector_bench.py infer --arm ours --batch 1 --steps 10 > gpurun_out/r02_infer_ours_b1.json 2> gpurun_out/r02_infer_ours_b1.err
