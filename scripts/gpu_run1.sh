set -x
nvidia-smi -L; nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_pytest1.log
timeout 600 python scripts/detector_bench.py logits > gpurun_out/r02_logits.json 2> gpurun_out/r02_logits.err
timeout 900 python scripts/detector_bench.py train --arm ours --steps 10 > gpurun_out/r02_train_ours_n1.json 2> gpurun_out/r02_train_ours_n1.err
timeout 900 python scripts/detector_bench.py train --arm pytorch --steps 4 --warmup 2 > gpurun_out/r02_train_pytorch_n1.json 2> gpurun_out/r02_train_pytorch_n1.err
timeout 900 python scripts/detector_bench.py train --arm pytorch --batch 4 --steps 4 --warmup 2 > gpurun_out/r02_train_pytorch_b4_n1.json 2> gpurun_out/r02_train_pytorch_b4_n1.err
timeout 900 python scripts/detector_bench.py infer --arm ours --steps 10 > gpurun_out/r02_infer_ours.json 2> gpurun_out/r02_infer_ours.err
timeout 900 python scripts/detector_bench.py infer --arm pytorch --steps 4 --warmup 2 > gpurun_out/r02_infer_pytorch.json 2> gpurun_out/r02_infer_pytorch.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err
tail -3 gpurun_out/r02_*.json gpurun_out/r02_pytest1.log
