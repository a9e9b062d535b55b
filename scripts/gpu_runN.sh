N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
timeout 600 $TR scripts/detector_bench.py train --arm ours --channels-last --steps 10 > gpurun_out/r02_train_ours_cl_n$N.json 2> gpurun_out/r02_train_ours_cl_n$N.err
timeout 600 $TR scripts/detector_bench.py train --arm ours --steps 10 > gpurun_out/r02_train_ours_n$N.json 2> gpurun_out/r02_train_ours_n$N.err
timeout 600 $TR scripts/detector_bench.py train --arm pytorch --batch 4 --steps 4 --warmup 2 > gpurun_out/r02_train_pytorch_b4_n$N.json 2> gpurun_out/r02_train_pytorch_b4_n$N.err
cat gpurun_out/r02_bench_n$N.json gpurun_out/r02_train_ours_cl_n$N.json gpurun_out/r02_train_ours_n$N.json gpurun_out/r02_train_pytorch_b4_n$N.json
tail -2 gpurun_out/r02_bench_n$N.err gpurun_out/r02_train_ours_cl_n$N.err
