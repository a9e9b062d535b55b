cd /root/repo
for mb in 8 32 100; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$((mb%10)) scripts/detector_bench.py train --channels-last --steps 10 --warmup 5 --bucket-mb $mb 2>/dev/null | tail -1 | cut -c1-60,200-420
done
