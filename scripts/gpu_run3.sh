set -x
python scripts/prof_one.py --bwd --cfg 8 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:selscan_bwd2_kernel -s 1 -c 1 -o gpurun_out/r02_bwd2_a python scripts/prof_one.py --bwd --cfg 8 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
