set -x
python bench.py --steps 2 --warmup 1 --no-train --no-cpu --no-e2e > gpurun_out/plain12.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-train --no-cpu --no-e2e > gpurun_out/ncu12.log 2>&1
python scripts/prof_one.py --bwd > gpurun_out/plain12b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:selscan_ -s 2 -c 3 -o gpurun_out/r02_final python scripts/prof_one.py --bwd > gpurun_out/ncu12b.log 2>&1
python scripts/prof_one.py --bwd --dtype bf16 > gpurun_out/plain12c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:selscan_ -s 2 -c 3 -o gpurun_out/r02_final_bf16 python scripts/prof_one.py --bwd --dtype bf16 > gpurun_out/ncu12c.log 2>&1
