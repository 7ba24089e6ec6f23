#!/bin/bash
# one gpurun call: GPU tests, bench lines (both arms), ncu launch list, ncu full capture, captures of the other sizes and of the
# one-kernel streaming tick, other configs
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_final.log 2>&1; tail -3 gpurun_out/gputest_final.log
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -c 400 gpurun_out/bench_r02_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rdif|beat_post" -c 4 -o gpurun_out/prof_r02 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdif -s 3 -c 1 -o gpurun_out/r02_n2048 -f python tools/quick_bench.py 2048 100 3600 x > gpurun_out/ncu_n2048.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdif -s 3 -c 1 -o gpurun_out/r02_n8192 -f python tools/quick_bench.py 8192 512 3600 x > gpurun_out/ncu_n8192.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdif -s 3 -c 1 -o gpurun_out/r02_n16384 -f python tools/quick_bench.py 16384 512 600 x > gpurun_out/ncu_n16384.log 2>&1
timeout 300 python tools/tick_bench.py > gpurun_out/tick_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:stream_tick -s 6 -c 1 -o gpurun_out/r02_tick -f python tools/tick_bench.py > gpurun_out/ncu_tick.log 2>&1
timeout 600 python tools/bench_configs.py > gpurun_out/configs_r02.json 2> gpurun_out/configs_r02.err
ls -la gpurun_out | tail -15
