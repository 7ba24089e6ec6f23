#!/bin/bash
# one gpurun call: ncu launch list and full capture of the bench command, captures of the other sizes and of the one-kernel
# streaming tick, each summarised on the box (the reports together exceed what gpurun copies back), other configs, then the
# bench lines of both arms with the fresh traffic record.  Results: gpurun_out/profiles_new/ (-> profiles/), gpurun_out/*.json
set -x
cd /root/repo
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rdif|beat_post" -c 4 -o $O/prof_r02 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/ncu_full.log 2>&1
python tools/summarize_profile.py r02 $O/launches_r02.csv $O/prof_r02.ncu-rep
python tools/summarize_kernel.py $O/prof_r02.ncu-rep profiles/r02_ncu_c2_phases.md "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras (-k regex:rdif|beat_post -c 4)"
for cfg in "2048 100 3600 x" "8192 512 3600 x" "16384 512 600 x"; do
  n=${cfg%% *}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdif -s 3 -c 1 -o $O/r02_n$n -f python tools/quick_bench.py $cfg > $O/ncu_n$n.log 2>&1
  python tools/summarize_kernel.py $O/r02_n$n.ncu-rep profiles/r02_ncu_n$n.md "python tools/quick_bench.py $cfg (-k regex:rdif -s 3 -c 1)"
  rm -f $O/r02_n$n.ncu-rep
done
timeout 300 python tools/tick_bench.py > $O/tick_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:stream_tick -s 6 -c 1 -o $O/r02_tick -f python tools/tick_bench.py > $O/ncu_tick.log 2>&1
python tools/summarize_kernel.py $O/r02_tick.ncu-rep profiles/r02_ncu_tick.md "python tools/tick_bench.py (-k regex:stream_tick -s 6 -c 1 --cache-control none: warm caches, as in a replayed graph)"
timeout 600 python tools/bench_configs.py > $O/configs_r02.json 2> $O/configs_r02.err
timeout 900 python bench.py > $O/bench_r02_n1.json 2> $O/bench_r02_n1.err
tail -c 300 $O/bench_r02_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r02_ref.json 2> $O/bench_r02_ref.err
mkdir -p $O/profiles_new
cp profiles/r02_launches.md profiles/r02_ncu_full.md profiles/r02_ncu_c2_phases.md profiles/r02_ncu_n2048.md profiles/r02_ncu_n8192.md profiles/r02_ncu_n16384.md profiles/r02_ncu_tick.md profiles/traffic.json $O/profiles_new/
du -sh $O; ls -la $O | tail -25
