set -x
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_r01_v7.csv &
SMI=$!
python bench.py --impl reference > gpurun_out/bench_r01_v7_reference.json 2> gpurun_out/bench_r01_v7_reference.err
python bench.py > gpurun_out/bench_r01_v7.json 2> gpurun_out/bench_r01_v7.err
python bench.py --channels-last > gpurun_out/bench_r01_v7_channels_last.json 2>/dev/null
for w in smpl_sweep maf_sampling eval_pass; do python bench.py --workload $w > gpurun_out/extra_v7_$w.json 2>/dev/null; done
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01_v7_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01_v7_smpl1536_loop256.csv python tools/profile_smpl.py --batch 1536 --loop-batch 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r01_v7 python tools/profile_smpl.py --batch 256 --loop-batch 256 > gpurun_out/p9.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:smpl_fused -o gpurun_out/prof_r01_v7_fused4096 python tools/profile_smpl.py --batch 4096 > gpurun_out/p10.log 2>&1
tail -c 600 gpurun_out/bench_r01_v7.json
