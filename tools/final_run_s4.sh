# final profiling pass of round 2 (last session): launch list of the default bench command's timed loop + ncu --set full of the
# fused SMPL kernel in both arithmetics at 4,096 bodies
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_s4_bench_steps2.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-eager --skip-sweep --skip-other --skip-train --skip-reduce-dim --skip-whole-loop --skip-e2e --skip-channels-last > gpurun_out/r02_bench_under_ncu.log 2>&1
for mode in bf16x3 3xtf32; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:smpl_fused -o gpurun_out/prof_r02_s4_fused4096_$mode -f python tools/profile_smpl.py --batch 4096 --gemm-mode $mode > gpurun_out/p_$mode.log 2>&1
  ncu -i gpurun_out/prof_r02_s4_fused4096_$mode.ncu-rep --page raw --csv > gpurun_out/prof_r02_s4_fused4096_$mode.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/prof_r02_s4_fused4096_$mode.csv > gpurun_out/r02_ncu_full_summary_s4_fused_4096bodies_$mode.txt 2>&1
done
tail -5 gpurun_out/r02_ncu_full_summary_s4_fused_4096bodies_3xtf32.txt
wc -l gpurun_out/r02_launches_s4_bench_steps2.csv
