timeout 200 python -m pytest tests -m gpu -x -q -k "pinned_host" 2>&1 | tail -5
for cs in 1 2; do
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu --skip-eager --skip-sweep --skip-other --skip-train --skip-reduce-dim --e2e-compute-streams $cs > gpurun_out/r02_bench_e2e_gather_cs$cs.json 2> gpurun_out/r02_bench_e2e_gather_cs$cs.err; tail -3 gpurun_out/r02_bench_e2e_gather_cs$cs.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/r02_bench_e2e_gather_cs$cs.json').read().strip().splitlines()[-1])
for k in ('e2e','e2e_copy_all','e2e_host_gather_channels_last','e2e_feat_resident'): print(k, json.dumps({a:b for a,b in (d.get(k) or {}).items() if a!='note'})[:500])
print(d['value'], d['ms_per_step'])"
done
