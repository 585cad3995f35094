# paired-tap staged NCHW sampler: parity tests, then configs[3] with the switch off / on (same smem per CTA) / on (same planes per CTA)
timeout 300 python -m pytest tests -m gpu -x -q -k "sampling or sample or maf" 2>&1 | tail -3
for cfg in "WHMR_STAGED_PAIR=0" "WHMR_STAGED_PAIR=1" "WHMR_STAGED_PAIR=1 WHMR_STAGED_PAIR_KEEP_CG=1" "WHMR_STAGED_PAIR=1 WHMR_STAGED_KB=24" "WHMR_STAGED_PAIR=1 WHMR_STAGED_KB=48"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --workload maf_sampling 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:(round(v['nchw_ms'],4), round(v['nchw_frac'],3)) for k,v in d['levels'].items()})"
done
