# WHMR_PDL bit mask sweep: 1 chain, 2 fused, 4 read-out, 8 projection, 16 sampling
for m in ${PDL_MASKS:-0 2 3 6 10 18 7 15 31}; do
  WHMR_PDL=$m timeout 300 python bench.py --steps 300 --skip-e2e --skip-cpu --skip-sweep --skip-parity > gpurun_out/pdl_$m.json 2>/dev/null
done
