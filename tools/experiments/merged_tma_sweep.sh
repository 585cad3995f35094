# merged TMA copies (one box per operand tile / per pose-feature chunk) on / off, both arithmetics, one and two issuers
for cfg in "WHMR_FUSED_MERGED_TMA=0" "WHMR_FUSED_MERGED_TMA=1" "WHMR_FUSED_MERGED_TMA=1 WHMR_FUSED_ISSUERS=2" \
           "WHMR_FUSED_MERGED_TMA=0 WHMR_GEMM_MODE=3xtf32" "WHMR_FUSED_MERGED_TMA=1 WHMR_GEMM_MODE=3xtf32" \
           "WHMR_FUSED_MERGED_TMA=1 WHMR_GEMM_MODE=3xtf32 WHMR_FUSED_ISSUERS_TF32=2"; do
  echo "== $cfg"
  env $cfg timeout 120 python tools/quick_smpl.py 37 256 4096 16384 2>&1 | tail -4
done
