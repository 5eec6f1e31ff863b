#!/bin/bash
# Instruction counts per object of libbuctd_b200.so (cuobjdump -sass): tcgen05 MMAs (UTCHMMA), TMEM loads / stores (LDTM / STTM),
# of which CTA-pair MMAs (UTCHMMA.2CTA, cta_group::2), TMA loads / stores (UTMALDG / UTMASTG; .2CTA loads complete on the
# leader's barrier), legacy tensor-core MMAs (HMMA: there must be none), MUFU.EX2.
#   bash tools/sass_summary.sh > profiles/sass_summary.txt
cd "$(dirname "$0")/../buctd_b200/csrc" || exit 1
printf "%-18s %9s %9s %7s %7s %8s %8s %6s %9s\n" object UTCHMMA ".2CTA" LDTM STTM UTMALDG UTMASTG HMMA MUFU.EX2
for o in api.o tapgemm.o tapgemm_chain.o pairgemm.o posatt_tc.o posatt_bwd.o gemm16.o ops.o decode.o loop_ops.o train_ops.o; do
  [ -f "$o" ] || continue
  s=$(/usr/local/cuda/bin/cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  h=$(echo "$s" | grep -v UTCHMMA | grep -c 'HMMA\.')
  printf "%-18s %9d %9d %7d %7d %8d %8d %6d %9d\n" "$o" "$(c UTCHMMA)" "$(c 'UTCHMMA.2CTA')" "$(c LDTM)" "$(c STTM)" "$(c UTMALDG)" "$(c UTMASTG)" "$h" "$(c 'MUFU.EX2')"
done
echo
echo "nvcc: $(/usr/local/cuda/bin/nvcc --version | tail -2 | head -1);  flags: -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo"
echo "built from git $(git rev-parse --short HEAD)"
