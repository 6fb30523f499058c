#!/bin/bash
# Tensor-core / TMA SASS mnemonics of every tcgen05 kernel of libtta.so (evidence for profiles/): counts per object
# file plus the first UTCHMMA / UTMALDG / LDTM lines with their addresses.
cd "$(dirname "$0")/../dnn-compression-tensor-admm_b200/csrc" || exit 1
for f in gram_tc gemm_tf32 lowrank2_fwd gemm_tma gemm_tn gemm_tc ttconv_tc; do
  [ -f $f.o ] || continue
  echo "== $f.o"
  cuobjdump -sass $f.o | grep -oE "\b(UTCHMMA|UTCQMMA|UTCMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|LDGSTS|UTCATOMSWS)[A-Z0-9_.]*" | sort | uniq -c | sort -rn
  cuobjdump -sass $f.o | grep -E "Function :|UTCHMMA|UTMALDG|LDTM|STTM|UTMASTG" | awk '/Function/ {fn=$0; n=0; print fn; next} n<6 {print; n++}'
done
