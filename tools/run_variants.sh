#!/bin/bash
# times every library variant under demovlp_b200/lib/variants with tools/run_fwd.py (B = $1, default 512)
cd "$(dirname "$0")/.."
B=${1:-512}
for so in demovlp_b200/lib/variants/*.so; do
  echo -n "$(basename $so .so): "
  DVLP_LIB=$PWD/$so timeout 120 python tools/run_fwd.py $B bwd 2>&1 | tail -1
done
