#!/bin/bash
# A/B at the metric's configuration on ONE box:  scripts/ab100m.sh SPEC...   SPEC = "-" (default build) | path/to/lib.so | VAR=val
mkdir -p gpurun_out
for spec in "$@"; do
  unset CUVS_B200_LIB
  echo "== $spec"
  case "$spec" in
    -)     python scripts/sweep_probes.py 100000000 16384 "48" 2>&1 | head -1 ;;
    *.so)  CUVS_B200_LIB=$PWD/$spec python scripts/sweep_probes.py 100000000 16384 "48" 2>&1 | head -1 ;;
    *=*)   env $spec python scripts/sweep_probes.py 100000000 16384 "48" 2>&1 | head -1 ;;
  esac
done
