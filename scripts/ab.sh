#!/bin/bash
# A/B of library builds on ONE box (box-to-box variance is ~15 %):  scripts/ab.sh LIB...   ("-" = the default in-tree build)
for lib in "$@"; do
  if [ "$lib" = "-" ]; then unset CUVS_B200_LIB; else export CUVS_B200_LIB=$PWD/$lib; fi
  echo "== $lib"
  python scripts/sweep_probes.py 100000000 16384 "48" 2>&1 | head -1
  python scripts/sweep_probes.py 10000000 1024 "64" 2>&1 | head -1
done
