#!/bin/bash
# A/B of env knobs on ONE box:  scripts/ab_env.sh "VAR=val VAR2=val" ...   ("-" = defaults); 100M headline shape
for kv in "$@"; do
  echo "== $kv"
  if [ "$kv" = "-" ]; then python scripts/sweep_probes.py 100000000 16384 "48" 2>&1 | head -1
  else env $kv python scripts/sweep_probes.py 100000000 16384 "48" 2>&1 | head -1; fi
done
