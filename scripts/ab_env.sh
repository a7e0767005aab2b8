#!/bin/bash
# A/B of env knobs on ONE box:  scripts/ab_env.sh N N_LISTS PROBES "VAR=val VAR2=val" ...   ("-" = defaults)
n=$1; nl=$2; pr=$3; shift 3
for kv in "$@"; do
  echo "== $kv"
  if [ "$kv" = "-" ]; then python scripts/sweep_probes.py $n $nl "$pr" 2>&1 | head -1
  else env $kv python scripts/sweep_probes.py $n $nl "$pr" 2>&1 | head -1; fi
done
