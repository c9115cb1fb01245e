#!/bin/bash
# parameter scan of the look-ahead Cholesky schedule (environment knobs of potrf_lookahead), sizes as arguments
sizes="${@:-4096 8192}"
for ob in 2 3 4; do for pdl in 24 40 80; do for big in 2048 4096 100000; do for sn in 30 60 120; do
  r=$(ABO_POTRF_OB=$ob ABO_POTRF_PDLTILES=$pdl ABO_POTRF_BIGREM=$big ABO_POTRF_SMALLNEXT=$sn python tools/potrf_probe.py $sizes 2>/dev/null | grep "^n=" | awk '{printf "%s %s ", $1, $3}')
  echo "OB=$ob PDL=$pdl BIGREM=$big SMALLNEXT=$sn : $r"
done; done; done; done
