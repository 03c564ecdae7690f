#!/bin/bash
# resident Cholesky: compact (rolled, phased) potrf vs the unrolled one
for u in 0 1; do
  echo "== DBA_CHOL_POTRF_UNROLLED=$u"
  DBA_CHOL_POTRF_UNROLLED=$u timeout 120 python tools/chol_timing.py 426 2>&1 | tail -31 | grep -E "column  [15]:|column 12|n=426|resident timing|potrf:" | head -20
done
