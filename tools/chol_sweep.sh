#!/bin/bash
# experiments on the resident Cholesky: DBA_CHOL_WARM bit 0 = warm the potrf code while idle, bit 1 = diagonal owner substitutes itself
for w in 0 1 2 3; do
  echo "== DBA_CHOL_WARM=$w"
  DBA_CHOL_WARM=$w timeout 120 python tools/chol_timing.py 426 2>&1 | tail -31 | grep -E "column  [15]:|column 12|n=426|resident timing"
done
DBA_CHOL_WARM=2 timeout 100 python tools/chol_timing.py 200 2>&1 | tail -1
