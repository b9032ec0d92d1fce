#!/bin/bash
# A/B timing of Gram tuning flags on the same box: alternates variants, 3 rounds
for r in 1 2 3; do
  for f in 0 4; do
    echo -n "flags=$f: "; SQFA_GRAM_FLAGS=$f python tools/exp_gram.py 2 2>&1 | grep -E "gram ks=512" | cut -c 1-40
  done
done
