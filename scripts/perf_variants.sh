#!/bin/bash
# usage: scripts/perf_variants.sh M   (runs scripts/perf_c4.py for every build/variants/*.so)
for so in build/variants/*.so; do echo "== $so"; OCCM_B200_LIB=$PWD/$so python scripts/perf_c4.py ${1:-128} 2>&1 | grep -E "RHS:|RK45|evals/s|Error|error" ; done
