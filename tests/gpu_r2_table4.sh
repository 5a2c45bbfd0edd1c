#!/bin/bash
# report.pdf Table 4 (quadratic mortar, 5 refinement cycles of the shipped parameter.txt) through the PRODUCT path:
# the stand-alone DarcyVT with MMMFE_TIE_RULE=dealii91 (deal.II's point location at the tie points, restated in
# host/darcy_vt.cpp).  Expected = the oracle under the same rule (tests/test_oracle_ties.py); cycles 0, 2, 4 are the
# published rows 1, 3, 5:   6.810e-01 1.347e+00 8.390e-01 2.133e+00 | 1.697e-01 3.511e-01 2.508e-01 2.820e-01 |
# 4.477e-02 8.589e-02 6.591e-02 9.201e-02.  GMRES counts of the current stop rule: 19 24 38 38 60.
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; mkdir -p "$ROOT/gpurun_out"
RUN=$(mktemp -d); mkdir -p $RUN/space-time-plots $RUN/time-step-plots
sed -e 's/^mortar_degree: .*/mortar_degree: 2/' -e 's/^num_refinement: .*/num_refinement: 5/' "$ROOT/parameter.txt" > $RUN/parameter.txt
cd $RUN
MMMFE_TIE_RULE=dealii91 timeout 120 "$ROOT/DarcyVT" > out.log 2>&1
OUT="$ROOT/gpurun_out/r2_table4.log"
{ grep -h "r_norm is\|GMRES converges\|Exception\|Aborting" out.log; tail -7 out.log; } > "$OUT" 2>&1
cat "$OUT"
ok=1
for row in "0 19 - 6.810e-01 - 1.347e+00 - 8.390e-01 - 2.133e+00" "2 38 [-0-9.]* 1.697e-01 [-0-9.]* 3.511e-01 [-0-9.]* 2.508e-01 [-0-9.]* 2.820e-01" \
           "4 60 [-0-9.]* 4.477e-02 [-0-9.]* 8.589e-02 [-0-9.]* 6.591e-02 [-0-9.]* 9.201e-02" "1 24 [-0-9.]* 4.181e-01" "3 38 [-0-9.]* 1.014e-01"; do
  grep -q -- "$row" out.log || { echo "MISSING: $row" | tee -a "$OUT"; ok=0; }
done
[ $ok = 1 ] && echo "Table 4 through the product path: OK" | tee -a "$OUT"
[ $ok = 1 ]
