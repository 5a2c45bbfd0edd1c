#!/bin/bash
# report.pdf Table 2 (linear mortar, all 5 refinement cycles of the shipped parameter.txt) through the stand-alone
# DarcyVT: error columns of every row in every printed digit; GMRES counts of the current stop rule 12 24 40 59 82.
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; mkdir -p "$ROOT/gpurun_out"
RUN=$(mktemp -d); mkdir -p $RUN/space-time-plots $RUN/time-step-plots
sed -e 's/^mortar_degree: .*/mortar_degree: 1/' -e 's/^num_refinement: .*/num_refinement: 5/' "$ROOT/parameter.txt" > $RUN/parameter.txt
cd $RUN
timeout 120 "$ROOT/DarcyVT" > out.log 2>&1
OUT="$ROOT/gpurun_out/r2_table2.log"
{ grep -h "GMRES converges\|Exception\|Aborting" out.log; tail -7 out.log; } > "$OUT" 2>&1
cat "$OUT"
ok=1
for row in "0 12 - 6.504e-01 - 1.209e+00 - 7.906e-01 - 7.979e-01" "1 24 [-0-9.]* 3.628e-01 [-0-9.]* 7.213e-01 [-0-9.]* 4.759e-01 [-0-9.]* 5.111e-01" \
           "2 40 [-0-9.]* 1.744e-01 [-0-9.]* 3.188e-01 [-0-9.]* 2.461e-01 [-0-9.]* 2.336e-01" "3 59 [-0-9.]* 8.625e-02 [-0-9.]* 1.460e-01 [-0-9.]* 1.245e-01 [-0-9.]* 1.201e-01" \
           "4 82 [-0-9.]* 4.288e-02 [-0-9.]* 6.928e-02 [-0-9.]* 6.248e-02 [-0-9.]* 6.112e-02"; do
  grep -q -- "$row" out.log || { echo "MISSING: $row" | tee -a "$OUT"; ok=0; }
done
[ $ok = 1 ] && echo "Table 2 through the product path: OK" | tee -a "$OUT"
[ $ok = 1 ]
