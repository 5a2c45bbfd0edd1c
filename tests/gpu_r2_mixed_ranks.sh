#!/bin/bash
# 2 GPUs, the stand-alone DarcyVT executable under a torchrun-style environment (RANK, WORLD_SIZE, LOCAL_RANK; NCCL id
# through the id file; system libnccl), checked against numbers the oracle produced on the CPU.
#  mixed:       rank 0 owns one 96^2 subdomain (persistent sweep kernel planned), rank 1 three 4^2 subdomains (no sweep
#               plan) -- the set-up cross-check's all-reduce must be joined by both ranks (it used to depend on a rank
#               having a plan).  Oracle: 18 iterations, r_norm 2.3620861579531436, errors 5.122e-01 1.057e+00 6.142e-01 1.697e+00
#  nonmatching: 2x2, 96^2 x 24 / 48^2 x 12, mortar 24 x 24 x 6 (configs[1] scaled down), two subdomains per rank, both
#               ranks on the persistent kernel.  Oracle: 95 iterations, r_norm 0.943484388230648,
#               errors 4.568e-02 1.858e-01 7.341e-02 2.123e-01
# usage: [NRANKS=1|2] gpu_r2_mixed_ranks.sh [mixed|nonmatching ...]
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; mkdir -p "$ROOT/gpurun_out"
CASES="${@:-mixed}"
OUT="$ROOT/gpurun_out/r2_mixed_ranks.log"; : > "$OUT"
port=29617
for CASE in $CASES; do
case $CASE in
  mixed)       MESH=$'mesh_pattern_sub_d0: 96 96 8\nmesh_pattern_sub_d1: 4 4 2\nmesh_pattern_sub_d2: 4 4 2\nmesh_pattern_sub_d3: 4 4 2\nmesh_pattern_mortar: 2 2 2'
               ITS=18; RNORM=2.36209; ERR="5.122e-01 - 1.057e+00 - 6.142e-01 - 1.697e+00" ;;
  nonmatching) MESH=$'mesh_pattern_sub_d0: 96 96 24\nmesh_pattern_sub_d1: 48 48 12\nmesh_pattern_sub_d2: 48 48 12\nmesh_pattern_sub_d3: 96 96 24\nmesh_pattern_mortar: 24 24 6'
               ITS="9[456]"; RNORM=0.943484;  # +-1: the stop test of this case flips with the summation order
               ERR="4.568e-02 - 1.858e-01 - 7.341e-02 - 2.123e-01" ;;
  *) echo "unknown case $CASE"; exit 2 ;;
esac
RUN=$(mktemp -d); mkdir -p $RUN/space-time-plots $RUN/time-step-plots
cat > $RUN/parameter.txt <<P
synthetic DarcyVT parameter file
----------------------------------------
positional format: keep one key/value per line
----------------------------------------
c0: 1.0

alpha: 1.0

space_degree: 0

mortar_degree: 1

num_refinement: 1

final_time: 0.5

tolerence: 1e-06

max_iteration: 500

need_plot_at_each_time_step(bool): 0

left_bc: D 0.0
bottom_bc: D 0.0
right_bc: D 0.0
top_bc: D 0.0

is_manufact_soln(bool): 1

$MESH
P
cd $RUN
port=$((port + 1))
N=${NRANKS:-2}
if [ "$N" = 1 ]; then
  timeout 70 "$ROOT/DarcyVT" > out0.log 2>&1; echo "(single rank)" > out1.log
else
  for r in $(seq 0 $((N - 1))); do
    RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r LOCAL_WORLD_SIZE=$N MASTER_PORT=$port timeout 70 "$ROOT/DarcyVT" > out$r.log 2>&1 &
  done
  wait
fi
{ echo "== $CASE, rank 0"; grep -h "r_norm is\|GMRES converges\|Exception\|Aborting" out0.log; tail -2 out0.log
  echo "== $CASE, rank 1"; tail -2 out1.log; ls space-time-plots | tr '\n' ' '; echo; } >> "$OUT" 2>&1
if grep -q "GMRES converges in $ITS iterations" out0.log && grep -q "r_norm is $RNORM" out0.log && grep -q -- "$ERR" out0.log; then
  echo "$CASE OK" >> "$OUT"
else
  echo "$CASE FAILED" >> "$OUT"
fi
done
cat "$OUT"
! grep -q FAILED "$OUT"
