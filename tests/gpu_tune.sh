#!/bin/bash
# tuning matrix of the persistent sweep kernel on configs[1] (run under gpurun); output: gpurun_out/tune_*.log
cd "$(dirname "$0")"
run() {  # name, env...
  name=$1; shift
  env MMMFE_SWEEP_AUTOTUNE=0 "$@" timeout 300 python gpu_perf_probe.py 1024 256 512 128 256 64 3 > ../gpurun_out/tune_$name.log 2>&1
  grep -E "^apply|sweep kernel:|sweep batches|cycles/CTA|rror" ../gpurun_out/tune_$name.log | sed "s/^/[$name] /"
}
for cfg in "$@"; do
  case $cfg in
    x:*)  # generic: x:<name>:VAR=VAL,VAR=VAL  (VAR without the MMMFE_ prefix)
      name=$(echo "$cfg" | cut -d: -f2); envs=$(echo "$cfg" | cut -d: -f3 | tr ',' ' ')
      args=(); for e in $envs; do args+=("MMMFE_$e"); done
      run "$name" "${args[@]}" ;;
    base) run base ;;
    k6) run k6 MMMFE_SWEEP_CHUNK=6144 ;;
    k8) run k8 MMMFE_SWEEP_CHUNK=8192 ;;
    k3) run k3 MMMFE_SWEEP_CHUNK=3072 ;;
    m256) run m256 MMMFE_SWEEP_MICRO_TILE=256 ;;
    m1024) run m1024 MMMFE_SWEEP_MICRO_TILE=1024 ;;
    c2) run c2 MMMFE_SWEEP_CTAS_PER_SM=2 MMMFE_SWEEP_CHUNK=2048 ;;
    c2k3) run c2k3 MMMFE_SWEEP_CTAS_PER_SM=2 MMMFE_SWEEP_CHUNK=3072 ;;
    c2s64) run c2s64 MMMFE_SWEEP_CTAS_PER_SM=2 MMMFE_SWEEP_CHUNK=3072 MMMFE_SUBTREE_CELLS=64 ;;
    t128c3) run t128c3 MMMFE_SWEEP_COMPUTE_THREADS=128 MMMFE_SWEEP_CTAS_PER_SM=3 MMMFE_SWEEP_CHUNK=2048 MMMFE_SUBTREE_CELLS=64 ;;
    c2k2) run c2k2 MMMFE_SWEEP_COMPUTE_THREADS=128 MMMFE_SWEEP_CTAS_PER_SM=2 MMMFE_SWEEP_CHUNK=2048 ;;
    l8) run l8 MMMFE_LEAF_CELLS=8 ;;
    l16) run l16 MMMFE_LEAF_CELLS=16 ;;
    l16s256) run l16s256 MMMFE_LEAF_CELLS=16 MMMFE_SUBTREE_CELLS=256 ;;
    t512) run t512 MMMFE_SWEEP_COMPUTE_THREADS=512 ;;
    sp300) run sp300 MMMFE_SWEEP_SPIN_NS=300 ;;
    sp1000) run sp1000 MMMFE_SWEEP_SPIN_NS=1000 ;;
    g2) run g2 MMMFE_SWEEP_GROUPS=2 ;;
    g4) run g4 MMMFE_SWEEP_GROUPS=4 ;;
    g2t512) run g2t512 MMMFE_SWEEP_GROUPS=2 MMMFE_SWEEP_COMPUTE_THREADS=512 ;;
    g4t512) run g4t512 MMMFE_SWEEP_GROUPS=4 MMMFE_SWEEP_COMPUTE_THREADS=512 ;;
    p2) run p2 MMMFE_SWEEP_PACK=2 ;;
    p4) run p4 MMMFE_SWEEP_PACK=4 ;;
    p4s64) run p4s64 MMMFE_SWEEP_PACK=4 MMMFE_SUBTREE_CELLS=64 ;;
    p2s64) run p2s64 MMMFE_SWEEP_PACK=2 MMMFE_SUBTREE_CELLS=64 ;;
    p4k6) run p4k6 MMMFE_SWEEP_PACK=4 MMMFE_SWEEP_CHUNK=6144 ;;
    p2k6) run p2k6 MMMFE_SWEEP_PACK=2 MMMFE_SWEEP_CHUNK=6144 ;;
    p4s256) run p4s256 MMMFE_SWEEP_PACK=4 MMMFE_SUBTREE_CELLS=256 ;;
    p4s256k6) run p4s256k6 MMMFE_SWEEP_PACK=4 MMMFE_SUBTREE_CELLS=256 MMMFE_SWEEP_CHUNK=6144 ;;
    p2ns) run p2ns MMMFE_SWEEP_PACK=2 MMMFE_SWEEP_SHARE_BLOBS=0 ;;
    conc) run conc MMMFE_SWEEP_CONCURRENT=1 ;;
    s64) run s64 MMMFE_SUBTREE_CELLS=64 ;;
    s256) run s256 MMMFE_SUBTREE_CELLS=256 ;;
  esac
done
