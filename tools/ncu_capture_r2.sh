#!/bin/bash
# GPU-box profiling recipe of round 2 (B200_PROFILING.md): plain run first, then (1) the launch list of the first ADMM
# iteration of bench.py's workload (all its kernels: stacked operators, one batched MG-PCG solve, coarse solve, interface
# block, MONITOR) and (2) `--set full` captures of one finest-level instance of each hot kernel of the batched solve.
#   usage: bash tools/ncu_capture_r2.sh <tag> [bench args...]
set -u
tag="${1:-r2}"; shift || true
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 0 --no-profile --no-e2e $*"
KRE='regex:k_level_pass|k_seg_|k_trip|k_rowmap|k_spmv|k_dense_gemv|k_gamma|k_moni|k_gather|k_scatter|k_axpy|k_sweep|k_tri_multi|k_tail_rhs'
# ncu cannot profile kernel nodes of graphs that contain conditional nodes: for the profiling runs the CG loop falls back to
# one graph launch per iteration (same kernels, same launch parameters), the staged LDLT solve to plain launches
export DDPCA_NO_WHILE_GRAPH=1 DDPCA_NO_LDLT_GRAPH=1
python bench.py $ARGS > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 300 gpurun_out/${tag}_plain.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 1500 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py $ARGS > gpurun_out/${tag}_ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/${tag}_$1 python bench.py $ARGS > gpurun_out/${tag}_ncu_$1.log 2>&1
  echo "capture $1 rc=$?"
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/${tag}_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/${tag}_$1.ncu-rep; fi
}
# first instances are on the finest level (the V-cycle starts there); FWD_FULL (post-smoothing) reaches it third
cap bwd       'k_level_pass<\(int\)2>'  0  1
cap fwd_zero  'k_level_pass<\(int\)0>'  0  1
cap fwd_full  'k_level_pass<\(int\)1>'  2  1
cap spmv      'k_level_pass<\(int\)4>'  0  1
cap resid     'k_level_pass<\(int\)3>'  0  1
cap restrict  'k_trip_spmv<\(int\)8'    0  1
cap prolong   'k_trip_spmv<\(int\)2'    2  1
du -sh gpurun_out; ls gpurun_out/ | grep ${tag} | head -50
