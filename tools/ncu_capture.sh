#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): plain run first, then (1) the launch list of two
# steady-state CG iterations and (2) `--set full` captures of one instance of each hot kernel.
# CSV exports are written next to the reports (gpurun_out/ returns at most 64 MiB).
#   usage: bash tools/ncu_capture.sh <tag> [bench args...]
set -u
tag="${1:-r1}"; shift || true
mkdir -p gpurun_out
# no warm-up solve: the first V-cycle's launches are at known positions (see cap calls below)
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-e2e $*"
KRE='regex:k_level_pass|k_sweep|k_spmv|k_resid|k_dense|k_update|k_dot|k_s_|k_jacobi'
# ncu cannot profile kernel nodes of graphs that contain conditional nodes: for the profiling runs the
# CG loop falls back to one graph launch per iteration (same kernels, same launch parameters)
export DDPCA_NO_WHILE_GRAPH=1
python bench.py $ARGS > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 300 gpurun_out/${tag}_plain.log; echo
# skip the set-up V-cycle and the first CG iteration (~45 launches with v2 kernels); two iterations follow
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 45 -c 100 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py $ARGS > gpurun_out/${tag}_ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/${tag}_$1 python bench.py $ARGS > gpurun_out/${tag}_ncu_$1.log 2>&1
  echo "capture $1 rc=$?"
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page source --csv > gpurun_out/${tag}_$1_source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/${tag}_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/${tag}_$1.ncu-rep; fi
}
# v2 kernels: one k_level_pass<MODE> launch per sweep / residual / product
cap spmv      'k_level_pass<\(int\)4>'  1  1
cap fwd_zero  'k_level_pass<\(int\)0>'  0  1
cap fwd_full  'k_level_pass<\(int\)1>'  2  1
cap bwd       'k_level_pass<\(int\)2>'  0  1
cap resid     'k_level_pass<\(int\)3>'  0  1
du -sh gpurun_out; ls gpurun_out/ | head -50
