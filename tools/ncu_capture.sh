#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): plain run first, then the launch list of two
# steady-state CG iterations, then full captures of the hot kernels.  Outputs in gpurun_out/.
#   usage: bash tools/ncu_capture.sh <tag> [bench args...]
set -u
tag="${1:-r1}"; shift || true
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --no-profile $*"
KRE='regex:k_sweep|k_spmv|k_resid|k_dense|k_update|k_dot|k_s_|k_jacobi'
python bench.py $ARGS > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 600 gpurun_out/${tag}_plain.log
# launches of the warm-up solve are skipped (-s); ~80 kernels per CG iteration
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 1800 -c 170 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py $ARGS > gpurun_out/${tag}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "$KRE" -s 1800 -c 85 \
    -o gpurun_out/${tag}_full python bench.py $ARGS > gpurun_out/${tag}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
