#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): plain run first, then (1) the launch list of two
# steady-state CG iterations and (2) `--set full` captures of one instance of each hot kernel.
# CSV exports are written next to the reports; reports that would break the 64 MiB return
# limit of gpurun_out/ are dropped after export.
#   usage: bash tools/ncu_capture.sh <tag> [bench args...]
set -u
tag="${1:-r1}"; shift || true
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --no-profile $*"
KRE='regex:k_sweep|k_spmv|k_resid|k_dense|k_update|k_dot|k_s_|k_jacobi'
python bench.py $ARGS > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 300 gpurun_out/${tag}_plain.log; echo
# launches of the warm-up solve are skipped (-s); ~80 kernels per CG iteration
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 1800 -c 170 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py $ARGS > gpurun_out/${tag}_ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 \
      -o gpurun_out/${tag}_$1 python bench.py $ARGS > gpurun_out/${tag}_ncu_$1.log 2>&1
  echo "capture $1 rc=$?"
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/${tag}_$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then ncu -i gpurun_out/${tag}_$1.ncu-rep --page source --csv > gpurun_out/${tag}_$1_source.csv 2>/dev/null; rm -f gpurun_out/${tag}_$1.ncu-rep; fi
}
# per CG iteration: 48 fwd-stage launches (8 colours x {pre,post} x 3 levels), first 8 = finest pre-smoothing
cap spmv      'k_spmv_group'       22   1
cap fwd_pre   'k_sweep_fwd_stage'  1056 2
cap fwd_post  'k_sweep_fwd_stage'  1096 2
cap bwd       'k_sweep_bwd_stage'  1056 2
cap resid     'k_resid_lower'      66   1
cap transfer  'k_spmvILi'          132  6
du -sh gpurun_out; ls -la gpurun_out/
