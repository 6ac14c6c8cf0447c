#!/bin/bash
# parity smoke (a few GPU tests) + bench line summary.  usage: bash tools/quick_bench.sh <tag> [bench args]
tag="${1:-q}"; shift || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mgpis.py -m gpu -x -q -k "vcycle or cg_solv_mc or spmv or g2" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python - "$tag" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("value %.1f M  ms/step %.2f  e2e %.1f M  launches %s  iters %s  parity %s" % (d["value"]/1e6, d["ms_per_step"], (d["e2e"]["value"] or 0)/1e6, d["gpu_launches"], d["config"]["cg_iterations_per_solve"], d["parity_rel_err_vs_reference"]))
print("roofline", d["roofline"])
for k, v in list(d["kernel_shares"].items())[:14]: print("  ", k, v)
PY
tail -3 gpurun_out/bench_${tag}.err
