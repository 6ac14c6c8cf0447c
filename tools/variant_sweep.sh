#!/bin/bash
# Times bench.py (device-resident MG-PCG) with every kernel-tuning build found in lib/var_*.so
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-profile > /dev/null 2>&1   # generate + cache workload
for lib in ddpca-admm_b200/lib/libddpca_b200.so ddpca-admm_b200/lib/var_*.so; do
  v=$(DDPCA_B200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
ks = d.get('kernel_shares') or {}
f = lambda k: (ks.get(k) or {}).get('GBps')
print('%.1f M DOFit/s  %.2f ms  fwd %s bwd %s spmv %s resid %s' % (d['value']/1e6, d['ms_per_step'], f('sweep_fwd@L3'), f('sweep_bwd@L3'), f('spmv@L3'), f('resid@L3')))")
  echo "$(basename $lib): $v"
done | tee gpurun_out/variants.txt
