#!/bin/bash
# Host half of the hierarchy set-up under AddressSanitizer + UBSan, without a device: builds a sanitized variant of the
# library (host code of csrc/mg.cu + csrc/plan.cpp: plans, permuted operators, v1 / v2 layouts, chunk tables, transfer
# forms) into /tmp and runs ddpca_mg_setup_dryrun over the golden hierarchies (single and batched, both orderings) and
# the bodies of the BLOCK fixture.  Any report goes to stderr; the checksums printed must equal tests/test_plan.py's.
#   bash tools/asan_dryrun.sh [extra.ddpk]
set -euo pipefail
root="$(cd "$(dirname "$0")/.." && pwd)"
so=/tmp/libddpca_b200_asan.so
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fopenmp,-O1,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
  -shared -o "$so" "$root/ddpca-admm_b200/csrc/mg.cu" "$root/ddpca-admm_b200/csrc/plan.cpp" -lgomp
export DDPCA_B200_LIB="$so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
cd "$root"
python - "$@" <<'PY'
import sys
sys.path.insert(0, "ddpca-admm_b200"); sys.path.insert(0, ".")
import ddpca_b200 as dd
from ddpca_b200 import ddpk
from tests.helpers import load_golden
for name in ("beam_2lev", "beam_3lev", "block_lagrange"):
    d, meta, A, P = load_golden(name)
    for mode in (dd.SMOOTH_MC, dd.SMOOTH_LEX):
        for ns in (1, 3):
            r = dd.setup_dryrun([(A, P)] * ns, mode)
            print(name, "mode", mode, "subs", ns, "bytes", r["device_bytes"], "v2 levels", r["v2_levels"], "checksum", r["checksum"])
d = ddpk.load("tests/golden/block_small.ddpk.gz")
hs, v = {}, 0
while f"body{v}.maxiLeve" in d:
    L = int(d[f"body{v}.maxiLeve"][0])
    hs.setdefault(L, []).append(([ddpk.get_csr(d, f"body{v}.consStif{l}") for l in range(L + 1)], [ddpk.get_csr(d, f"body{v}.realProl{l}") for l in range(L)]))
    v += 1
for L, lst in hs.items():
    for mode in (dd.SMOOTH_MC, dd.SMOOTH_LEX):
        r = dd.setup_dryrun(lst, mode)
        print("block_small:", len(lst), "bodies,", L + 1, "levels, mode", mode, "bytes", r["device_bytes"], "v2 levels", r["v2_levels"])
for path in sys.argv[1:]:
    dd_ = ddpk.load(path, copy=False)
    r = dd.setup_dryrun([ddpk.get_hierarchy(dd_)], dd.SMOOTH_MC)
    print(path, "bytes", r["device_bytes"], "v2 levels", r["v2_levels"], "checksum", r["checksum"])
print("sanitized dry run finished")
PY
