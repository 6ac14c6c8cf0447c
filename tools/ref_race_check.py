#!/usr/bin/env python
"""Is the REFERENCE's own host set-up deterministic on this box?  Runs a reference-built driver N times with
the box's default OpenMP threads (and once more with OMP_NUM_THREADS=1), tallies exit codes, the md5 of the
operator dump and the ADMM iteration count.  Written to root-cause the round-1 flaky parity test: on the
16-core GPU box `beam_admm --glob 1 --doma 8,1,1` was seen to abort inside Eigen's setFromTriplets (index out
of range) -- a data race in the reference's nested-OpenMP set-up, not in the device path."""
import collections
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cases = {
    "beam_admm g1 8x1x1": ["beam_admm", "--glob", "1", "--musc", "1", "--doma", "8,1,1", "--ref-iters", "0"],
    "block_admm g2 2x2x2": ["block_admm", "--glob", "2", "--musc", "1", "--divi", "2,2,2", "--ref-iters", "0"],
}
out = {"cores": os.cpu_count()}
for name, cmd in cases.items():
    for threads in (None, "1"):
        tally = collections.Counter()
        for i in range(n if threads is None else 3):
            tmp = tempfile.mkdtemp(prefix="race_")
            dump = os.path.join(tmp, "d.ddpk")
            env = dict(os.environ)
            if threads:
                env["OMP_NUM_THREADS"] = threads
            p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", cmd[0])] + cmd[1:] + ["--out", dump], cwd=tmp, env=env, capture_output=True)
            key = {"rc": p.returncode}
            if p.returncode == 0:
                key["iters"] = json.loads(p.stdout.decode().strip().splitlines()[-1]).get("ref_iterNumbReco")
                key["md5"] = hashlib.md5(open(dump, "rb").read()).hexdigest()[:10]
            else:
                key["err"] = p.stderr.decode()[-200:].strip().splitlines()[-1][:160] if p.stderr else ""
            tally[json.dumps(key, sort_keys=True)] += 1
            subprocess.run(["rm", "-rf", tmp])
        out[f"{name} threads={threads or 'default'}"] = dict(tally)
        print(name, threads, dict(tally), flush=True)
print(json.dumps(out, indent=1))
