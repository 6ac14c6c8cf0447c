#!/bin/bash
# Regenerates the committed golden fixtures from the UNTOUCHED reference
# (compiled by oracle/Makefile into oracle/_ref/ from /root/reference).
# Run in the build container only (needs oracle/_ref/*): bash tests/golden/make_golden.sh
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="$here/../../oracle/_ref"
tmp="$(mktemp -d)"
cd "$tmp"   # the reference writes Beam/ Block/ result directories into the cwd
# 2-level and 3-level BEAM (no domain decomposition) hierarchies + reference
# MGPIS::MULT_VCYC / CG_SOLV / OUTP_SUB1 results (examples/BEAM.h:403-421)
"$ref/beam_nodd" --glob 1 --divi 8,2,2 --jacobi 1 --extra 1 --out "$tmp/beam_2lev.ddpk" > "$here/beam_2lev.json"
"$ref/beam_nodd" --glob 2 --divi 4,2,2 --jacobi 1 --extra 1 --out "$tmp/beam_3lev.ddpk" > "$here/beam_3lev.json"
gzip -9 -n -c "$tmp/beam_2lev.ddpk" > "$here/beam_2lev.ddpk.gz"
gzip -9 -n -c "$tmp/beam_3lev.ddpk" > "$here/beam_3lev.ddpk.gz"
rm -rf "$tmp"
ls -la "$here"
python "$here/make_block_fixture.py"
# BLOCK on the dual-mortar path (examples/BLOCK.cpp:96-102 -> MCONTACT::LAGRANGE(1), MCONTACT.h:2847-3701): the
# condensed system of the first active-set step, its rebuilt 3-level hierarchy and the reference's own
# `mgpi.BiCGSTAB_SOLV(1, F, U_1)` result (:3561-3562)
tmp="$(mktemp -d)"
cd "$tmp"
# (--skew 0.2 adds a non-symmetric variant of the same system solved by the reference class: lagrange_tap.h, SKEW_VARIANT)
"$ref/block_lagrange" --glob 2 --divi 1,1,1 --skew 0.2 --out "$tmp/block_lagrange.ddpk" > "$here/block_lagrange.json"
gzip -9 -n -c "$tmp/block_lagrange.ddpk" > "$here/block_lagrange.ddpk.gz"
rm -rf "$tmp"
