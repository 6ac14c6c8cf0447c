"""SURVEY.md §8 row f-3 on the GPU box: the reference's dual-mortar path MCONTACT::LAGRANGE(1)
(MCONTACT.h:2847-3701) with its solve -- `mgpi.ESTABLISH(); mgpi.BiCGSTAB_SOLV(1, F, U_1)` (:3561-3562) -- on the
device through the MGPIS overlay, against the untouched reference.

  python tools/lagrange_bench.py [--block-glob 3] [--cylinder] [--ref-cylinder] > gpurun_out/lagrange.json

BLOCK (patch test, one active-set step): overlay binary and pure-reference binary both run here.  CYLINDER_1 (Hertzian
contact, locaLeve 5: 504 036 rows, two active-set steps): the overlay binary runs here and is compared with
tests/golden/cylinder_lagrange.json (the reference's run in the build container, 3 min; --ref-cylinder repeats it on
this box).  Prints one JSON object; exits non-zero if a comparison fails."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_lagrange_golden import run  # noqa: E402

BIN = os.path.join(ROOT, "ddpca-admm_b200", "host", "_bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def compare(gpu, ref, tol=1e-8):
    bad = []
    if gpu["error"] or not gpu["converged"]:
        bad.append("overlay run failed or did not converge")
    if gpu["active_set_steps"] != ref["active_set_steps"]:
        bad.append(f"active-set steps {gpu['active_set_steps']} != {ref['active_set_steps']}")
    if gpu["unconverged_constraints"] != ref["unconverged_constraints"]:
        bad.append(f"active-set changes {gpu['unconverged_constraints']} != {ref['unconverged_constraints']}")
    err = max(abs(a - b) / b for a, b in zip(gpu["disp_norm"], ref["disp_norm"]))
    if err > tol:
        bad.append(f"displacement norms differ by {err:.2e}")
    lerr = 0.0
    for a, b in zip(gpu["resuLagr"], ref["resuLagr"]):
        if a["rows"] != b["rows"] or a["status_counts"] != b["status_counts"]:
            bad.append(f"{a['file']}: active set differs ({a['status_counts']} != {b['status_counts']})")
        elif b["normal_norm"] > 0:
            lerr = max(lerr, abs(a["normal_norm"] - b["normal_norm"]) / b["normal_norm"])
    if lerr > tol:
        bad.append(f"multiplier norms differ by {lerr:.2e}")
    return {"max_rel_err_disp_norm": err, "max_rel_err_multiplier_norm": lerr, "failures": bad}


def summary(r):
    keys = ("impl", "rows", "converged", "active_set_steps", "bicgstab_iters", "unconverged_constraints", "establish_s", "bicgstab_s", "total_s")
    return {k: r[k] for k in keys}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--block-glob", type=int, default=3)
    ap.add_argument("--cylinder", action="store_true")
    ap.add_argument("--ref-cylinder", action="store_true")
    a = ap.parse_args()
    out, failed = {}, False
    if a.block_glob > 0:
        args = ["--glob", str(a.block_glob), "--divi", "2,2,2"]
        ref = run(os.path.join(REF, "block_lagrange"), args, "Block")
        gpu = run(os.path.join(BIN, "block_lagrange_b200"), args, "Block")
        c = compare(gpu, ref)
        out["BLOCK"] = {"args": args, "b200": summary(gpu), "reference": summary(ref), **c,
                        "solve_speedup": sum(ref["establish_s"] + ref["bicgstab_s"]) / sum(gpu["establish_s"] + gpu["bicgstab_s"])}
        failed |= bool(c["failures"])
    if a.cylinder:
        args = ["--loca", "5"]
        if a.ref_cylinder:
            ref, where = run(os.path.join(REF, "cylinder_lagrange"), args, "Cylinder"), "this box"
        else:
            ref, where = json.load(open(os.path.join(ROOT, "tests", "golden", "cylinder_lagrange.json"))), "build container (tests/golden/cylinder_lagrange.json)"
        gpu = run(os.path.join(BIN, "cylinder_lagrange_b200"), args, "Cylinder")
        c = compare(gpu, ref)
        out["CYLINDER_1"] = {"args": args, "b200": summary(gpu), "reference": summary(ref), "reference_run_on": where, **c,
                             "solve_speedup": sum(ref["establish_s"] + ref["bicgstab_s"]) / sum(gpu["establish_s"] + gpu["bicgstab_s"])}
        failed |= bool(c["failures"])
    print(json.dumps(out))
    sys.exit(1 if failed else 0)
