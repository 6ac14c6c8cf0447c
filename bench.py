#!/usr/bin/env python
"""bench.py -- the DDPCA-ADMM hot path on B200: MG-PCG DOF*iter/s inside the ADMM contact / domain-decomposition
solve, ADMM solve wall time, strong scaling over 1/2/4/8 GPUs (BASELINE.json metric).

Workload (fixed, whatever --gpus says: STRONG scaling): the reference's BEAM example with domain decomposition
(examples/BEAM.h:390-609), domaNumb 8x2x1 = 16 subdomains of ~217 k DOF joined by 22 tied interfaces, globLeve 3 on
a coarsest mesh of 128x8x2 elements (synthetic refinement of the example's 64x4x2: 3.47 M DOF, 4 multigrid levels
per subdomain), macroscopic coarse problem of 76 860 rows.  The reference's own host C++ does
mesh, contact search, MULTIGRID::STIF_MATR/CONSTRAINT and MCONTACT::ESTABLISH (set-up, out of scope of the GPU
path) through the prebuilt driver oracle/_ref/beam_admm, outside the timed region; the same process then runs
the untouched reference loop for a few iterations on this box's host cores (the CPU baseline).

A "step" is one complete MCONTACT::CONTACT_ANALYSIS (MCONTACT.h:2493-2723): zero initial state, ADMM iterations
until MONITOR reports convergence; every iteration solves all subdomains with MG-PCG to 1e-14 (one batched solve
per rank), exchanges interface traces and updates auxiliary variables and multipliers.  With N ranks the 16
subdomains are partitioned over the ranks (ddpca_b200/partition.py); per iteration the ranks all-reduce the coarse
right-hand side and the MONITOR sums and swap interface traces pairwise (NCCL).

  value : sum over subdomains and ADMM iterations of (n_L * CG iterations) / max-over-ranks device time, state in HBM
  e2e   : the same with, per step, the load vectors copied from pinned host memory and the displacements read back
  admm  : solve wall time, ADMM iterations/s, upload time, launches per iteration, parity against the reference
  roofline     : dominant kernel of the batched MG-PCG, algorithmic bytes / CUDA-event time per launch
  cpu_baseline : the untouched reference (oracle/_ref/beam_admm) on the host cores, first iterations of the same solve

`--impl reference` times the reference's own loop on the same configuration (all host cores) and prints the same
line with "impl": "reference"; a step there is one ADMM iteration (a bounded sample of the solve).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ddpca-admm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

REF_BEAM_DD = os.path.join(ROOT, "oracle", "_ref", "beam_admm")
METRIC = "MG-PCG DOF*iter/s inside the ADMM solve"
UNIT = "DOF*iter/s"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_cache_dir():
    d = os.path.join(tempfile.gettempdir(), "ddpca_bench_cache")
    os.makedirs(d, exist_ok=True)
    return d


def workload_name(args):
    return f"BEAM DD domaNumb={args.doma.replace(',', 'x')} globLeve={args.glob}" + (f" diviNumb={args.divi.replace(',', 'x')}" if args.divi else "") + f" muscSett={args.musc}"


def run_reference_driver(args, ref_iters, out):
    """oracle/_ref/beam_admm: reference host set-up (+ dump) and `ref_iters` iterations of the untouched loop."""
    if not os.access(REF_BEAM_DD, os.X_OK):
        raise SystemExit("oracle/_ref/beam_admm is missing: run __graft_entry__.build() in the build container")
    cmd = [REF_BEAM_DD, "--glob", str(args.glob), "--doma", args.doma, "--musc", str(args.musc), "--ref-iters", str(ref_iters)]
    if args.divi:
        cmd += ["--divi", args.divi]
    cmd += ["--out", out] if out else ["--nomat"]
    t0 = time.time()
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference is a host OpenMP code and gets the whole box
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = os.environ.get("DDPCA_REF_THREADS", str(os.cpu_count() or 1))
    txt = subprocess.check_output(cmd, cwd=workload_cache_dir(), env=env).decode()
    meta = json.loads(txt.strip().splitlines()[-1])
    meta["driver_wall_s"] = time.time() - t0
    return meta


def generate_workload(args):
    """Reference host set-up once per box (cached in /tmp), with the first `--cpu-iters` reference iterations."""
    tag = f"beam_dd_g{args.glob}_{args.doma.replace(',', 'x')}" + (f"_d{args.divi.replace(',', 'x')}" if args.divi else "") + f"_m{args.musc}_r{args.cpu_iters}"
    out = os.path.join(workload_cache_dir(), tag + ".ddpk")
    if os.path.exists(out) and os.path.exists(out + ".json"):
        return out, json.load(open(out + ".json"))
    meta = run_reference_driver(args, args.cpu_iters, out + ".tmp")
    os.replace(out + ".tmp", out)
    json.dump(meta, open(out + ".json", "w"))
    return out, meta


def ref_rate(meta, w, k):
    """DOF*iter/s of the reference over its ADMM iterations [w, w+k): per-iteration wall times and CG iteration sums
    come from the driver (admm_hook.h); CG counts are summed over the CG_SOLV calls, DOF = mean subdomain size (the
    reference does not say which call belongs to which subdomain; sizes differ by < 5 %)."""
    ts, cg, calls = meta["ref_iter_s"], meta["ref_cg_iters"], meta["ref_cg_calls"]
    n = len(ts)
    w = min(w, max(0, n - 1))
    k = max(1, min(k, n - w))
    dof = sum(meta["body_dof"]) / len(meta["body_dof"])
    t = sum(ts[w:w + k])
    return {"value": dof * sum(cg[w:w + k]) / t, "seconds": t, "first": w, "count": k, "cg_iters": sum(cg[w:w + k]), "cg_calls": sum(calls[w:w + k]),
            "admm_iter_per_s": k / t}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm
        return {"sm_mhz": statistics.median(busy or sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """Reference arm: the untouched MCONTACT::CONTACT_ANALYSIS on this box's host cores, same configuration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    need = args.warmup + args.steps
    meta = run_reference_driver(args, need, None)
    r = ref_rate(meta, args.warmup, args.steps)
    cores = meta.get("omp_max_threads", os.cpu_count())
    sample = (f"{workload_name(args)}: ADMM iterations {r['first']}..{r['first'] + r['count'] - 1} of the untouched reference loop "
              f"({r['cg_calls']} MGPIS::CG_SOLV calls, {r['cg_iters']} CG iterations, {r['seconds']:.2f} s), OpenMP over subdomains, {cores} threads"
              + ("" if r["count"] == args.steps else f"; the reference converged after {len(meta['ref_iter_s'])} iterations, fewer than warmup+steps"))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / r["count"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "bodies": meta["bodies"], "interfaces": meta["interfaces"], "body_dof": meta["body_dof"],
                   "step": "one ADMM iteration of the reference loop (bounded sample of the solve)", "rel_tol": 1e-14},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "admm": {"admm_iter_per_s": r["admm_iter_per_s"], "iterations_timed": r["count"], "cores": cores,
                 "solve_wall_s_extrapolated": None if meta.get("ref_iterNumbReco") is None else sum(meta["ref_iter_s"])},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--glob", type=int, default=3, help="BEAM globLeve (3: 0.93 M DOF in all)")
    ap.add_argument("--doma", default="8,2,1", help="BEAM domaNumb (subdomains per direction)")
    ap.add_argument("--divi", default="128,8,2", help="BEAM diviNumb of the coarsest mesh (the example's own is 64,4,2: a quarter of the DOF)")
    ap.add_argument("--musc", type=int, default=1, help="coarse-space correction: 1 macroscopic problem, 2 interface-eliminated, 3 both, 0 none")
    ap.add_argument("--cpu-iters", type=int, default=4, help="reference iterations run on the host cores for the CPU baseline")
    ap.add_argument("--smoother", default="mc", choices=["mc", "lex"])
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    # host side of the upload (planning, layouts: OpenMP inside libddpca_b200): an even share of the cores per rank
    # (torchrun's default of one thread per worker would serialise it); must be set before the library is loaded
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))

    import numpy as np
    import torch

    import ddpca_b200 as dd
    from ddpca_b200 import ddpk
    from ddpca_b200.comm import TorchComm
    from ddpca_b200.partition import partition_bodies

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- set-up (untimed): the reference's host C++ builds every operator; upload to HBM ------------------
    if rank == 0:
        path, meta = generate_workload(args)
    barrier()
    if rank != 0:
        path, meta = generate_workload(args)
    d = ddpk.load(path, copy=False)
    nb, ni = int(d["nbody"][0]), int(d["niface"][0])
    contBody = [[int(x) for x in d[f"if{ts}.contBody"]] for ts in range(ni)]
    nlev = [int(d[f"body{v}.maxiLeve"][0]) + 1 for v in range(nb)]
    weights = [len(d[f"body{v}.consStif{nlev[v] - 1}.val"]) for v in range(nb)]
    body_rank = partition_bodies(weights, contBody, world)
    comm = TorchComm(dev) if world > 1 else None
    stream = comm.stream if comm is not None else torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    t0 = time.time()
    mc = dd.MCONTACT.from_ddpk(d, device=local, smoother=dd.SMOOTH_MC if args.smoother == "mc" else dd.SMOOTH_LEX,
                               body_rank=body_rank if world > 1 else None, rank=rank, comm=comm)
    if comm is None:
        from ddpca_b200.lib import check, load_library
        import ctypes as C

        check(load_library().ddpca_admm_set_stream(mc._h, C.c_void_p(stream.cuda_stream)))
    torch.cuda.synchronize()
    upload_s = time.time() - t0
    mine = [v for v in range(nb) if body_rank[v] == rank]
    dof_local = sum(mc.body_dof[v] for v in mine)

    def solve(consForc=None):
        mc.reset(consForc)
        mc.CONTACT_ANALYSIS()
        return mc.iterNumbReco + 1

    # ---- device-resident timing: `value` -----------------------------------------------------
    for _ in range(args.warmup):
        solve()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    mc.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dof_iters, cg_total, admm_iters = 0.0, 0, 0
    e0.record(stream)
    for _ in range(args.steps):
        admm_iters += solve()
        dof_iters += mc.cg_dof_iters
        cg_total += mc.cg_iters
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = mc.launch_count()
    clocks = sampler.stop()
    disp_final = mc.resuDisp

    # ---- end-to-end: per step the load vectors come from pinned host memory and the displacements go back ---------
    pin_in = {v: torch.from_numpy(np.ascontiguousarray(d[f"body{v}.consForc"])).pin_memory() for v in mine}
    pin_out = {v: torch.empty(mc.nfull[v], dtype=torch.float64).pin_memory() for v in mine}
    h2d = sum(8 * t.numel() for t in pin_in.values())
    d2h = sum(8 * t.numel() for t in pin_out.values())
    import ctypes as C

    from ddpca_b200.lib import check, load_library

    lib = load_library()

    def solve_e2e():
        it = solve({v: pin_in[v].data_ptr() for v in mine})
        for v in mine:
            check(lib.ddpca_admm_get_disp(mc._h, C.c_int(v), C.c_void_p(pin_out[v].data_ptr())))
        return it

    e2e_steps = 0 if args.no_e2e else args.steps
    for _ in range(0 if args.no_e2e else min(args.warmup, 1)):
        solve_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_dof_iters = 0.0
    e2.record(stream)
    for _ in range(e2e_steps):
        solve_e2e()
        e2e_dof_iters += mc.cg_dof_iters
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # ---- max over ranks, whole-job aggregate --------------------------------------------------------
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    work = torch.tensor([dof_iters, e2e_dof_iters, float(launches), float(cg_total), float(h2d), float(d2h)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_max, ms_e2e_max = t.tolist()
    value = work[0].item() / (ms_max * 1e-3)
    e2e_value = work[1].item() / (ms_e2e_max * 1e-3) if e2e_steps else None

    # ---- parity: the state after the reference's own first iterations (dumped by the driver) ------------------
    k_ref = int(d["ref.first_iters"][0]) if "ref.first_iters" in d else 0
    err = 0.0
    if k_ref:
        mc.reset()
        for tc in range(k_ref):
            mc.MONITOR(tc, mc.step(tc))
        disp = mc.resuDisp
        for v in mine:
            r = d[f"ref.resuDisp{v}"]
            err = max(err, float(np.linalg.norm(disp[v] - r) / np.linalg.norm(r)))
    et = torch.tensor([err], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)

    # ---- roofline of the dominant kernel class (CUDA events around every launch of the batched solves) ----
    roofline = None
    shares = {}
    if not args.no_profile:
        if rank == 0:
            mc.profile(True)
        mc.reset()
        for tc in range(2):   # two ADMM iterations = two batched solves of ~20 CG iterations each; all ranks take part
            mc.MONITOR(tc, mc.step(tc))
        if rank == 0:
            prof = mc.profile_get(max(nlev))
            mc.profile(False)
            tot = sum(v[0] for v in prof.values())
            peak, peak_src = hbm_peak()
            for (kname, lvl), (kms, kn, kb) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
                shares[f"{kname}@L{lvl}"] = {"share": round(kms / tot, 4), "launches": kn, "avg_us": round(1e3 * kms / kn, 2),
                                             "GBps": round(kb / (kms * 1e-3) / 1e9, 1) if kms > 0 else None}
            (kname, lvl), (kms, kn, kb) = max(prof.items(), key=lambda kv: kv[1][0])
            ach = kb / (kms * 1e-3) / 1e9
            traffic = None
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "NCU_TRAFFIC.json")))
                traffic = tr.get(f"beam_dd_g{args.glob}_{args.doma.replace(',', 'x')}_d{args.divi.replace(',', 'x')}_n{world}", {}).get(f"{kname}@L{lvl}")
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": f"{kname}@L{lvl}", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                        "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": kb / kn, "avg_launch_us": 1e3 * kms / kn,
                        "share_of_solve_kernels": round(kms / tot, 4),
                        "note": "per-launch CUDA events on rank 0's batched MG-PCG (all its subdomains in one launch), host-polled loop without look-ahead"}

    if rank == 0:
        steps = max(1, args.steps)
        solve_s = ms_max * 1e-3 / steps
        cpu = ref_rate(meta, 1, args.cpu_iters) if meta.get("ref_cg_iters") else None
        cpu_baseline = None
        if cpu:
            cores = meta.get("omp_max_threads", os.cpu_count())
            cpu_baseline = {"value": cpu["value"], "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": f"the same workload: ADMM iterations {cpu['first']}..{cpu['first'] + cpu['count'] - 1} of the untouched reference loop right after its set-up "
                                      f"({cpu['cg_calls']} MGPIS::CG_SOLV calls, {cpu['cg_iters']} CG iterations, {cpu['seconds']:.2f} s), OpenMP over subdomains, {cores} threads",
                            "admm_iter_per_s": cpu["admm_iter_per_s"]}
        its = admm_iters // steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "bodies": nb, "interfaces": ni, "body_dof": mc.body_dof, "levels_per_body": nlev,
                       "total_dof": int(sum(mc.body_dof)), "step": "one MCONTACT::CONTACT_ANALYSIS to convergence (MG-PCG of every subdomain to 1e-14 in every ADMM iteration)",
                       "smoother": args.smoother, "body_rank": body_rank, "exchange": "per iteration: all-reduce coarse RHS + MONITOR sums, pairwise swap of interface traces (NCCL)" if world > 1 else "none (one rank)",
                       "l2_policy": f"inputs larger than L2: the rank's finest operators ({sum(weights[v] for v in mine) * 9.33 / 1e6:.0f} MB) are streamed ~5 times per CG iteration vs 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(work[4].item()), "d2h_bytes_per_step": int(work[5].item()),
                    "ms_per_step": ms_e2e_max / steps if e2e_steps else None},
            "gpu_launches": int(work[2].item()),
            "admm": {"solve_wall_s": solve_s, "admm_iterations": its, "admm_iter_per_s": its / solve_s, "upload_s": round(upload_s, 2), "upload_breakdown_s": getattr(mc, "upload_times", None),
                     "mgpcg_dof_iter_per_s": value, "cg_iterations_per_solve": int(work[3].item()) // steps,
                     "gpu_launches_per_admm_iteration": round(work[2].item() / max(1, admm_iters) / 1.0, 1),
                     "reference_set_up_s": round(meta.get("driver_wall_s", 0.0) - sum(meta.get("ref_iter_s", [])), 1),
                     "parity_rel_err_resuDisp_vs_reference_after_first_iterations": et.item(), "parity_iterations_compared": k_ref,
                     "reference_admm_iter_per_s": cpu["admm_iter_per_s"] if cpu else None,
                     "reference_solve_wall_s_at_same_iteration_count": (its / cpu["admm_iter_per_s"]) if cpu else None,
                     "reference_cores": meta.get("omp_max_threads")},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kernel_shares": shares,
        }
        print(json.dumps(line), flush=True)
    mc.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
