#!/usr/bin/env python
"""bench.py -- MG-PCG DOF*iter/s of the DDPCA-ADMM hot path on B200 (BASELINE.json metric).

A "step" is one complete MGPIS::CG_SOLV(1, consForc) (MGPIS.h:163-225) on one subdomain
hierarchy: x0 = 0, V-cycle-preconditioned CG down to ||r|| <= 1e-14 ||b||.  The hierarchy is
produced by the reference's own host C++ (mesh, TRANSFER, STIF_MATR, CONSTRAINT -- setup, out
of scope of the GPU path) through the prebuilt driver oracle/_ref/beam_nodd, outside the timed
region.  With N ranks every rank owns one such subdomain (subdomains are independent in the
solve phase of an ADMM iteration, MCONTACT.h:2511-2538): weak scaling, no data-path collective.

  value : sum over ranks of (DOF * CG iterations * steps) / max-over-ranks device time, operands in HBM
  e2e   : the same through MGPIS.CG_SOLV with HOST (pinned) buffers, H2D + D2H inside the timing
  roofline     : dominant kernel class, algorithmic bytes / CUDA-event time
  cpu_baseline : the untouched reference (oracle/_ref) timed on this box's host cores

`--impl reference` times the reference's own CPU implementation (oracle/_ref/beam_nodd
--bench-steps) and prints the same line with "impl": "reference".
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

REF_BEAM = os.path.join(ROOT, "oracle", "_ref", "beam_nodd")
METRIC = "MG-PCG DOF*iter/s"
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


def generate_workload(glob, ref_solve):
    """Run the reference's host setup (and, optionally, its CPU solve) once; cache in /tmp."""
    cdir = workload_cache_dir()
    out = os.path.join(cdir, f"beam_nodd_g{glob}.ddpk")
    meta_p = out + ".json"
    if os.path.exists(out) and os.path.exists(meta_p):
        meta = json.load(open(meta_p))
        if (not ref_solve) or "cg_mg_s" in meta:
            return out, meta
    if not os.access(REF_BEAM, os.X_OK):
        raise SystemExit("oracle/_ref/beam_nodd is missing: run __graft_entry__.build() in the build container")
    t0 = time.time()
    txt = subprocess.check_output([REF_BEAM, "--glob", str(glob), "--out", out + ".tmp", "--solve", "1" if ref_solve else "0"], cwd=cdir).decode()
    meta = json.loads(txt.strip().splitlines()[-1])
    meta["generate_wall_s"] = time.time() - t0
    os.replace(out + ".tmp", out)
    json.dump(meta, open(meta_p, "w"))
    return out, meta


REF_BLOCK = os.path.join(ROOT, "oracle", "_ref", "block_admm")


def generate_admm_workload(glob, divi):
    """Reference host set-up of the BLOCK example (mesh, contact search, MCONTACT::ESTABLISH) plus the
    untouched reference ADMM loop for the CPU baseline / golden results; cached in /tmp."""
    cdir = workload_cache_dir()
    tag = f"block_g{glob}" + (f"_d{divi.replace(',', 'x')}" if divi else "")
    out = os.path.join(cdir, tag + ".ddpk")
    if os.path.exists(out) and os.path.exists(out + ".json"):
        return out, json.load(open(out + ".json"))
    if not os.access(REF_BLOCK, os.X_OK):
        raise SystemExit("oracle/_ref/block_admm is missing: run __graft_entry__.build() in the build container")
    cmd = [REF_BLOCK, "--glob", str(glob), "--musc", "1", "--out", out + ".tmp", "--ref-iters", "0"]
    if divi:
        cmd += ["--divi", divi]
    t0 = time.time()
    for attempt in range(3):   # the driver ends through a watcher thread (_exit); a rare teardown race aborts it
        try:
            txt = subprocess.check_output(cmd, cwd=cdir).decode()
            break
        except subprocess.CalledProcessError:
            if attempt == 2:
                raise
    meta = json.loads(txt.strip().splitlines()[-1])
    meta["generate_wall_s"] = time.time() - t0
    os.replace(out + ".tmp", out)
    json.dump(meta, open(out + ".json", "w"))
    return out, meta


def admm_leg(args, rank, world, local, dist, torch):
    """ADMM contact solve of the BLOCK example on `world` GPUs: bodies are partitioned over the ranks,
    three all-reduces per iteration.  Strong scaling: the problem is fixed."""
    import numpy as np

    import ddpca_b200 as dd
    from ddpca_b200 import ddpk
    from ddpca_b200.comm import TorchComm
    from ddpca_b200.partition import partition_bodies

    if rank == 0:
        path, meta = generate_admm_workload(args.admm_glob, args.admm_divi)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        path, meta = generate_admm_workload(args.admm_glob, args.admm_divi)
    d = ddpk.load(path)
    nb, ni = int(d["nbody"][0]), int(d["niface"][0])
    contBody = [[int(x) for x in d[f"if{ts}.contBody"]] for ts in range(ni)]
    weights = [len(d[f"body{v}.consStif{int(d[f'body{v}.maxiLeve'][0])}.val"]) for v in range(nb)]
    body_rank = partition_bodies(weights, contBody, world)
    comm = TorchComm(torch.device("cuda", local)) if world > 1 else None
    t0 = time.time()
    mc = dd.MCONTACT.from_ddpk(d, device=local, body_rank=body_rank if world > 1 else None, rank=rank, comm=comm)
    upload_s = time.time() - t0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t1 = time.time()
    mc.CONTACT_ANALYSIS()
    torch.cuda.synchronize()
    solve_s = time.time() - t1
    tt = torch.tensor([solve_s], device=torch.device("cuda", local), dtype=torch.float64)
    wk = torch.tensor([mc.cg_dof_iters, float(mc.cg_iters), float(mc.launch_count())], device=torch.device("cuda", local), dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(wk, op=dist.ReduceOp.SUM)
    solve_s = tt.item()
    iters = mc.iterNumbReco + 1
    disp = mc.resuDisp
    err = 0.0
    for v in range(nb):
        if disp[v] is not None and f"ref.resuDisp{v}" in d:
            r = d[f"ref.resuDisp{v}"]
            err = max(err, float(np.linalg.norm(disp[v] - r) / np.linalg.norm(r)))
    et = torch.tensor([err], device=torch.device("cuda", local), dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    out = {
        "workload": f"BLOCK domaNumb=1x1x1 globLeve={args.admm_glob}: {nb} bodies {meta['body_dof']}, {ni} interfaces (2 frictionless contact, 6 tied), macroscopic problem {meta.get('globCoup_rows')} rows",
        "n_gpus": world, "body_rank": body_rank, "scaling": "strong",
        "admm_iterations": iters, "reference_admm_iterations": meta.get("ref_iterNumbReco", -2) + 1,
        "solve_wall_s": solve_s, "upload_s": upload_s, "admm_iter_per_s": iters / solve_s,
        "mgpcg_dof_iter_per_s": wk[0].item() / solve_s, "cg_iterations_total": int(wk[1].item()), "gpu_launches": int(wk[2].item()),
        "max_rel_err_resuDisp_vs_reference": et.item(),
        "cpu_baseline": {"solve_wall_s": meta.get("ref_admm_s"), "kind": "reference", "cores": os.cpu_count(),
                         "note": "untouched MCONTACT::CONTACT_ANALYSIS, OpenMP over bodies/interfaces (nested, dynamic); bodies < 50 000 DOF use host LDLT (MCONTACT.h:2527)"},
    }
    mc.close()
    return out


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
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """Reference arm: the untouched reference's MGPIS::CG_SOLV on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    glob = args.ref_glob
    cdir = workload_cache_dir()
    t0 = time.time()
    txt = subprocess.check_output([REF_BEAM, "--glob", str(glob), "--solve", "0", "--bench-steps", str(args.steps), "--bench-warmup", str(args.warmup)], cwd=cdir).decode()
    meta = json.loads(txt.strip().splitlines()[-1])
    n = meta["levels"][-1][0]
    v = meta["bench_dof_iter_per_s"]
    sample = f"BEAM no-DD globLeve={glob} ({n} DOF, {len(meta['levels'])} levels): {args.steps} MGPIS::CG_SOLV(1,.) calls, {meta['bench_iters'] // max(1, args.steps)} CG iterations each"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * meta["bench_s"] / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BEAM no-DD single-subdomain MG-PCG (bounded sample: globLeve={glob}, {n} DOF; per-thread throughput of the reference is size-independent)",
                   "levels": meta["levels"], "rel_tol": 1e-14},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample,
                         "note": "the reference's MG-PCG is single-threaded per subdomain (Eigen nbThreads()==-1); one subdomain => one core"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--glob", type=int, default=3, help="BEAM globLeve of the GPU workload (3: 861 696 DOF, 4 levels)")
    ap.add_argument("--ref-glob", type=int, default=2, help="BEAM globLeve of the bounded CPU sample for --impl reference")
    ap.add_argument("--smoother", default="mc", choices=["mc", "lex"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--admm", action="store_true", help="also run the ADMM contact solve (BLOCK example) and report it under \"admm\"")
    ap.add_argument("--admm-glob", type=int, default=2, help="BLOCK globLeve of the ADMM leg (2: 3 x 45 725 DOF + 6 plates)")
    ap.add_argument("--admm-divi", default="", help="BLOCK coarsest divisions a,b,c (default 6,6,6)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    import ddpca_b200 as dd
    from ddpca_b200 import ddpk

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

    # ---- setup (untimed): reference host C++ builds the hierarchy; upload to HBM ----------------
    want_ref = (rank == 0) and not args.no_cpu_baseline
    if rank == 0:
        path, meta = generate_workload(args.glob, want_ref)
    barrier()
    if rank != 0:
        path, meta = generate_workload(args.glob, False)
    d = ddpk.load(path)
    A, P = ddpk.get_hierarchy(d)
    b_host = np.ascontiguousarray(d["consForc"])
    n = A[-1].shape[0]
    t0 = time.time()
    mg = dd.MGPIS.from_hierarchy(A, P, device=local, smoother=dd.SMOOTH_MC if args.smoother == "mc" else dd.SMOOTH_LEX)
    establish_s = time.time() - t0
    # a real (non-null) torch stream: the library launches on it, torch events bracket it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    mg.set_stream(stream.cuda_stream)
    b_dev = torch.from_numpy(b_host).to(dev)
    x_dev = torch.empty_like(b_dev)

    # ---- device-resident timing: `value` -----------------------------------------------------
    for _ in range(args.warmup):
        mg.CG_SOLV_dev(1, b_dev.data_ptr(), x_dev.data_ptr())
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    mg.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters_total = 0
    e0.record(stream)
    for _ in range(args.steps):
        iters_total += mg.CG_SOLV_dev(1, b_dev.data_ptr(), x_dev.data_ptr())
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = mg.launch_count()
    clocks = sampler.stop()
    x_gpu = x_dev.cpu().numpy()

    # ---- end-to-end through the public host-buffer API: `e2e` --------------------------------------
    b_pin = torch.from_numpy(b_host).pin_memory()
    x_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    e2e_steps = 0 if args.no_e2e else args.steps
    for _ in range(0 if args.no_e2e else min(args.warmup, 2)):
        mg.CG_SOLV(1, b_pin.numpy())
    barrier()
    import ctypes as C

    from ddpca_b200.lib import check, load_library

    lib = load_library()
    it_c, res_c, tol_c = C.c_long(), C.c_double(), C.c_double()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_iters = 0
    e2.record(stream)
    for _ in range(e2e_steps):
        check(lib.ddpca_mg_pcg(mg._h, C.c_int(1), C.c_void_p(b_pin.data_ptr()), C.c_void_p(x_pin.data_ptr()), C.c_double(1e-14), C.c_long(n),
                               C.byref(it_c), C.byref(res_c), C.byref(tol_c)))
        e2e_iters += it_c.value
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # ---- max over ranks, whole-job aggregate --------------------------------------------------------
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    work = torch.tensor([float(n) * iters_total, float(n) * e2e_iters, float(launches)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_max, ms_e2e_max = t.tolist()
    value = work[0].item() / (ms_max * 1e-3)
    e2e_value = work[1].item() / (ms_e2e_max * 1e-3) if e2e_steps else None

    # ---- roofline of the dominant kernel class (CUDA events around every launch, same workload) ----
    roofline = None
    shares = {}
    if rank == 0 and not args.no_profile:
        mg.profile(True)
        for _ in range(max(1, min(3, args.steps))):
            mg.CG_SOLV_dev(1, b_dev.data_ptr(), x_dev.data_ptr())
        prof = mg.profile_get()
        mg.profile(False)
        tot = sum(v[0] for v in prof.values())
        peak, peak_src = hbm_peak()
        best = max(prof.items(), key=lambda kv: kv[1][0])
        for (kname, lvl), (kms, kn, kb) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            shares[f"{kname}@L{lvl}"] = {"share": round(kms / tot, 4), "launches": kn, "avg_us": round(1e3 * kms / kn, 2),
                                         "GBps": round(kb / (kms * 1e-3) / 1e9, 1) if kms > 0 else None}
        (kname, lvl), (kms, kn, kb) = best
        ach = kb / (kms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel on this workload, from the
        # committed `ncu --set full` capture (profiles/NCU_TRAFFIC.json); null for kernels / workloads not captured
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "NCU_TRAFFIC.json")))
            traffic = tr.get(f"beam_g{args.glob}", {}).get(f"{kname}@L{lvl}")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": f"{kname}@L{lvl}", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                    "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": kb / kn, "avg_launch_us": 1e3 * kms / kn,
                    "share_of_step": round(kms / tot, 4)}

    admm = None
    if args.admm:
        admm = admm_leg(args, rank, world, local, dist, torch)
    if rank == 0:
        parity = None
        if "cg_mg_x" in d:
            parity = float(np.linalg.norm(x_gpu - d["cg_mg_x"]) / np.linalg.norm(d["cg_mg_x"]))
        cpu_baseline = None
        if "cg_mg_s" in meta:
            cpu_baseline = {"value": meta["dof_iter_per_s"], "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": f"the same workload once: untouched reference MGPIS::CG_SOLV(1,.) on {n} DOF, {meta['cg_mg_iters']} iterations, {meta['cg_mg_s']:.2f} s on one host core (its MG-PCG is single-threaded per subdomain)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BEAM no-DD single-subdomain MG-PCG per GPU: globLeve={args.glob}, {n} DOF, {len(A)} levels, V(1,1) SGS ({args.smoother}), rel_tol 1e-14",
                       "levels": [[a.shape[0], a.nnz] for a in A], "subdomains_per_gpu": 1,
                       "l2_policy": f"inputs larger than L2: finest operator {12 * A[-1].nnz / 1e6:.0f} MB streamed several times per iteration vs 126 MB L2",
                       "cg_iterations_per_solve": iters_total // max(1, args.steps), "reference_cg_iterations": meta.get("cg_mg_iters"),
                       "establish_s": round(establish_s, 2)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n, "ms_per_step": ms_e2e_max / args.steps},
            "gpu_launches": int(work[2].item()),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kernel_shares": shares,
            "parity_rel_err_vs_reference": parity,
        }
        if admm is not None:
            line["admm"] = admm
        print(json.dumps(line), flush=True)
    mg.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
