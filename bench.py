#!/usr/bin/env python
"""bench.py -- the hot path on N GPUs of one node.

Workload (BASELINE.json configs[3]): (H2O)_16 / cc-pVDZ RHF, 400 Cartesian basis functions,
192 shells, 1.72e8 unique shell quartets, 3.2e9 unique contracted ERIs; the shell-quartet
list is sharded over the ranks (strong scaling: the molecule is fixed).

One *step* = one pass of the hot path over the whole shard:
    (1) every unique contracted ERI of the shard recomputed by the shell-class kernels into
        the packed store (qbx_eri_recompute_async), then
    (2) one RHF Fock build: J/K digestion of the packed store (qbx_fock_build_device), which for
        N > 1 ends in the library's own NCCL all-reduce of the partial G matrices (qbx_comm_init).
`value` = unique contracted ERIs of the whole job / step time (contracted ERIs/s); the Fock
build alone (stored mode, the per-SCF-iteration cost) is reported as `fock_build_ms`.
Inputs are resident in HBM; the packed store (10.7 GB at N = 1 with the default 1e-12 Schwarz
screening, 25.7 GB unscreened) is far larger than L2.

`e2e` = the same metric through the C ABI with HOST buffers, every step: qbx_basis_create
(basis H2D) -> qbx_eri_store (Schwarz bounds, task lists, all ERIs) -> qbx_fock_build
(densities H2D, G D2H) -> qbx_basis_destroy.  Nothing is carried over between steps except
freed device blocks in the library's allocation pool (qbx.h: qbx_pool_trim); one untimed
warm-up step, then the median of three.

`--impl reference` times the CPU restatement of the reference's algorithm (oracle/) on the
host cores on a bounded sample of the same workload's unique contracted ERIs.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "contracted_eris_per_sec"
UNIT = "ERIs/s"


def workload(name):
    import quiqbox_b200 as qb
    from molecules import benzene, h2o, water_cluster
    if name.startswith("w"):
        n = int(name[1:])
        nuc, xyz = water_cluster(n)
        label = f"(H2O)_{n} cc-pVDZ RHF"
    elif name == "benzene":
        nuc, xyz = benzene()
        label = "benzene cc-pVDZ RHF"
    else:
        nuc, xyz = h2o()
        label = "H2O cc-pVDZ RHF"
    bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
    return label, nuc, xyz, bs


def unique_count(n):
    m = n * (n + 1) // 2
    return m * (m + 1) // 2


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.kill()                 # a lingering nvidia-smi poller slows every CUDA API call
            self.proc.wait()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores():
    """Cores this process may run on (the box's, not what a launcher put into OMP_NUM_THREADS)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_threads():
    """Set the oracle's OpenMP team to all host cores and return the team size that really runs.
    torch.distributed.run exports OMP_NUM_THREADS=1: without this the N >= 2 reference arm of round 1 ran on one
    thread while reporting every core."""
    import oracle
    return int(oracle.lib().orc_set_threads(host_cores()))


def bench_config(label, n, screen):
    """The workload description: identical in the GPU arm and in the reference arm."""
    return {"workload": label, "basis": "cc-pVDZ", "scf": "RHF", "nbf": n, "unique_eris_unscreened": unique_count(n),
            "screen_tol": screen}


def cpu_sample(bs, seconds, parallel=True):
    """Time the CPU oracle on uniformly sampled unique function quartets of the workload."""
    import oracle
    import quiqbox_b200 as qb
    ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
    n = ob.nbf
    rng = np.random.RandomState(12345)
    done, t_used, batch = 0, 0.0, 4096
    ob.eri_list(rng.randint(0, n, size=(256, 4)), parallel)            # warm-up
    while t_used < seconds:
        idx = rng.randint(0, n, size=(batch, 4))
        t0 = time.perf_counter()
        ob.eri_list(idx, parallel)
        t_used += time.perf_counter() - t0
        done += batch
    return done / t_used, done, t_used


def cpu_fock_sample(n_small=120, n_target=400, reps=3):
    """The second half of the metric on the CPU: the reference's getGcore (HartreeFock.jl:305-319, restated in
    orc_getGcore: 2 N^4 multiply-adds over the DENSE tensor, threads over the (mu, nu) pairs) timed on a dense
    N = 120 tensor -- benzene/cc-pVDZ size, 1.66 GB; the cost does not depend on the values, so the tensor is
    synthetic -- and scaled by (N_target / 120)^4 to the workload (whose dense tensor, 204.8 GB, the reference
    could not even allocate).  Returns (seconds at n_small, extrapolated seconds at n_target)."""
    import oracle
    rng = np.random.RandomState(7)
    H = np.empty(n_small ** 4)
    blk = rng.uniform(-1, 1, n_small ** 2)
    for i in range(n_small ** 2):                                      # cheap fill; values are irrelevant to the timing
        H[i * n_small ** 2:(i + 1) * n_small ** 2] = blk
    H = H.reshape((n_small,) * 4, order="F")
    D = rng.uniform(-1, 1, (n_small, n_small)); D = (D + D.T) / 2
    oracle.getGcore(H, 2 * D, D)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.getGcore(H, 2 * D, D)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return t, t * (n_target / n_small) ** 4


def run_reference(args):
    """Reference arm: the CPU restatement of the reference's algorithm (the reference is pure
    Julia and cannot be installed here: no julia binary, no network -- see DESIGN.md), on ALL host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    label, nuc, xyz, bs = workload(args.workload)
    threads = cpu_threads()
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(bs, per_step)
    rates, tot_t, tot_n = [], 0.0, 0
    for _ in range(args.steps):
        r, n, t = cpu_sample(bs, per_step)
        rates.append(r); tot_t += t; tot_n += n
    v = tot_n / tot_t
    f_small, f_target = cpu_fock_sample(n_target=len(bs))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(label, len(bs), args.screen),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{tot_n} uniformly sampled contracted ERIs of {label} per run, OpenMP over quartets "
                                       f"({threads} threads, set explicitly), per-primitive-component Obara-Saika as in the reference",
                             "fock_build_s": f_target, "fock_build_s_measured_n120": f_small,
                             "fock_build_rule": "orc_getGcore (= HartreeFock.jl:305-319) on a dense N = 120 tensor, x (N/120)^4"},
            "secondary_metric": {"metric": "rhf_fock_build_seconds_per_iter", "value": f_target, "unit": "s", "higher_is_better": False,
                                 "note": "extrapolated from N = 120, see cpu_baseline.fock_build_rule"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_synthetic(args):
    """BASELINE.json configs[4]: synthetic shell-quartet batches per class (ss|ss) .. (dd|dd), contraction degree
    K in {1, 2, 3, 4, 6, 9}, ERI throughput as absolute numbers and as a fraction of the measured FP64 roofline
    (qbx_prim_batch; tools/sweep_synthetic.py prints the same sweep as a table).  One JSON line; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_synthetic
    peak, rows = sweep_synthetic.sweep()
    tot_v = sum(r["eris_per_sec"] * r["seconds"] for r in rows); tot_t = sum(r["seconds"] for r in rows)
    tot_f = sum(r["model_flops"] for r in rows)
    worst = min(rows, key=lambda r: r["frac_of_fp64_peak"] if r["K"] >= 3 else 9)
    best = max(rows, key=lambda r: r["frac_of_fp64_peak"])
    print(json.dumps({"metric": METRIC, "value": tot_v / tot_t, "unit": UNIT, "n_gpus": 1, "steps": 1, "warmup": 1,
                      "ms_per_step": 1e3 * tot_t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": {"workload": "synthetic shell-quartet batches, 21 classes x K in {1,2,3,4,6,9}",
                                                      "generator": "qbx_prim_batch, seed 42, centres in a 10-bohr cube, exponents log-uniform in [0.1, 1e3]"},
                      "roofline": {"bound": "fp64", "achieved": tot_f / tot_t * 1e-12, "peak": peak, "unit": "TFLOP/s",
                                   "frac": tot_f / tot_t * 1e-12 / peak, "traffic": None, "kernel": "all 126 (class, K) launches",
                                   "peak_source": "measured in this run (qbx_fp64_peak)", "worst_K>=3": worst, "best": best},
                      "per_class": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="w16", help="w<N> water cluster | benzene | h2o | synthetic (per-class sweep, configs[4])")
    ap.add_argument("--screen", type=float, default=1e-12)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-batched", action="store_true",
                    help="also time the shell-batched algorithm on the host cores (tools/cpu_shell_batched.py, SURVEY.md 8d mode ii)")
    args = ap.parse_args()
    if args.workload == "synthetic":
        return run_synthetic(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import quiqbox_b200 as qb
    from quiqbox_b200 import lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.init(local)
    lib = L.load()
    if world > 1:
        # the collective lives inside the boundary: qbx_fock_build(_device) ends in the library's own ncclAllReduce
        # (include/qbx.h: qbx_comm_init); torch.distributed only carries the 128-byte NCCL id and the timing barriers
        from quiqbox_b200.parallel import LibComm
        LibComm(rank, world)
    stream = torch.cuda.Stream()            # a real (non-default) stream shared by torch, NCCL and libqbx
    torch.cuda.set_stream(stream)
    L.check(lib.qbx_set_stream(C.c_void_p(stream.cuda_stream)))

    label, nuc, xyz, bs = workload(args.workload)
    n = len(bs)
    t0 = time.perf_counter()
    db = qb.DeviceBasis(bs)
    eri = qb.DeviceERI(db, mode="stored", screen_tol=args.screen, rank=rank, nranks=world)
    setup_s = time.perf_counter() - t0
    info = db.info()

    rng = np.random.RandomState(3)
    Dh = rng.uniform(-1, 1, (n, n)); Dh = (Dh + Dh.T) / (2 * n)
    dDJ = torch.from_numpy(2 * Dh).cuda(); dDK = torch.from_numpy(Dh).cuda()
    dG = torch.zeros(n * n, dtype=torch.float64, device="cuda")

    def step():
        L.check(lib.qbx_eri_recompute_async(db.handle))
        L.check(lib.qbx_fock_build_device(db.handle, 1, C.c_void_p(dDJ.data_ptr()), C.c_void_p(dDK.data_ptr()),
                                          C.c_void_p(dG.data_ptr()), C.c_void_p(stream.cuda_stream)))     # incl. the all-reduce

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    st0 = db.stats()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clk:
        barrier()
        ev[0].record(stream)
        for _ in range(args.steps):
            step()
        ev[1].record(stream)
        barrier()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    st1 = db.stats()

    # Fock build alone (stored-mode digestion + all-reduce): the per-SCF-iteration cost
    fe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    fe[0].record(stream)
    for _ in range(args.steps):
        L.check(lib.qbx_fock_build_device(db.handle, 1, C.c_void_p(dDJ.data_ptr()), C.c_void_p(dDK.data_ptr()),
                                          C.c_void_p(dG.data_ptr()), C.c_void_p(stream.cuda_stream)))
    fe[1].record(stream)
    barrier()
    fock_ms = fe[0].elapsed_time(fe[1]) / args.steps

    # ERI recompute alone (class kernels overlapped on side streams, as in the step)
    ee = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ee[0].record(stream)
    for _ in range(args.steps):
        L.check(lib.qbx_eri_recompute_async(db.handle))
    ee[1].record(stream)
    barrier()
    eri_ms = ee[0].elapsed_time(ee[1]) / args.steps

    # per-class kernel times: one serialised recompute with CUDA events around every class launch
    cls = np.zeros((21, 6))
    L.check(lib.qbx_class_stats(db.handle, L.ptr(cls)))
    peak = C.c_double()
    L.check(lib.qbx_fp64_peak(C.byref(peak)))

    # max over ranks / sums over ranks
    vals = torch.tensor([ms, fock_ms, float(info["n_values"]), float(info["n_quartets"]), float(info["n_prim_quartets"]),
                         float(info["model_flops"]), float(info["stored_bytes"]), eri_ms],
                        dtype=torch.float64, device="cuda")
    mx, sm = vals.clone(), vals.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ms_max, fock_max, eri_ms_max = mx[0].item(), mx[1].item(), mx[7].item()
    tot_values, tot_quartets, tot_primq, tot_flops, tot_bytes = (sm[i].item() for i in (2, 3, 4, 5, 6))

    # ---------------- e2e through the C ABI with host buffers (every step: ingest -> ERIs -> Fock -> G on host)
    e2e = None
    if not args.no_e2e:
        L.check(lib.qbx_set_stream(None))
        eri = None
        db.close()
        mod = qb.MultiOrbitalData.from_orbitals(bs)
        arrs = [np.ascontiguousarray(a) for a in (mod.cen, mod.xpn, mod.ang, mod.bf_off, mod.bf_prim, mod.bf_w)]
        DJh, DKh, Gh = np.asfortranarray(2 * Dh), np.asfortranarray(Dh), np.zeros(n * n)
        h2d = sum(a.nbytes for a in arrs) + DJh.nbytes + DKh.nbytes
        times = []
        for it in range(1 + max(2, min(args.steps, 3))):
            barrier()
            t0 = time.perf_counter()
            h = C.c_void_p()
            L.check(lib.qbx_basis_create(mod.nprim, L.ptr(arrs[0]), L.ptr(arrs[1]), L.ptr(arrs[2]), mod.nbf, L.ptr(arrs[3]),
                                         L.ptr(arrs[4]), L.ptr(arrs[5]), C.byref(h)))
            t1 = time.perf_counter()
            L.check(lib.qbx_eri_store(h, args.screen, 0, rank, world))
            t2 = time.perf_counter()
            L.check(lib.qbx_fock_build(h, 1, L.ptr(DJh), L.ptr(DKh), L.ptr(Gh)))
            phases = [t1 - t0, t2 - t1, time.perf_counter() - t2]          # (the Fock build includes the library's all-reduce)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            lib.qbx_basis_destroy(h)
            if it > 0:
                times.append(dt)
        t = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pool = np.zeros(3, dtype=np.int64)
        L.check(lib.qbx_pool_trim(L.ptr(pool)))
        e2e = {"value": tot_values / t.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "pool": {"reused_blocks": int(pool[0]), "driver_allocations": int(pool[1]), "idle_bytes": int(pool[2])},
               "d2h_bytes_per_step": int(Gh.nbytes), "seconds_per_step": t.item(),
               "phase_seconds_rank0": {"qbx_basis_create": phases[0], "qbx_eri_store": phases[1], "qbx_fock_build": phases[2]},
               "path": "qbx_basis_create -> qbx_eri_store(stored) -> qbx_fock_build (host D, host G) -> qbx_basis_destroy"}

    # ---------------- the caller of the hot path: a whole runHartreeFock on the workload, SCF step on the device
    # (wall-time breakdown init / Fock / eigen / extrapolation)
    runhf = None
    if not args.no_e2e:
        try:
            from quiqbox_b200.parallel import LibComm
            comm = None
            if world > 1:
                comm = LibComm.__new__(LibComm); comm.rank, comm.size = rank, world        # communicator already initialised above
            cfg = qb.HFconfig(initial=":CoreH")
            walls = []
            for _ in range(2):                                 # the first call also loads cuBLAS / cuSOLVER and fills the pool
                tm = {}
                t0 = time.perf_counter()
                hdb = qb.DeviceBasis(bs)
                r = qb.runHartreeFock((nuc, xyz), hdb, cfg, mode="stored", screen_tol=args.screen, comm=comm, device_scf=True, timings=tm)
                torch.cuda.synchronize()
                walls.append(time.perf_counter() - t0)
                if len(walls) < 2:
                    hdb.close()
            wall = walls[-1]
            runhf = {"energy_hartree": float(sum(r.energy)), "converged": bool(r.converged), "steps": int(r.steps),
                     "fock_builds": int(r.fockBuilds), "wall_seconds": wall, "wall_seconds_first_call_in_process": walls[0],
                     "setup_seconds (basis, one-electron matrices, Schwarz, task lists, all ERIs)": wall - tm["scf_loop_seconds"] - tm["guess_seconds"],
                     "guess_seconds": tm["guess_seconds"], "scf_loop_seconds": tm["scf_loop_seconds"],
                     "device_fock_seconds": tm["device_fock_seconds"], "device_eigen_density_seconds": tm["device_eigen_seconds"],
                     "device_energy_residual_seconds": tm["device_step_seconds"] - tm["device_fock_seconds"] - tm["device_eigen_seconds"],
                     "host_extrapolation_and_launch_seconds": tm["scf_loop_seconds"] - tm["device_step_seconds"],
                     "initial": ":CoreH", "scf": "defaults (DD -> ADIIS -> DIIS to 1e-9)", "step_on_device": True}
            hdb.close()
        except Exception as exc:                               # a reported extra, never the reason for a missing bench line
            runhf = {"unavailable": str(exc)[-300:]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # dominant kernel = the class kernel with the largest device time
        k = int(np.argmax(cls[:, 1]))
        code = int(cls[k, 0])
        kflops = cls[k, 4] / cls[k, 1] * 1e-12 if cls[k, 1] > 0 else 0.0
        all_flops = cls[:, 4].sum() / (eri_ms * 1e-3) * 1e-12       # all classes, overlapped, this rank
        digest_gbs = (info["stored_bytes"] / (fock_ms * 1e-3)) * 1e-9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj["workload"] == label and world == tj["n_gpus"]:
                traffic = tj["dram_bytes_per_launch"].get(f"eri_class_kernel<{code // 1000},{code // 100 % 10},{code // 10 % 10},{code % 10}>")
        except Exception:
            pass
        def kernel_name(code):
            la, lb, lc, ld = code // 1000, code // 100 % 10, code // 10 % 10, code % 10
            nacc = ((la + lb + 1) * (la + lb + 2) * (la + lb + 3) // 6 - la * (la + 1) * (la + 2) // 6) * \
                   ((lc + ld + 1) * (lc + ld + 2) * (lc + ld + 3) // 6 - lc * (lc + 1) * (lc + 2) // 6)
            if nacc >= 180:                                   # QBX_COOP_ACC: warp-cooperative kernels
                return f"eri_coop2_kernel<{la},{lb},{lc},{ld}>"
            if lb == 0 and lc == 0 and ld == 0 and la <= 1 and os.environ.get("QBX_GC", "1") != "0":
                return f"eri_group_kernel<{la}>"              # ket-side general-contraction sharing
            return f"eri_class_kernel<{la},{lb},{lc},{ld}>"
        try:
            if traffic is None and tj["workload"] == label and world == tj["n_gpus"]:
                traffic = tj["dram_bytes_per_launch"].get(kernel_name(code))
        except Exception:
            pass
        switches = {k: os.environ[k] for k in ("QBX_GC", "QBX_COOP_MIN_ACC", "QBX_DIGEST_SPREAD", "QBX_POOL_GB") if k in os.environ}
        per_class = [{"class": f"({int(r[0]) // 1000}{int(r[0]) // 100 % 10}|{int(r[0]) // 10 % 10}{int(r[0]) % 10})",
                      "ms": r[1] * 1e3, "quartets": r[2], "prim_quartets": r[3],
                      "tflops_model": (r[4] / r[1] * 1e-12) if r[1] > 0 else 0.0, "kernel": kernel_name(int(r[0]))} for r in cls if r[2] > 0]
        line = {
            "metric": METRIC, "value": tot_values / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(label, n, args.screen),
            "workload_stats": {"shells": info["nshell"], "unique_shell_quartets": tot_quartets, "unique_eris": tot_values,
                               "prim_quartets": tot_primq,
                               "l2": "inputs larger than L2 (packed store %.1f GB/rank)" % (info["stored_bytes"] * 1e-9),
                               "parallelism": f"shell-quartet shards x{world}", "setup_seconds": setup_s,
                               "switches": switches or "defaults"},
            "eri_ms": eri_ms_max, "fock_build_ms": fock_max, "fock_build_s_per_iter": fock_max * 1e-3,
            "secondary_metric": {"metric": "rhf_fock_build_seconds_per_iter", "value": fock_max * 1e-3, "unit": "s",
                                 "higher_is_better": False, "note": "stored-mode J/K digestion + all-reduce, device-timed, max over ranks"},
            "gpu_launches": int(round((st1["launches"] - st0["launches"]))),
            "roofline": {"bound": "fp64", "kernel": kernel_name(code),
                         "achieved": kflops, "peak": peak.value, "unit": "TFLOP/s", "frac": kflops / peak.value if peak.value else None,
                         "traffic": traffic, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/traffic.json); the kernel is FP64-bound, its DRAM traffic is the packed output plus the task list",
                         "peak_source": "measured in this run (qbx_fp64_peak, DFMA chains); MEASURED_PEAKS.json has no FP64 entry",
                         "all_eri_kernels_achieved": all_flops, "all_eri_kernels_frac": all_flops / peak.value if peak.value else None,
                         "work_model": "SURVEY.md 8(d): flops = prim_quartets*(prim+acc) + quartets*hrr"},
            "roofline_digest": {"bound": "hbm", "kernel": "digest_kernel<*> (stored-mode Fock build)", "achieved": digest_gbs,
                                "peak": hbm_peak, "unit": "GB/s", "frac": digest_gbs / hbm_peak, "traffic": None,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "per_class": per_class,
            "clocks": clk.summary(),
        }
        if e2e:
            line["e2e"] = e2e
        if runhf:
            line["runhf"] = runhf
        if world == 1 and args.cpu_seconds > 0:
            threads = cpu_threads()
            v, nd, tu = cpu_sample(bs, args.cpu_seconds)
            f_small, f_target = cpu_fock_sample(n_target=n)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{nd} uniformly sampled contracted ERIs of {label} in {tu:.1f} s, OpenMP over "
                                              f"quartets ({threads} threads); per-primitive-component algorithm of the reference",
                                    "fock_build_s": f_target, "fock_build_s_measured_n120": f_small,
                                    "fock_build_rule": "orc_getGcore (= HartreeFock.jl:305-319) on a dense N = 120 tensor, x (N/120)^4"}
        if world == 1 and args.cpu_batched:
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import cpu_shell_batched
                line["cpu_baseline_batched"] = cpu_shell_batched.run()
            except Exception as exc:                       # a reported extra, never the reason for a missing bench line
                line["cpu_baseline_batched"] = {"unavailable": str(exc)[-300:]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
