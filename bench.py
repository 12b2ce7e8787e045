#!/usr/bin/env python
"""bench.py -- atom-steps/s of the MD-Bench short-range force path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU build (oracle/_ref)

A "step" is one complete MD run of the workload: `ntimes` (200) velocity-Verlet timesteps of the
Cu-FCC LJ system starting from the generated lattice, i.e. exactly what the reference's
"Performance: ... million atom updates per second" line measures (verletlist/main.c:337-338),
plus -- inside the timed region, which makes our number conservative -- the state reset, ghost
setup, first neighbor build and first force that the reference excludes from TOTAL.
Workload at N GPUs: Cu FCC 128^3 unit cells (8 388 608 atoms) PER GPU (weak scaling towards
BASELINE config 5, 256^3 = 67M atoms at 8 GPUs), LJ sigma=eps=1, cutoff 2.5, skin 0.3, rebuild every
20, verletlist full neighbor lists, DP.
"""
import argparse
import importlib
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "atom-steps/sec (LJ Cu FCC, verletlist, full neighbor lists)"
METRIC_CP = "atom-steps/sec (LJ Cu FCC, clusterpair %dx%d)"
UNIT = "atom-steps/s"
# algorithmic work of the LJ full-list force kernel per atom-step
FLOP_PER_ATOM_STEP = 1429.0           # 8*L + 15*C, L = 76.035 listed, C = 54.74 inside the cutoff
BYTES_PER_ATOM_STEP = {"dp": 365.0, "sp": 335.0}
# clusterpair, decomposed runs only (a single domain counts its pairs live): 8 flop per evaluated atom pair + 15 per pair inside
# the cutoff, single-domain counts of the same lattice over 200 steps (4x4: 13.7 cluster pairs and 55.6 in-cutoff pairs per atom)
CP_FLOP_PER_ATOM_STEP = {4: 2589.3, 8: 2995.0}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.lines, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [q.strip() for q in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref, built from /root/reference by oracle/Makefile)
def run_reference_binary(nx, ntimes, threads, variant="vl_dp_aos"):
    """returns dict(atom_steps, total_s, perf) parsed from the reference's own report, or None"""
    from refbind import ref_binary
    exe = ref_binary(variant)
    if exe is None:
        return None
    try:
        with open("/proc/cpuinfo") as f:
            if "avx512f" not in f.read():
                return None
    except OSError:
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_SCHEDULE="static", OMP_PROC_BIND="close")
    t0 = time.perf_counter()
    try:
        out = subprocess.run([exe, "-nx", str(nx), "-ny", str(nx), "-nz", str(nx), "-n", str(ntimes)],
                             env=env, capture_output=True, text=True, timeout=900).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    wall = time.perf_counter() - t0
    m = re.search(r"TOTAL ([0-9.]+)s FORCE ([0-9.]+)s NEIGH ([0-9.]+)s", out)
    p = re.search(r"Performance: ([0-9.]+) million atom updates per second", out)
    a = re.search(r"System: (\d+) atoms", out)
    if not (m and p and a):
        return None
    return dict(atoms=int(a.group(1)), ntimes=ntimes, total_s=float(m.group(1)), force_s=float(m.group(2)),
                neigh_s=float(m.group(3)), perf=float(p.group(1)) * 1e6, wall_s=wall)


def run_port_oracle(nx, ntimes):
    """fallback CPU baseline: the oracle restatement (single thread, scalar)"""
    from portbind import OracleVL
    o = OracleVL(True)
    o.configure(nx=nx, ntimes=ntimes)
    o.setup(create=True)
    t0 = time.perf_counter()
    o.run(ntimes)
    dt = time.perf_counter() - t0
    n = o.geti("Natoms")
    return dict(atoms=n, ntimes=ntimes, total_s=dt, perf=n * ntimes / dt, wall_s=dt)


def cpu_baseline(sample_nx=64, sample_steps=40):
    thr = host_threads()
    r = run_reference_binary(sample_nx, sample_steps, thr)
    if r is not None:
        return {"value": r["perf"], "unit": UNIT, "cores": thr, "kind": "reference",
                "sample": "reference MDBench (GCC -Ofast AVX512 OpenMP static) Cu FCC %d^3 = %d atoms x %d steps; TOTAL %.2fs FORCE %.2fs NEIGH %.2fs"
                          % (sample_nx, r["atoms"], sample_steps, r["total_s"], r["force_s"], r["neigh_s"])}
    r = run_port_oracle(32, 20)
    return {"value": r["perf"], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle restatement (scalar C, 1 thread) Cu FCC 32^3 x 20 steps; reference binary not runnable on this host"}


def reference_arm(args, rank, world):
    """The reference's own CPU build (oracle/_ref/MDBench-vl_dp_aos: GCC -Ofast AVX-512, OpenMP) on the host cores of the box.
    Each bench step is a BOUNDED SAMPLE of the GPU arm's workload: the same per-GPU box (nx = --ref-nx, default the GPU arm's
    128; default 64^3 = 1 048 576 atoms, see --ref-nx) for --ref-ntimes (20) timesteps = one rebuild interval instead of
    200 timesteps (ten of them); the metric is normalised per atom-step.  `config` says what this arm ran; `ms_per_step` is the wall time of the reference process
    (start-up, lattice generation and first list build included) while `value` uses the TOTAL the reference itself reports
    (verletlist/main.c:337-338), exactly like its "Performance" line."""
    if rank != 0:
        return 0
    thr = host_threads()
    total_steps = args.steps + args.warmup
    nx, nt = (args.ref_nx or args.nx), args.ref_ntimes
    runs = []
    kind = "reference"
    for i in range(total_steps):
        r = run_reference_binary(nx, nt, thr)
        if r is None:
            kind = "port"
            r = run_port_oracle(32, 20)
            thr = 1
        if i >= args.warmup:
            runs.append(r)
    atom_steps = sum(r["atoms"] * r["ntimes"] for r in runs)
    tsum = sum(r["total_s"] for r in runs)
    val = atom_steps / tsum
    if kind != "reference":
        nx, nt = 32, 20
    sample = ("reference MDBench-vl_dp_aos Cu FCC %d^3 x %d timesteps per bench step (GPU arm: %d^3 per GPU x %d timesteps)"
              % (nx, nt, args.nx, args.ntimes)) if kind == "reference" \
        else "oracle restatement Cu FCC 32^3 x 20 steps per bench step"
    cfg = workload_config(args, world)
    cfg["workload"] = ("Cu FCC %dx%dx%d unit cells (%d atoms) on the host CPU, LJ sigma=eps=1 rc=2.5 skin=0.3, reneigh 20, verletlist full "
                       "neighbor lists, DP, %d timesteps per bench step: a bounded sample of the GPU arm's workload (%d^3 unit cells "
                       "per GPU, %d timesteps per bench step); atom-steps/s is size-normalised" % (nx, nx, nx, 4 * nx ** 3, nt, args.nx, args.ntimes))
    cfg.update({"nx_per_gpu": nx, "ntimes": nt, "ntimes_gpu_arm": args.ntimes, "nx_per_gpu_gpu_arm": args.nx,
                "global_box": "%dx%dx%d unit cells = %d atoms" % (nx, nx, nx, 4 * nx ** 3), "parallelism": "%d OpenMP threads" % thr})
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(r["wall_s"] for r in runs) / max(1, len(runs)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": thr, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, world, grid=(1, 1, 1)):
    return {"workload": "Cu FCC %dx%dx%d unit cells per GPU (%d atoms/GPU), LJ sigma=eps=1 rc=2.5 skin=0.3, reneigh 20, "
                        "%s, %s, %d timesteps per bench step (BASELINE config 1 physics at config 5 per-GPU size)"
                        % (args.nx, args.nx, args.nx, 4 * args.nx ** 3,
                           ("clusterpair 4x%d %s cluster-pair lists" % (args.cluster_n, "half" if args.half else "full"))
                           if getattr(args, "scheme", "verletlist") == "clusterpair"
                           else "verletlist %s neighbor lists" % ("half" if args.half else "full"),
                           args.precision.upper(), args.ntimes),
            "nx_per_gpu": args.nx, "ntimes": args.ntimes, "precision": args.precision,
            "l2": ("inputs larger than L2 (cluster-pair list %.1f GB + cluster data %.2f GB per GPU vs 126 MB L2)"
                   % (args.nx ** 3 * 100 * 4 / 1e9, 4 * args.nx ** 3 * 36 / 1e9))
            if getattr(args, "scheme", "verletlist") == "clusterpair" else
            "inputs larger than L2 (neighbor list %.1f GB + positions %.2f GB per GPU vs 126 MB L2)"
            % (4 * args.nx ** 3 * 100 * 4 / 1e9, 4 * args.nx ** 3 * 24 / 1e9),
            "global_box": "%dx%dx%d unit cells = %d atoms" % (args.nx * grid[0], args.nx * grid[1], args.nx * grid[2],
                                                             4 * args.nx ** 3 * grid[0] * grid[1] * grid[2]),
            "parallelism": "1 domain" if grid == (1, 1, 1) else
            "spatial decomposition %dx%dx%d bricks over %d GPU(s), ghost exchange %s"
            % (grid[0], grid[1], grid[2], world,
               "per step by peer stores into the neighbor GPUs' arrays over NVLink (CUDA IPC + flag handshake), per rebuild by NCCL send/recv"
               if world > 1 else "by device copies")}


def clusterpair_secondary(m, args, local, stream, steps=2):
    """the same box with the reference's OTHER scheme, as BASELINE config 2 runs it (clusterpair 4x4, SP, full lists):
    device-resident throughput (state restore + cluster build + list build + first force inside the timed region, like the
    primary) and the force kernel's time per launch against the measured FP32 FMA peak"""
    import torch
    P = m.default_params(precision=m.SP, layout=m.AOS, nx=args.nx, ny=args.nx, nz=args.nx, ntimes=args.ntimes)
    c = m.ClusterSimulation(P, cluster_n=4, device=local)
    c.setStream(stream)
    natoms = c.createAtom()
    c.setup(adjust=True)
    c.updateSingleAtoms()
    c.saveState()
    cp0, in0 = c.countPairs()

    def one():
        c.restoreState()
        c.setup(adjust=False)
        return c.run(args.ntimes)

    one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rec, _ = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cp1, in1 = c.countPairs()
    c.setTiming(True)
    c.resetKernelStats()
    _, tm = one()
    ks = c.kernelStats()
    c.setTiming(False)
    f_ms = ks["force_ms"] / max(1, ks["force_launches"])
    flop = (8.0 * 0.5 * (cp0 + cp1) * 16 + 15.0 * 0.5 * (in0 + in1)) / natoms
    peak = m.measure_fma_peak(m.SP, local)
    ach = flop * natoms / (f_ms * 1e-3) * 1e-12
    traffic, ncu_pipes = None, None
    try:
        te = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("clusterpair/sp/%d" % args.nx, {})
        if "k_cp_force_lj_sp_duo<4, 1" in te.get("kernel", ""):   # a capture of the kernel that runs: <N = 4, FI = true, ..>
            traffic = te.get("bytes")
            ncu_pipes = dict(te.get("ncu") or {}, source=te.get("source"), commit=te.get("commit"))
    except Exception:
        pass
    out = {"metric": METRIC_CP % (4, 4), "config": "BASELINE config 2 physics (clusterpair 4x4, SP, full lists) at %d^3 unit cells" % args.nx,
           "value": natoms * args.ntimes * steps / (ms * 1e-3), "unit": UNIT, "dtype": "f32", "steps": steps,
           "ms_per_step": ms / steps,
           "roofline": {"kernel": "k_cp_force_lj_sp_duo<4, FI> (LJ tile force, two lanes per i-cluster, packed FP32, with finalIntegrate(n) + initialIntegrate(n+1) in its epilogue)", "bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak if peak else None, "traffic": traffic, "ncu_capture": ncu_pipes, "ms_per_launch": f_ms, "flop_per_atom_step": flop,
                        "force_share_of_step": ks["force_ms"] / (tm["TOTAL"] * 1e3) if tm["TOTAL"] else None,
                        "neigh_ms_per_rebuild": ks["neigh_ms"] / max(1, ks["neigh_launches"])},
           "thermo_final": {"step": int(rec[-1][0]), "T": float(rec[-1][1]), "P": float(rec[-1][2])}}
    c.close()
    return out



def verletlist_secondary(m, args, local, stream, name, precision="dp", half=0, eam=False, steps=2):
    """one more verletlist configuration at the same per-GPU box: `steps` timed bench steps (state restore + setup + `ntimes`
    timesteps, device-resident) and the force kernel's time per launch against its roofline.  Algorithmic work per atom-step
    (SURVEY 8d): LJ full 8 L + 15 C; LJ half: every local-local pair once + 6 flop per reaction update; EAM: pass 1 8 L + 14 C,
    pass 2 8 L + 40 C (L listed, C in-cutoff pairs per atom, counted live)."""
    import numpy as np
    import torch
    dp = precision == "dp"
    kw = dict(precision=m.DP if dp else m.SP, layout=m.AOS, nx=args.nx, ny=args.nx, nz=args.nx, ntimes=args.ntimes, half_neigh=half)
    if eam:
        kw["force_field"] = m.FF_EAM
    sim = m.Simulation(m.default_params(**kw), device=local)
    sim.setStream(stream)
    if eam:   # BASELINE config 4 physics: Cu_u3 funcfl tables (committed fixture made from the reference's data/Cu_u3.eam)
        g = np.load(os.path.join(ROOT, "tests", "golden", "eam_cu_nx5.npz"))
        sim.setEam(int(g["funcfl_nrho"]), float(g["funcfl_drho"]), int(g["funcfl_nr"]), float(g["funcfl_dr"]), float(g["funcfl_cut"]),
                   float(g["funcfl_mass"]), g["funcfl_frho"], g["funcfl_zr"], g["funcfl_rhor"])
    natoms = sim.createAtom()
    sim.setup(adjust=True)
    sim.saveState()
    l0, c0 = sim.countPairs()

    def one():
        sim.restoreState()
        sim.setup(adjust=False)
        return sim.run(args.ntimes)

    one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rec, _ = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    l1, c1 = sim.countPairs()
    sim.setTiming(True)
    sim.resetKernelStats()
    _, tm = one()
    ks = sim.kernelStats()
    sim.setTiming(False)
    sim.close()
    f_ms = ks["force_ms"] / max(1, ks["force_launches"])
    L, C = 0.5 * (l0 + l1) / natoms, 0.5 * (c0 + c1) / natoms     # stored entries / in-cutoff entries per atom
    T = 8 if dp else 4
    if eam:
        flop, nbytes, kernel = 16 * L + 54 * C, 2 * 4 * L + 8 + 8 * T, "k_eam_density_v3 + k_eam_ghost_fp_v3 + k_eam_force_v3 (three passes, one force call)"
    elif half:
        flop, nbytes, kernel = 8 * L + 21 * C, 4 * L + 4 + 9 * T, "k_force_lj_half_v2 (reaction forces by RED.ADD)"
    else:
        flop, nbytes, kernel = 8 * L + 15 * C, 4 * L + 4 + 6.3 * T, "k_force_lj_full_fi<%s> (fused force + integrate)" % ("double" if dp else "float")
    peak = m.measure_fma_peak(m.DP if dp else m.SP, local)
    ach = flop * natoms / (f_ms * 1e-3) * 1e-12
    return {"name": name, "metric": METRIC if not eam else "atom-steps/sec (EAM Cu FCC, verletlist, three passes)",
            "config": "%s at %d^3 unit cells (%d atoms), %s, %d timesteps per bench step" % (name, args.nx, natoms, precision.upper(), args.ntimes),
            "value": natoms * args.ntimes * steps / (ms * 1e-3), "unit": UNIT, "dtype": "f64" if dp else "f32", "steps": steps,
            "ms_per_step": ms / steps,
            "roofline": {"kernel": kernel, "bound": "fp64" if dp else "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": ach / peak if peak else None, "traffic": None, "ms_per_launch": f_ms, "flop_per_atom_step": flop,
                         "algorithmic_bytes_per_atom_step": nbytes, "hbm_gbs": nbytes * natoms / (f_ms * 1e-3) * 1e-9,
                         "pairs_per_atom": {"stored": L, "in_cutoff": C},
                         "force_share_of_step": ks["force_ms"] / (tm["TOTAL"] * 1e3) if tm["TOTAL"] else None,
                         "neigh_ms_per_rebuild": ks["neigh_ms"] / max(1, ks["neigh_launches"])},
            "thermo_final": {"step": int(rec[-1][0]), "T": float(rec[-1][1]), "P": float(rec[-1][2])}}


def reference_cuda(args):
    """the reference's OWN CUDA variant (TOOLCHAIN=NVCC verletlist, rebuilt for sm_100 from the reference sources
    by oracle/Makefile ref-cuda, NUM_THREADS=128) on the same GPU with the same command line as the GPU arm's workload."""
    from refbind import ref_binary
    exe = ref_binary("vl_%s_aos-cuda" % args.precision)
    if exe is None:
        return {"unavailable": "oracle/_ref/MDBench-vl_%s_aos-cuda not built" % args.precision}
    nx = args.ref_nx or args.nx   # its host-side lattice generation, binning and ghost setup take 150 s at 128^3
    cmd = [exe, "-nx", str(nx), "-ny", str(nx), "-nz", str(nx), "-n", str(args.ntimes)]
    t0 = time.perf_counter()
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout
    except (OSError, subprocess.TimeoutExpired) as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}
    wall = time.perf_counter() - t0
    mm = re.search(r"TOTAL ([0-9.]+)s FORCE ([0-9.]+)s NEIGH ([0-9.]+)s", out)
    pp = re.search(r"Performance: ([0-9.]+) million atom updates per second", out)
    if not (mm and pp):
        return {"unavailable": "no report line in the output of %s" % " ".join(cmd)}
    nreb = max(1, args.ntimes // 20)
    return {"value": float(pp.group(1)) * 1e6, "unit": UNIT, "binary": "oracle/_ref/" + os.path.basename(exe), "atoms": 4 * nx ** 3,
            "command": " ".join(os.path.basename(c) if c == exe else c for c in cmd),
            "total_s": float(mm.group(1)), "force_s": float(mm.group(2)), "neigh_s": float(mm.group(3)),
            "force_ms_per_call": 1e3 * float(mm.group(2)) / (args.ntimes + 1), "neigh_ms_per_rebuild": 1e3 * float(mm.group(3)) / nreb,
            "wall_s": wall, "note": "the reference's kernels (forceCuda.cu / neighborCuda.cu) with its host-side binning and ghost setup; "
                                    "unmodified sources, compiled for sm_100"}


def parity_check(m, dist, rank, world, local, stream):
    """BASELINE config 1 exactly as stated (Cu FCC 32^3, 200 timesteps, DP) through the SAME path the timed run uses -- one
    domain at N = 1, one brick per GPU with the peer-store halo / NCCL migration at N > 1 -- against the reference's own
    numbers (tests/golden/thermo_lj.json, generated from the reference build): final T and P to rel 1e-10."""
    gold = [q for q in json.load(open(os.path.join(ROOT, "tests", "golden", "thermo_lj.json")))
            if q["variant"] == "vl_dp_aos" and q["nx"] == 32 and q["half"] == 0][0]
    P = m.default_params(precision=m.DP, nx=32, ny=32, nz=32, ntimes=200)
    if world > 1:
        uid = [m.dd_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim = m.Decomposition(P, m.dd_grid(world), nprocs=world, proc=rank, nccl_id=uid[0], device=local)
        path = "%dx%dx%d bricks, one per GPU: peer-store halo, NCCL migration" % m.dd_grid(world)
    else:
        sim = m.Simulation(P, device=local)
        path = "one domain"
    sim.setStream(stream)
    sim.createAtom()
    sim.setup(adjust=True)
    rec, _ = sim.run(200)
    sim.close()
    T, Pr = float(rec[-1][1]), float(rec[-1][2])
    rt, rp = abs(T - gold["T_full"]) / gold["T_full"], abs(Pr - gold["P_full"]) / gold["P_full"]
    printed = all(abs(a[1] - b[1]) <= 6e-7 * b[1] and abs(a[2] - b[2]) <= 6e-7 * b[2] for a, b in zip(rec, gold["records"]))
    return {"case": "BASELINE config 1: Cu FCC 32^3 (131072 atoms), LJ, DP, 200 timesteps, verletlist full lists", "path": path,
            "T": T, "P": Pr, "T_ref": gold["T_full"], "P_ref": gold["P_full"], "rel_T": rt, "rel_P": rp, "tol": 1e-10,
            "printed_records_equal": bool(printed and len(rec) == len(gold["records"])),
            "ok": bool(rt <= 1e-10 and rp <= 1e-10 and printed), "reference": "tests/golden/thermo_lj.json (reference build vl_dp_aos)"}


def parity_check_cp(m, dist, rank, world, local, stream):
    """BASELINE config 2 exactly as stated (Cu FCC 32^3, clusterpair 4x4, SP, 200 timesteps) through the SAME path the timed
    clusterpair run uses -- one domain at N = 1, one brick per GPU with ghost clusters over NCCL at N > 1 -- against the reference's
    scalar 4x4 build (tests/golden/thermo_cp.json, variant cpref44_sp): every thermo record to the stated SP tolerance, rel 1e-4."""
    gold = [q for q in json.load(open(os.path.join(ROOT, "tests", "golden", "thermo_cp.json"))) if q["variant"] == "cpref44_sp"][0]
    P = m.default_params(precision=m.SP, nx=32, ny=32, nz=32, ntimes=200)
    if world > 1:
        uid = [m.dd_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim = m.Decomposition(P, m.dd_grid(world), nprocs=world, proc=rank, nccl_id=uid[0], device=local, cluster_n=4)
        path = "%dx%dx%d bricks, one per GPU: ghost clusters and migration over NCCL" % m.dd_grid(world)
    else:
        sim = m.ClusterSimulation(P, cluster_n=4, device=local)
        path = "one domain"
    sim.setStream(stream)
    sim.createAtom()
    sim.setup(adjust=True)
    rec, _ = sim.run(200)
    sim.close()
    worst = max(max(abs(a[1] - b[1]) / b[1], abs(a[2] - b[2]) / b[2]) for a, b in zip(rec, gold["records"]))
    return {"case": "BASELINE config 2: Cu FCC 32^3 (131072 atoms), LJ, clusterpair 4x4, SP, 200 timesteps", "path": path,
            "T": float(rec[-1][1]), "P": float(rec[-1][2]), "T_ref": gold["records"][-1][1], "P_ref": gold["records"][-1][2],
            "worst_rel_over_records": float(worst), "tol": 1e-4, "ok": bool(len(rec) == len(gold["records"]) and worst <= 1e-4),
            "reference": "tests/golden/thermo_cp.json (reference build cpref44_sp: computeForceLJRef, M = N = 4)"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=128, help="unit cells per GPU per dimension")
    ap.add_argument("--ntimes", type=int, default=200)
    ap.add_argument("--precision", default="dp", choices=["dp", "sp"])
    ap.add_argument("--half", type=int, default=0)
    ap.add_argument("--ref-nx", type=int, default=64,
                    help="box of the reference arm and of reference_cuda (0 = the GPU arm's --nx; at 128 the reference needs 107 s per "
                         "bench step, 87 s of it serial start-up, i.e. 14 min for the default --steps 5 --warmup 3)")
    ap.add_argument("--ref-ntimes", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sort", action="store_true", help="A/B: re-sort the atoms by bin at every rebuild (SORT_ATOMS)")
    ap.add_argument("--opt", action="append", default=[], help="name=value for mdb_setOption (A/B)")
    ap.add_argument("--global-nx", type=int, default=0,
                    help="STRONG scaling (SURVEY 8d: 256^3 on 1/2/4/8 GPUs): the WHOLE box is G^3 unit cells whatever the GPU count")
    ap.add_argument("--bricks", default=None, help="gx,gy,gz: run a decomposed box on ONE GPU (debug / A-B)")
    ap.add_argument("--scheme", default="verletlist", choices=["verletlist", "clusterpair"],
                    help="OPT_SCHEME of the reference; clusterpair = GROMACS-style 4 x N cluster pairs (one GPU)")
    ap.add_argument("--cluster-n", type=int, default=4, choices=[4, 8])
    ap.add_argument("--no-parity", action="store_true", help="skip the config-1 parity run that precedes the timed region")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary measurement (BASELINE config 2 physics: clusterpair 4x4 SP at the same box)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = importlib.import_module("md-bench_b200")
    dp = args.precision == "dp"
    decomposed = world > 1 or args.bricks is not None
    stream = torch.cuda.current_stream().cuda_stream
    cp = args.scheme == "clusterpair"
    if args.global_nx and not decomposed:
        args.nx = args.global_nx
    parity = None
    if not args.no_parity and args.bricks is None and (not cp or (args.cluster_n == 4 and not dp and not args.half)):
        try:
            parity = (parity_check_cp if cp else parity_check)(m, dist, rank, world, local, stream)
        except Exception as e:   # reported, never hidden: a failed check makes the line say so
            parity = {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}
    if cp and not decomposed:
        grid = (1, 1, 1)
        P = m.default_params(precision=m.DP if dp else m.SP, layout=m.AOS, nx=args.nx, ny=args.nx, nz=args.nx,
                             ntimes=args.ntimes, half_neigh=args.half)
        sim = m.ClusterSimulation(P, cluster_n=args.cluster_n, device=local)
    elif decomposed:
        # spatial decomposition: one brick per GPU (or --bricks gx,gy,gz on one GPU), ghosts over NCCL/NVLink
        grid = tuple(int(v) for v in args.bricks.split(",")) if args.bricks else m.dd_grid(world)
        uid = [m.dd_unique_id() if (rank == 0 and world > 1) else None]
        if world > 1:
            dist.broadcast_object_list(uid, src=0)
        gbox = (args.global_nx,) * 3 if args.global_nx else (args.nx * grid[0], args.nx * grid[1], args.nx * grid[2])
        if any(gbox[k] % grid[k] for k in range(3)):
            raise SystemExit("bench.py: --global-nx must be a multiple of the brick grid %s" % (grid,))
        P = m.default_params(precision=m.DP if dp else m.SP, nx=gbox[0], ny=gbox[1], nz=gbox[2], ntimes=args.ntimes,
                             half_neigh=args.half)
        # clusterpair: ghost CLUSTERS from the neighbor bricks (cp_dd.cuh), NCCL send/recv between the processes
        sim = m.Decomposition(P, grid, nprocs=world, proc=rank, nccl_id=uid[0], device=local,
                              cluster_n=args.cluster_n if cp else 0)
    else:
        grid = (1, 1, 1)
        P = m.default_params(precision=m.DP if dp else m.SP, layout=m.AOS, nx=args.nx, ny=args.nx, nz=args.nx,
                             ntimes=args.ntimes, half_neigh=args.half)
        sim = m.Simulation(P, device=local)
    sim.setStream(stream)
    if args.sort and not cp:
        sim.setOption("sort_atoms", 1)
    for kv in args.opt:
        k, v = kv.split("=")
        sim.setOption(k, float(v))
    natoms = sim.createAtom()          # atoms of the WHOLE job
    nlocal0 = sim.counts()["Nlocal"]   # atoms on this rank
    sim.setup(adjust=True)
    sim.saveState()
    listed0, inside0 = (0, 0) if decomposed else sim.countPairs()

    def one_step():
        sim.restoreState()
        sim.setup(adjust=False)
        return sim.run(args.ntimes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    sim.resetKernelStats()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        rec, _ = one_step()
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = sim.kernelStats()["launches"]
    listed1, inside1 = (0, 0) if decomposed else sim.countPairs()
    value = natoms * args.ntimes * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (LJ force): CUDA-event time per launch, measured live ----
    fused_force = not cp and not args.half and not any(
        kv.split("=")[0] in ("fuse_force", "fuse_integrate") for kv in args.opt)
    sim.setTiming(True)
    sim.resetKernelStats()
    _, tm = one_step()
    ks = sim.kernelStats()
    sim.setTiming(False)
    f_ms = ks["force_ms"] / max(1, ks["force_launches"])
    per_launch_atoms = nlocal0 / (grid[0] * grid[1] * grid[2] // world)   # atoms one force launch covers
    peak_tf = m.measure_fma_peak(m.DP if dp else m.SP, local)
    flop_per_atom = FLOP_PER_ATOM_STEP
    if cp:
        # 8 flop per evaluated atom pair (every listed cluster pair = M x N atom pairs) + 15 more per pair
        # inside the cutoff; counts taken live from the current list (mean of the first and the last list of a run)
        flop_per_atom = (8.0 * 0.5 * (listed0 + listed1) * 4 * args.cluster_n + 15.0 * 0.5 * (inside0 + inside1)) / natoms
        if decomposed:   # no pair counter across bricks: the single-domain count of the same lattice and run length
            flop_per_atom = CP_FLOP_PER_ATOM_STEP[args.cluster_n]
    ach_tf = flop_per_atom * per_launch_atoms / (f_ms * 1e-3) * 1e-12
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    ach_gbs = BYTES_PER_ATOM_STEP[args.precision] * per_launch_atoms / (f_ms * 1e-3) * 1e-9
    traffic, traffic_src, ncu_pipes = None, None, None
    try:   # DRAM bytes of one launch from the committed ncu --set full capture of this very configuration, if there is one
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # decomposed verletlist runs launch the brick variant of the fused kernel (<.., XY, ZG>): its own capture
        te = tj.get("%s/%s/%d%s" % (args.scheme, args.precision, args.nx, "/brick" if (decomposed and not cp) else ""))
        if te and not args.half and (not decomposed or not cp) and (not cp or args.cluster_n == 4):
            # only a capture of the kernel that actually runs counts (the fused kernels replaced the session-3 ones)
            if cp or not fused_force or "_fi" in te.get("kernel", ""):
                traffic, traffic_src = te["bytes"], te["source"]
                ncu_pipes = dict(te.get("ncu") or {}, commit=te.get("commit"))   # the commit the capture was taken at
    except Exception:
        pass
    roofline = {"kernel": (("k_cp_force_lj_sp_duo<%d, FI> (two lanes per i-cluster, packed FP32, integrate halves in the epilogue)" % args.cluster_n)
                           if (not dp and not args.half) else
                           "k_cp_force_%s<%s,%d,%s>" % ("jl" if args.half else "lj", "double" if dp else "float", args.cluster_n, "half" if args.half else "full")) if cp
                else ("k_force_lj_full_fi<%s> (LJ force with finalIntegrate(n) + initialIntegrate(n+1) in its epilogue; flop and byte counts are the force part only)"
                      % ("double" if dp else "float")) if fused_force
                else "k_force_lj_%s<%s>" % ("half" if args.half else "full", "double" if dp else "float"),
                "bound": "fp64" if dp else "fp32", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach_tf / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                # pipe utilisations of the SAME committed ncu --set full capture the traffic comes from (not measured in this run)
                "ncu_capture": ncu_pipes,
                "algorithmic_bytes_per_launch": BYTES_PER_ATOM_STEP[args.precision] * per_launch_atoms if not cp else None,
                "peak_source": "measured in this run: FMA issue micro-benchmark (md-bench_b200/csrc/peaks.cu)",
                "ms_per_launch": f_ms, "flop_per_atom_step": flop_per_atom,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "bytes_per_atom_step": BYTES_PER_ATOM_STEP[args.precision], "peak_source": hbm_src},
                "force_share_of_step": ks["force_ms"] / (tm["TOTAL"] * 1e3) if tm["TOTAL"] else None,
                "neigh_ms_per_rebuild": ks["neigh_ms"] / max(1, ks["neigh_launches"]),
                "halo_ms_per_step": (ks["comm_ms"] / args.ntimes) if decomposed else None,
                "pairs_per_atom": None if decomposed else
                {"cluster_pairs_per_atom_t0": listed0 / natoms, "in_cutoff_t0": inside0 / natoms,
                 "cluster_pairs_per_atom_end": listed1 / natoms, "in_cutoff_end": inside1 / natoms} if cp else
                {"listed_t0": listed0 / natoms, "in_cutoff_t0": inside0 / natoms,
                 "listed_end": listed1 / natoms, "in_cutoff_end": inside1 / natoms}}

    # ---- e2e: same metric through the C ABI with HOST buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        sim.restoreState()
        real = np.float64 if dp else np.float32
        tdt = torch.float64 if dp else torch.float32
        ksteps = max(1, min(args.steps, 3))
        if decomposed:
            cap = nlocal0 + nlocal0 // 8 + 1024     # atoms migrate between ranks during a run
            htag = torch.empty(nlocal0, dtype=torch.int32, pin_memory=True)
            hx = torch.empty((3, nlocal0), dtype=tdt, pin_memory=True)
            hv = torch.empty((3, nlocal0), dtype=tdt, pin_memory=True)
            otag = torch.empty(cap, dtype=torch.int32, pin_memory=True)
            ox = torch.empty((3, cap), dtype=tdt, pin_memory=True)
            ov = torch.empty((3, cap), dtype=tdt, pin_memory=True)
            sim.get_into("x", htag.numpy(), hx.numpy())
            sim.get_into("v", None, hv.numpy())
        elif cp and not decomposed:   # positions AoS, velocities SoA (clusterpair/atom.h:66-92)
            hx = torch.empty((natoms, 3), dtype=tdt, pin_memory=True)
            hv = torch.empty((3, natoms), dtype=tdt, pin_memory=True)
            ox = torch.empty((natoms, 3), dtype=tdt, pin_memory=True)
            ov = torch.empty((3, natoms), dtype=tdt, pin_memory=True)
            sim.get("x", out=hx.numpy())
            sim.get("v", out=hv.numpy())
        else:
            hx = torch.empty((natoms, 3), dtype=tdt, pin_memory=True)
            hv = torch.empty((natoms, 3), dtype=tdt, pin_memory=True)
            ox = torch.empty((natoms, 3), dtype=tdt, pin_memory=True)
            ov = torch.empty((natoms, 3), dtype=tdt, pin_memory=True)
            sim.get("x", out=hx.numpy())
            sim.get("v", out=hv.numpy())
        for it in range(1 + ksteps):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            if decomposed:
                sim.setAtoms(htag.numpy(), hx.numpy(), hv.numpy())   # H2D from pinned host memory
                sim.setup(adjust=False)
                rec_e, _ = sim.run(args.ntimes)                      # thermo records come back D2H
                assert sim.counts()["Nlocal"] <= cap
                sim.get_into("x", otag.numpy(), ox.numpy())          # D2H of the final state
                sim.get_into("v", None, ov.numpy())
            elif cp and not decomposed:
                sim.setAtomsRaw(hx.numpy(), hv.numpy())
                sim.setup(adjust=False)
                rec_e, _ = sim.run(args.ntimes)   # ends with updateSingleAtoms: the atom arrays hold the final state
                sim.get("x", out=ox.numpy())
                sim.get("v", out=ov.numpy())
            else:
                sim.setAtoms(hx.numpy(), hv.numpy())
                sim.setup(adjust=False)
                rec_e, _ = sim.run(args.ntimes)
                sim.get("x", out=ox.numpy())
                sim.get("v", out=ov.numpy())
        barrier()
        te = time.perf_counter() - t0
        tt = torch.tensor([te], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
        nb = natoms * 3 * np.dtype(real).itemsize + (natoms * 4 if decomposed else 0)   # whole job
        e2e = {"value": natoms * args.ntimes * ksteps / te, "unit": UNIT, "h2d_bytes_per_step": 2 * nb,
               "d2h_bytes_per_step": 2 * nb + 8 * 3 * len(rec_e), "steps": ksteps,
               "final_T": float(rec_e[-1][1])}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline()
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": "failed: %s" % e}

    # ---- secondary: BASELINE config 2 physics (clusterpair 4x4, SP) at the same per-GPU box, N = 1 only ----
    secondary, secondaries, ref_cuda = None, None, None
    if world == 1 and not decomposed and not cp and not args.no_secondary:
        sim.close()
        sim = None
        try:
            secondary = clusterpair_secondary(m, args, local, stream)
        except Exception as e:   # the primary line must not depend on the secondary measurement
            secondary = {"error": "%s: %s" % (type(e).__name__, e)}
        # the other BASELINE configurations' physics at the same box (configs 1 SP, 3 half lists, 4 EAM), each with its roofline
        secondaries = []
        for kw in (dict(name="verletlist LJ full lists SP", precision="sp"),
                   dict(name="verletlist LJ half lists DP (BASELINE config 3 physics)", precision="dp", half=1),
                   dict(name="verletlist EAM DP (BASELINE config 4 physics, Cu_u3 tables)", precision="dp", eam=True),
                   dict(name="verletlist EAM SP", precision="sp", eam=True)):
            try:
                secondaries.append(verletlist_secondary(m, args, local, stream, **kw))
            except Exception as e:
                secondaries.append({"name": kw["name"], "error": "%s: %s" % (type(e).__name__, e)})
        if rank == 0 and not args.half:
            ref_cuda = reference_cuda(args)

    if rank == 0:
        cfg = workload_config(args, world, grid)
        if args.global_nx:   # strong scaling: the box is fixed, the per-GPU share shrinks with the GPU count
            g = args.global_nx
            cfg["workload"] = ("STRONG scaling: Cu FCC %d^3 unit cells (%d atoms) in total over %d GPU(s), " % (g, 4 * g ** 3, world)
                               + cfg["workload"].split("LJ sigma", 1)[1].join(["LJ sigma", ""]))
            cfg["global_box"] = "%dx%dx%d unit cells = %d atoms" % (g, g, g, 4 * g ** 3)
            cfg["nx_per_gpu"] = [g // grid[0], g // grid[1], g // grid[2]]
        line = {"metric": (METRIC_CP % (4, args.cluster_n)) if cp else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.global_nx else "weak",
                "vs_baseline": None, "dtype": "f64" if dp else "f32", "data": "synthetic", "config": cfg,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "secondary": secondary,
                "secondaries": secondaries, "reference_cuda": ref_cuda,
                "gpu_launches": int(launches), "clocks": clk,
                "thermo_final": {"step": int(rec[-1][0]), "T": float(rec[-1][1]), "P": float(rec[-1][2])}}
        print(json.dumps(line))
    if sim is not None:
        sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
