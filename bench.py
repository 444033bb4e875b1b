#!/usr/bin/env python
"""bench.py - throughput of the SPARTACUS-Surface solver hot path on B200.

Metric (BASELINE.json): column·g·layers / s, SW+LW, on the synthetic
vegetated-urban canopy of SURVEY.md §8(d): 1,048,576 columns x 16 layers,
nreg = 3, one SW and one LW interval (g = 2), FP64.  One "step" = one radsurf
pass over all columns of the rank.  Columns are independent, so ranks are
independent shards (no collective in the path); per-rank work is fixed and the
run reports weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams 2|4] [--columns C]
  python bench.py --impl reference ...   # CPU oracle (restatement of the Fortran) on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLAY = 16
FULL_COLUMNS = 1 << 20
UNIT = "column*g*layers/s"
METRIC = "column_g_layers_per_s_SW+LW"

# Algorithmic FLOPs per (column, layer) (SURVEY.md section 8d / App. C closed form).  "full": every layer
# charged the full-size (nreg = 3) count, the agreed numerator of SURVEY 8d; "seg": layers that solve only
# a sub-block of their regions (no vegetation in the layer: clear region only) charged with the order they
# actually solve - the reference does the same - while the sweeps stay full size.
def _layer_flops(n, d):
    N = 2 * n + d
    sw = 25 * n**3 + (38.67 + 8 * d) * n**3 + 2 / 3 * N**3 + 2 * N * N * d + 4 * n * N * d + 8 * d * n * n
    lw = 25 * n**3 + 34.67 * n**3 + 16 * n * n
    return sw, lw


def flop_table(ns, f_full, f_clear, f_veg):
    """per (column, layer): kernel family -> (flops "full" count, flops segment-aware)"""
    n, d, m = 3 * ns, 3, 4 * ns
    sw_full, lw_full = _layer_flops(n, d)
    sw_c, lw_c = _layer_flops(ns, 1)
    sw_v, lw_v = _layer_flops(2 * ns, 2)
    sw_sweep = 10.67 * n**3 + 22 * n * n + 4 * m * m + 6 * n * n * d
    lw_sweep = 10.67 * n**3 + 26 * n * n + 4 * m * m
    return {"sw_layer": (sw_full, f_full * sw_full + f_clear * sw_c + f_veg * sw_v),
            "lw_layer": (lw_full, f_full * lw_full + f_clear * lw_c + f_veg * lw_v),
            "sw_sweep": (sw_sweep, sw_sweep), "lw_sweep": (lw_sweep, lw_sweep)}


# Compulsory HBM bytes per (column, layer): 25 input + 40 output doubles + per-column terms (section 8d)
ALGO_BYTES_PER_COL_LAYER = 540.0


def sweep_bytes(ns, f_full, f_clear, f_veg):
    """Bytes a sweep kernel must move per (column, layer) given the layer / sweep kernel split (nreg = 3,
    urban): the layer matrices of the solved sub-block once in the upward and once in the fused downward
    sweep, interface state (a_above, d_above / source_above, LU factors) written and read once, geometry
    block, flux outputs and per-layer inputs."""
    def one(nr):
        n, d = nr * ns, nr
        sw_up = 2 * n * n + 2 * n * d + d * d                # R, T, S_dn, S_up, E
        sw_dn = 3 * n * n + 3 * n * d + 2 * d * d            # + int_diff, int_dir_diff, int_dir
        lw_up = 2 * n * n + n                                # R, T, source
        lw_dn = 3 * n * n + 2 * n + 10                       # + int_flux, int_flux_source, bookkeeping
        return sw_up + sw_dn, lw_up + lw_dn
    n, d = 3 * ns, 3
    sw_if, lw_if = 2 * n * n + n * d, 2 * n * n + n         # per interface, written once and read once
    uvg = 1 + 8                                              # segment (upward), geometry block (downward)
    sw = lw = 0.0
    for f, nr in ((f_full, 3), (f_clear, 1), (f_veg, 2)):
        a, b = one(nr)
        sw += f * a
        lw += f * b
    return {"sw_sweep": 8.0 * (sw + 2 * sw_if + uvg + 26 + 8), "lw_sweep": 8.0 * (lw + 2 * lw_if + uvg + 14 + 6)}


def make_config(streams):
    from spartacus_surface_b200 import config_type
    cfg = config_type(do_sw=True, do_lw=True, nsw=1, nlw=1, n_vegetation_region_urban=2,
                      n_vegetation_region_forest=2, n_stream_sw_urban=streams, n_stream_lw_urban=streams,
                      n_stream_sw_forest=streams, n_stream_lw_forest=streams)
    return cfg


def allocate_outputs(cfg, ncol, ntotlay, device=None, pinned=False):
    from spartacus_surface_b200 import canopy_flux_type, boundary_conds_out_type
    bc = boundary_conds_out_type().allocate(ncol, 1, 1, device=device)
    fl = [canopy_flux_type().allocate(cfg, ncol, ntotlay, 1, use_direct=d, do_save_flux_profile=False, device=device)
          for d in (True, True, False, False)]
    return bc, fl


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and throttle reasons of the samples that arrived in [t_begin, t_end]
        (the timed region); if the region was too short to catch one, of the samples since the
        sampler started (it starts before the warm-up steps, so those are under the same load)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if t_begin is None or t_begin <= t <= (t_end or t) + 0.15]
        window = "timed region" if inside else "warm-up and timed region"
        for r in (inside or [r for _, r in self.rows]):
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def host_threads():
    """Host cores this process may use (torchrun sets OMP_NUM_THREADS=1: ask the scheduler instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """CPU arm: the oracle (C++ restatement of the Fortran solver; the Fortran itself cannot be
    built: no Fortran compiler in the image) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from spartacus_surface_b200.synthetic import make_synthetic
    cfg = make_config(args.streams).consolidate(oracle_lib.legendre_gauss_init)
    ncol = args.cpu_columns
    cp, sw, lw = make_synthetic(cfg, ncol, NLAY)
    bc, fl = allocate_outputs(cfg, ncol, cp.ntotlay)
    threads = host_threads()
    solver = oracle_lib.make_solver(nthreads=threads, nblocksize=16)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc = solver(cfg, cp, sw, lw, bc, None, None, *fl)
        dt = time.perf_counter() - t0
        assert rc == 0
        if it >= args.warmup:
            times.append(dt)
    units = ncol * NLAY * 2
    total = sum(times)
    value = units * len(times) / total
    sample = f"{ncol} of {FULL_COLUMNS} synthetic columns per step (same generator and seed), OpenMP blocks of 16 columns"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.columns),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def head_to_host(obj, nc, ncol):
    """Host copy of the first `nc` columns of an API object whose double members live on the device."""
    import copy
    import numpy as np
    out = copy.copy(obj)
    for k, v in vars(obj).items():
        if type(v).__module__.startswith("torch"):
            rows = nc if v.shape[0] == ncol else nc * NLAY
            setattr(out, k, np.ascontiguousarray(v[:rows].cpu().numpy()))
        elif isinstance(v, np.ndarray) and v.shape and v.shape[0] == ncol:
            setattr(out, k, np.ascontiguousarray(v[:nc]))
    if hasattr(out, "ncol"):
        out.ncol, out.ntotlay = nc, nc * NLAY
    return out


def workload_config(args, ncol):
    return {"workload": f"synthetic vegetated-urban canopy, {ncol} columns x {NLAY} layers x (1 SW + 1 LW) "
                        f"intervals per GPU, nreg=3, {args.streams} streams per hemisphere",
            "columns_per_gpu": ncol, "layers": NLAY, "streams": args.streams, "nreg": 3,
            "parallelism": "independent column shards, no collective",
            "l2": "inputs+outputs per step (>= 9 GB at full size) exceed the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2, 3, 4, 8])
    ap.add_argument("--columns", type=int, default=FULL_COLUMNS, help="columns per GPU")
    ap.add_argument("--cpu-columns", type=int, default=131072, help="columns of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the generic kernels")
    ap.add_argument("--split", action="store_true", help="split layer / sweeps kernels instead of the column-resident ones")
    ap.add_argument("--fused-sort-group", type=int, default=None, help="tuning: -1 no column ordering, 0 whole chunk")
    ap.add_argument("--opt", action="append", default=[], help="tuning: name=value passed to ssb200_set_option")
    ap.add_argument("--sort-group", type=int, default=None, help="tuning: column ordering group (0 = whole chunk)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    from spartacus_surface_b200 import radsurf
    from spartacus_surface_b200._lib import load
    from spartacus_surface_b200.synthetic import make_synthetic, to_host

    lib = load()
    if not torch.cuda.is_available() or lib.ssb200_device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    lib.ssb200_set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: whatever libraries print there (the NCCL version
    # banner under torchrun) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    if args.generic:
        lib.ssb200_set_option(b"fast_kernels", 0)
    if args.split:
        lib.ssb200_set_option(b"fused_kernels", 0)
    if args.fused_sort_group is not None:
        lib.ssb200_set_option(b"fused_sort", 0 if args.fused_sort_group < 0 else 1)
        lib.ssb200_set_option(b"fused_sort_group", max(args.fused_sort_group, 0))
    for kv in args.opt:
        k, v = kv.split("=")
        assert lib.ssb200_set_option(k.encode(), int(v)) == 0, kv
    if args.sort_group is not None:
        lib.ssb200_set_option(b"sort_columns", 0 if args.sort_group < 0 else 1)
        lib.ssb200_set_option(b"sort_group", max(args.sort_group, 0))

    cfg = make_config(args.streams).consolidate()
    ncol = args.columns
    cp, sw, lw = make_synthetic(cfg, ncol, NLAY, col_offset=rank * ncol, device=device)
    bc, fl = allocate_outputs(cfg, ncol, cp.ntotlay, device=device)
    stream = torch.cuda.current_stream().cuda_stream
    status = torch.zeros(1, dtype=torch.int32, device=device)

    def step():
        rc = radsurf(cfg, cp, sw, lw, bc, None, None, *fl, stream=stream)
        assert rc == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = lib.ssb200_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    launches = lib.ssb200_kernel_launch_count() - launches0
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    units_per_rank = ncol * NLAY * 2
    value = world * units_per_rank * args.steps / (ms_max * 1e-3)

    # ---- conservation residuals and parity on a subsample (reported, rank 0) --------------------
    res = {}
    names = ("sw_norm_dir", "sw_norm_diff", "lw_internal", "lw_norm")
    for name, f in zip(names, fl):
        r = torch.zeros(ncol, dtype=torch.float64, device=device)
        s = f.as_struct()
        cps = cp.as_struct()
        rc = lib.ssb200_canopy_flux_check_device(C.byref(s), C.byref(cps), C.c_void_p(r.data_ptr()), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
        denom = f.top_dn[:, 0].abs().clamp_min(1e-30) if name != "lw_internal" else f.top_net[:, 0].abs().clamp_min(1e-30)
        res[name] = float((r.abs() / denom).max().item())
    nonfinite = int(sum((~torch.isfinite(getattr(f, k))).sum().item() for f in fl
                        for k in ("ground_net", "top_net", "clear_air_abs", "veg_abs", "wall_net", "roof_net")))

    # ---- end to end through the host-pointer C ABI entry (pinned host buffers), every rank at once ----
    e2e = None
    if args.e2e_steps > 0:
        hcp, hsw, hlw = to_host(cp), to_host(sw), to_host(lw)
        pin = []
        for obj in (hcp, hsw, hlw):
            for k, v in list(vars(obj).items()):
                if isinstance(v, np.ndarray) and v.dtype == np.float64:
                    tt = torch.from_numpy(v).pin_memory()
                    pin.append(tt)
                    setattr(obj, k, tt.numpy())
        hbc, hfl = allocate_outputs(cfg, ncol, hcp.ntotlay)
        for obj in [hbc] + hfl:
            for k, v in list(vars(obj).items()):
                if isinstance(v, np.ndarray) and v.dtype == np.float64:
                    tt = torch.from_numpy(v).pin_memory()
                    pin.append(tt)
                    setattr(obj, k, tt.numpy())
        h2d = sum(v.nbytes for obj in (hcp, hsw, hlw) for v in vars(obj).values()
                  if isinstance(v, np.ndarray) and v.dtype == np.float64)
        d2h = sum(v.nbytes for obj in [hbc] + hfl for v in vars(obj).values()
                  if isinstance(v, np.ndarray) and v.dtype == np.float64)
        radsurf(cfg, hcp, hsw, hlw, hbc, None, None, *hfl)  # warm-up (allocates the device mirrors)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rc = radsurf(cfg, hcp, hsw, hlw, hbc, None, None, *hfl)  # returns when the outputs are on the host
            assert rc == 0
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * units_per_rank * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
               "ms_per_step": 1e3 * dt / args.e2e_steps,
               "note": "ssb200_radsurf with pinned host arrays on every rank at once: H2D of every input, kernels, "
                       "D2H of every output; host wall clock, max over ranks; bytes summed over ranks"}
        # the device-resident result equals the host-path result (same kernels)
        dev_top = fl[0].top_net[:4096, 0].cpu().numpy()
        assert np.array_equal(dev_top, hfl[0].top_net[:4096, 0]), "device and host entry disagree"
        del pin, hcp, hsw, hlw, hbc, hfl

    out = None
    if rank == 0:
        # ---- per-kernel times (library CUDA events on the launch stream), separate pass ---------
        lib.ssb200_set_profiling(1)
        step()
        torch.cuda.synchronize()
        tms, cnt = (C.c_double * 5)(), (C.c_int64 * 5)()
        lib.ssb200_last_kernel_times_ms(tms)
        lib.ssb200_last_kernel_counts(cnt)
        lib.ssb200_set_profiling(0)
        fam = ["sw_layer", "sw_sweep", "lw_layer", "lw_sweep", "surface"]
        kt = {fam[i]: {"ms": tms[i], "launches": int(cnt[i])} for i in range(5)}
        fp64_peak = lib.ssb200_measure_fp64_peak_tflops(1 << 15)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # segment mix of the layer problems (which sub-block of regions a layer solves), from the inputs
        vfr, bfr = cp.veg_fraction, cp.building_fraction
        min_veg = cfg.min_vegetation_fraction
        f_clear = float((vfr <= min_veg).double().mean().item())
        f_veg = float(((vfr > min_veg) & ((1.0 - bfr - vfr) <= min_veg)).double().mean().item())
        f_full = 1.0 - f_clear - f_veg
        fl_tab = flop_table(args.streams, f_full, f_clear, f_veg)
        fam_bytes = sweep_bytes(args.streams, f_full, f_clear, f_veg)
        traffic_tab = {}
        try:  # DRAM bytes per (column, layer) of each kernel family from the committed ncu --set full capture
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))
        except (OSError, ValueError):
            pass
        step_ms = ms_max / args.steps
        roofline_kernels = {}
        for k in ("sw_layer", "lw_layer", "sw_sweep", "lw_sweep"):
            if kt[k]["launches"] == 0 or kt[k]["ms"] <= 0.0:
                continue  # column-resident kernels: the whole pass is one launch, booked as *_layer
            n_l = max(1, kt[k]["launches"])
            per_ms = kt[k]["ms"] / n_l
            units = ncol * NLAY / n_l  # (column, layer) pairs of one launch (one chunk of columns)
            tf_full = fl_tab[k][0] * units / (per_ms * 1e-3) / 1e12
            tf_seg = fl_tab[k][1] * units / (per_ms * 1e-3) / 1e12
            tr = traffic_tab.get(f"{k}_s{args.streams}")
            ent = {"avg_launch_ms": per_ms, "launches_per_step": kt[k]["launches"],
                   "column_layers_per_launch": units,
                   "algorithmic_flops_per_column_layer": fl_tab[k][0],
                   "algorithmic_flops_per_column_layer_segment_aware": fl_tab[k][1],
                   "fp64_tflops": tf_full, "fp64_frac": tf_full / fp64_peak if fp64_peak > 0 else None,
                   "fp64_frac_segment_aware": tf_seg / fp64_peak if fp64_peak > 0 else None,
                   "traffic": tr["dram_bytes_per_column_layer"] * units if tr else None}
            if k in fam_bytes:
                gb = fam_bytes[k] * units / (per_ms * 1e-3) / 1e9
                ent.update({"bound": "hbm", "algorithmic_bytes_per_column_layer": fam_bytes[k], "gbs": gb,
                            "hbm_frac": gb / hbm_peak})
                if ent["traffic"]:
                    ent["dram_gbs_from_ncu_traffic"] = ent["traffic"] / (per_ms * 1e-3) / 1e9
            else:
                ent["bound"] = "fp64"
            roofline_kernels[k] = ent
        dom = max(roofline_kernels, key=lambda k: kt[k]["ms"])
        e = roofline_kernels[dom]
        if e["bound"] == "fp64":
            roofline = {"bound": "fp64", "kernel": dom, "achieved": e["fp64_tflops"], "peak": fp64_peak,
                        "unit": "TFLOP/s", "frac": e["fp64_frac"], "traffic": e["traffic"],
                        "frac_segment_aware": e["fp64_frac_segment_aware"],
                        "flops_note": "achieved = SURVEY 8(d) algorithmic flops per (column, layer) x pairs per launch / "
                                      "launch time; SURVEY charges every layer the full nreg=3 count - "
                                      "frac_segment_aware charges layers without vegetation the order they solve"}
        else:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": e["gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": e["hbm_frac"], "traffic": e["traffic"],
                        "bytes_note": "algorithmic bytes: layer matrices of the solved sub-block read in the upward and "
                                      "in the fused downward sweep, interface state written and read once, "
                                      "geometry block, flux outputs (DESIGN.md section 4.3)"}
        roofline.update({"avg_launch_ms": e["avg_launch_ms"], "launches_per_step": e["launches_per_step"],
                         "segment_mix": {"all_regions": f_full, "clear_only": f_clear, "vegetated_only": f_veg},
                         "fp64_peak_source": "measured in this run: register-resident independent DFMA chains on "
                                             "all SMs (ssb200_measure_fp64_peak_tflops); MEASURED_PEAKS.json has "
                                             "no FP64 entry",
                         "hbm_peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "traffic_source": "profiles/r01_dram_traffic.json (ncu dram__bytes_read+write per kernel "
                                           "family, scaled to the launch)" if e["traffic"] else None})
        all_full = sum(v[0] for v in fl_tab.values()) * ncol * NLAY
        all_seg = sum(v[1] for v in fl_tab.values()) * ncol * NLAY
        roofline["whole_step"] = {"bound": "fp64", "achieved": all_full / (step_ms * 1e-3) / 1e12,
                                  "frac": all_full / (step_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None,
                                  "frac_segment_aware": all_seg / (step_ms * 1e-3) / 1e12 / fp64_peak
                                  if fp64_peak > 0 else None,
                                  "algorithmic_flops_per_column_layer": sum(v[0] for v in fl_tab.values()),
                                  "note": "SURVEY 8(d) flop count of the reference formulation / step time"}
        gbs = ALGO_BYTES_PER_COL_LAYER * ncol * NLAY / (step_ms * 1e-3) / 1e9
        roofline_hbm = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                        "note": "compulsory input+output bytes of the whole path only (540 B per column-layer)"}

        # ---- CPU baseline beside it (oracle on the host cores, bounded sample) -------------------
        cpu = None
        parity_err = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            import parity
            from spartacus_surface_b200.radsurf_canopy_flux import ALL_FIELDS
            nc = min(args.cpu_columns, ncol)
            # the CPU sample uses the very arrays the GPU solved (first nc columns, copied from HBM)
            ccp, csw, clw = head_to_host(cp, nc, ncol), head_to_host(sw, nc, ncol), head_to_host(lw, nc, ncol)
            cbc, cfl = allocate_outputs(cfg, nc, ccp.ntotlay)
            cpu_threads = host_threads()
            solver = oracle_lib.make_solver(nthreads=cpu_threads)
            solver(cfg, ccp, csw, clw, cbc, None, 256, *cfl)  # touch pages / warm caches
            t0 = time.perf_counter()
            rc = solver(cfg, ccp, csw, clw, cbc, None, None, *cfl)
            dt = time.perf_counter() - t0
            assert rc == 0
            cpu = {"value": nc * NLAY * 2 / dt, "unit": UNIT, "cores": cpu_threads,
                   "kind": "port",
                   "sample": f"first {nc} of the {ncol} columns, one pass, OpenMP dynamic blocks of 16 columns "
                             f"({dt:.1f} s); C++ restatement of the reference (no Fortran compiler in the image)"}
            # parity of the GPU result on the same columns (max over fields of err relative to field max)
            got = {n: {k: getattr(f, k)[:nc * (NLAY if getattr(f, k).shape[0] != ncol else 1)].cpu().numpy()
                       for k in ALL_FIELDS if getattr(f, k) is not None} for n, f in zip(names, fl)}
            exp = {n: {k: getattr(f, k) for k in ALL_FIELDS if getattr(f, k) is not None} for n, f in zip(names, cfl)}
            errs = parity.field_errors(got, exp)
            flux_fields = [k for k in errs if "sunlit" not in k[1]]
            # verdict with the tolerance of tests/parity.py on a 4096-column subsample (needs the no-FMA oracle)
            ns_ = min(4096, nc)
            scp, ssw, slw = head_to_host(cp, ns_, ncol), head_to_host(sw, ns_, ncol), head_to_host(lw, ns_, ncol)
            outs = []
            for nofma in (False, True):
                sbc, sfl = allocate_outputs(cfg, ns_, scp.ntotlay)
                oracle_lib.make_solver(nthreads=cpu_threads, nofma=nofma)(cfg, scp, ssw, slw, sbc, None, None, *sfl)
                outs.append({n: {k: getattr(f, k) for k in ALL_FIELDS if getattr(f, k) is not None}
                             for n, f in zip(names, sfl)})
            gsub = {n: {k: v[:ns_ * (NLAY if v.shape[0] != nc else 1)] for k, v in f.items()} for n, f in got.items()}
            ok, worst, lines = parity.check(gsub, outs[0], outs[1])
            parity_err = {"within_tolerance": bool(ok), "max_err_over_bound": worst,
                          "tolerance": "per field max(1e-9, 50 x oracle FMA/no-FMA sensitivity), tests/parity.py",
                          "max_rel_err_fluxes": max(errs[k] for k in flux_fields),
                          "max_rel_err_sunlit_fractions": max([errs[k] for k in errs if "sunlit" in k[1]] or [0.0]),
                          "columns": nc, "definition": "max|gpu-oracle| / max|oracle| per field"}

        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, ncol),
            "roofline": roofline, "roofline_kernels": roofline_kernels, "roofline_hbm": roofline_hbm,
            "kernel_times_one_step": kt,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "conservation_max_abs_residual_over_top_flux": res, "nonfinite_outputs": nonfinite,
            "parity_vs_oracle": parity_err, "library": lib.ssb200_version().decode(),
            "kernels": "generic" if args.generic else ("register-resident, split layer / sweeps" if args.split
                                                       else "column-resident (fused) where available"),
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
