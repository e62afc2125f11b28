#!/usr/bin/env python
"""
bench.py -- ray-surface intercepts/s of the fused sequential raytrace on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8d cfg 2): spherical
concave grating spectrograph with a constant line-spacing ruling,
100 x 100 field x 100 x 100 pupil = 1e8 rays x 64 wavelengths, 3 surfaces
(object, grating, sensor), fp64.

One "step" is one pass of the hot path over the whole batch: 64 launches of the
fused trace kernel, one per wavelength slab of 1e8 rays.

* ``value``     device-resident: the 1e8-ray slab is a dense structure of arrays
                in HBM (81 B/ray read, 81 B/ray written = 162 algorithmic bytes
                per ray, HBM-bound; every launch streams 16 GB, far more than L2).
* ``e2e``       the same metric through the public API with HOST buffers:
                ``SequentialSystem.image_rays`` (separable wavelength / field /
                pupil axes from host memory -> fused trace + detector binning ->
                detector planes back on the host); copies inside the timed region.
* ``--impl reference``  the reference's CPU algorithm (the NumPy oracle port:
                the reference itself cannot be imported without named_arrays /
                astropy) on all host cores, on a bounded sample of the workload.

Launch:  python bench.py --gpus N --steps K --warmup W        (N = 1)
         python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""

from __future__ import annotations
import argparse
import ctypes as C
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "ray-surface intercepts/sec"
UNIT = "intercepts/s"
N_SURFACES = 3
BYTES_PER_RAY = 162  # 10 fp64 + 1 mask byte in, the same out
FLOP_PER_RAY = 410  # algorithmic flop of cfg 2 (119 + 165 + 126)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num-field", type=int, default=100)
    ap.add_argument("--num-pupil", type=int, default=100)
    ap.add_argument("--num-wavelength", type=int, default=64)
    ap.add_argument("--cpu-sample-rays", type=int, default=20_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-rays", type=int, default=20_000_000, help="rays of the host-array e2e sample")
    ap.add_argument("--strong", default="cfg3,cfg5", help="strong-scaling image simulations to run (cfg3,cfg5 or none)")
    ap.add_argument("--strong-steps", type=int, default=2)
    ap.add_argument("--only-strong", action="store_true", help="skip the cfg 2 sections (development)")
    return ap.parse_args()


def workload_name(args) -> str:
    return (
        f"cfg2 spherical concave grating, constant ruling spacing: "
        f"{args.num_field}x{args.num_field} field x {args.num_pupil}x{args.num_pupil} pupil "
        f"x {args.num_wavelength} wavelengths, {N_SURFACES} surfaces"
    )


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores, bounded sample
# ---------------------------------------------------------------------------
def cpu_reference(args, sample_rays: int, repeats: int = 3) -> dict:
    """
    Time ``oracle.raytrace.propagate_rays`` on `sample_rays` rays of the workload on the host cores, the way
    BASELINE.md section 4 prescribes: whole-array fp64 NumPy per operator step, Snell's law through the numba
    ``guvectorize`` kernel the reference itself uses (``optika/materials/_snells_law.py:294-366``; restated
    in ``oracle/snell_numba.py``).  Two ways of using the cores are timed and the faster one is reported:

    * ``as the reference runs``: one Python thread over the whole array, numba's ``target="parallel"`` pool
      for the Snell step (everything else in the reference is single-threaded NumPy);
    * ``split``: the rays cut into one chunk per core, every chunk traced by its own thread (NumPy and the
      serial numba kernel release the GIL) -- more parallelism than the reference has.
    """
    import configs
    from concurrent.futures import ThreadPoolExecutor
    from oracle import raytrace as ora

    cores = os.cpu_count() or 1
    n_pupil = max(2, int(round((sample_rays / 100) ** 0.5)))
    system = configs.spherical_grating(num_field=10, num_pupil=n_pupil, num_wavelength=1)
    _, rays = system._input(None, None, None, None, False, False)
    r0, _ = configs.flatten_rays(rays)
    r0 = {k: np.ascontiguousarray(v.reshape(-1)) for k, v in r0.items()}
    n = r0["px"].size
    surfaces = system.surfaces_all
    chunks = np.array_split(np.arange(n), cores)

    def work(idx):
        sub = {k: v[idx[0] : idx[-1] + 1] for k, v in r0.items()}
        return ora.propagate_rays(surfaces, sub)

    timings = {}
    try:
        have_numba = ora.use_numba_snell("parallel") == "parallel"
        numba_threads = None
        if have_numba:
            from oracle import snell_numba

            numba_threads = snell_numba.threads()
            ora.propagate_rays(surfaces, {k: v[:1000] for k, v in r0.items()})  # compile outside the timing
            best = None
            for _ in range(repeats):
                t0 = time.perf_counter()
                ora.propagate_rays(surfaces, r0)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            timings["as the reference runs (NumPy on one thread + numba parallel Snell)"] = best
        ora.use_numba_snell("serial" if have_numba else None)
        best = None
        with ThreadPoolExecutor(max_workers=cores) as pool:
            list(pool.map(work, [c[:100] for c in chunks]))
            for _ in range(repeats):
                t0 = time.perf_counter()
                list(pool.map(work, chunks))
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        timings[f"rays split over {cores} threads" + (" + numba serial Snell" if have_numba else " (no numba: NumPy Snell)")] = best
    finally:
        ora.use_numba_snell(None)
    mode, best = min(timings.items(), key=lambda kv: kv[1])
    return dict(
        value=n * N_SURFACES / best,
        unit=UNIT,
        cores=cores,
        kind="port",
        sample=f"{n} rays x {N_SURFACES} surfaces of the cfg2 system (one wavelength), NumPy oracle port with the "
        f"reference's numba Snell kernel, {mode}, best of {repeats}",
        seconds=best,
        modes={k: n * N_SURFACES / v for k, v in timings.items()},
        numba_threads=numba_threads,
    )


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    for _ in range(args.warmup):
        cpu_reference(args, max(20000, args.cpu_sample_rays // 20))
    values = []
    seconds = []
    for _ in range(args.steps):
        r = cpu_reference(args, args.cpu_sample_rays)
        values.append(r["value"])
        seconds.append(r["seconds"])
    value = float(np.mean(values))
    line = dict(
        impl="reference",
        metric=METRIC,
        value=value,
        unit=UNIT,
        n_gpus=args.gpus,
        steps=args.steps,
        warmup=args.warmup,
        ms_per_step=1e3 * float(np.mean(seconds)),
        higher_is_better=True,
        scaling="weak",
        vs_baseline=None,
        dtype="f64",
        data="synthetic",
        config=dict(workload=workload_name(args), sample=r["sample"]),
        cpu_baseline=dict(value=value, unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        wall_s=time.perf_counter() - t_start,
    )
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """
    SM clock, power and throttle reasons DURING the timed region.  NVML is polled from a thread
    every 10 ms (the timed region of three steps is under half a second: `nvidia-smi -lms` often
    has not produced its first line by then); `nvidia-smi` remains the fallback.
    """

    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index: int):
        self.rows = []
        self.samples = []  # (sm MHz, power W, reasons bitmask) from NVML
        self.proc = None
        self.thread = None
        self.gpu_index = gpu_index
        self.nvml = None
        self.handle = None
        self.sm_max = None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = None
            try:
                # the CUDA device by its UUID: NVML numbers the GPUs of the box, CUDA the visible ones
                import torch

                uuid = getattr(torch.cuda.get_device_properties(gpu_index), "uuid", None)
                if uuid is not None:
                    name = str(uuid)
                    name = name if name.startswith("GPU-") else "GPU-" + name
                    for candidate in (name, name.encode()):
                        try:
                            self.handle = pynvml.nvmlDeviceGetHandleByUUID(candidate)
                            break
                        except Exception:
                            continue
            except Exception:
                self.handle = None
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons"
        )
        while True:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                power = float(nv.nvmlDeviceGetPowerUsage(self.handle)) / 1e3
                self.samples.append((sm, power, int(reasons(self.handle))))
            except Exception:
                pass
            if self._stop.wait(0.01):
                break

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop.set()
            if self.thread is not None:
                self.thread.join(timeout=2)
            if not self.samples:
                return dict(sm_mhz=None, sm_max_mhz=self.sm_max, reasons=["no samples"])
            mask = 0
            for _, _, bits in self.samples:
                mask |= bits
            return dict(
                sm_mhz=float(np.median([s[0] for s in self.samples])),
                sm_max_mhz=self.sm_max,
                power_w_max=float(np.max([s[1] for s in self.samples])),
                samples=len(self.samples),
                reasons=sorted(name for name, bit in self.REASON_BITS.items() if mask & bit),
                source="NVML, 10 ms polling during the timed region",
            )
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(
            sm_mhz=float(np.median(sm)),
            sm_max_mhz=float(np.max(smax)),
            power_w_max=float(np.max(power)),
            samples=len(sm),
            reasons=sorted(reasons),
            source="nvidia-smi -lms 100 during the timed region",
        )



# ---------------------------------------------------------------------------
# strong scaling of the image simulations north_star names (cfg 3: 1e9 rays; cfg 5: 1e10 rays into
# eight 4096 x 4096 images): the WHOLE job is fixed, every rank draws, traces and bins a slab of the ray
# grid on chip, and the planes are summed over the ranks and read back to the host while the next
# configuration is traced (optika_b200.distributed.ImagePipeline).  Collective and read-back are
# inside the timed region.
# ---------------------------------------------------------------------------
def strong_scaling(name, system, grid_spec, n_surfaces, flop_per_ray, device, rank, world, steps, warmup):
    import torch
    import torch.distributed as dist
    from optika_b200 import _engine, distributed, named as na

    wavelength, field, pupil, axes = grid_spec
    t_setup = time.perf_counter()
    grids = system.ray_grids(
        1.0, wavelength, field, pupil, axes[0], axes[1:3], axes[3:5],
        normalized_field=False, normalized_pupil=False, random=True, seed=0,
    )
    w = np.asarray(wavelength.ndarray, dtype=float)
    w_edges = np.array([w.min(), w.max()])  # integrate=True: one spectral bin (_sequential.py:1189-1196)
    compiled = system._compiled_local
    ex, ey = system.sensor.pixel_edges()
    leading = tuple(compiled.shape.values())
    n_rays = sum(g.size for g in grids)
    stream = torch.cuda.current_stream(device)

    def make(local, counts=True, moments=True):
        image = _engine.DeviceImage.zeros(
            w_edges, ex, ey, device, leading=leading, moments=moments, counts=counts, fused=True, pad_to=1 if local else world
        )
        return distributed.ImagePipeline(image, device, local=local, rezero=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def run(pipeline, shard, n_steps, n_warm, collective=True):
        """ms per exposure (device, this rank), stage times, host planes of the last exposure."""
        marks = []

        def on_launch(c, phase):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            marks.append(e)

        totals, stages, planes, wall = [], [], None, []
        for k in range(n_warm + n_steps):
            timed = k >= n_warm
            pipeline.timing = timed
            marks.clear()
            if collective:
                barrier()
            else:
                torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            # (the planes are zero: allocated so, and handed back zeroed by the pipeline after every exposure)
            planes = system.collect_grids(
                grids, w_edges, device=device, pipeline=pipeline, shard=shard, on_launch=on_launch if timed else None
            )
            # collect_grids returned after pipeline.finish(): side stream drained, all ranks done
            stream.wait_stream(pipeline.side)
            e1.record(stream)
            torch.cuda.synchronize(device)
            wall.append(time.perf_counter() - t0)
            if timed:
                totals.append(e0.elapsed_time(e1))
                st = pipeline.stage_ms()
                st["ms_trace"] = sum(marks[i].elapsed_time(marks[i + 1]) for i in range(0, len(marks), 2))
                stages.append(st)
        mean = lambda key: float(np.mean([st[key] for st in stages]))  # noqa: E731
        return (
            float(np.mean(totals)), dict(ms_trace=mean("ms_trace"), ms_reduce=mean("ms_reduce"), ms_d2h=mean("ms_d2h")),
            planes, float(np.mean(wall[n_warm:])),
        )

    pipeline = make(local=(world == 1))
    transport = pipeline.transport
    setup_s = time.perf_counter() - t_setup
    ms, stages, planes, wall = run(pipeline, shard=True, n_steps=steps, n_warm=warmup)
    counts_sharded = np.array(planes["counts"]) if rank == 0 and world > 1 else None
    binned = int(planes["counts"].sum()) if rank == 0 else 0
    flux_total = float(planes["flux"].sum()) if rank == 0 else 0.0
    d2h_total = sum(int(v.nbytes) for v in planes.values())
    t = torch.tensor([ms, stages["ms_trace"], stages["ms_reduce"], stages["ms_d2h"], wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_trace, ms_reduce, ms_d2h, ms_wall = [float(v) for v in t.tolist()]
    out = dict(
        workload=name,
        rays=n_rays,
        surfaces=n_surfaces,
        n_gpus=world,
        scaling="strong",
        ms_total=ms_total,
        ms_wall=ms_wall,
        ms_trace=ms_trace,
        ms_reduce=ms_reduce,
        ms_d2h=ms_d2h,
        stage_note="ms_total: CUDA events on the launch stream around zeroing, all launches, the reduce-scatter and "
                   "the device -> host copies (max over ranks); ms_trace / ms_reduce / ms_d2h: busy time of each stage "
                   "(max over ranks); reduce and copy of configuration k run on a side stream under the trace of k + 1",
        intercepts_per_s=n_rays * n_surfaces / (ms_total * 1e-3),
        fp64_tflops_algorithmic=flop_per_ray * n_rays / (ms_total * 1e-3) / 1e12,
        images=f"{int(np.prod(leading)) if leading else 1} x {len(ex) - 1} x {len(ey) - 1} pixels, planes flux / flux cos / counts",
        d2h_bytes=d2h_total,
        d2h_bytes_per_rank=d2h_total // world,
        note="this object times THREE planes per image (flux, flux x cos(incidence), hit counts): the counts exist so that "
             "the sharded + reduced image can be compared with the one-GPU image count for count; at N = 8 their read-back "
             "(3.2 GB for cfg 5) is bound by what the host can ingest, not by the trace. `two_planes` / `one_plane` time "
             "the same exposure with the planes SequentialSystem.image needs: flux x cos only for sensor materials that "
             "use the angle of incidence, flux alone for the default IdealSensorMaterial",
        collective=({
            "peer": "every rank pulls its 1/N slice of every other rank's planes over NVLink with the copy engines "
                    "(CUDA IPC peer memory, interprocess events), adds the N pieces, and copies the sum to a shared "
                    "page-locked host buffer over its own PCIe link",
            "nccl": "reduce_scatter (NCCL), one per dtype and configuration; every rank copies its 1/N slice to a "
                    "shared page-locked host buffer over its own PCIe link",
        }.get(transport, transport)) if world > 1 else "none (one GPU)",
        transport=transport,
        shard_axis=(["wavelength", "field_x", "field_y", "pupil_x", "pupil_y"][distributed.best_shard_axis(grids[0].count, world)]
                    if world > 1 else None),
        rays_binned=binned,
        binned_fraction=binned / n_rays if rank == 0 else None,
        flux_total=flux_total,
        setup_s=setup_s,
        steps=steps,
        warmup=warmup,
    )
    if world > 1:
        # the same job on rank 0 ALONE, same run (no collective; the other ranks wait): the denominator of
        # the strong-scaling efficiency, and the image the sharded + reduced one must equal count for count
        n1 = None
        if rank == 0:
            single = make(local=True)
            ms1, st1, planes1, wall1 = run(single, shard=False, n_steps=max(1, min(steps, 2)), n_warm=1, collective=False)
            n1 = dict(ms_total=ms1, ms_wall=wall1 * 1e3, **st1)
            n1["counts_equal"] = bool(np.array_equal(planes1["counts"], counts_sharded))
            n1["flux_max_rel_diff"] = float(
                np.max(np.abs(planes1["flux"] - planes["flux"])) / max(float(np.max(np.abs(planes1["flux"]))), 1e-300)
            )
            del planes1
            single.close()
            del single
        dist.barrier()
        if rank == 0:
            out["n1_same_run"] = n1
            out["efficiency_vs_n1"] = n1["ms_total"] / (world * ms_total)
            out["counts_equal_n1"] = n1["counts_equal"]
    del planes
    pipeline.close()
    del pipeline
    torch.cuda.empty_cache()

    # The same exposure with fewer planes read back (the hit counts above exist for the exact comparison between N
    # ranks and one).  `two_planes`: flux and flux x cos(incidence), what `SequentialSystem.image` needs for a sensor
    # material that depends on the angle of incidence; `one_plane`: flux only, what it reads back for the default
    # IdealSensorMaterial (optika/sensors/materials/_materials.py:1566-1643 ignores the direction).
    def variant(moments, planes_text):
        lean = make(local=(world == 1), counts=False, moments=moments)
        ms2, st2, planes2, _ = run(lean, shard=True, n_steps=steps, n_warm=warmup)
        d2h2 = sum(int(v.nbytes) for v in planes2.values())
        t2 = torch.tensor([ms2, st2["ms_trace"], st2["ms_reduce"], st2["ms_d2h"]], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        res = dict(zip(("ms_total", "ms_trace", "ms_reduce", "ms_d2h"), [float(v) for v in t2.tolist()]))
        res.update(d2h_bytes=d2h2, planes=planes_text, intercepts_per_s=n_rays * n_surfaces / (res["ms_total"] * 1e-3))
        del planes2
        lean.close()
        del lean
        torch.cuda.empty_cache()
        if world > 1:
            n1 = None
            if rank == 0:
                single = make(local=True, counts=False, moments=moments)
                ms1, st1, planes1, _ = run(single, shard=False, n_steps=max(1, min(steps, 2)), n_warm=1, collective=False)
                n1 = dict(ms_total=ms1, **st1)
                del planes1
                single.close()
                del single
            dist.barrier()
            if rank == 0:
                res["n1_same_run"] = n1
                res["efficiency_vs_n1"] = n1["ms_total"] / (world * res["ms_total"])
            torch.cuda.empty_cache()
        return res

    two = variant(True, "flux / flux cos (SequentialSystem.image with a sensor material that uses the angle of incidence)")
    out["one_plane"] = variant(False, "flux (SequentialSystem.image with the default IdealSensorMaterial)")
    out["two_planes"] = two
    return out


def strong_scaling_configs(device, rank, world, steps, warmup, which):
    import configs
    from optika_b200 import named as na, units as u

    axes = ("wavelength", "field_x", "field_y", "pupil_x", "pupil_y")

    def grid(w_lo, w_hi, n_w, half_field, n_field, half_pupil, n_pupil):
        wavelength = na.ScalarArray(np.linspace(w_lo, w_hi, n_w + 1), axes[0])
        field = na.Cartesian2dVectorArray(
            na.ScalarArray(np.linspace(-half_field[0], half_field[0], n_field[0] + 1), axes[1]),
            na.ScalarArray(np.linspace(-half_field[1], half_field[1], n_field[1] + 1), axes[2]),
        )
        pupil = na.Cartesian2dVectorArray(
            na.ScalarArray(np.linspace(-half_pupil[0], half_pupil[0], n_pupil[0] + 1), axes[3]),
            na.ScalarArray(np.linspace(-half_pupil[1], half_pupil[1], n_pupil[1] + 1), axes[4]),
        )
        return wavelength, field, pupil, axes

    out = {}
    if "cfg3" in which:
        # cfg 3: EUV slitless spectrograph (octagonal field stop, toroidal VLS grating), 1e9 rays, 8 wavelength cells
        deg = u.deg
        out["cfg3_strong"] = strong_scaling(
            "cfg3 toroidal VLS spectrograph: 8 wavelength x 100x100 field x 112x112 pupil cells = 1.0e9 stratified "
            "random rays, 4 surfaces, 2048x1024 sensor",
            configs.toroidal_vls(6, 12, 3),
            grid(25 * u.nm, 35 * u.nm, 8, (0.2 * deg, 0.2 * deg), (100, 100), (22.0, 22.0), (112, 112)),
            4, 771, device, rank, world, steps, warmup,
        )
    if "cfg5" in which:
        # cfg 5: full detector image simulation: one field cell per pixel, 10 x 8 pupil cells each, 8 tilts
        system = configs.telescope_4k(num_tilt=8, num_pixel=4096)
        half = float(np.arctan(0.5 * 4096 * 15e-3 / 3200.0))
        out["cfg5_strong"] = strong_scaling(
            "cfg5 misaligned telescope, full detector image: 8 tilts x 4096x4096 field cells (one per pixel) x 10x8 "
            "pupil cells = 1.07e10 stratified random rays, 6 surfaces, 4096x4096 sensor",
            system,
            grid(499 * u.nm, 501 * u.nm, 1, (half, half), (4096, 4096), (160.0, 160.0), (10, 8)),
            6, 791, device, rank, world, steps, warmup,
        )
    return out

# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import optika_b200 as optika
    from optika_b200 import _engine, _lib, named as na, units as u
    import configs

    lib = _lib.lib()
    stream = torch.cuda.current_stream(device)

    which_strong = [w for w in args.strong.split(",") if w in ("cfg3", "cfg5")]
    if args.only_strong:
        strong = strong_scaling_configs(device, rank, world, args.strong_steps, 1, which_strong)
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, unit=UNIT, n_gpus=world, config=strong)), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the system and its separable grid; rank r traces its own pupil slab (weak scaling)
    nf, npup, nw = args.num_field, args.num_pupil, args.num_wavelength
    system = configs.spherical_grating(num_field=nf, num_pupil=npup, num_wavelength=nw)
    grid = system.grid_input
    if world > 1:
        # shard by pupil slab: each rank owns the same number of pupil_x cells over its own strip
        lo, hi = -45.0 + 90.0 * rank / world, -45.0 + 90.0 * (rank + 1) / world
        grid.pupil = na.Cartesian2dVectorLinearSpace(
            start=na.Cartesian2dVectorArray(lo, -45.0), stop=na.Cartesian2dVectorArray(hi, 45.0),
            axis=na.Cartesian2dVectorArray("pupil_x", "pupil_y"), num=npup, centers=True,
        )
    compiled = system._compiled
    n_slab = nf * nf * npup * npup
    wavelengths = np.asarray(grid.wavelength.ndarray, dtype=np.float64)

    # ---- FP64 peak (not in MEASURED_PEAKS.json): DFMA micro-benchmark
    fp64_peak = C.c_double(0.0)
    _lib.check(lib.optk_measure_fp64_peak(C.byref(fp64_peak), stream.cuda_stream))
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        hbm_peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm_peak = 6650.0
        peak_source = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"

    # ---- dense structure-of-arrays slab resident in HBM
    fx = torch.as_tensor(np.array(grid.field.x.ndarray), device=device)
    fy = torch.as_tensor(np.array(grid.field.y.ndarray), device=device)
    px = torch.as_tensor(np.array(grid.pupil.x.ndarray), device=device)
    py = torch.as_tensor(np.array(grid.pupil.y.ndarray), device=device)
    shape4 = (nf, nf, npup, npup)

    def dense(t, pos):
        view = [1, 1, 1, 1]
        view[pos] = -1
        return t.reshape(view).expand(shape4).contiguous().reshape(-1)

    dx = (-torch.cos(fy).reshape(1, -1, 1, 1) * torch.sin(fx).reshape(-1, 1, 1, 1)).expand(shape4).contiguous().reshape(-1)
    dy = (-torch.sin(fy)).reshape(1, -1, 1, 1).expand(shape4).contiguous().reshape(-1)
    dz = (torch.cos(fy).reshape(1, -1, 1, 1) * torch.cos(fx).reshape(-1, 1, 1, 1)).expand(shape4).contiguous().reshape(-1)
    fields_in = dict(
        px=dense(px, 2), py=dense(py, 3), pz=torch.zeros(n_slab, dtype=torch.float64, device=device),
        dx=dx, dy=dy, dz=dz,
        intensity=torch.ones(n_slab, dtype=torch.float64, device=device),
        attenuation=torch.zeros(n_slab, dtype=torch.float64, device=device),
        index_refraction=torch.ones(n_slab, dtype=torch.float64, device=device),
    )
    mask_in = torch.ones(n_slab, dtype=torch.uint8, device=device)
    w_dense = [torch.full((n_slab,), float(w), dtype=torch.float64, device=device) for w in wavelengths]
    out = {name: torch.empty(n_slab, dtype=torch.float64, device=device) for name in _lib.FIELDS}
    mask_out = torch.empty(n_slab, dtype=torch.uint8, device=device)

    rin = [_lib.RaysIn() for _ in range(nw)]
    rout = _lib.RaysOut()
    for f, name in enumerate(_lib.FIELDS):
        rout.field[f] = out[name].data_ptr()
    rout.unvignetted = mask_out.data_ptr()
    for k in range(nw):
        r = rin[k]
        r.n_axes = 1
        r.dims[0] = n_slab
        for f, name in enumerate(_lib.FIELDS):
            t = w_dense[k] if name == "wavelength" else fields_in[name]
            r.field[f] = t.data_ptr()
            r.stride[f][0] = 1
        r.unvignetted = mask_in.data_ptr()
        r.mask_stride[0] = 1

    launches = 0

    def step():
        nonlocal launches
        for k in range(nw):
            _lib.check(
                lib.optk_trace(
                    compiled.handle, 0, C.byref(rin[k]), C.byref(rout), 0, N_SURFACES, 1, 0, 0,
                    None, None, None, stream.cuda_stream,
                )
            )
            launches += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- value: device-resident
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    timed_launches = launches
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / args.steps
    rays_per_step = n_slab * nw * world
    value = rays_per_step * N_SURFACES / (ms_per_step * 1e-3)

    ms_per_launch = ms_total / timed_launches
    achieved_gbs = BYTES_PER_RAY * n_slab / (ms_per_launch * 1e-3) / 1e9
    achieved_tflops = FLOP_PER_RAY * n_slab / (ms_per_launch * 1e-3) / 1e12

    # which kernel served the dense launches: the run-time specialised one (csrc/jit.cu) when NVRTC is
    # there, else the bulk-copy pipeline; and what a pure copy with the same 22-stream SoA access
    # pattern reaches in this run (MEASURED_PEAKS' figure is a two-stream torch copy)
    specialised = int(lib.optk_jit_compiled()) > 0
    kernel_name = (
        "optk_jit_kernel (optk::trace_body compiled at run time for this surface list; dense SoA in, dense SoA out)"
        if specialised else "optk::trace_kernel_tma<128, 4> (bulk-copy pipeline; dense SoA in, dense SoA out)"
    )
    soa_copy = C.c_double(0.0)
    _lib.check(lib.optk_measure_soa_copy(n_slab, C.byref(soa_copy), stream.cuda_stream))

    # DRAM traffic of the same kernel from the committed `ncu --set full` capture, scaled to
    # this launch size (profiles/r01_traffic.json; null when the capture is absent)
    traffic = None
    traffic_file = ROOT / "profiles" / "r01_traffic.json"
    if traffic_file.exists():
        cap = json.loads(traffic_file.read_text())
        traffic = cap["dram_bytes_per_launch"] / cap["rays_per_launch"] * n_slab

    # sanity: the traced rays are physical (most of them reach the sensor)
    unv_frac = float(mask_out.float().mean().item())

    # parity of the TIMED kernel's own output, in the run: 1e5 rays spread over the last slab it wrote, against
    # the oracle on the same inputs (positions / directions 1e-9, masks exact up to enumerated edge rays)
    parity_in_run = None
    if rank == 0:
        import parity
        from oracle import raytrace as ora

        m_sample = min(n_slab, 100_000)
        idx = (torch.arange(m_sample, device=device, dtype=torch.int64) * (n_slab - 1)) // max(m_sample - 1, 1)
        sample = {name: (w_dense[-1] if name == "wavelength" else fields_in[name])[idx].cpu().numpy() for name in _lib.FIELDS}
        sample["unvignetted"] = np.ones(len(idx), dtype=bool)
        want = ora.propagate_rays(system.surfaces_all, sample, extended=True)
        got = {name: out[name][idx].cpu().numpy() for name in _lib.FIELDS}
        got["unvignetted"] = mask_out[idx].cpu().numpy().astype(bool)
        report = parity.compare_states(got, want, system.surfaces_all)  # raises on a failure
        parity_in_run = True
        parity_report = dict(rays=int(len(idx)), **{k: v for k, v in report.items() if k != "edge_rays"})

    # ---- e2e: public API, host buffers in, detector planes out (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        del w_dense, out
        torch.cuda.empty_cache()
        edges = na.ScalarArray(np.array([wavelengths.min() - 1e-9, wavelengths.max() + 1e-9]), "wavelength")
        h2d = 8 * (nw + 2 * npup + 3 * nf * nf + 5)  # the separable axes the engine uploads
        image_host = {}

        breakdown = os.environ.get("OPTK_BENCH_BREAKDOWN") is not None  # per-stage wall times to stderr

        def e2e_step():
            t_a = time.perf_counter()
            image = system.image_rays(edges, device=device, counts=True, **configs.PHYSICAL)
            if breakdown:
                torch.cuda.synchronize()
                t_b = time.perf_counter()
            if world > 1:
                dist.all_reduce(image.flux)
                dist.all_reduce(image.moment_real)
                dist.all_reduce(image.counts)
            if breakdown:
                torch.cuda.synchronize()
                t_c = time.perf_counter()
            if rank == 0 or world == 1:
                image_host.update(image.to_host(pinned=True))
            if breakdown:
                print(f"[rank {rank}] e2e step: trace+bin {t_b - t_a:.4f} s, all_reduce {t_c - t_b:.4f} s, "
                      f"read-back {time.perf_counter() - t_c:.4f} s", file=sys.stderr, flush=True)
            return image

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, args.steps)
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = sum(int(v.numel() * v.element_size()) for v in image_host.values()) if image_host else 0
        binned = int(image_host["counts"].sum().item()) if image_host else 0
        e2e = dict(
            value=rays_per_step * N_SURFACES / float(tt.item()),
            unit=UNIT,
            h2d_bytes_per_step=int(h2d),
            d2h_bytes_per_step=int(d2h),
            api="SequentialSystem.image_rays (fused trace + detector binning), host grids in, host image planes out",
            seconds_per_step=float(tt.item()),
            rays_binned=binned,
        )

    # ---- the seam itself with HOST ray arrays: propagate_rays(surfaces, rays) -> rays, every ray
    #      crossing PCIe both ways (optk_trace_host: slabs on three streams).  Bounded sample.
    host_rays = None
    if not args.no_e2e and rank == 0:
        n_host = min(n_slab, args.host_rays)
        pin = lambda dt: torch.empty(n_host, dtype=dt, pin_memory=True)  # noqa: E731
        hin = {name: pin(torch.float64) for name in _lib.FIELDS}
        for name in _lib.FIELDS:
            src = torch.full((n_host,), float(wavelengths[0]), dtype=torch.float64) if name == "wavelength" \
                else fields_in[name][:n_host].cpu()
            hin[name].copy_(src)
        hout = {name: pin(torch.float64) for name in _lib.FIELDS}
        hmask = pin(torch.uint8)
        hr, ho = _lib.RaysIn(), _lib.RaysOut()
        hr.n_axes = 1
        hr.dims[0] = n_host
        for f, name in enumerate(_lib.FIELDS):
            hr.field[f] = hin[name].data_ptr()
            hr.stride[f][0] = 1
            ho.field[f] = hout[name].data_ptr()
        hr.unvignetted = None
        ho.unvignetted = hmask.data_ptr()

        def host_step():
            _lib.check(
                lib.optk_trace_host(
                    compiled.handle, 0, C.byref(hr), C.byref(ho), 0, N_SURFACES, 1, 0, 0, None, None, None, 1 << 22, 1
                )
            )

        host_step()
        t0 = time.perf_counter()
        for _ in range(2):
            host_step()
        dt_host = (time.perf_counter() - t0) / 2
        host_rays = dict(
            value=n_host * N_SURFACES / dt_host,
            unit=UNIT,
            rays=n_host,
            h2d_bytes_per_step=80 * n_host,
            d2h_bytes_per_step=81 * n_host,
            pcie_gbytes_per_s_each_way=80 * n_host / dt_host / 1e9,
            api="optk_trace_host (propagate_rays with host ray arrays in and out, pinned, 4M-ray slabs)",
            unvignetted_fraction=float(hmask.float().mean().item()),
        )
        del hin, hout, hmask

    # ---- cfg 2 "generated on chip": stratified random rays drawn, traced and binned inside
    #      one launch per wavelength cell (optk_trace_grid); device-timed, no ray ever in HBM
    on_chip = None
    if not args.no_e2e:
        from optika_b200 import _grid

        def cell_edges(lo, hi, n):
            return np.linspace(lo, hi, n + 1)

        lo_w, hi_w = float(wavelengths.min()), float(wavelengths.max())
        if hi_w == lo_w:
            hi_w = lo_w * (1 + 1e-6)
        half_field = 0.05 * u.deg
        ray_grid = _grid.RayGrid(
            [cell_edges(lo_w, hi_w, nw), cell_edges(-half_field, half_field, nf), cell_edges(-half_field, half_field, nf),
             cell_edges(-45.0, 45.0, npup * world), cell_edges(-45.0, 45.0, npup)],
            jitter=True, seed=0,
        ).shard(rank, world, axis=3)
        ex, ey = system.sensor.pixel_edges()
        chip_image = _engine.DeviceImage.zeros(
            np.array([lo_w, hi_w]), ex, ey, device, moments=True, counts=True
        )
        local = system._compiled_local
        chip_launches = 0

        def chip_step():
            nonlocal chip_launches
            before = _grid.LAUNCHES
            # as few launches as the 2^31 - 1 ray limit allows: every launch then spans many wavelength cells and
            # the strided CTA order spreads the resident CTAs over all of them (one 1e8-ray launch per cell piles
            # the reductions of a whole launch onto the ~2000 pixels of one 0.7 nm band)
            _grid.trace_grid(local, ray_grid, image=chip_image, write_rays=False, device=device)
            chip_launches = _grid.LAUNCHES - before

        chip_step()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        chip_step()
        c1.record(stream)
        barrier()
        tc = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        on_chip = dict(
            value=ray_grid.size * world * N_SURFACES / (float(tc.item()) * 1e-3),
            unit=UNIT,
            ms_per_step=float(tc.item()),
            launches_per_step=chip_launches,
            api="optk_trace_grid: Philox-jittered vertex grid -> trace -> detector planes, one fused kernel",
            rays_binned_fraction=float(chip_image.counts.sum().item()) / (2.0 * ray_grid.size),
        )

    # ---- strong scaling of the cfg 3 / cfg 5 image simulations (collective + read-back timed)
    strong = {}
    if which_strong and not args.no_e2e:
        del fields_in, mask_in, mask_out
        torch.cuda.empty_cache()
        strong = strong_scaling_configs(device, rank, world, args.strong_steps, 1, which_strong)

    # ---- CPU baseline beside it (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(args, args.cpu_sample_rays)
        cpu.pop("seconds", None)

    if rank == 0:
        line = dict(
            metric=METRIC,
            value=value,
            unit=UNIT,
            n_gpus=world,
            steps=args.steps,
            warmup=args.warmup,
            ms_per_step=ms_per_step,
            higher_is_better=True,
            scaling="weak",
            vs_baseline=None,
            dtype="f64",
            data="synthetic",
            config=dict(
                workload=workload_name(args),
                rays_per_step=rays_per_step,
                launches_per_step=nw,
                l2="inputs exceed L2: every launch streams 16.2 GB of dense rays through HBM",
                unvignetted_fraction=unv_frac,
                generated_on_chip=on_chip,
                **strong,
                sharding="pupil slab per rank, no data-path collective" if world > 1 else "single GPU",
            ),
            roofline=dict(
                bound="hbm",
                achieved=achieved_gbs,
                peak=hbm_peak,
                unit="GB/s",
                frac=achieved_gbs / hbm_peak,
                traffic=traffic,
                kernel=kernel_name,
                copy_same_pattern_gbs=soa_copy.value,
                note=(
                    "peak is the copy bandwidth MEASURED_PEAKS.json records (torch copy_, two streams); a fraction "
                    "slightly above 1 means this kernel's 22-stream pattern sustains more than that copy does "
                    "(copy_same_pattern_gbs: a pure copy with this kernel's pattern, same run; nominal HBM3e 8 TB/s)"
                ),
                algorithmic_bytes_per_launch=BYTES_PER_RAY * n_slab,
                ms_per_launch=ms_per_launch,
                peak_source=peak_source,
                fp64=dict(
                    achieved=achieved_tflops,
                    peak=fp64_peak.value / 1e12,
                    unit="TFLOP/s",
                    frac=achieved_tflops / (fp64_peak.value / 1e12) if fp64_peak.value else None,
                    algorithmic_flop_per_ray=FLOP_PER_RAY,
                    peak_source="optk_measure_fp64_peak (DFMA micro-benchmark, this run)",
                ),
            ),
            cpu_baseline=cpu,
            e2e=dict(e2e, rays_to_host=host_rays) if e2e is not None else None,
            gpu_launches=timed_launches,
            clocks=clocks,
            parity_in_run=parity_in_run,
            parity=parity_report,
        )
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
