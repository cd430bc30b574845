#!/usr/bin/env python
"""bench.py - Msamples/s of the Mode S receive hot path (IQ demod + preamble detect + slice + CRC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 28]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic 1090 MHz IQ at 4 Msps, 2^28 complex samples (2 GiB) per GPU,
1000 DF11/DF17 bursts at mixed SNR on complex Gaussian noise. A step = one pass of the whole receive
chain over that buffer (fresh stream each step). At N > 1 every rank owns an independent channel of the
same size (weak scaling, no data-path collective; NCCL only fans the buffers out from rank 0).

  value        whole-job Msamples/s with the input resident in HBM (max-over-ranks device time, CUDA events)
  e2e          same through the public API from PINNED HOST memory: H2D of the IQ + D2H of the frames inside
  roofline     scan kernel: 8 algorithmic bytes/sample / its measured duration vs the measured HBM peak
  cpu_baseline the reference's own C++ (oracle/_ref: unmodified preamble_impl/slicer_impl/modes_crc behind the
               restated GNU Radio front end) on a bounded sample, all host threads
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 4e6
THRESHOLD_DB = 7.0
N_BURSTS = 1000
NOISE_SIGMA = 0.01


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        top = sorted(sm)[len(sm) // 2:] if sm else []          # under-load half
        return {"sm_mhz": float(np.median(top)) if top else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_bursts(n, seed):
    """Host-side burst overlays of the scene (sparse): list of (first_sample, complex64 array)."""
    from gr_air_modes_b200 import synth
    rng = np.random.Generator(np.random.PCG64(seed))
    spc = RATE / 2e6
    span = 240 * spc + 8
    starts = np.sort(rng.uniform(span, n - 2 * span, N_BURSTS))
    out = []
    for s in starts:
        df = (11, 17)[int(rng.integers(0, 2))]
        frame = synth.make_frame(df, rng)
        snr = rng.uniform(6.0, 30.0)
        amp = float(np.sqrt(2.0 * NOISE_SIGMA ** 2 * 10 ** (snr / 10)))
        b = synth.Burst(float(s), frame, amp, float(rng.uniform(0, 2 * np.pi)))
        out.append(synth.burst_waveform(b, spc) + (frame,))
    return out


def make_device_scene(n, seed, device):
    import torch
    g = torch.Generator(device=device); g.manual_seed(1000 + seed)
    iq = torch.empty(2 * n, device=device, dtype=torch.float32)
    step = 1 << 26
    for a in range(0, 2 * n, step):
        m = min(step, 2 * n - a)
        iq[a:a + m] = torch.randn(m, device=device, generator=g) * NOISE_SIGMA
    frames = []
    for n0, w, frame in make_bursts(n, seed):
        t = torch.from_numpy(np.ascontiguousarray(w).view(np.float32)).to(device)
        iq[2 * n0: 2 * n0 + t.numel()] += t
        frames.append(frame.hex())
    return iq, frames


def bind_near_gpu(index):
    """Pin this process to the CPUs next to GPU `index` (NVML's ideal affinity) so that the pinned staging buffer of
    the end-to-end leg lands on the NUMA node the GPU's PCIe root hangs off. Returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:                                    # CUDA's device order need not be NVML's: go through the PCI address
            import torch
            pr = torch.cuda.get_device_properties(index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
        except Exception:                       # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1} & os.sched_getaffinity(0)
        if not cpus:
            return "nvml gave no usable cpu set"
        os.sched_setaffinity(0, cpus)
        return "%d cpus near gpu %d" % (len(cpus), index)
    except Exception as e:                      # noqa: BLE001 - optional tuning only
        return "not bound (%s)" % type(e).__name__


def _tame_malloc():
    """Keep multi-MB buffers on the heap instead of mmap/munmap per call: the per-call allocations are an artefact
    of the oracle's whole-buffer driver (GNU Radio keeps persistent ring buffers), and first-touch page faults
    would otherwise dominate the CPU arm."""
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 1 << 30)            # M_MMAP_THRESHOLD
        libc.mallopt(-1, (1 << 31) - 1)      # M_TRIM_THRESHOLD
    except Exception:
        pass


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cpu_oracle as co
    from gr_air_modes_b200 import synth
    _tame_malloc()
    port = co.Port()
    kind = "reference" if co.ref_available() else "port"
    ref = co.Ref() if kind == "reference" else None
    cores = max(1, min(len(os.sched_getaffinity(0)), 256))       # every host thread we are allowed to use
    n_slice = 1 << (23 if cores <= 64 else 22)
    sc = synth.make_scene(RATE, n_slice, max(1, int(N_BURSTS * n_slice / (1 << args.log2n))), 7, noise_sigma=NOISE_SIGMA)
    slices = [sc.iq] * cores                                     # read-only input shared by the threads

    def work(iq):
        bb, avg = port.frontend(iq, RATE, True, co.MA_GR_FLOAT, 4096)     # GNU Radio's fp32 running-sum schedule
        r = (ref.run_streams(bb, avg, RATE, THRESHOLD_DB) if ref else port.run_streams(bb, avg, RATE, THRESHOLD_DB))
        return len(r.msgs)

    with ThreadPoolExecutor(cores) as ex:
        for _ in range(max(args.warmup, 2)):
            list(ex.map(work, slices))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            msgs = list(ex.map(work, slices))
        dt = time.perf_counter() - t0
    total = args.steps * cores * n_slice
    val = total / dt / 1e6
    sample = "%d threads x 2^%d-sample cuts of the 4 Msps scene per step (GR fp32 moving averages + %s scan/slice/CRC)" % (
        cores, int(np.log2(n_slice)), "unmodified reference" if ref else "oracle port")
    line = {"impl": "reference", "metric": "Msamples/s IQ demod+slice+CRC", "value": val, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic 1090 MHz IQ @ 4 Msps, DF11/DF17 bursts at mixed SNR (configs[1]), bounded CPU sample",
                       "rate_sps": RATE, "threshold_db": THRESHOLD_DB, "use_pmf": True},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "msgs_per_step": int(sum(msgs))}
    print(json.dumps(line))


def cpu_baseline_leg(log2n):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cpu_oracle as co
    from gr_air_modes_b200 import synth
    _tame_malloc()
    port = co.Port()
    kind = "reference" if co.ref_available() else "port"
    ref = co.Ref() if kind == "reference" else None
    cores = max(1, min(len(os.sched_getaffinity(0)), 256))       # every host thread we are allowed to use
    n_slice = 1 << (23 if cores <= 64 else 22)
    sc = synth.make_scene(RATE, n_slice, max(1, int(N_BURSTS * n_slice / (1 << log2n))), 7, noise_sigma=NOISE_SIGMA)
    reps = max(1, int(round(2 * (1 << 29) / (cores * n_slice))))  # ~2^30 samples in total: 10-30 s of CPU work

    def work(k):
        bb, avg = port.frontend(sc.iq, RATE, True, co.MA_GR_FLOAT, 4096)
        r = (ref.run_streams(bb, avg, RATE, THRESHOLD_DB) if ref else port.run_streams(bb, avg, RATE, THRESHOLD_DB))
        return len(r.msgs)

    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(work, range(cores)))
        list(ex.map(work, range(cores)))
        t0 = time.perf_counter()
        for _ in range(reps):
            list(ex.map(work, range(cores)))
        dt = time.perf_counter() - t0
    val = reps * cores * n_slice / dt / 1e6
    return {"value": val, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "sample": "%d passes x %d threads over a 2^%d-sample cut of the same 4 Msps scene; GR fp32 moving averages + "
                      "%s preamble/slicer/CRC" % (reps, cores, int(np.log2(n_slice)), "unmodified reference" if ref else "oracle port")}


def time_shard_arm(args, rank, local_rank, world, device):
    """Secondary multi-GPU mode: one recording, `world` time spans with halos, the scan-loop state
    (16 bytes) handed down the ranks between the dense and the sparse stages. Strong scaling; not the headline."""
    import torch
    import torch.distributed as dist
    import gr_air_modes_b200 as am
    from gr_air_modes_b200 import shard
    n = 1 << args.log2n
    plan = shard.time_shard_plan(n, world, am.query_geometry(RATE, THRESHOLD_DB, True))
    active = rank < len(plan)
    sp = plan[rank] if active else None
    t_f0 = time.perf_counter()
    whole = sent = None
    if rank == 0:
        whole, sent = make_device_scene(n, 0, device)
        for r in range(1, len(plan)):
            dist.send(whole[2 * plan[r].first_sample: 2 * plan[r].end].contiguous(), dst=r)
        iq = whole[: 2 * sp.end]
    elif active:
        iq = torch.empty(2 * (sp.end - sp.first_sample), dtype=torch.float32, device=device)
        dist.recv(iq, src=0)
    torch.cuda.synchronize()
    fanout_s = time.perf_counter() - t_f0
    q = am.msg_queue()
    rx = am.rx_path(RATE, THRESHOLD_DB, q, use_pmf=True, device=local_rank)
    rx._ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    recv, send = shard.dist_state_exchange(rank, device)

    gather = shard.dist_all_gather6(world, device) if world > 1 else (lambda v: [list(v)])
    last = len(plan) - 1
    fallbacks = [0]

    def step():
        if args.chain:
            if not active:
                return
            rx.seek(sp.first_sample, sp.first_decision)
            rx.process(iq, flush=sp.flush, collect=False)
            entry = recv()[:2] if rank else (0, 0)
            rx.resolve(entry)
            if not sp.flush:
                send(rx.walk_state() + (0,))
            return
        mine = [0] * 6                                   # speculative resolution: one all-gather, no chain
        if active:
            rx.seek(sp.first_sample, sp.first_decision)
            rx.process(iq, flush=sp.flush, collect=False)
            if rank < last:
                rx.resolve(None)
                s = rx.walk_summary()
                mine = [s.pos, s.p, s.first_real, s.first_packet, s.exact_span, s.frames_passed]
        entries, _, bad = shard.compose_entries(plan, gather(mine))
        fallbacks[0] += bad < last
        if active and rank >= bad:
            rx.resolve(entries[rank] if rank == bad else recv()[:2])
            if not sp.flush:
                send(rx.walk_state() + (0,))

    rx.defer_resolve(True)
    for _ in range(args.warmup):
        step()
    launches0 = rx.stats().kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    rx._ctx.join()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    launches = rx.stats().kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_max = shard.max_over_ranks(ms, world, device)
    # ---- check: the spans' messages, concatenated in rank order, are the one-shot run's
    if args.chain:
        mine = shard.process_time_sharded(rx, iq, sp, recv, send) if active else 0
    else:
        mine = shard.process_time_sharded_speculative(rx, iq if active else None, plan, rank, gather, recv, send)
    msgs = q.strings()
    q.flush()
    allm = [None] * world
    if world > 1:
        dist.all_gather_object(allm, msgs)
    else:
        allm = [msgs]
    if rank == 0:
        rx.defer_resolve(False)
        rx.reset()
        rx._slicer._first = True                  # a fresh slicer's first message has 6 digits (slicer_impl.cc:192)
        rx.process(whole, flush=True)
        ref_msgs = q.strings()
        joined = [m for part in allm for m in part]
        line = {"metric": "Msamples/s IQ demod+slice+CRC", "value": n * args.steps / (ms_max * 1e-3) / 1e6,
                "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "mode": "time-shard/" + ("chain" if args.chain else "speculative"),
                "fallback_steps": fallbacks[0],
                "config": {"workload": "ONE synthetic 4 Msps recording of 2^%d samples cut into %d time spans with halos"
                                       % (args.log2n, len(plan)), "parallelism": "time-shard x%d" % len(plan),
                           "fanout_s": round(fanout_s, 3), "timing": "host clock between barrier+synchronize (the state "
                           "hand-over is host-driven)", "spans": [repr(x) for x in plan]},
                "identical_to_one_shot": joined == ref_msgs, "msgs": len(ref_msgs),
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--time-shard", action="store_true",
                    help="secondary mode: ONE 2^log2n-sample recording cut into --gpus spans (strong scaling)")
    ap.add_argument("--chain", action="store_true", help="--time-shard: plain hand-over chain instead of speculative resolution")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gr_air_modes_b200 as am
    from gr_air_modes_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    n = 1 << args.log2n
    if args.time_shard:
        time_shard_arm(args, rank, local_rank, world, device)
        return

    # ---- input: rank 0 synthesises every channel and fans it out over NCCL (timed separately)
    t_f0 = time.perf_counter()
    iq, sent = shard.fan_out(lambda ch: make_device_scene(n, ch, device), rank, world, device, 2 * n)
    torch.cuda.synchronize()
    fanout_s = time.perf_counter() - t_f0

    q = am.msg_queue()
    rx = am.rx_path(RATE, THRESHOLD_DB, q, use_pmf=True, device=local_rank)
    rx._ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    rx._ctx.call("amb_enable_timing", 1)

    def step():
        rx.reset()
        rx.process(iq, flush=True, collect=False)

    for _ in range(args.warmup):
        step()
    nmsg = rx.drain()
    got = {m.split()[0] for m in q.strings()}
    decoded = len(got & set(sent))
    launches0 = rx.stats().kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    rx._ctx.join()                 # the sparse kernels of the last step run on the library's second stream
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    scan_ms = rx._ctx.scan_times_ms(min(args.steps, 64))
    launches = rx.stats().kernel_launches - launches0
    rx.drain(); q.flush()
    ms_max = shard.max_over_ranks(ms, world, device)
    scan_avg = shard.max_over_ranks(float(np.mean(scan_ms)), world, device)

    # ---- end to end through the public API: pinned host IQ -> H2D -> chain -> frames D2H
    all_cpus = os.sched_getaffinity(0)
    numa = bind_near_gpu(local_rank)          # the pinned buffer is first-touched on the GPU's NUMA node
    host = torch.empty(2 * n, dtype=torch.float32, pin_memory=True)
    host.copy_(iq)
    torch.cuda.synchronize()
    d2h = 0
    for _ in range(1):
        rx.reset(); rx.process(host, flush=True)
    q.flush()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        rx.reset()
        rx.process(host, flush=True)          # H2D copy + kernels + frame read-back (drain) inside
        d2h = len(rx.frames) * 80 + 32
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.e2e_steps
    e2e_ms = shard.max_over_ranks(e2e_ms, world, device)
    q.flush()
    clocks = sampler.stop() if rank == 0 else None     # sampled across both timed regions (device + end to end)
    os.sched_setaffinity(0, all_cpus)                  # the CPU baseline leg below gets every host thread back

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ach = 8.0 * n / (scan_avg * 1e-3) / 1e9
        line = {
            "metric": "Msamples/s IQ demod+slice+CRC", "value": world * n * args.steps / (ms_max * 1e-3) / 1e6,
            "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic 1090 MHz IQ @ 4 Msps, 2^%d complex samples per GPU, %d DF11/DF17 bursts at mixed SNR "
                                   "(BASELINE configs[1]); one independent channel per GPU" % (args.log2n, N_BURSTS),
                       "rate_sps": RATE, "threshold_db": THRESHOLD_DB, "use_pmf": True, "samples_per_gpu": n,
                       "l2": "input (%.1f GiB per GPU) is larger than L2, no flush needed" % (8 * n / 2 ** 30),
                       "parallelism": "channel-per-gpu x%d" % world, "fanout_s": round(fanout_s, 3),
                       "e2e_steps": args.e2e_steps, "msgs_per_step": nmsg, "bursts_decoded": decoded},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "kernel": "amb_scan_kernel<2,true>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 8 * n, "scan_ms": scan_avg},
            "e2e": {"value": world * n / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * n,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "h2d_gbps": 8 * n / (e2e_ms * 1e-3) / 1e9,
                    "host_numa_binding": numa},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        tr = os.path.join(ROOT, "profiles", "scan_traffic.json")
        if os.path.exists(tr):
            try:
                line["roofline"]["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_2p%d" % args.log2n)
            except Exception:
                pass
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(args.log2n)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
