#!/usr/bin/env python
"""bench.py — QFT state-vector benchmark (BASELINE.json: "QFT-33 wall time; effective state-vector HBM GB/s
vs roofline; 1/2/4/8 GPU").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--qubits n]

One "step" = one full QFT-n simulation: reset the register to |x>, apply all n + n(n-1)/2 gates.
N = 1 runs QFT-33 (128 GiB complex-f64 state, BASELINE configs[3]); N > 1 runs the weak-scaling
sweep QFT-(33 + log2 N) with the state sharded on the top log2 N qubits (configs[4]).

Reported metric: effective state-vector GB/s = gates * 32 B * 2^n / wall time — the bandwidth an
unfused one-sweep-per-gate engine (the reference's structure, src/circuit/simulation.rs:37-56) would
need; it is a throughput comparable between the CPU reference and this engine.  The QFT wall time
itself is `ms_per_step`; `roofline` carries the honest per-fused-pass HBM figure.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

X_INIT = 0x123456789  # initial basis state


def qft_gates(n):
    import quantr_b200 as qb
    from helpers import qft_circuit
    return qft_circuit(qb.Circuit, qb.Gate, n).get_gates()


def n_qft_gates(n):
    return n + n * (n - 1) // 2


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, flag in zip(names, s[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(n, threads=1):
    """The restated reference CPU path (oracle, faithful per-gate hash-map rebuild, single thread like the
    reference, README.md:96) on a bounded QFT-n sample.  Returns (seconds, GB/s-effective)."""
    import numpy as np
    from helpers import OracleCircuit, encode_gates, orc, qb, qft_circuit
    c = qft_circuit(OracleCircuit, qb.Gate, n)
    enc = encode_gates(c.circuit_gates, n)
    reg = np.zeros(1 << n, dtype=np.complex128)
    reg[X_INIT & ((1 << n) - 1)] = 1.0
    t0 = time.perf_counter()
    orc.simulate(n, enc.ops, enc.n_ops, reg, mode="faithful")
    dt = time.perf_counter() - t0
    return dt, n_qft_gates(n) * 32.0 * (1 << n) / dt / 1e9


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Rust crate cannot be built here)."""
    if rank != 0:
        return
    n_sample = args.ref_qubits
    for _ in range(args.warmup):
        cpu_baseline(max(8, n_sample - 3))
    times, vals = [], []
    for _ in range(args.steps):
        dt, v = cpu_baseline(n_sample)
        times.append(dt)
        vals.append(v)
    value = sum(vals) / len(vals)
    n_full = args.qubits
    sample = (f"QFT-{n_sample} ({n_qft_gates(n_sample)} gates) per step: the full QFT-{n_full} is out of reach for the reference "
              f"algorithm (2^{n_full}-entry hash map); faithful C++ restatement, 1 thread (the reference is single-threaded)")
    line = {
        "impl": "reference", "metric": "qft_effective_state_vector_gbps", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"QFT-{n_full} complex f64 (bounded sample: QFT-{n_sample})", "qubits": n_full, "sample_qubits": n_sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def setup_peer_exchange(state, dist, world, rank, device):
    """Direct NVLink exchange: all-gather the shards' IPC handles (torch is plumbing only) and map the peers.  Every
    rank ends in the same mode: if any rank cannot export or import, all of them stay on / return to the NCCL transport.
    Returns True when the peer-memory transport is active."""
    import torch
    ok = 1
    try:
        handle = state.peer_export()
    except Exception as e:  # e.g. CUDA IPC unavailable in this container
        ok, handle = 0, bytes(64)
        print(f"[bench] rank {rank}: peer export failed ({e}); falling back to the NCCL transport", file=sys.stderr)
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(device)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    exported = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(exported, op=dist.ReduceOp.MIN)
    imported = False
    if int(exported.item()):
        try:
            state.peer_import([bytes(h.cpu().numpy().tobytes()) for h in allh])
            imported = True
        except Exception as e:  # e.g. no peer access between two GPUs
            ok = 0
            print(f"[bench] rank {rank}: peer import failed ({e}); falling back to the NCCL transport", file=sys.stderr)
    else:
        ok = 0
    agree = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    if int(agree.item()):
        return True
    if imported:
        state.peer_import([])
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="total qubits (default 33 + log2(gpus))")
    ap.add_argument("--ref-qubits", type=int, default=16, help="bounded sample size of the CPU reference arm")
    ap.add_argument("--cpu-baseline-qubits", type=int, default=17)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=0)
    ap.add_argument("--shots", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    g = max(0, world.bit_length() - 1)
    if not args.qubits:
        args.qubits = 33 + g
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import quantr_b200 as qb
    from quantr_b200 import _ffi as F
    from quantr_b200.circuit import encode_gates

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quantr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        lib = F.load_library()
        import ctypes as C
        buf = C.create_string_buffer(128)
        if rank == 0:
            F.check(lib.qsv_nccl_unique_id(buf, 128))
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().numpy().tobytes())

    n = args.qubits
    n_local = n - g
    x = X_INIT & ((1 << n) - 1)
    gates = qft_gates(n)
    enc = encode_gates(gates, n)
    n_gates = n_qft_gates(n)

    state = qb.DeviceState(n, local_rank, rank=rank, world=world, nccl_id=nccl_id)
    exchange_path = "nccl send/recv through staging"
    if world > 1 and not os.environ.get("QSV_NCCL_EXCHANGE"):
        if setup_peer_exchange(state, dist, world, rank, torch.device("cuda", local_rank)):
            exchange_path = "in-place swap kernels over peer-mapped memory (NVLink loads/stores)"
    if args.tile_bits:
        state.set_option("tile_bits", args.tile_bits)
    if args.low_bits:
        state.set_option("low_bits", args.low_bits)
    # sharded: the register starts as a basis state, so the scheduler may park the last-targeted qubits in the rank id
    plan = qb.Plan(n, enc, n_local=n_local, tile_bits=args.tile_bits, low_bits=args.low_bits, free_layout=world > 1)
    pstats = plan.stats()
    pdesc = plan.describe()
    passes = pstats["n_passes"]
    state.set_option("timing", 1)  # per-pass CUDA events inside the library -> qsv_stats.device_ms / exchange_ms

    import ctypes as C
    dev_ptr, stream_ptr = C.c_void_p(), C.c_void_p()
    F.check(state.lib.qsv_device_pointer(state.handle, C.byref(dev_ptr), C.byref(stream_ptr)))
    ext = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local_rank))

    def barrier():
        state.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        state.init_basis(x)
        return state.run_plan(plan)

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: exactly K steps, device-timed on the library's stream -------------------------
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    passes_ms = exchange_ms = 0.0
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev_start.record(ext)
        for k in range(args.steps):
            state.init_basis(x)
            st = state.run_plan(plan)
            passes_ms += st["device_ms"]
            exchange_ms += st["exchange_ms"]
        ev_end.record(ext)
        barrier()
    total_ms = ev_start.elapsed_time(ev_end)
    if world > 1:
        t = torch.tensor([total_ms, passes_ms, exchange_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, passes_ms, exchange_ms = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = total_ms / args.steps
    value = n_gates * 32.0 * float(1 << n) / (ms_per_step * 1e-3) / 1e9

    # ---- verification outside the timed region: closed-form QFT amplitudes + norm -------------------
    rng = np.random.default_rng(1234)  # same indices on every rank: qsv_gather is collective on sharded handles
    idx = rng.integers(0, 1 << n, size=4096, dtype=np.uint64)
    got = state.gather(idx)
    rev = np.zeros_like(idx)
    for b in range(n):
        rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(n - 1 - b)
    ph = np.array([(int(r) * x) % (1 << n) for r in rev], dtype=np.float64) * (2 * np.pi / float(1 << n))
    expect = (np.cos(ph) + 1j * np.sin(ph)) / np.sqrt(float(1 << n))
    max_err = float(np.max(np.abs(got - expect)))
    norm = state.norm_sqr()

    # ---- e2e: through the reference-facing C-ABI calls with HOST buffers -----------------------------
    # qsv_init_basis + qsv_apply(host qsv_op[]: lowering, scheduling, schedule H2D, launches)
    # + qsv_sample(host uniforms -> host indices) + qsv_download(4096 amplitudes)
    e2e = None
    if world == 1:
        uniforms = np.random.default_rng(7).random(args.shots)
        def e2e_step():
            state.init_basis(x)
            state.apply(enc)
            idxs = state.sample(uniforms)
            amps = state.download(0, 4096)
            return idxs, amps
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        plan_bytes = sum(p["bytes"] for p in pdesc["passes"])
        e2e = {"value": n_gates * 32.0 * float(1 << n) / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": int(plan_bytes + 8 * args.shots), "d2h_bytes_per_step": int(8 * args.shots + 16 * 4096 + 8),
               "path": "qsv_init_basis + qsv_apply(host ops) + qsv_sample(host uniforms) + qsv_download(4096 amps)"}

    peak, peak_src = read_peaks()
    bytes_per_pass = 32.0 * float(1 << n_local)
    # QSV_FUSED_INIT (opt-in): the first pass synthesises its input instead of reading it - it moves half the bytes
    fused_init = int(os.environ.get("QSV_FUSED_INIT", "0") or 0)
    passes_bytes = bytes_per_pass * (passes - 0.5 if fused_init and passes else passes)
    achieved = passes_bytes * args.steps / (passes_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and n_local == 33:  # the committed capture is per launch at 2^33 amplitudes per GPU
        try:
            traffic = json.load(open(tpath)).get("pass_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "pass_kernel (fused tile pass)", "algorithmic_bytes_per_launch": passes_bytes / passes if passes else bytes_per_pass,
                "avg_launch_ms": passes_ms / (passes * args.steps), "peak_source": peak_src, "per_gpu": True}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            nb = args.cpu_baseline_qubits
            dt, v = cpu_baseline(nb)
            cpu = {"value": v, "unit": "GB/s", "cores": 1, "kind": "port", "seconds": dt,
                   "sample": f"QFT-{nb} ({n_qft_gates(nb)} gates), faithful C++ restatement of simulation.rs:64-135 (per-gate hash-map rebuild), "
                             f"1 thread; QFT-{n} itself is unreachable for that algorithm"}
        line = {
            "metric": "qft_effective_state_vector_gbps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"QFT-{n} complex f64, |x=0x{x:x}> -> {n_gates} gates (H + CRk, no final swaps)", "qubits": n,
                       "local_qubits": n_local, "state_bytes_per_gpu": 16 << n_local, "l2_policy": "state >> 126 MB L2 (no flush needed)", "init": ("fused into the first pass (mode %d)" % fused_init) if fused_init else "memset + set_amp before the first pass",
                       "tile_bits": pdesc["tile_bits"], "low_bits": pdesc["low_bits"], "parallelism": f"shard{world}"},
            "qft_wall_time_ms": ms_per_step, "fused_passes": passes, "passes_per_gate": passes / n_gates,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int((passes + (0 if fused_init else 1) + (pstats["n_exchanges"] * (world - 1) if exchange_path.startswith("in-place") else 0)) * args.steps),
            "exchange": None if world == 1 else {
                "remaps_per_step": pstats["n_exchanges"], "bytes_sent_per_gpu_per_step": pstats["exchange_bytes"],
                "ms_per_step": exchange_ms / args.steps,
                "achieved_GBps_per_direction": (pstats["exchange_bytes"] * args.steps / (exchange_ms * 1e-3) / 1e9) if exchange_ms > 0 else None,
                "path": exchange_path, "nvlink_peak_GBps_per_direction": 770.0, "peak_source": "B200_PROFILING.md measured peer copy"},
            "clocks": clocks.summary(), "max_abs_err_vs_closed_form": max_err, "norm_sqr": norm,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # importers unmap their peers before any exporter frees its shard (CUDA IPC teardown order)
        state.peer_import([])
        dist.barrier()
    state.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
