#!/usr/bin/env python
"""bench.py — QFT state-vector benchmark (BASELINE.json: "QFT-33 wall time; effective state-vector HBM GB/s
vs roofline; 1/2/4/8 GPU").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload qft|grover] [--qubits n]

One "step" = one full simulation of the workload: reset the register to its initial basis state, apply every gate.
  qft     N = 1: QFT-33 (128 GiB complex-f64 register, BASELINE configs[3]); N > 1: the weak-scaling sweep QFT-(33 + log2 N)
          with the register sharded on log2 N qubits (configs[4]).
  grover  Grover search built from the reference's native gates (H, X, CNot, Toffoli, CZ with an ancilla V-chain,
          tests/grovers.rs:75-155 writes the oracle with Custom multi-CNOTs instead), 33 + log2 N qubits.

Headline `value` = wall time of one step in ms (device-timed, max over ranks; lower is better).  Next to it:
`roofline` (achieved HBM GB/s per fused pass against the measured copy peak, with every pass listed), `effective_hbm_gbps`
(algorithmic pass bytes of all ranks / wall time), `gate_equivalent_gbps` (what a one-sweep-per-gate engine - the reference's
structure, src/circuit/simulation.rs:37-56 - would have to stream; informational, may exceed the HBM peak), `e2e` (the same
step through the reference-facing C-ABI calls with host buffers), `sampling` (K6/K7), and at N = 1 `small_configs`
(BASELINE configs 1-2 through the C ABI) and `same_config` (QFT-16 on the GPU and on the CPU reference path: the one
same-size ratio).  At N > 1 `sharded_parity_max_abs_err` compares a sharded random circuit with the CPU oracle.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
NVLINK_PEAK = 770.0   # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)


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


# ---- workloads ---------------------------------------------------------------------------------------------------

def qft_workload(n):
    import quantr_b200 as qb
    from helpers import qft_circuit
    x = X_INIT & ((1 << n) - 1)
    gates = qft_circuit(qb.Circuit, qb.Gate, n).get_gates()
    return {"name": f"QFT-{n} complex f64, |x=0x{x:x}> -> {n_qft_gates(n)} gates (H + CRk, no final swaps)", "gates": gates, "basis": x,
            "n_gates": n_qft_gates(n), "check": "closed_form_qft"}


def grover_workload(n, iterations):
    from workloads import grover_circuit
    import quantr_b200 as qb
    c, info = grover_circuit(qb.Circuit, qb.Gate, n, iterations=iterations)
    gates = c.get_gates()
    n_gates = sum(1 for g in gates if g.kind != 0)
    return {"name": f"Grover-{n}: {info['search']} search + {info['ancilla']} V-chain ancilla wires" + (f" + {info['idle']} idle" if info['idle'] else "") +
                    f", marked item 0x{info['marked']:x}, {iterations} iteration(s) of oracle + diffusion from native H/X/Toffoli/CZ ({n_gates} gates; "
                    f"the optimal count would be ~{int(0.785 * 2 ** (info['search'] / 2))})",
            "gates": gates, "basis": 0, "n_gates": n_gates, "check": "grover", "info": info}


# ---- CPU reference path ------------------------------------------------------------------------------------------

def cpu_qft_seconds(n):
    """The restated reference CPU path (oracle, faithful per-gate hash-map rebuild, single thread like the reference,
    README.md:96) on QFT-n.  Returns seconds."""
    import numpy as np
    from helpers import OracleCircuit, encode_gates, orc, qb, qft_circuit
    c = qft_circuit(OracleCircuit, qb.Gate, n)
    enc = encode_gates(c.circuit_gates, n)
    reg = np.zeros(1 << n, dtype=np.complex128)
    reg[X_INIT & ((1 << n) - 1)] = 1.0
    t0 = time.perf_counter()
    orc.simulate(n, enc.ops, enc.n_ops, reg, mode="faithful")
    return time.perf_counter() - t0


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm on the host cores (oracle port; the Rust crate cannot be built
    here).  The reference is single-threaded by construction (README.md:96), so one thread is all it can use.  Each step is
    a bounded sample of the workload: QFT-<ref_qubits>; the full size needs a 2^n-entry hash map."""
    if rank != 0:
        return
    n_sample = args.ref_qubits
    for _ in range(args.warmup):
        cpu_qft_seconds(max(8, n_sample - 3))
    times = [cpu_qft_seconds(n_sample) for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    n_full = args.qubits
    sample = (f"QFT-{n_sample} ({n_qft_gates(n_sample)} gates) per step: the full QFT-{n_full} is out of reach for the reference "
              f"algorithm (2^{n_full}-entry hash map); faithful C++ restatement, 1 thread (the reference is single-threaded)")
    line = {
        "impl": "reference", "metric": "qft_wall_time_ms", "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"QFT-{n_full} complex f64 (bounded sample: QFT-{n_sample}, 2^{n_full - n_sample} times fewer amplitudes)", "qubits": n_full,
                   "sample_qubits": n_sample},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": "port", "sample": sample},
        "same_config": {"workload": f"QFT-{n_sample}", "cpu_reference_ms": ms},
        "gate_equivalent_gbps": n_qft_gates(n_sample) * 32.0 * (1 << n_sample) / (ms * 1e-3) / 1e9,
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- multi-GPU plumbing (torch.distributed is rendezvous only) -----------------------------------------------------

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


def sharded_parity(qb, F, dist, torch, rank, world, local_rank, nccl_id, use_peers):
    """Outside the timed region: a sharded random circuit over every gate kind (n = 16, uploaded register, controls and
    phases on the qubits held in the rank id, several remaps) against the CPU oracle.  Returns (max-abs error, remaps)."""
    import numpy as np
    from helpers import OracleCircuit, encode_gates, orc, random_any_gate_circuit
    n = 16
    g = world.bit_length() - 1
    rng = np.random.default_rng(2024)  # same circuit and register on every rank
    c = random_any_gate_circuit(OracleCircuit, qb.Gate, n, 160, rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    s = qb.DeviceState(n, local_rank, rank=rank, world=world, nccl_id=nccl_id)
    if use_peers:
        setup_peer_exchange(s, dist, world, rank, torch.device("cuda", local_rank))
    nl = n - g
    s.upload(reg[rank << nl:(rank + 1) << nl], first=rank << nl)
    stats = s.apply(enc)
    got = s.gather(np.arange(1 << n, dtype=np.uint64))
    err = None
    if rank == 0:
        ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
        err = float(np.max(np.abs(got - ref)))
    s.peer_import([]) if use_peers else None
    dist.barrier()
    s.close()
    return err, int(stats["n_exchanges"])


# ---- N = 1 extras: BASELINE configs 1-2 through the C ABI, the same-size ratio, sampling ------------------------------

def small_configs(qb, F, device):
    """Wall time per call sequence (host clock around the C-ABI calls, pre-encoded qsv_op arrays, results in host buffers)."""
    import numpy as np
    from golden import reference_vectors as rv
    from helpers import qft_circuit
    from quantr_b200 import states as st
    from quantr_b200.circuit import encode_gates
    out = {}
    # config 1: examples/grovers.rs, 19 gates on 3 qubits, then measure_all(500)
    c = rv.build_example_grovers(qb.Circuit, qb.Gate, st)
    enc = encode_gates(c.get_gates(), 3)
    s = qb.DeviceState(3, device)
    u = np.random.default_rng(1).random(500)

    def c1():
        s.init_basis(0)
        s.apply(enc)
        return s.sample(u)
    for _ in range(20):
        idx = c1()
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        idx = c1()
    out["config1_grover3_simulate_measure_all_500_us"] = (time.perf_counter() - t0) / reps * 1e6
    counts = np.bincount(idx.astype(np.int64), minlength=8)
    out["config1_bins_110_111"] = [int(counts[6]), int(counts[7])]
    out["config1_other_bins_empty"] = bool(counts[:6].sum() == 0)
    s.close()
    # config 2: QFT-16 from |0xACE1>, full state back to the host
    n, x = 16, 0xACE1
    enc = encode_gates(qft_circuit(qb.Circuit, qb.Gate, n).get_gates(), n)
    s = qb.DeviceState(n, device)

    def c2():
        s.init_basis(x)
        s.apply(enc)
        return s.download()
    for _ in range(10):
        amps = c2()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        amps = c2()
    t_hit = (time.perf_counter() - t0) / reps
    from helpers import qft_expected
    out["config2_qft16_simulate_get_state_us"] = t_hit * 1e6
    out["config2_max_abs_err_vs_closed_form"] = float(np.max(np.abs(amps - qft_expected(n, x))))
    # first call of a circuit (no cached plan): lowering + scheduling + schedule upload included
    t_miss = []
    for k in range(5):
        s.set_option("low_bits", 3 + (k % 2))  # a different cache key each time
        t0 = time.perf_counter()
        c2()
        t_miss.append(time.perf_counter() - t0)
    out["config2_qft16_first_call_us"] = min(t_miss) * 1e6
    s.close()
    out["note"] = ("host wall time around qsv_init_basis + qsv_apply + qsv_sample / qsv_download with pre-encoded ops; repeated calls hit "
                   "the handle's plan cache, first_call includes lowering, scheduling and the schedule upload")
    return out


def config3_random_layered(qb, device, n=30, depth=100):
    """BASELINE configs[2]: random layered circuit (H/Rx/Ry/Rz column + n/3 CNot/Toffoli per layer), n = 30, depth 100,
    from |0>: device time of one qsv_apply (plan cached by a first, untimed call), passes, and the norm as a sanity check;
    parity of this generator is in tests/ (oracle at n <= 28, three schedules at n = 30)."""
    from helpers import OracleCircuit, random_layered_circuit
    from quantr_b200.circuit import encode_gates
    c = random_layered_circuit(OracleCircuit, qb.Gate, n, depth, seed=30)
    enc = encode_gates(list(c.circuit_gates), n)
    s = qb.DeviceState(n, device)
    t0 = time.perf_counter()
    s.init_basis(0)
    st = s.apply(enc)
    s.synchronize()
    first_ms = (time.perf_counter() - t0) * 1e3
    s.init_basis(0)
    s.synchronize()
    t0 = time.perf_counter()
    st = s.apply(enc)
    s.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    norm = s.norm_sqr()
    s.close()
    bytes_per_pass = 32.0 * float(1 << n)
    return {"workload": f"random layered circuit, {n} qubits, depth {depth} ({st['n_gates']} gates: H/Rx/Ry/Rz/CNot/Toffoli)", "ms": ms, "first_call_ms": first_ms,
            "fused_passes": st["n_passes"], "rounds": st["n_rounds"], "passes_per_layer": st["n_passes"] / depth,
            "avg_pass_GBps": bytes_per_pass * st["n_passes"] / (ms * 1e-3) / 1e9, "norm_sqr": norm,
            "note": "host wall time around one qsv_apply + synchronize with the plan cached (first_call_ms includes lowering and scheduling of the 4,000 gates)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="qft", choices=["qft", "grover"])
    ap.add_argument("--grover-iterations", type=int, default=1)
    ap.add_argument("--qubits", type=int, default=0, help="total qubits (default 33 + log2(gpus))")
    ap.add_argument("--ref-qubits", type=int, default=16, help="bounded sample size of the CPU reference arm")
    ap.add_argument("--cpu-baseline-qubits", type=int, default=16)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=0)
    ap.add_argument("--shots", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip small_configs / sharded parity (profiling runs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    g = max(0, world.bit_length() - 1)
    if not args.qubits:
        args.qubits = 33 + g
    if args.impl == "reference":
        run_reference(args, rank)
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
    def new_nccl_id():  # one ncclUniqueId per communicator (= per sharded handle): made on rank 0, broadcast by torch
        lib = F.load_library()
        buf = C.create_string_buffer(128)
        if rank == 0:
            F.check(lib.qsv_nccl_unique_id(buf, 128))
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        nccl_id = new_nccl_id()

    n = args.qubits
    n_local = n - g
    wl = qft_workload(n) if args.workload == "qft" else grover_workload(n, args.grover_iterations)
    x, n_gates = wl["basis"], wl["n_gates"]
    enc = encode_gates(wl["gates"], n)

    use_peers = world > 1 and not os.environ.get("QSV_NCCL_EXCHANGE")
    parity_err = parity_remaps = None
    if world > 1 and not args.no_extras:
        parity_err, parity_remaps = sharded_parity(qb, F, dist, torch, rank, world, local_rank, new_nccl_id(), use_peers)

    state = qb.DeviceState(n, local_rank, rank=rank, world=world, nccl_id=nccl_id)
    exchange_path = "nccl send/recv through staging"
    if use_peers and setup_peer_exchange(state, dist, world, rank, torch.device("cuda", local_rank)):
        exchange_path = "in-place swap kernels over peer-mapped memory (NVLink loads/stores)"
    if args.tile_bits:
        state.set_option("tile_bits", args.tile_bits)
    if args.low_bits:
        state.set_option("low_bits", args.low_bits)
    # sharded: the register starts as a basis state, so the scheduler may park the last-targeted qubits in the rank id
    plan = qb.Plan(n, enc, n_local=n_local, tile_bits=args.tile_bits, low_bits=args.low_bits, free_layout=True)  # every step starts from qsv_init_basis
    pstats = plan.stats()
    pdesc = plan.describe()
    passes = pstats["n_passes"]

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

    # ---- timed region: exactly K steps, device-timed on the library's stream, no host synchronisation inside ---------
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev_start.record(ext)
        for k in range(args.steps):
            step()
        ev_end.record(ext)
        barrier()
    total_ms = ev_start.elapsed_time(ev_end)
    overlapped_remaps = state.get_info("overlapped_exchanges") if world > 1 else 0  # of the last timed step

    # ---- the same workload with the prefix folding switched off (N = 1; context for the headline: every stage then runs as
    # part of a pass over the whole register) -------------------------------------------------------------------------
    unfolded = None
    if world == 1 and pdesc.get("prefix_ops", 0) and not args.no_extras:
        try:
            os.environ["QSV_FOLD_PREFIX"] = "0"
            try:
                plan_nf = qb.Plan(n, enc, n_local=n_local, tile_bits=args.tile_bits, low_bits=args.low_bits, free_layout=True)
            finally:
                del os.environ["QSV_FOLD_PREFIX"]
            for _ in range(2):
                state.init_basis(x)
                state.run_plan(plan_nf)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(args.steps):
                state.init_basis(x)
                state.run_plan(plan_nf)
            e1.record(ext)
            barrier()
            unfolded = {"ms_per_step": e0.elapsed_time(e1) / args.steps, "fused_passes": plan_nf.stats()["n_passes"],
                        "note": "QSV_FOLD_PREFIX=0: no gate is folded into the initial amplitudes"}
            plan_nf.close()
        except Exception as e:  # context only: never fatal for the headline line
            unfolded = {"error": str(e)}
        state.init_basis(x)
        state.run_plan(plan)  # the register holds the headline plan's result again for the checks below

    # ---- per-pass device times (CUDA events around every launch, inside the library): a second, separately timed loop --
    state.set_option("timing", 1)
    pass_ms = None
    exchange_ms = 0.0
    for k in range(args.steps):
        st = step()
        per = [ms for (kind, _), ms in zip(plan.steps(), state.last_step_ms()) if kind == "pass"]
        pass_ms = per if pass_ms is None else [a + b for a, b in zip(pass_ms, per)]
        exchange_ms += st["exchange_ms"]
    state.set_option("timing", 0)
    pass_ms = [p / args.steps for p in pass_ms]
    exchange_ms /= args.steps
    barrier()
    if world > 1:
        t = torch.tensor([total_ms, exchange_ms] + pass_ms, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, exchange_ms, pass_ms = float(t[0]), float(t[1]), [float(v) for v in t[2:]]
    ms_per_step = total_ms / args.steps

    # ---- verification outside the timed region ------------------------------------------------------------------
    rng = np.random.default_rng(1234)  # same indices on every rank: qsv_gather is collective on sharded handles
    idx = rng.integers(0, 1 << n, size=4096, dtype=np.uint64)
    got = state.gather(idx)
    norm = state.norm_sqr()
    check = {}
    if wl["check"] == "closed_form_qft":
        rev = np.zeros_like(idx)
        for b in range(n):
            rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(n - 1 - b)
        ph = np.array([(int(r) * x) % (1 << n) for r in rev], dtype=np.float64) * (2 * np.pi / float(1 << n))
        expect = (np.cos(ph) + 1j * np.sin(ph)) / np.sqrt(float(1 << n))
        check["max_abs_err_vs_closed_form"] = float(np.max(np.abs(got - expect)))
    else:
        from workloads import grover_expected_amplitudes
        info = wl["info"]
        marked_amp, other_amp, probe = grover_expected_amplitudes(info, args.grover_iterations)
        got2 = state.gather(np.array(probe["indices"], dtype=np.uint64))
        check["max_abs_err_vs_closed_form"] = float(np.max(np.abs(got2 - np.array(probe["expect"]))))
        check["grover_marked_probability"] = float(abs(marked_amp) ** 2)

    # ---- e2e: the same step through the reference-facing C-ABI calls with HOST buffers ------------------------------
    # qsv_init_basis + qsv_apply(host qsv_op[]: lowering, scheduling, schedule H2D, launches; the first call of the loop
    # is timed too) + qsv_sample(host uniforms -> host indices) + qsv_gather(4096 host indices -> host amplitudes; the
    # whole register does not fit the host, SURVEY.md 7.2 hard part 4)
    uniforms = np.random.default_rng(7).random(args.shots)

    def e2e_step():
        state.init_basis(x)
        state.apply(enc)
        idxs = state.sample(uniforms)
        amps = state.gather(idx)
        return idxs, amps
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e = {"value": e2e_ms, "unit": "ms", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(enc.nbytes + 8 * args.shots + 8 * 4096), "d2h_bytes_per_step": int(8 * args.shots + 16 * 4096),
           "path": "qsv_init_basis + qsv_apply(host qsv_op[]) + qsv_sample(host uniforms -> host indices) + qsv_gather(4096 host indices -> host "
                   "amplitudes; the 128 GiB register itself cannot be brought to a host)"}

    # ---- sampling (K6 + K7), device work measured by host clock around the synchronous C-ABI call --------------------
    state.init_basis(x)
    state.run_plan(plan)
    barrier()
    t0 = time.perf_counter()
    state.sample(uniforms)          # K6 (block sums + scan) + K7
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    state.sample(uniforms)          # prefix cached: K7 + copies only
    t_second = time.perf_counter() - t0
    k6_ms = max(0.0, (t_first - t_second) * 1e3)
    sampling = {"shots": args.shots, "k6_ms": k6_ms, "k6_GBps": (16.0 * (1 << n_local) / (k6_ms * 1e-3) / 1e9) if k6_ms > 0 else None,
                "k7_ms": t_second * 1e3, "note": "K6 = prob_block_sums + scan (one read of the shard, cached until the register changes); K7 = sample_shots incl. H2D/D2H of the shots"}

    # ---- the global-qubit remap in isolation (N > 1): H on wire 0, a qubit held in the rank id, on the finished register ----
    exchange_probe = None
    if world > 1 and args.workload == "qft":
        probe = qb.Circuit.new(n)
        probe.add_gate(qb.Gate.H, 0)
        penc = encode_gates(probe.get_gates(), n)
        state.init_basis(x)
        state.run_plan(plan)
        state.set_option("timing", 1)
        st = state.apply(penc)
        state.set_option("timing", 0)
        t = torch.tensor([st["exchange_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ex_ms = float(t[0])
        exchange_probe = {"circuit": "H on wire 0 (held in the rank id) applied to the QFT result: one remap + one pass",
                          "remaps": st["n_exchanges"], "bytes_sent_per_gpu": st["exchange_bytes"], "ms": ex_ms,
                          "achieved_GBps_per_direction": (st["exchange_bytes"] / (ex_ms * 1e-3) / 1e9) if ex_ms > 0 else None,
                          "nvlink_peak_GBps_per_direction": NVLINK_PEAK, "frac_of_peak": (st["exchange_bytes"] / (ex_ms * 1e-3) / 1e9 / NVLINK_PEAK) if ex_ms > 0 else None}

    peak, peak_src = read_peaks()
    bytes_per_pass = 32.0 * float(1 << n_local)
    # fused initialisation (default; QSV_FUSED_INIT=0 turns it off): the first pass does not read the register
    fused_init = int(os.environ.get("QSV_FUSED_INIT", "2") or 0)
    steps_desc = plan.steps()
    first_is_pass = bool(steps_desc) and steps_desc[0][0] == "pass"
    per_pass = []
    for i, ms in enumerate(pass_ms):
        nbytes = bytes_per_pass * (0.5 if (i == 0 and fused_init and first_is_pass) else 1.0)
        gbps = nbytes / (ms * 1e-3) / 1e9
        per_pass.append({"ms": ms, "algorithmic_bytes": nbytes, "GBps": gbps, "frac": gbps / peak, "rounds": len(pdesc["passes"][i]["rounds"]),
                         "kind": "write-only (fused basis initialisation)" if nbytes < bytes_per_pass else "read+write"})
    rw = [p for p in per_pass if p["kind"] == "read+write"] or per_pass
    rw_bytes, rw_ms = sum(p["algorithmic_bytes"] for p in rw), sum(p["ms"] for p in rw)
    achieved = rw_bytes / (rw_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and n_local == 33:
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("pass_kernel_dram_bytes_per_launch"), tj.get("source", "static: profiles/traffic.json (ncu capture, not this run)")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "pass_kernel_tma (fused tile pass; read+write launches)", "algorithmic_bytes_per_launch": rw_bytes / len(rw),
                "avg_launch_ms": rw_ms / len(rw), "min_pass_frac": min(p["frac"] for p in rw), "per_pass": per_pass,
                "peak_source": peak_src, "per_gpu": True}
    all_bytes = sum(p["algorithmic_bytes"] for p in per_pass) * world
    n_swap_launches = pstats["n_exchanges"] * (world - 1) if exchange_path.startswith("in-place") else 0
    launches_per_step = passes + (0 if (fused_init and first_is_pass) else 2) + n_swap_launches

    if rank == 0:
        cpu = same = small = config3 = None
        if world == 1 and not args.no_extras:
            small = small_configs(qb, F, local_rank)
            state.close()  # 128 GiB back before the 16 GiB register of config 3 is allocated
            state = None
            try:
                config3 = config3_random_layered(qb, local_rank)
            except Exception as e:  # reported, never fatal for the headline line
                config3 = {"error": str(e)}
        if not args.no_cpu_baseline:
            nb = args.cpu_baseline_qubits
            dt = cpu_qft_seconds(nb)
            cpu = {"value": dt * 1e3, "unit": "ms", "cores": 1, "kind": "port", "seconds": dt,
                   "sample": f"QFT-{nb} ({n_qft_gates(nb)} gates, 2^{n - nb} times fewer amplitudes than the GPU step), faithful C++ restatement of "
                             f"simulation.rs:64-135 (per-gate hash-map rebuild), 1 thread; QFT-{n} itself is unreachable for that algorithm"}
            if small and nb == 16:
                same = {"workload": "QFT-16 from |0xACE1> (BASELINE configs[1]), full state back on the host",
                        "gpu_e2e_ms": small["config2_qft16_simulate_get_state_us"] / 1e3, "gpu_e2e_first_call_ms": small["config2_qft16_first_call_us"] / 1e3,
                        "cpu_reference_ms": dt * 1e3, "speedup": dt * 1e3 / (small["config2_qft16_simulate_get_state_us"] / 1e3)}
        line = {
            "metric": "qft_wall_time_ms" if args.workload == "qft" else "grover_wall_time_ms", "value": ms_per_step, "unit": "ms", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "qubits": n, "local_qubits": n_local, "state_bytes_per_gpu": 16 << n_local,
                       "l2_policy": "state >> 126 MB L2 (no flush needed)",
                       "init": ("fused into the first pass (mode %d): zero tiles are written by bulk tensor stores, the register is not read" % fused_init)
                       if (fused_init and first_is_pass) else "memset + set_amp before the first pass",
                       "tile_bits": pdesc["tile_bits"], "low_bits": pdesc["low_bits"], "parallelism": f"shard{world}"},
            "fused_passes": passes, "passes_per_gate": passes / n_gates,
            "effective_hbm_gbps": all_bytes / (ms_per_step * 1e-3) / 1e9,
            "gate_equivalent_gbps": n_gates * 32.0 * float(1 << n) / (ms_per_step * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "sampling": sampling, "small_configs": small, "config3": config3, "same_config": same,
            "exchange": None if world == 1 else {
                "remaps_per_step": pstats["n_exchanges"], "bytes_sent_per_gpu_per_step": pstats["exchange_bytes"],
                "ms_per_step": exchange_ms, "exposed_ms": max(0.0, ms_per_step - sum(pass_ms)),
                "pipelined_remaps_per_step": overlapped_remaps,
                "note": "ms_per_step = the remaps run on their own (separately timed loop); exposed_ms = timed step - sum of the passes' own times: "
                        "what the remaps add to the step when they run slice by slice next to their neighbouring passes",
                "achieved_GBps_per_direction": (pstats["exchange_bytes"] / (exchange_ms * 1e-3) / 1e9) if exchange_ms > 0 else None,
                "path": exchange_path, "nvlink_peak_GBps_per_direction": NVLINK_PEAK, "peak_source": "B200_PROFILING.md measured peer copy"},
            "exchange_probe": exchange_probe, "prefix_ops_folded_into_initial_state": pdesc.get("prefix_ops", 0), "without_prefix_folding": unfolded,
            "prefix": {"lowered_ops": pdesc.get("prefix_ops", 0), "local_qubits": pdesc.get("prefix_local_bits", 0), "rank_qubits": g,
                       "note": "leading gates on the top qubits of the basis state act on a product state: they are applied, on every step, to its "
                               "2^(rank_qubits + local_qubits) non-zero amplitudes only (host up to 14 local qubits, a device sub-register beyond; its "
                               "kernel launches are not in gpu_launches); the first pass synthesises its tiles from that table. QSV_FOLD_PREFIX=0 turns it off"},
            "sharded_parity_max_abs_err": parity_err, "sharded_parity_remaps": parity_remaps,
            "clocks": clocks.summary(), "norm_sqr": norm,
        }
        line.update(check)
        print(json.dumps(line), flush=True)
    if world > 1:
        # importers unmap their peers before any exporter frees its shard (CUDA IPC teardown order)
        state.peer_import([])
        dist.barrier()
    if state is not None:
        state.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
