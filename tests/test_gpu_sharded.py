"""Two-GPU parity: a sharded register (NCCL global-qubit exchange over NVLink) against the CPU oracle.
Skipped on boxes with fewer than two GPUs; CPU coverage of the same path is in tests/test_sharded.py."""
import os
import sys

import numpy as np
import pytest

from helpers import OracleCircuit, encode_gates, orc, qb, qft_circuit, qft_expected, random_any_gate_circuit

pytestmark = pytest.mark.gpu
G = qb.Gate


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, result_dir):
    import ctypes as C

    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from quantr_b200 import _ffi as F
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = F.load_library()

    def new_nccl_id():  # one ncclUniqueId per communicator, made on rank 0 and broadcast by the caller (here: gloo)
        buf = C.create_string_buffer(128)
        if rank == 0:
            F.check(lib.qsv_nccl_unique_id(buf, 128))
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        return bytes(t.numpy().tobytes())

    nccl_id = new_nccl_id()
    results = {}
    # QFT: one remap, free initial layout
    n, x = 16, 0xACE1
    enc = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
    s = qb.DeviceState(n, rank, rank=rank, world=world, nccl_id=nccl_id)
    s.set_option("tile_bits", 8)
    s.init_basis(x)
    stats = s.apply(enc)
    allidx = np.arange(1 << n, dtype=np.uint64)
    results["qft"] = s.gather(allidx)
    results["qft_exchanges"] = stats["n_exchanges"]
    results["qft_norm"] = s.norm_sqr()
    u = np.random.default_rng(5).random(20000)
    results["qft_samples"] = s.sample(u)
    # the same without folding the leading rank-qubit stages into the initial amplitudes: one global-qubit remap
    os.environ["QSV_FOLD_PREFIX"] = "0"
    s.set_option("tile_bits", 9)  # another plan-cache key: the handle would otherwise re-run the folded plan
    s.init_basis(x)
    stats = s.apply(enc)
    results["qft_unfolded"] = s.gather(allidx)
    results["qft_unfolded_exchanges"] = stats["n_exchanges"]
    del os.environ["QSV_FOLD_PREFIX"]
    s.close()
    # random circuit on an uploaded register (canonical layout, several remaps)
    n2 = 14
    rng = np.random.default_rng(99)
    c = random_any_gate_circuit(OracleCircuit, G, n2, 120, rng)
    enc2 = encode_gates(c.circuit_gates, n2)
    reg = rng.normal(size=1 << n2) + 1j * rng.normal(size=1 << n2)
    reg /= np.linalg.norm(reg)
    s2 = qb.DeviceState(n2, rank, rank=rank, world=world, nccl_id=new_nccl_id())
    s2.set_option("tile_bits", 7)
    # this register exchanges through peer-mapped memory (NVLink loads/stores) instead of NCCL send/recv
    mine = torch.frombuffer(bytearray(s2.peer_export()), dtype=torch.uint8).clone()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    s2.peer_import([bytes(h.numpy().tobytes()) for h in allh])
    nl = n2 - (world.bit_length() - 1)
    s2.upload(reg[rank << nl:(rank + 1) << nl], first=rank << nl)
    st2 = s2.apply(enc2)
    results["rand"] = s2.gather(np.arange(1 << n2, dtype=np.uint64))
    results["rand_exchanges"] = st2["n_exchanges"]
    results["rand_layout_identity"] = s2.layout() == list(range(n2))
    # remapped layout: qsv_download is collective and un-permutes on the device; canonical layout: every rank reads its shard
    if results["rand_layout_identity"]:
        results["rand_download"] = results["rand"]
        results["rand_download_shard"] = s2.download(rank << nl, 1 << nl) if rank == 0 else np.zeros(1)
    else:
        results["rand_download"] = s2.download(0, 1 << n2)
    s2.peer_import([])  # importers unmap their peers before any exporter frees its shard (CUDA IPC teardown order)
    dist.barrier()
    s2.close()
    if rank == 0:
        np.savez(os.path.join(result_dir, "out.npz"), **results)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_register_matches_oracle(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npz")
    n, x = 16, 0xACE1
    assert int(out["qft_exchanges"]) == 0  # the first stage (the rank qubit) is folded into the ranks' initial amplitudes
    assert np.max(np.abs(out["qft"] - qft_expected(n, x))) < 1e-12
    assert int(out["qft_unfolded_exchanges"]) == 1
    assert np.max(np.abs(out["qft_unfolded"] - qft_expected(n, x))) < 1e-12
    assert abs(float(out["qft_norm"]) - 1.0) < 1e-12
    counts = np.bincount(out["qft_samples"].astype(np.int64), minlength=1 << n)
    assert counts.sum() == 20000 and counts.max() <= 6  # uniform distribution over 65536 outcomes
    n2 = 14
    rng = np.random.default_rng(99)
    c = random_any_gate_circuit(OracleCircuit, G, n2, 120, rng)
    enc2 = encode_gates(c.circuit_gates, n2)
    reg = rng.normal(size=1 << n2) + 1j * rng.normal(size=1 << n2)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n2, enc2.ops, enc2.n_ops, reg, mode="dense")
    assert int(out["rand_exchanges"]) >= 1
    assert np.max(np.abs(out["rand"] - ref)) < 1e-12
    assert np.max(np.abs(out["rand_download"] - ref)) < 1e-12
    if bool(out["rand_layout_identity"]):
        assert np.max(np.abs(out["rand_download_shard"] - ref[:1 << (n2 - 1)])) < 1e-12


def _worker_pipelined(rank, world, port, result_dir):
    """Random circuit on an uploaded register through peer memory, pipelined TMA kernel forced onto a small register
    (QSV_ASYNC=2, set by the parent): the remaps run slice by slice against their neighbouring passes."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ctypes as C
    from quantr_b200 import _ffi as F
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = F.load_library()
    buf = C.create_string_buffer(128)
    if rank == 0:
        F.check(lib.qsv_nccl_unique_id(buf, 128))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    n = 19
    rng = np.random.default_rng(4242)
    c = random_any_gate_circuit(OracleCircuit, G, n, 150, rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    s = qb.DeviceState(n, rank, rank=rank, world=world, nccl_id=bytes(t.numpy().tobytes()))
    s.set_option("tile_bits", 11)
    mine = torch.frombuffer(bytearray(s.peer_export()), dtype=torch.uint8).clone()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    s.peer_import([bytes(h.numpy().tobytes()) for h in allh])
    nl = n - (world.bit_length() - 1)
    results = {}
    for label, overlap, k in (("serial", 0, 2), ("pipelined2", 1, 1), ("pipelined4", 1, 2), ("pipelined8", 1, 3)):
        s.set_option("overlap", overlap)
        s.set_option("exchange_slices_log2", k)
        s.init_basis(0)  # back to the canonical layout: the previous run left the qubits remapped
        s.upload(reg[rank << nl:(rank + 1) << nl], first=rank << nl)
        st = s.apply(enc)
        results[label] = s.gather(np.arange(1 << n, dtype=np.uint64))
        results[label + "_exchanges"] = st["n_exchanges"]
        results[label + "_overlapped"] = s.get_info("overlapped_exchanges")
    s.peer_import([])
    dist.barrier()
    s.close()
    if rank == 0:
        np.savez(os.path.join(result_dir, "out.npz"), **results)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_two_gpu_pipelined_exchange_matches_oracle(tmp_path, monkeypatch):
    """north_star: global-qubit swaps over NVLink overlapped with local fused passes (state_api.cu run_overlapped)."""
    import torch.multiprocessing as mp
    monkeypatch.setenv("QSV_ASYNC", "2")
    world = 2
    port = 31700 + (os.getpid() % 2000)
    mp.spawn(_worker_pipelined, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npz")
    n = 19
    rng = np.random.default_rng(4242)
    c = random_any_gate_circuit(OracleCircuit, G, n, 150, rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
    assert int(out["serial_exchanges"]) >= 1 and int(out["serial_overlapped"]) == 0
    assert np.max(np.abs(out["serial"] - ref)) < 1e-12
    for label in ("pipelined2", "pipelined4", "pipelined8"):
        assert int(out[label + "_overlapped"]) >= 1, label
        assert np.max(np.abs(out[label] - ref)) < 1e-12, label


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_in_library_multi_gpu_handle_matches_oracle(tmp_path):
    """'single-process, multi-device inside the library, invisible to the caller' - one handle from
    qsv_create_multi, no launcher, no NCCL id or IPC handles in the caller's hands; the reference-facing Circuit API on top
    of it through QSV_DEVICES (Circuit::simulate, src/circuit.rs:364-388, knows nothing about devices)."""
    n = 15
    rng = np.random.default_rng(77)
    c = random_any_gate_circuit(OracleCircuit, G, n, 140, rng)
    enc = encode_gates(c.circuit_gates, n)
    reg = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    reg /= np.linalg.norm(reg)
    ref = orc.simulate(n, enc.ops, enc.n_ops, reg, mode="dense", threads=4)
    s = qb.DeviceState(n, devices=[0, 1])
    assert s.get_info("devices") == 2 and s.get_info("n_local_qubits") == n - 1
    s.set_option("tile_bits", 7)
    # uploaded register (any range: here two pieces that straddle the shard boundary), random circuit with remaps
    cut = (1 << (n - 1)) + 37
    s.upload(reg[:cut], first=0)
    s.upload(reg[cut:], first=cut)
    st = s.apply(enc)
    assert st["n_exchanges"] >= 1
    allidx = np.arange(1 << n, dtype=np.uint64)
    assert np.max(np.abs(s.gather(allidx) - ref)) < 1e-12
    assert np.max(np.abs(s.download(0, 1 << n) - ref)) < 1e-12  # un-permuting download if the run left the qubits remapped
    assert np.max(np.abs(s.download(1000, 5000) - ref[1000:6000])) < 1e-12
    assert abs(s.norm_sqr() - 1.0) < 1e-12
    # sampling over the whole register: empirical distribution against |ref|^2
    shots = 200000
    idx = s.sample(np.random.default_rng(5).random(shots)).astype(np.int64)
    probs = np.abs(ref) ** 2
    top = np.argsort(probs)[-8:]
    counts = np.bincount(idx, minlength=1 << n)
    for t in top:
        assert abs(counts[t] / shots - probs[t]) < 6 * np.sqrt(probs[t] / shots) + 1e-4
    # basis state + QFT: the leading stage on the rank qubit is folded into the shards' initial amplitudes
    x = 0x2ACE
    encq = encode_gates(qft_circuit(OracleCircuit, G, n).circuit_gates, n)
    s.init_basis(x)
    stq = s.apply(encq)
    assert stq["n_exchanges"] == 0
    assert np.max(np.abs(s.gather(allidx) - qft_expected(n, x))) < 1e-12
    # checkpoint: one file per shard
    path = str(tmp_path / "ck")
    s.save(path)
    assert os.path.exists(path + ".r0") and os.path.exists(path + ".r1")
    s.init_basis(0)
    s.load(path)
    assert np.max(np.abs(s.gather(allidx) - qft_expected(n, x))) < 1e-12
    s.close()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_circuit_api_over_two_gpus_via_qsv_devices(monkeypatch):
    """The drop-in surface itself: Circuit.simulate / get_state / measure_all with QSV_DEVICES=0,1."""
    monkeypatch.setenv("QSV_DEVICES", "0,1")
    n = 12
    got = qft_circuit(qb.Circuit, G, n, 0xACE).simulate()
    want = qft_circuit(OracleCircuit, G, n, 0xACE).simulate().get_state().take().get_amplitudes()
    assert got._state.get_info("devices") == 2
    assert np.max(np.abs(got.get_state().take().get_amplitudes() - want)) < 1e-12
    bins = got.measure_all(2000).take()
    assert sum(bins.values()) == 2000
