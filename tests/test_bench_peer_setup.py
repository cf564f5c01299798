"""bench.py's peer-exchange set-up must leave every rank in the same transport mode, whatever fails where: the remap
protocols differ (swap kernels between two barriers vs grouped send/recv), so a split decision would deadlock.
Exercised with two gloo processes and fake registers."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeState:
    def __init__(self, rank, fail_export_on, fail_import_on):
        self.rank, self.fail_export_on, self.fail_import_on = rank, fail_export_on, fail_import_on
        self.imported = None

    def peer_export(self):
        if self.rank == self.fail_export_on:
            raise RuntimeError("no IPC")
        return bytes([self.rank + 1]) * 64

    def peer_import(self, handles):
        if handles and self.rank == self.fail_import_on:
            raise RuntimeError("no peer access")
        self.imported = list(handles)


def _worker(rank, world, port, case, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fail_export_on, fail_import_on = case
    st = FakeState(rank, fail_export_on, fail_import_on)
    active = bench.setup_peer_exchange(st, dist, world, rank, torch.device("cpu"))
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write(f"{int(active)} {len(st.imported) if st.imported is not None else -1}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,expect_active", [((-1, -1), True), ((1, -1), False), ((-1, 0), False), ((0, 1), False)])
def test_every_rank_ends_in_the_same_transport_mode(case, expect_active, tmp_path):
    import torch.multiprocessing as mp
    port = 29650 + abs(hash(case)) % 200
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    results = [open(tmp_path / f"r{r}.txt").read().split() for r in range(2)]
    assert all(int(a) == int(expect_active) for a, _ in results), results
    for r, (_, n_imported) in enumerate(results):
        if expect_active:
            assert int(n_imported) == 2          # both handles mapped
        else:
            assert int(n_imported) in (-1, 0)    # never imported, or dropped again
