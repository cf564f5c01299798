"""Model check of the pipelined pass kernel's buffer ring (quantr_b200/csrc/pass_kernel_tma.cu; first written for its cp.async predecessor).

kG compute groups consume tiles k = g, g + kG, ... from a ring of kNB shared-memory buffers (tile k lives in buffer
k % kNB).  The consumer of tile k refills its buffer with tile k + kNB; a consumer learns that its tile has landed from
the buffer's mbarrier, polled by *parity*.  The model walks random interleavings of the groups and of the asynchronous
load completions and checks that every consumer reads exactly its own, completely loaded tile.

It documents a race found on hardware in round 1: without the per-buffer issue counter a group that runs
two tiles ahead of the group refilling its next buffer sees the parity of the phase before last and consumes a stale
buffer; with the counter (`issued[]`, st.release / ld.acquire in the kernel) no interleaving does.
"""
import random

import pytest


class Buffer:
    def __init__(self):
        self.completed = 0      # mbarrier phases completed so far
        self.content = None     # tile held (None while a load is in flight)
        self.issued = 0         # loads issued into this buffer (the kernel's issued[] counter)
        self.in_flight = None   # tile being loaded


def parity_wait_passes(buf, phase):
    """mbarrier.try_wait.parity: true iff the phase with this parity is not the one in progress."""
    return (buf.completed & 1) != (phase & 1)


def run(seed, n_buffers, n_groups, n_tiles, guarded):
    rng = random.Random(seed)
    bufs = [Buffer() for _ in range(n_buffers)]

    def issue(k):
        b = bufs[k % n_buffers]
        assert b.in_flight is None, "two loads in flight into one buffer"
        b.in_flight, b.content = k, None
        b.issued = k // n_buffers + 1

    for k in range(min(n_buffers, n_tiles)):   # prologue
        issue(k)
    nxt = list(range(n_groups))                # next tile of every group
    while True:
        actors = [("load", i) for i, b in enumerate(bufs) if b.in_flight is not None]
        for g in range(n_groups):
            k = nxt[g]
            if k >= n_tiles:
                continue
            b = bufs[k % n_buffers]
            if guarded and b.issued < k // n_buffers + 1:
                continue                       # wait_issued() spins
            if parity_wait_passes(b, k // n_buffers):
                actors.append(("consume", g))
        if not actors:
            if all(k >= n_tiles for k in nxt):
                return None
            return f"deadlock with next tiles {nxt}"
        kind, i = rng.choice(actors)
        if kind == "load":
            b = bufs[i]
            b.content, b.in_flight = b.in_flight, None
            b.completed += 1
        else:
            k = nxt[i]
            b = bufs[k % n_buffers]
            if b.content != k:
                return f"group {i} consumed buffer {k % n_buffers} holding {b.content} (in flight {b.in_flight}) instead of tile {k}"
            if k + n_buffers < n_tiles:
                issue(k + n_buffers)           # release + refill (after the round's loads, before its arithmetic)
            nxt[i] = k + n_groups


@pytest.mark.parametrize("n_buffers,n_groups", [(6, 4), (3, 2)])
def test_guarded_ring_never_consumes_a_stale_buffer(n_buffers, n_groups):
    for seed in range(400):
        assert run(seed, n_buffers, n_groups, 97, guarded=True) is None


def test_unguarded_parity_wait_is_racy():
    """The model reproduces the hardware failure: parity alone lets a group overtake the refill of its buffer."""
    failures = [run(seed, 6, 4, 97, guarded=False) for seed in range(400)]
    assert any(f is not None for f in failures)
