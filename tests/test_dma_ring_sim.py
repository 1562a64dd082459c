"""CPU: discrete-event simulation of the copy-engine ring protocol (landiff_b200/dma_ring.py + RingAttention._attention_dma).

The REAL host code runs for every rank of a ring (PeerRing.push / wait_arrival / release and the hop schedule); only the
three C-ABI calls it makes (stream wait-value, stream write-value, async copy), torch's CUDA streams / events and the
attention launch are replaced by queue entries.  A simulator then executes the per-stream queues in random
interleavings with the semantics of the real primitives (in-order streams, wait blocks its stream until the flag is
reached, a copy moves whatever the source holds WHEN IT EXECUTES) and checks, for rings of 2, 3, 4 and 8 ranks over
several back-to-back calls (layers) whose K|V buffer is overwritten each call:
  * no deadlock under any explored interleaving,
  * every attention reads exactly the shard the ring schedule prescribes, of the right call — i.e. no receive buffer
    and no local K|V buffer is overwritten before its last reader has run, and nothing is read before it has landed.
"""
import random

import pytest
import torch

from landiff_b200 import dma_ring, parallel


class Sim:
    def __init__(self):
        self.queues = {}        # stream id -> list of ops
        self.flags = {}         # address -> uint32 value
        self.tensors = {}       # data_ptr -> tensor (copy targets / sources)
        self.tokens_done = set()
        self.log = []           # (rank, call, hop, shard, shard_call)
        self.next_token = 0

    def stream(self, sid):
        self.queues.setdefault(sid, [])
        return sid

    def enqueue(self, sid, op):
        self.queues[sid].append(op)

    def run(self, rng):
        while True:
            live = [s for s, q in self.queues.items() if q]
            if not live:
                return
            runnable = [s for s in live if self._ready(self.queues[s][0])]
            assert runnable, "deadlock: " + "; ".join(f"stream {s}: {self.queues[s][0][0]}" for s in live)
            s = rng.choice(runnable)
            self._exec(self.queues[s].pop(0))

    def _ready(self, op):
        kind = op[0]
        if kind == "wait_geq":
            return ((self.flags.get(op[1], 0) - op[2]) & 0xFFFFFFFF) < 0x80000000
        if kind == "wait_token":
            return op[1] in self.tokens_done
        return True

    def _exec(self, op):
        kind = op[0]
        if kind == "write":
            self.flags[op[1]] = op[2]
        elif kind == "copy":
            dst, src = self.tensors[op[1]], self.tensors[op[2]]
            dst.view(-1).copy_(src.view(-1))
        elif kind == "record":
            self.tokens_done.add(op[1])
        elif kind == "call":
            op[1]()


class FakeLib:
    """The three C-ABI entry points PeerRing uses, as queue entries."""

    def __init__(self, sim):
        self.sim = sim

    def ld_stream_wait_geq_u32(self, addr, value, stream):
        self.sim.enqueue(stream, ("wait_geq", addr, value))
        return 0

    def ld_stream_write_u32(self, addr, value, stream):
        self.sim.enqueue(stream, ("write", addr, value))
        return 0

    def ld_copy_async(self, dst, src, nbytes, stream):
        self.sim.enqueue(stream, ("copy", dst, src))
        return 0


class FakeStream:
    def __init__(self, sim, sid):
        self.sim, self.cuda_stream = sim, sim.stream(sid)

    def wait_event(self, ev):
        self.sim.enqueue(self.cuda_stream, ("wait_token", ev.token))


class FakeEvent:
    sim = None

    def __init__(self, *a, **k):
        self.token = None

    def record(self, stream):
        FakeEvent.sim.next_token += 1
        self.token = FakeEvent.sim.next_token
        FakeEvent.sim.enqueue(stream.cuda_stream, ("record", self.token))


def build_ring(sim, sp):
    """Per rank: a RingAttention in dma mode whose PeerRing is wired to its neighbours through simulated memory."""
    lib = FakeLib(sim)
    shape = (2, 4)
    recv = [[torch.zeros(shape) for _ in range(2)] for _ in range(sp)]
    flag_base = [1000 * (r + 1) for r in range(sp)]          # fake device addresses of each rank's flag array
    rings = []
    for r in range(sp):
        for t in recv[r]:
            sim.tensors[t.data_ptr()] = t
        pr = object.__new__(dma_ring.PeerRing)
        pr.lib, pr.device, pr.nbytes, pr.shape, pr.dtype = lib, "cpu", 2 * 4 * 4, shape, torch.float32
        pr._flags_ptr = flag_base[r]
        down, up = (r + 1) % sp, (r - 1) % sp
        pr._down_recv = [t.data_ptr() for t in recv[down]]
        pr._down_flags, pr._up_flags = flag_base[down], flag_base[up]
        pr.recv = recv[r]
        pr.next_id, pr.last_sent = 1, [0, 0]
        ra = object.__new__(parallel.RingAttention)
        ra.layout = parallel.Layout(sp, r, 1, sp)
        ra.group, ra.device, ra.transport = None, "cpu", "dma"
        ra.comm_stream = FakeStream(sim, f"comm{r}")
        ra.compute_stream = FakeStream(sim, f"compute{r}")
        ra._bufs, ra._peer = {}, {}
        ra._peer_ring = lambda kv, pr=pr: pr
        kv = torch.zeros(shape)
        sim.tensors[kv.data_ptr()] = kv
        ws = dict(q=torch.zeros(1, 1, 1, 64), kv=kv, attn=torch.zeros(1, 1, 64))
        rings.append((ra, ws))
    return rings


def simulate(sp, seed, monkeypatch, attention_dma=None, n_calls=5):
    """Host phase for every rank (all calls enqueued up front), then one random interleaving.  Returns the read log;
    raises AssertionError on deadlock."""
    from landiff_b200 import ops

    sim = Sim()
    FakeEvent.sim = sim
    rings = build_ring(sim, sp)
    current = {}
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: current["ra"].compute_stream)
    monkeypatch.setattr(dma_ring, "check", lambda rc, what: None)

    def fake_attention(q, k, v, out=None, lse=None, out_f32=None, variant=0):
        ra, call, hop = current["ra"], current["call"], current["hop_counter"][0]
        current["hop_counter"][0] += 1
        r = ra.layout.rank

        def read(k=k, r=r, call=call, hop=hop):
            sim.log.append((r, call, hop, int(k.view(-1)[0]), int(k.view(-1)[1])))

        sim.enqueue(ra.compute_stream.cuda_stream, ("call", read))

    monkeypatch.setattr(ops, "attention", fake_attention)
    monkeypatch.setattr(ops, "attention_merge", lambda *a, **k: None)
    # each call starts with the "QKV GEMM" overwriting the local K|V buffer with (rank, call); ranks are interleaved
    # arbitrarily on the host, calls stay in order
    order = [(c, r) for c in range(n_calls) for r in range(sp)]
    random.Random(seed).shuffle(order)
    order.sort(key=lambda cr: cr[0])
    for call, r in order:
        ra, ws = rings[r]
        current.update(ra=ra, call=call, hop_counter=[0])

        def produce(kv=ws["kv"], r=r, call=call):
            kv.view(-1)[0], kv.view(-1)[1] = float(r), float(call)

        sim.enqueue(ra.compute_stream.cuda_stream, ("call", produce))
        (attention_dma or parallel.RingAttention._attention_dma)(ra, ws, 0)
    sim.run(random.Random(1000 + seed))
    return sim.log


def reads_are_correct(log, sp, n_calls=5):
    return len(log) == sp * n_calls * sp and all(s == (r - h) % sp and sc == c for r, c, h, s, sc in log)


@pytest.mark.parametrize("sp", [2, 3, 4, 8])
def test_protocol_is_safe_and_live_under_random_interleavings(sp, monkeypatch):
    for seed in range(12):
        log = simulate(sp, seed, monkeypatch)
        for r, call, hop, shard, shard_call in log:
            assert shard == (r - hop) % sp and shard_call == call, \
                f"sp={sp} seed={seed}: rank {r} call {call} hop {hop} read shard {shard} of call {shard_call}"
        assert reads_are_correct(log, sp)


MUTATIONS = {
    "no guard of the local K|V buffer at the end of a call":
        ("    done = torch.cuda.Event()\n    done.record(self.comm_stream)\n    compute.wait_event(done)\n", ""),
    "release before the forwarding copy has read the buffer": ("                compute.wait_event(fwd)\n", ""),
    "attention does not wait for the arrival": ("            pr.wait_arrival(cur_j, cur_T, compute)\n", "            pass\n"),
    "forward does not wait for the arrival":
        ("                pr.wait_arrival(cur_j, cur_T, self.comm_stream)   # forward as soon as it has landed\n",
         "                pass\n"),
    "buffers are never released": ("            pr.release(cur_j, cur_T, compute)\n", "            pass\n"),
}


@pytest.mark.parametrize("name", sorted(MUTATIONS))
def test_every_guard_of_the_schedule_is_necessary(name, monkeypatch):
    """Mutation check: removing any single wait / release from RingAttention._attention_dma makes some interleaving
    read the wrong shard or deadlock — so the simulation above really exercises those guards."""
    import inspect
    import textwrap

    old, new = MUTATIONS[name]
    src = textwrap.dedent(inspect.getsource(parallel.RingAttention._attention_dma))
    assert old in src, "the mutation no longer matches the source; update MUTATIONS"
    ns = {}
    exec(src.replace(old, new), dict(vars(parallel), torch=torch), ns)
    broken = 0
    for sp in (2, 4):
        for seed in range(16):
            try:
                broken += not reads_are_correct(simulate(sp, seed, monkeypatch, attention_dma=ns["_attention_dma"]), sp)
            except AssertionError:      # deadlock
                broken += 1
    assert broken > 0, name


def test_simulator_catches_a_missing_release_wait():
    """Sanity of the harness: without the sender's wait on free[j] a fast sender overwrites an unread buffer."""
    sim = Sim()
    a, b = torch.zeros(2), torch.zeros(2)
    sim.tensors[1], sim.tensors[2] = a, b
    s1, s2 = sim.stream("send"), sim.stream("recv")
    seen = []
    for T in (1, 2):
        sim.enqueue(s1, ("call", lambda T=T: a.fill_(T)))
        sim.enqueue(s1, ("copy", 2, 1))
        sim.enqueue(s1, ("write", 10, T))
        sim.enqueue(s2, ("wait_geq", 10, T))
        sim.enqueue(s2, ("call", lambda: seen.append(int(b[0]))))
    bad = 0
    for seed in range(40):
        for q in sim.queues.values():
            q[:] = list(q)
        sim2 = Sim()
        sim2.tensors, sim2.queues = sim.tensors, {k: list(v) for k, v in sim.queues.items()}
        a.zero_(); b.zero_(); seen.clear()
        sim2.run(random.Random(seed))
        bad += seen != [1, 2]
    assert bad > 0
