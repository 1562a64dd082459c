"""CPU: discrete-event simulation of the copy-engine exchange protocols (landiff_b200/dma_ring.py): the per-layer K|V
all-gather (RingAttention._attention_dma) and, at the end of the file, the per-step output exchange (OutputGather.gather).

The REAL host code runs for every rank of a sequence-parallel group (PeerGather.push_all / kernel_shards / release_all
and the per-layer schedule); only the three C-ABI calls it makes (stream wait-value, stream write-value, async copy),
torch's CUDA streams / events and the attention launch are replaced by queue entries.  The multi-shard attention kernel
is modelled as what it does: for every shard in order, wait for the shard's arrival flag (the in-kernel poll blocks that
kernel, hence its stream), then read it.  A simulator then executes the per-stream queues in random interleavings with
the semantics of the real primitives (in-order streams, a wait blocks its stream until the flag is reached, a copy moves
whatever the source holds WHEN IT EXECUTES) and checks, for groups of 2, 3 and 4 ranks over several back-to-back calls
(layers) whose K|V buffer is overwritten each call:
  * no deadlock under any explored interleaving,
  * every attention reads exactly the shards the schedule prescribes (local, rank-1, rank-2, ...), of the right call —
    i.e. no receive buffer and no local K|V buffer is overwritten before its last reader has run, and nothing is read
    before it has landed.
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


def build_group(sim, sp):
    """Per rank: a RingAttention in dma mode whose PeerGather is wired to its peers through simulated memory."""
    lib = FakeLib(sim)
    shape = (2, 4)
    recv = [[torch.zeros(shape) for _ in range(sp - 1)] for _ in range(sp)]
    flag_base = [1000 * (r + 1) for r in range(sp)]          # fake device addresses of each rank's flag page
    ranks = []
    for r in range(sp):
        for t in recv[r]:
            sim.tensors[t.data_ptr()] = t
        pg = object.__new__(dma_ring.PeerGather)
        pg.lib, pg.device, pg.nbytes, pg.shape, pg.dtype = lib, "cpu", 2 * 4 * 4, shape, torch.float32
        pg.n, pg.me = sp, r
        pg._flags_ptr = flag_base[r]
        pg._peer_recv = {d: recv[(r + d) % sp][d - 1].data_ptr() for d in range(1, sp)}
        pg._peer_flags = {d: flag_base[(r + d) % sp] for d in range(1, sp)}
        pg.recv = recv[r]
        pg.next_id, pg.last_sent = 1, 0
        ra = object.__new__(parallel.RingAttention)
        ra.layout = parallel.Layout(sp, r, 1, sp)
        ra.group, ra.device, ra.transport = None, "cpu", "dma"
        ra.comm_stream = FakeStream(sim, f"comm{r}")
        ra.compute_stream = FakeStream(sim, f"compute{r}")
        ra._bufs, ra._peer = {}, {}
        ra._peer_gather = lambda kv, pg=pg: pg
        kv = torch.zeros(shape)
        sim.tensors[kv.data_ptr()] = kv
        ws = dict(q=torch.zeros(1, 1, 1, 64), kv=kv, attn=torch.zeros(1, 1, 64))
        ranks.append((ra, ws, pg))
    return ranks


def simulate(sp, seed, monkeypatch, attention_dma=None, n_calls=5, mutate=None):
    """Host phase for every rank (all calls enqueued up front), then one random interleaving.  Returns the read log;
    raises AssertionError on deadlock."""
    from landiff_b200 import ops

    sim = Sim()
    FakeEvent.sim = sim
    ranks = build_group(sim, sp)
    current = {}
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: current["ra"].compute_stream)
    monkeypatch.setattr(dma_ring, "check", lambda rc, what: None)

    def fake_attention_shards(q, shards, out=None, variant=0, **kw):
        ra, call = current["ra"], current["call"]
        r = ra.layout.rank
        for j, sh in enumerate(shards):
            k = sh[0]
            if len(sh) >= 5 and sh[3]:
                sim.enqueue(ra.compute_stream.cuda_stream, ("wait_geq", sh[3], sh[4]))   # the kernel's poll of the flag

            def read(k=k, r=r, call=call, j=j):
                sim.log.append((r, call, j, int(k.view(-1)[0]), int(k.view(-1)[1])))

            sim.enqueue(ra.compute_stream.cuda_stream, ("call", read))

    monkeypatch.setattr(ops, "attention_shards", fake_attention_shards)
    if mutate is not None:
        for _, _, pg in ranks:
            mutate(pg)
    # each call starts with the "QKV GEMM" overwriting the local K|V buffer with (rank, call); ranks are interleaved
    # arbitrarily on the host, calls stay in order
    order = [(c, r) for c in range(n_calls) for r in range(sp)]
    random.Random(seed).shuffle(order)
    order.sort(key=lambda cr: cr[0])
    for call, r in order:
        ra, ws, _ = ranks[r]
        current.update(ra=ra, call=call)

        def produce(kv=ws["kv"], r=r, call=call):
            kv.view(-1)[0], kv.view(-1)[1] = float(r), float(call)

        sim.enqueue(ra.compute_stream.cuda_stream, ("call", produce))
        (attention_dma or parallel.RingAttention._attention_dma)(ra, ws, 0)
    sim.run(random.Random(1000 + seed))
    return sim.log


def reads_are_correct(log, sp, n_calls=5):
    return len(log) == sp * n_calls * sp and all(s == (r - j) % sp and sc == c for r, c, j, s, sc in log)


@pytest.mark.parametrize("sp", [2, 3, 4])
def test_protocol_is_safe_and_live_under_random_interleavings(sp, monkeypatch):
    for seed in range(12):
        log = simulate(sp, seed, monkeypatch)
        for r, call, j, shard, shard_call in log:
            assert shard == (r - j) % sp and shard_call == call, \
                f"sp={sp} seed={seed}: rank {r} call {call} shard slot {j} read shard {shard} of call {shard_call}"
        assert reads_are_correct(log, sp)


def _strip_flags(pg):
    orig = pg.kernel_shards
    pg.kernel_shards = lambda kv, T: [sh[:3] for sh in orig(kv, T)]


def _never_release(pg):
    pg.release_all = lambda T, stream: None


def _forget_previous_transfer(pg):
    orig = pg.push_all

    def push(src, comm):
        T = orig(src, comm)
        pg.last_sent = 0        # the next push will not wait for the consumption of this one
        return T

    pg.push_all = push


SOURCE_MUTATIONS = {
    "no guard of the local K|V buffer at the end of a call":
        ("    done = torch.cuda.Event()\n    done.record(self.comm_stream)\n    compute.wait_event(done)\n", ""),
    "pushes do not wait for the QKV GEMM": ("    self.comm_stream.wait_event(ready)\n", ""),
}
OBJECT_MUTATIONS = {
    "the kernel does not wait for the arrival flags": _strip_flags,
    "receive buffers are never released": _never_release,
    "a sender does not wait for the consumption of its previous shard": _forget_previous_transfer,
}


@pytest.mark.parametrize("name", sorted(SOURCE_MUTATIONS) + sorted(OBJECT_MUTATIONS))
def test_every_guard_of_the_schedule_is_necessary(name, monkeypatch):
    """Mutation check: removing any single wait / release makes some interleaving read the wrong shard or deadlock — so
    the simulation above really exercises those guards."""
    import inspect
    import textwrap

    attention_dma, mutate = None, None
    if name in SOURCE_MUTATIONS:
        old, new = SOURCE_MUTATIONS[name]
        src = textwrap.dedent(inspect.getsource(parallel.RingAttention._attention_dma))
        assert old in src, "the mutation no longer matches the source; update SOURCE_MUTATIONS"
        ns = {}
        exec(src.replace(old, new), dict(vars(parallel), torch=torch), ns)
        attention_dma = ns["_attention_dma"]
    else:
        mutate = OBJECT_MUTATIONS[name]
    broken = 0
    for sp in (2, 4):
        for seed in range(16):
            try:
                broken += not reads_are_correct(simulate(sp, seed, monkeypatch, attention_dma=attention_dma, mutate=mutate), sp)
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


# ---------------------------------------------------------------------------------------------------------------------
# The per-step OUTPUT exchange (parallel.OutputGather): the same PeerGather protocol over the whole world, consumed by a
# plain kernel behind stream-level waits instead of in-kernel polls.

def build_world(sim, world, cfg):
    lib = FakeLib(sim)
    sp = world // cfg
    rows_local, tok_rows = (1 if cfg == 2 else 2), 2
    shape = (rows_local, tok_rows, 64)
    recv = [[torch.zeros(shape, dtype=torch.bfloat16) for _ in range(world - 1)] for _ in range(world)]
    flag_base = [5000 * (r + 1) for r in range(world)]
    out = []
    for r in range(world):
        for t in recv[r]:
            sim.tensors[t.data_ptr()] = t
        pg = object.__new__(dma_ring.PeerGather)
        pg.lib, pg.device, pg.shape, pg.dtype = lib, "cpu", shape, torch.bfloat16
        pg.nbytes = rows_local * tok_rows * 64 * 2
        pg.n, pg.me = world, r
        pg._flags_ptr = flag_base[r]
        pg._peer_recv = {d: recv[(r + d) % world][d - 1].data_ptr() for d in range(1, world)}
        pg._peer_flags = {d: flag_base[(r + d) % world] for d in range(1, world)}
        pg.recv = recv[r]
        pg.next_id, pg.last_sent = 1, 0
        og = object.__new__(parallel.OutputGather)
        og.layout = parallel.Layout(world, r, cfg, sp)
        og.rows_local, og.tok_rows = rows_local, tok_rows
        og.pg = pg
        og.comm_stream = FakeStream(sim, f"ocomm{r}")
        og._sent = None
        og.compute_stream = FakeStream(sim, f"ocompute{r}")
        local = torch.zeros(shape, dtype=torch.bfloat16)
        sim.tensors[local.data_ptr()] = local
        out.append((og, local))
    return out


def simulate_output_exchange(world, cfg, seed, monkeypatch, gather=None, mutate=None, n_steps=4):
    from landiff_b200 import ops

    sim = Sim()
    FakeEvent.sim = sim
    ranks = build_world(sim, world, cfg)
    current = {}
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: current["og"].compute_stream)
    monkeypatch.setattr(dma_ring, "check", lambda rc, what: None)

    def owner_of(addr):
        for base, t in sim.tensors.items():
            if base <= addr < base + t.numel() * t.element_size():
                return t
        raise AssertionError(f"block address {addr} is not inside any simulated buffer")

    def fake_unpatchify_blocks(blocks, out):
        og, step = current["og"], current["step"]
        r = og.layout.rank

        def read(blocks=list(blocks), r=r, step=step):
            for addr, row, g0, count in blocks:
                t = owner_of(addr)
                sim.log.append((r, step, row, g0, int(t.view(-1)[0]), int(t.view(-1)[1])))

        sim.enqueue(og.compute_stream.cuda_stream, ("call", read))

    monkeypatch.setattr(ops, "unpatchify_blocks", fake_unpatchify_blocks)
    if mutate is not None:
        for og, _ in ranks:
            mutate(og)
    order = [(s, r) for s in range(n_steps) for r in range(world)]
    random.Random(seed).shuffle(order)
    order.sort(key=lambda sr: sr[0])
    n_total, text_len = 4 * (world // cfg) * 3, 2      # token counts only feed blocks_of's bookkeeping
    for step, r in order:
        og, local = ranks[r]
        current.update(og=og, step=step)

        def final_gemm(local=local, r=r, step=step):      # the final linear writes this rank's block: tag (rank, step)
            local.view(-1)[0], local.view(-1)[1] = float(r), float(step)

        sim.enqueue(og.compute_stream.cuda_stream, ("call", final_gemm))
        (gather or parallel.OutputGather.gather)(og, local, None, n_total, text_len)
    sim.run(random.Random(2000 + seed))
    return sim.log


def output_reads_are_correct(log, world, cfg, n_steps=4):
    """Every rank, every step: one block per (row, shard) of the output, each carrying the tag of the rank that owns it and
    of THIS step."""
    sp = world // cfg
    rows_local = 1 if cfg == 2 else 2
    if len(log) != world * n_steps * world * rows_local:
        return False
    for r, step, row, g0, src_rank, src_step in log:
        if src_step != step:
            return False
        if cfg == 2 and src_rank // sp != row:
            return False
    # every (row, shard) exactly once per (rank, step)
    seen = {}
    for r, step, row, g0, src_rank, _ in log:
        seen.setdefault((r, step), []).append((row, src_rank % sp))
    return all(sorted(v) == sorted((row, s) for row in range(2) for s in range(sp)) for v in seen.values())


@pytest.mark.parametrize("world,cfg", [(2, 1), (4, 2), (8, 2), (4, 1)])
def test_output_exchange_is_safe_and_live(world, cfg, monkeypatch):
    for seed in range(8):
        log = simulate_output_exchange(world, cfg, seed, monkeypatch)
        assert output_reads_are_correct(log, world, cfg), f"world={world} cfg={cfg} seed={seed}"


def _no_arrival_wait(og):
    og.pg.wait_all = lambda T, stream: None


def _no_release(og):
    og.pg.release_all = lambda T, stream: None


OUTPUT_SOURCE_MUTATIONS = {   # (old, new) on the dedented source of OutputGather.gather
    "the next final GEMM may overwrite the block while it is being pushed": ("    compute.wait_event(self._sent)", "    pass"),
    "the push does not wait for the final GEMM": ("    self.comm_stream.wait_event(ready)\n", ""),
}
OUTPUT_OBJECT_MUTATIONS = {
    "the scatter kernel does not wait for the arrival flags": _no_arrival_wait,
    "receive buffers are never released": _no_release,
}


@pytest.mark.parametrize("name", sorted(OUTPUT_SOURCE_MUTATIONS) + sorted(OUTPUT_OBJECT_MUTATIONS))
def test_every_guard_of_the_output_exchange_is_necessary(name, monkeypatch):
    import inspect
    import textwrap

    gather, mutate = None, None
    if name in OUTPUT_SOURCE_MUTATIONS:
        old, new = OUTPUT_SOURCE_MUTATIONS[name]
        src = textwrap.dedent(inspect.getsource(parallel.OutputGather.gather))
        assert old in src, "the mutation no longer matches the source; update OUTPUT_SOURCE_MUTATIONS"
        ns = {}
        exec(src.replace(old, new), dict(vars(parallel), torch=torch), ns)
        gather = ns["gather"]
    else:
        mutate = OUTPUT_OBJECT_MUTATIONS[name]
    broken = 0
    for world, cfg in ((2, 1), (4, 2)):
        for seed in range(16):
            try:
                broken += not output_reads_are_correct(
                    simulate_output_exchange(world, cfg, seed, monkeypatch, gather=gather, mutate=mutate), world, cfg)
            except AssertionError:      # deadlock
                broken += 1
    assert broken > 0, name
