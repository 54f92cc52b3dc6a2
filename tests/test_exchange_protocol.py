"""CPU: the ordering protocol of dissect_b200/exchange.py (SM-free packed-chunk exchange) under a stream simulator.

PeerExchange only ever enqueues four kinds of stream operations: wait-until-word >= value, copy (data or a 4-byte
sequence number), and — through the engine — "the row gather reads gather buffer k" bracketed by events.  Here the
CUDA calls are replaced by recorders, every rank's streams become FIFO queues, and a scheduler executes runnable queue
heads in random order.  Checked for many interleavings: no deadlock; every gather of chunk s sees exactly the shards of
chunk s from every rank (nothing stale, nothing overwritten early); buffers are reused across many chunks."""
import random

import pytest

from dissect_b200.exchange import NBUF, PeerExchange

SEQ_BASE = 0x7000_0000          # fake address of seqsrc: value s lives at SEQ_BASE + 4 s
FLAG_BASE = 0x1000_0000         # rank p's flag words at FLAG_BASE + p * 0x1000
BUF_BASE = 0x2000_0000          # rank p's buffer k at BUF_BASE + p * 0x100_0000 + k * 0x10_0000
SRC_BASE = 0x5000_0000


class _Ptr:
    def __init__(self, addr):
        self.addr = addr

    def data_ptr(self):
        return self.addr


class _Stream:
    def __init__(self, name):
        self.name = name
        self.cuda_stream = name


class _Engine:
    """What dgrm_push_bed_device / dgrm_release_inputs do to stream order (dgrm_api.cu): the decode stream waits for
    an event on the producer stream, gathers from the buffer, and release_inputs makes the consumer stream wait for
    the decode stream."""

    def __init__(self, sim, rank):
        self.sim, self.rank, self.nev = sim, rank, 0

    def push_bed_device(self, ptr, pitch, nrows, row_keep=None, stream=None):
        ev = (self.rank, "push", self.nev)
        self.nev += 1
        self.sim.enqueue(self.rank, stream, ("record", ev))
        self.sim.enqueue(self.rank, "prep", ("wait_event", ev))
        self.sim.enqueue(self.rank, "prep", ("gather", ptr, row_keep))

    def release_inputs(self, stream):
        ev = (self.rank, "rel", self.nev)
        self.nev += 1
        self.sim.enqueue(self.rank, "prep", ("record", ev))
        self.sim.enqueue(self.rank, stream, ("wait_event", ev))


class _SimExchange(PeerExchange):
    def __init__(self, sim, rank, world, shard, pitch):   # no CUDA: fill in what __init__ would have mapped
        self.sim = sim
        self.rank, self.world, self.shard, self.pitch = rank, world, shard, pitch
        self.flags = _Ptr(FLAG_BASE + rank * 0x1000)
        self.seqsrc = _Ptr(SEQ_BASE)
        self.bufs = [_Ptr(BUF_BASE + rank * 0x100_0000 + k * 0x10_0000) for k in range(NBUF)]
        self.peer_buf = [[BUF_BASE + p * 0x100_0000 + k * 0x10_0000 for k in range(NBUF)] for p in range(world)]
        self.peer_flag = [FLAG_BASE + p * 0x1000 for p in range(world)]
        self.copy_stream, self.arrive_stream, self.rel_stream = _Stream("copy"), _Stream("arrive"), _Stream("rel")
        self.seq_post = self.seq_cons = 0
        self._trace = None

    def _wait_geq(self, stream, addr, value):
        self.sim.enqueue(self.rank, stream.cuda_stream, ("wait_mem", addr, value))

    def _copy(self, stream, dst, src, nbytes):
        self.sim.enqueue(self.rank, stream.cuda_stream, ("copy", dst, src, nbytes))


class _Sim:
    def __init__(self, world, shard, pitch, seed):
        self.world, self.shard, self.pitch = world, shard, pitch
        self.queues = {}            # (rank, stream) -> list of ops
        self.mem = {}               # flag word address -> value
        self.events = set()
        self.slots = {}             # (rank, k, source rank) -> chunk sequence number held
        self.gathered = {r: [] for r in range(world)}
        self.rng = random.Random(seed)

    def enqueue(self, rank, stream, op):
        self.queues.setdefault((rank, stream), []).append(op)

    def _runnable(self, op):
        if op[0] == "wait_mem":
            return self.mem.get(op[1], 0) - op[2] >= 0
        if op[0] == "wait_event":
            return op[1] in self.events
        return True

    def _execute(self, rank, op):
        kind = op[0]
        if kind == "record":
            self.events.add(op[1])
        elif kind == "copy":
            _, dst, src, nbytes = op
            if nbytes == 4:                                  # a sequence number into somebody's flag word
                assert SEQ_BASE <= src
                self.mem[dst] = (src - SEQ_BASE) // 4
            else:                                            # a shard into slot `rank` of somebody's buffer
                owner = (dst - BUF_BASE) // 0x100_0000
                k = ((dst - BUF_BASE) % 0x100_0000) // 0x10_0000
                off = (dst - BUF_BASE) % 0x10_0000
                assert off == rank * self.shard * self.pitch and nbytes <= self.shard * self.pitch
                self.slots[(owner, k, rank)] = (src - SRC_BASE) // 0x1000      # chunk tag carried by the source address
        elif kind == "gather":
            _, ptr, want = op
            k = ((ptr - BUF_BASE) % 0x100_0000) // 0x10_0000
            got = [self.slots.get((rank, k, q)) for q in range(self.world)]
            assert got == [want] * self.world, f"rank {rank} gathered chunk {want} but buffer {k} holds {got}"
            self.gathered[rank].append(want)

    def run(self):
        while True:
            heads = [(key, q[0]) for key, q in self.queues.items() if q]
            if not heads:
                return
            ready = [(key, op) for key, op in heads if self._runnable(op)]
            assert ready, f"deadlock: {[(k, op[:3]) for k, op in heads]}"
            key, op = self.rng.choice(ready)
            self.queues[key].pop(0)
            self._execute(key[0], op)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("seed", range(6))
def test_exchange_protocol_is_safe_and_live(world, seed):
    shard, pitch, nchunks = 4, 16, 11                        # 11 chunks over 3 buffers: every buffer is reused 3 times
    sim = _Sim(world, shard, pitch, seed)
    xs = [_SimExchange(sim, r, world, shard, pitch) for r in range(world)]
    engines = [_Engine(sim, r) for r in range(world)]
    rng = random.Random(1000 + seed)
    # hosts run at different speeds: interleave the ranks' host programs arbitrarily (each posts up to NBUF ahead, as
    # bench.py does), then let the simulated GPUs drain their queues in random order
    progress = [[0, 0] for _ in range(world)]                # [posted, consumed] per rank
    while any(c < nchunks for _, c in progress):
        r = rng.randrange(world)
        posted, consumed = progress[r]
        if consumed >= nchunks:
            continue
        if posted < min(nchunks, consumed + NBUF):
            s = posted + 1
            xs[r].post(SRC_BASE + s * 0x1000, shard)         # the source address tags the chunk
            progress[r][0] += 1
        else:
            xs[r].consume(engines[r], consumed + 1)          # row_keep doubles as "which chunk do I expect"
            progress[r][1] += 1
        if rng.random() < 0.3:
            # GPUs make progress while the hosts are still enqueuing (partial drains must not trip the checks)
            for _ in range(rng.randrange(20)):
                ready = [(k, q[0]) for k, q in sim.queues.items() if q and sim._runnable(q[0])]
                if not ready:
                    break
                key, op = sim.rng.choice(ready)
                sim.queues[key].pop(0)
                sim._execute(key[0], op)
    sim.run()
    for r in range(world):
        assert sim.gathered[r] == list(range(1, nchunks + 1))
