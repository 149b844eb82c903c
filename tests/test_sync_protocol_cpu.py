"""Exhaustive interleaving check of the two synchronisation protocols of the fused TV-L1 kernel
(denseflow_b200/csrc/tvl1_fused.cu), as small state machines explored over EVERY schedule:

* lane barrier + convergence partials (`grid_barrier`, `job.partials`): every CTA of a lane must sum the partials of
  the SAME check, otherwise the CTAs disagree on convergence, run different numbers of barriers and the lane never
  completes.  One set of partials is not enough (a CTA one barrier ahead overwrites its slot while a slower CTA is still
  summing — the hang this round found on the GPU); two sets, alternating per check, are.
* neighbour-warp progress counters of the tile loop (`wait_ge` / `signal`, order P0 | P1 D0 | P2 D1 | P3 D2 | D3): a warp
  reads the dual row of the warp above and the first primal row of the warp below; both must be the version of the
  iteration it is in, under every interleaving the waits allow.

This models the protocol, not the arithmetic: shared locations carry version numbers, a read states which version it
needs.  It is a design regression test (no GPU, no product code involved)."""
from collections import deque


def explore(programs, init_mem):
    """programs[t] = list of steps; a step is ('write', loc, val) | ('read', loc, want) | ('add', loc, k) |
    ('wait_ge', loc, val).  Breadth-first over all interleavings; returns (violations, deadlocked, states)."""
    locs = sorted(init_mem, key=repr)
    idx = {l: i for i, l in enumerate(locs)}
    start = (tuple(0 for _ in programs), tuple(init_mem[l] for l in locs))
    seen, todo = {start}, deque([start])
    violations, deadlocks = [], 0
    while todo:
        pcs, mem = todo.popleft()
        moved = False
        done = True
        for t, prog in enumerate(programs):
            if pcs[t] == len(prog):
                continue
            done = False
            op, loc, val = prog[pcs[t]]
            m = list(mem)
            if op == "wait_ge":
                if mem[idx[loc]] < val:
                    continue
            elif op == "write":
                m[idx[loc]] = val
            elif op == "add":
                m[idx[loc]] += val
            elif op == "read":
                if mem[idx[loc]] != val:
                    violations.append((t, pcs[t], loc, mem[idx[loc]], val))
            moved = True
            nxt = (pcs[:t] + (pcs[t] + 1,) + pcs[t + 1:], tuple(m))
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
        if not moved and not done:
            deadlocks += 1
    return violations, deadlocks, len(seen)


def lane_programs(G, checks, nsets):
    progs = []
    for b in range(G):
        p = []
        for c in range(checks):
            s = c % nsets
            p.append(("write", ("partial", s, b), c))   # job.partials[part_sel * kPartialSet + bid] = bs
            p.append(("add", "counter", 1))              # atomicAdd(counter, 1)
            p.append(("wait_ge", "counter", (c + 1) * G))  # spin until the arrival count reaches the epoch
            for i in range(G):                            # every CTA sums all partials in the same order
                p.append(("read", ("partial", s, i), c))
        progs.append(p)
    mem = {"counter": 0}
    for s in range(nsets):
        for b in range(G):
            mem[("partial", s, b)] = -1
    return progs, mem


def test_single_buffered_partials_can_be_overwritten_before_they_are_summed():
    v, dead, _ = explore(*lane_programs(G=2, checks=3, nsets=1))
    assert v, "the checker must see the race the GPU hit"
    assert dead == 0


def test_double_buffered_partials_are_safe_under_every_interleaving():
    for G, checks in ((2, 5), (3, 4)):
        v, dead, states = explore(*lane_programs(G, checks, nsets=2))
        assert not v and dead == 0, (G, checks, v[:3])
        assert states > 100  # the search really branched


def warp_programs(W, iters, guard_p0=True):
    """Progress counter of warp w: 2*it+1 after the first primal row of iteration `it`, 2*it+2 after its last dual row."""
    progs = []
    for w in range(W):
        p = []
        for it in range(iters):
            if w > 0:
                if guard_p0:
                    p.append(("wait_ge", ("prog", w - 1), 2 * it))     # D3 of the previous iteration of the warp above
                p.append(("read", ("p_bot", w - 1), it))               # its last dual row, as left by iteration it-1
            p.append(("write", ("u_row0", w), it + 1))                  # P0: new u of this warp's first row
            p.append(("write", ("prog", w), 2 * it + 1))                # signal
            if w < W - 1:
                p.append(("wait_ge", ("prog", w + 1), 2 * it + 1))      # P0 of the warp below, this iteration
                p.append(("read", ("u_row0", w + 1), it + 1))           # D3 needs the NEW u of the row below
            p.append(("write", ("p_bot", w), it + 1))                   # D3: this warp's last dual row
            p.append(("write", ("prog", w), 2 * it + 2))                # signal
        progs.append(p)
    mem = {}
    for w in range(W):
        mem[("prog", w)] = 0      # tile loaded
        mem[("u_row0", w)] = 0    # version 0 = the loaded tile
        mem[("p_bot", w)] = 0
    return progs, mem


def test_neighbour_warp_flags_order_every_cross_warp_read():
    for W, iters in ((2, 4), (3, 3), (4, 2)):
        v, dead, states = explore(*warp_programs(W, iters))
        assert not v and dead == 0, (W, iters, v[:3])
        assert states > 100


def test_the_checker_sees_a_missing_wait():
    v, _, _ = explore(*warp_programs(3, 3, guard_p0=False))
    assert v


def chunk_programs(G, chunks, buffers):
    """Tile chunks of a level: a CTA reads its neighbours' flow (halo) from buffer `cur`, writes its own tile to the other
    buffer, arrives at the lane barrier, flips `cur`.  Version c = written in chunk c-1 (0: level start)."""
    progs = []
    for b in range(G):
        p = []
        for c in range(chunks):
            cur, nxt = c % buffers, (c + 1) % buffers
            for nb in (b - 1, b + 1):
                if 0 <= nb < G:
                    p.append(("read", ("u", cur, nb), c))
            p.append(("write", ("u", nxt, b), c + 1))
            p.append(("add", "counter", 1))
            p.append(("wait_ge", "counter", (c + 1) * G))
        progs.append(p)
    mem = {"counter": 0}
    for s in range(buffers):
        for b in range(G):
            mem[("u", s, b)] = 0
    return progs, mem


def test_ping_pong_flow_planes_need_exactly_two_buffers():
    v, dead, _ = explore(*chunk_programs(G=3, chunks=4, buffers=2))
    assert not v and dead == 0, v[:3]
    v, _, _ = explore(*chunk_programs(G=3, chunks=4, buffers=1))
    assert v  # in place: a neighbour's halo is overwritten before it is read
