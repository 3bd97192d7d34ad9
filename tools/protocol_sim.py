"""Functional simulation of the tcgen05 kernel's mbarrier protocol (nm_mlp_tc.cu) on the real layer program.

Agents (producer, 4 MMA issuers, 2 epilogue sets, front-end) are generators that yield when they would block; a
round-robin scheduler runs them until everyone finishes (ok) or nobody can move (deadlock -> prints who waits on what).
Timing is not modelled, only ordering / phase correctness.  Commits are modelled as completing immediately.

    NM_TC_POLICY=1 python tools/protocol_sim.py [--tiles 3] [--stages 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: too many arrivals in phase {self.phase}"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, k):            # hardware semantics: try_wait.parity(k & 1) -- ONE parity bit, so a waiter that is two
        return (self.phase & 1) != (k & 1)   # phases away from the barrier's current phase aliases (this is what is modelled)


def simulate(prog, tiles, NS, verbose=False, armed_counter=True, fe_emit=False):
    L = [prog.layers[i] for i in range(prog.n_layers)]
    w_full = [Bar(f"w_full{i}", 1) for i in range(NS)]
    w_empty = [Bar(f"w_empty{i}", 1) for i in range(NS)]
    pe_full = [Bar(f"pe_full{i}", 1) for i in range(2)]
    pe_empty = [Bar(f"pe_empty{i}", 4) for i in range(2)]
    chunk = [Bar(f"chunk_ready{i}", 1) for i in range(4)]       # one arrival per epilogue set (4 warps act together)
    d_full = [Bar(f"d_full{i}", 4) for i in range(4)]
    kb_free = [Bar(f"kb_free{i}", 4) for i in range(4)]
    dir_full, dir_empty = Bar("dir_full", 1), Bar("dir_empty", 4)
    emit_done = [Bar(f"emit_done{i}", 1) for i in range(2)]      # fe_emit (mode 1): the front end answers chunk_ready[2..3] per layer
    uses_dir = any(L[i].pe_src == 2 for i in range(len(L)))
    waiting = {}
    armed = [0]
    errors = []

    def wait(me, bar, k):
        while not bar.done(k):
            waiting[me] = f"{bar.name} completion #{k} (phase now {bar.phase}, pending {bar.pending})"
            yield
        waiting.pop(me, None)

    def producer():
        g = 0
        for t in range(tiles):
            for b in range(prog.n_blocks):
                slot, rnd = g % NS, g // NS
                if rnd > 0:
                    yield from wait("producer", w_empty[slot], rnd - 1)
                w_full[slot].arrive()
                g += 1
                armed[0] = g

    def fe_emit_tile(T):
        for li in range(len(L)):
            gl = T * len(L) + li
            for n in (2, 3):
                yield from wait("frontend", chunk[n], gl)
                emit_done[n - 2].arrive()

    def frontend():
        for t in range(tiles):
            buf = t & 1
            if t >= 2:
                yield from wait("frontend", pe_empty[buf], t // 2 - 1)
            pe_full[buf].arrive()
            if fe_emit and t > 0:
                yield from fe_emit_tile(t - 1)
            if uses_dir:
                if t >= 1:
                    yield from wait("frontend", dir_empty, t - 1)
                dir_full.arrive()
        if fe_emit:
            yield from fe_emit_tile(tiles - 1)

    def issuer(w):
        me = f"issuer{w}"
        g = 0
        gl = 0
        for t in range(tiles):
            buf = t & 1
            yield from wait(me, pe_full[buf], t // 2)
            for li, Lp in enumerate(L):
                waited = -1

                def pass_group(gr):
                    nonlocal waited
                    while waited < gr:
                        waited += 1
                        if gl > 0:
                            yield from wait(me, chunk[waited], gl - 1)
                        if (Lp.none_d >> (4 * w + waited)) & 1:
                            d_full[waited].arrive()
                        if (Lp.none_k >> (4 * w + waited)) & 1:
                            kb_free[waited].arrive()
                for b in range(Lp.blk_begin, Lp.blk_end):
                    B = prog.blocks[b]
                    slot, rnd = g % NS, g // NS
                    if (B.flags >> 4) == w:
                        yield from pass_group(B.group)
                        if B.src == 2:
                            yield from wait(me, dir_full, t)
                        while armed_counter and armed[0] <= g:
                            waiting[me] = f"armed counter > {g}"
                            yield
                        yield from wait(me, w_full[slot], rnd)
                        if w_full[slot].phase != rnd + 1:
                            errors.append(f"{me}: consumed slot {slot} for block {g} (round {rnd}) while the barrier had completed {w_full[slot].phase} rounds")
                        w_empty[slot].arrive()
                        if B.flags & 1:
                            d_full[B.nc].arrive()
                        if B.flags & 2:
                            kb_free[B.kb].arrive()
                    g += 1
                yield from pass_group(3)
                gl += 1
            pe_empty[buf].arrive()
            dir_empty.arrive()

    def epilogue(s):
        me = f"epi_set{s}"
        gl = 0
        for t in range(tiles):
            for li, Lp in enumerate(L):
                for nn in range(2):
                    n = s + 2 * nn
                    yield from wait(me, d_full[n], gl)
                    yield from wait(me, kb_free[n], gl)
                    if fe_emit and nn == 1 and gl > 0:
                        yield from wait(me, emit_done[n - 2], gl - 1)
                    chunk[n].arrive()
                gl += 1

    agents = {"producer": producer(), "frontend": frontend(), **{f"issuer{w}": issuer(w) for w in range(4)},
              **{f"epi_set{s}": epilogue(s) for s in range(2)}}
    alive = dict(agents)
    steps = 0
    while alive:
        progressed = False
        for name in list(alive):
            before = (tuple(b.phase for b in w_full + w_empty + pe_full + pe_empty + chunk + d_full + kb_free + emit_done + [dir_full, dir_empty]),
                      tuple(b.pending for b in w_full + w_empty + pe_full + pe_empty + chunk + d_full + kb_free + emit_done + [dir_full, dir_empty]))
            try:
                next(alive[name])
            except StopIteration:
                del alive[name]
                progressed = True
                continue
            after = (tuple(b.phase for b in w_full + w_empty + pe_full + pe_empty + chunk + d_full + kb_free + emit_done + [dir_full, dir_empty]),
                     tuple(b.pending for b in w_full + w_empty + pe_full + pe_empty + chunk + d_full + kb_free + emit_done + [dir_full, dir_empty]))
            progressed |= before != after
        steps += 1
        if not progressed:
            return False, dict(waiting, errors=errors[:3])
    return (not errors), dict(errors=errors[:3])


def simulate_pair(prog, tiles, NS, ghost=False):
    """Two CTAs of a cluster sharing ONE weight stream (nm_mlp_tc.cu, P.cluster == 2): rank 0's producer multicasts every
    stage into both rings; each CTA's producer arms its own w_full (modelled as a second arrival: arm + copy completion); a
    stage is released into BOTH CTAs' w_empty barriers (count 2) by the consuming issuer of each CTA.  ghost=True: rank 1 has
    one tile less and its issuers run a ghost iteration for the last round (wait for the stage, release it).  Only the ring
    protocol and the tile-level barriers that gate it are modelled per CTA; returns (ok, info)."""
    L = [prog.layers[i] for i in range(prog.n_layers)]
    uses_dir = any(L[i].pe_src == 2 for i in range(len(L)))
    waiting, errors = {}, []

    class Cta:
        def __init__(self, r):
            self.r = r
            self.w_full = [Bar(f"c{r}.w_full{i}", 2) for i in range(NS)]
            self.w_empty = [Bar(f"c{r}.w_empty{i}", 2) for i in range(NS)]
            self.pe_full = [Bar(f"c{r}.pe_full{i}", 1) for i in range(2)]
            self.pe_empty = [Bar(f"c{r}.pe_empty{i}", 4) for i in range(2)]
            self.chunk = [Bar(f"c{r}.chunk{i}", 1) for i in range(4)]
            self.d_full = [Bar(f"c{r}.d_full{i}", 4) for i in range(4)]
            self.kb_free = [Bar(f"c{r}.kb_free{i}", 4) for i in range(4)]
            self.dir_full, self.dir_empty = Bar(f"c{r}.dir_full", 1), Bar(f"c{r}.dir_empty", 4)
            self.armed = 0
            self.real = tiles - (1 if (ghost and r == 1) else 0)

        def bars(self):
            return self.w_full + self.w_empty + self.pe_full + self.pe_empty + self.chunk + self.d_full + self.kb_free + [self.dir_full, self.dir_empty]
    C = [Cta(0), Cta(1)]

    def wait(me, bar, k):
        while not bar.done(k):
            waiting[me] = f"{bar.name} completion #{k} (phase now {bar.phase}, pending {bar.pending})"
            yield
        waiting.pop(me, None)

    def producer(c):
        me = f"c{c.r}.producer"
        g = 0
        for t in range(tiles):                       # both ranks: the SAME number of rounds (iter_exists)
            for b in range(prog.n_blocks):
                slot, rnd = g % NS, g // NS
                if rnd > 0:
                    yield from wait(me, c.w_empty[slot], rnd - 1)
                c.w_full[slot].arrive()              # arm (expect_tx)
                if c.r == 0:                         # the multicast copy lands in both rings
                    C[0].w_full[slot].arrive()
                    C[1].w_full[slot].arrive()
                g += 1
                c.armed = g
                yield

    def frontend(c):
        me = f"c{c.r}.frontend"
        for t in range(c.real):
            buf = t & 1
            if t >= 2:
                yield from wait(me, c.pe_empty[buf], t // 2 - 1)
            c.pe_full[buf].arrive()
            if uses_dir:
                if t >= 1:
                    yield from wait(me, c.dir_empty, t - 1)
                c.dir_full.arrive()

    def issuer(c, w):
        me = f"c{c.r}.issuer{w}"
        g = gl = 0
        for t in range(tiles):
            is_ghost = t >= c.real
            buf = t & 1
            if not is_ghost:
                yield from wait(me, c.pe_full[buf], t // 2)
            for li, Lp in enumerate(L):
                waited = -1

                def pass_group(gr):
                    nonlocal waited
                    while waited < gr:
                        waited += 1
                        if gl > 0:
                            yield from wait(me, c.chunk[waited], gl - 1)
                        if (Lp.none_d >> (4 * w + waited)) & 1:
                            c.d_full[waited].arrive()
                        if (Lp.none_k >> (4 * w + waited)) & 1:
                            c.kb_free[waited].arrive()
                for b in range(Lp.blk_begin, Lp.blk_end):
                    B = prog.blocks[b]
                    slot, rnd = g % NS, g // NS
                    if (B.flags >> 4) == w:
                        if not is_ghost:
                            yield from pass_group(B.group)
                            if B.src == 2:
                                yield from wait(me, c.dir_full, t)
                        while c.armed <= g:
                            waiting[me] = f"armed counter > {g}"
                            yield
                        yield from wait(me, c.w_full[slot], rnd)
                        if c.w_full[slot].phase != rnd + 1:
                            errors.append(f"{me}: consumed slot {slot} for block {g} (round {rnd}) while the barrier had completed {c.w_full[slot].phase} rounds")
                        C[0].w_empty[slot].arrive()      # multicast commit: both CTAs' barriers
                        C[1].w_empty[slot].arrive()
                        if not is_ghost:
                            if B.flags & 1:
                                c.d_full[B.nc].arrive()
                            if B.flags & 2:
                                c.kb_free[B.kb].arrive()
                    g += 1
                if not is_ghost:
                    yield from pass_group(3)
                    gl += 1
            if not is_ghost:
                c.pe_empty[buf].arrive()
                c.dir_empty.arrive()

    def epilogue(c, s_):
        me = f"c{c.r}.epi_set{s_}"
        gl = 0
        for t in range(c.real):
            for li, Lp in enumerate(L):
                for nn in range(2):
                    n = s_ + 2 * nn
                    yield from wait(me, c.d_full[n], gl)
                    yield from wait(me, c.kb_free[n], gl)
                    c.chunk[n].arrive()
                gl += 1

    agents = {}
    for c in C:
        agents[f"c{c.r}.producer"] = producer(c)
        agents[f"c{c.r}.frontend"] = frontend(c)
        for w in range(4):
            agents[f"c{c.r}.issuer{w}"] = issuer(c, w)
        for s_ in range(2):
            agents[f"c{c.r}.epi{s_}"] = epilogue(c, s_)
    alive = dict(agents)
    allbars = C[0].bars() + C[1].bars()
    while alive:
        progressed = False
        for name in list(alive):
            before = (tuple(b.phase for b in allbars), tuple(b.pending for b in allbars), C[0].armed, C[1].armed)
            try:
                next(alive[name])
            except StopIteration:
                del alive[name]
                progressed = True
                continue
            progressed |= before != (tuple(b.phase for b in allbars), tuple(b.pending for b in allbars), C[0].armed, C[1].armed)
        if not progressed:
            return False, dict(waiting, errors=errors[:3])
    return (not errors), dict(errors=errors[:3])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=3)
    ap.add_argument("--stages", type=int, default=5)
    a = ap.parse_args()
    from oracle import nerf_oracle as O
    from test_host_logic import debug_pack
    for arch in (dict(), dict(num_layers=4, hidden_size=128, num_encoding_fn_xyz=6), dict(num_layers=3, hidden_size=128, use_viewdirs=False)):
        cfg = O.NetCfg(**{**O.NetCfg().__dict__, **arch})
        for sigma_only in (False, True):
            prog, _ = debug_pack(cfg, O.init_weights(cfg, 1), sigma_only)
            for ns in sorted({2, 3, a.stages}):
                ok, who = simulate(prog, a.tiles, ns)
                ok0, who0 = simulate(prog, a.tiles, ns, armed_counter=False)
                print(f"arch={arch} sigma_only={sigma_only} stages={ns}: {'ok' if ok else 'FAIL ' + str(who)}"
                      f"   [without the armed counter: {'ok' if ok0 else 'FAIL ' + str(who0)[:160]}]")
