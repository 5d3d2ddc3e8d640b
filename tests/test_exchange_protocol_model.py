"""A model of the persistent kernel's cross-GPU exchange protocol (tfdiffeq_b200/csrc/b2ode_fused.cu: control_allreduce,
remote_gather, Mailbox::fused_part / fused_hw), run under random schedules on the CPU.

What is modelled: R ranks x G_r blocks; every block, per exchange `seq`, stores a 16-byte partial tagged `seq % 15 + 1` into slot
[seq & 1][rank][block] of EVERY peer's mailbox and into its own GPU's slot array, bumps a local arrival counter (relaxed: not
ordered with the partial), waits for the counter, reads the local slots (re-reading stale tags) and polls the peers' slots
until every tag matches.  Stores become visible after arbitrary delays; only stores of one thread to one address stay
ordered, and a kernel boundary flushes a rank's stores.  Ranks run ahead of each other as far as the protocol lets them; a
group is reused for solves whose grids shrink and grow; a solve has 0..5 exchanges.  A slot that holds no partial holds
poison, and a solve with a smaller grid than the previous writer of a buffer poisons what it no longer writes before its
first exchange on that buffer.

What is checked: a reader never accepts a partial of another exchange (same tag, different sequence number), nothing
dead-locks, and every rank completes every solve.  The negative controls show that the model sees the failures the protocol
is built against: without poison a regrown grid accepts a stale partial; with poison written to BOTH buffers at the start
of a solve a slow peer loses a partial it has not read yet.
"""
import random

import pytest

POISON = (None, None)


def tag_of(seq):
    return seq % 15 + 1          # b2ode_pay16.cuh: never 0, so cleared memory never validates


class Violation(Exception):
    pass


class Model:
    def __init__(self, ranks, solves, rng, poison=True, poison_both_at_start=False, max_blocks=6):
        self.R, self.rng, self.poison, self.both = ranks, rng, poison, poison_both_at_start
        self.solves = solves                      # list of (grid per rank, number of exchanges)
        self.S = max_blocks
        init = POISON if poison else (0, -1)      # zero-filled memory reads as tag 0
        # mailbox[d][par][src][slot] = (tag, true sequence number)
        self.mail = [[[[init] * self.S for _ in range(ranks)] for _ in range(2)] for _ in range(ranks)]
        self.local = [[[(0, -1)] * self.S for _ in range(2)] for _ in range(ranks)]
        self.ctr = [0] * ranks
        self.hw = [[0, 0] for _ in range(ranks)]
        self.ll_seq = [0] * ranks
        self.solve_of = [0] * ranks               # index of the solve a rank is in
        self.pending = []                         # (rank of the writer, kind, address, value): in flight
        self.blocks = {}                          # (rank, block) -> state dict
        self.done_blocks = [0] * ranks
        for r in range(ranks):
            self._launch(r)

    # ---- stores ------------------------------------------------------------------------------------------------
    def _store(self, writer, kind, addr, val):
        self.pending.append((writer, kind, addr, val))

    def _deliver(self, i):
        writer, kind, addr, val = self.pending.pop(i)
        if kind == "mail":
            d, par, src, slot = addr
            self.mail[d][par][src][slot] = val
        elif kind == "local":
            r, par, slot = addr
            self.local[r][par][slot] = val
        else:
            self.ctr[addr] += 1

    def _deliverable(self):
        seen, out = set(), []
        for i, (w, kind, addr, _) in enumerate(self.pending):      # per (writer thread, address): program order
            key = (w, kind, addr)
            if key not in seen:
                seen.add(key)
                out.append(i)
        return out

    def _flush(self, rank):
        while True:
            idx = [i for i in self._deliverable() if self.pending[i][0][0] == rank]
            if not idx:
                return
            self._deliver(idx[0])

    # ---- kernels -----------------------------------------------------------------------------------------------
    def _launch(self, r):
        k = self.solve_of[r]
        if k >= len(self.solves):
            return
        grids, n_ex = self.solves[k]
        self.ctr[r] = 0                                                       # workspace cleared by the host before the launch:
        self.local[r] = [[(0, -1)] * self.S for _ in range(2)]                # arrival counter and the intra-GPU slot array
        self.done_blocks[r] = 0
        for b in range(grids[r]):
            self.blocks[(r, b)] = dict(e=0, phase="send", base=self.ll_seq[r], hw=tuple(self.hw[r]), n_ex=n_ex, G=grids[r])

    def _finish_block(self, r, b):
        st = self.blocks.pop((r, b))
        self.done_blocks[r] += 1
        if b == 0:                                                            # block 0 records what the solve left behind
            self.ll_seq[r] = st["base"] + st["n_ex"]
            if st["n_ex"] >= 1:
                self.hw[r][(st["base"] + 1) & 1] = st["G"]
            if st["n_ex"] >= 2:
                self.hw[r][(st["base"] + 2) & 1] = st["G"]
        if self.done_blocks[r] == st["G"]:                                    # kernel boundary: its stores are performed
            self._flush(r)
            self.solve_of[r] += 1
            self._launch(r)

    def _step_block(self, r, b):
        """One atomic action of block (r, b); returns False if it could not move."""
        st = self.blocks[(r, b)]
        grids = self.solves[self.solve_of[r]][0]
        if st["phase"] == "send":
            if st["e"] == st["n_ex"]:
                self._finish_block(r, b)
                return True
            st["e"] += 1
            seq = st["base"] + st["e"]
            par, tag, w = seq & 1, tag_of(seq), (r, b)
            if self.poison:
                pars = [par] if not self.both else ([0, 1] if st["e"] == 1 else [])
                if st["e"] <= 2 or self.both:
                    for pp in pars:
                        for slot in range(st["G"] + b, st["hw"][pp], st["G"]):
                            for d in range(self.R):
                                if d != r:
                                    self._store(w, "mail", (d, pp, r, slot), POISON)
            for d in range(self.R):
                if d != r:
                    self._store(w, "mail", (d, par, r, b), (tag, seq))
            self._store(w, "local", (r, par, b), (tag, seq))
            self._store(w, "ctr", r, 1)
            st["phase"] = "wait"
            return True
        seq = st["base"] + st["e"]
        par, tag = seq & 1, tag_of(seq)
        if self.ctr[r] < st["e"] * st["G"]:
            return False
        for slot in range(st["G"]):                                           # local partials: stale tags are re-read
            t, s = self.local[r][par][slot]
            if t != tag:
                return False
            if s != seq:
                raise Violation("rank %d block %d took local partial of exchange %d for %d" % (r, b, s, seq))
        for src in range(self.R):                                             # the comm warp's gather
            if src == r:
                continue
            for slot in range(grids[src]):
                t, s = self.mail[r][par][src][slot]
                if t != tag:
                    return False
                if s != seq:
                    raise Violation("rank %d block %d took rank %d's partial of exchange %s for %d (slot %d)" % (r, b, src, s, seq, slot))
        st["phase"] = "send"
        return True

    def run(self, max_steps=200000):
        for _ in range(max_steps):
            if not self.blocks and not self.pending:
                return
            movable = list(self.blocks.keys())
            self.rng.shuffle(movable)
            deliver = self._deliverable()
            if deliver and (not movable or self.rng.random() < 0.45):
                self._deliver(self.rng.choice(deliver))
                continue
            for key in movable:
                if self._step_block(*key):
                    break
            else:
                if not deliver:
                    raise Violation("dead-lock: %d blocks wait and nothing is in flight" % len(self.blocks))
                self._deliver(self.rng.choice(deliver))
        raise Violation("no termination within the step budget")


def _random_solves(rng, ranks, n, max_blocks):
    out = []
    for _ in range(n):
        grids = [rng.randint(1, max_blocks) for _ in range(ranks)]
        out.append((grids, rng.randint(0, 5)))
    return out


@pytest.mark.parametrize("ranks", [2, 3, 4])
def test_protocol_holds_under_random_schedules(ranks):
    for seed in range(60):
        rng = random.Random(1000 * ranks + seed)
        solves = _random_solves(rng, ranks, 7, 6)
        m = Model(ranks, solves, rng)
        m.run()
        assert all(s == len(solves) for s in m.solve_of)


def test_sequence_numbers_wrap_the_tag():
    # 30 exchanges between a shrink and the regrowth: the old tag AND the old buffer come round again
    solves = [([5, 5], 3), ([1, 2], 5), ([1, 1], 5), ([2, 1], 5), ([1, 2], 5), ([1, 1], 5), ([2, 2], 4), ([5, 5], 4), ([2, 2], 1), ([5, 4], 3)]
    for seed in range(40):
        m = Model(2, solves, random.Random(seed))
        m.run()


def _finds_violation(**kw):
    for seed in range(400):
        rng = random.Random(seed)
        # the tag comes round every 15 exchanges and the buffer every 2: 30 exchanges after the large solve a stale slot matches
        solves = [([5, 5], 2)] + [([1, 1], rng.randint(3, 5)) for _ in range(rng.randint(5, 9))] + [([5, 5], 3)]
        try:
            Model(2, solves, rng, **kw).run()
        except Violation:
            return True
    return False


def test_model_sees_a_stale_partial_without_poison():
    assert _finds_violation(poison=False)


def test_model_sees_the_lost_partial_when_both_buffers_are_poisoned_at_once():
    assert _finds_violation(poison=True, poison_both_at_start=True)
