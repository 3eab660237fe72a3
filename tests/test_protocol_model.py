"""Model check (CPU) of the two NVLink protocols used inside the kernels — the mailbox all-reduce
(csrc/pk_device.cuh: pk_grid_reduce) and the halo push fused into the SpMV kernel (csrc/pk_spmv.cu: k_spmv_tma<HALO>) — under random
interleavings of the ranks' steps.  The model keeps the features the correctness argument rests on: two banks selected
by sequence parity, payload stores ordered before the flag store, readers that only read after seeing their flags, and
no other synchronisation between ranks.  Checked: every rank obtains exactly the sum / halo of the right sequence number
(no stale or overwritten bank is ever read), for thousands of random schedules, including ranks that run far ahead."""
import random

import pytest


def _allreduce_rank(me, P, mbox, n_rounds, values, results):
    """Generator = one rank's last block; every `yield` is a point where another rank may run."""
    for seq in range(1, n_rounds + 1):
        bank = seq & 1
        v = values[me][seq]
        for p in range(P):                          # payload stores to every mailbox (own included)
            mbox[p][bank][me]["payload"] = (seq, v)
            yield
        for p in range(P):                          # fence, then flags
            mbox[p][bank][me]["flag"] = seq
            yield
        for p in range(P):                          # wait for the P flags addressed to me
            while mbox[me][bank][p]["flag"] != seq:
                yield
        total = 0
        for p in range(P):                          # read in rank order
            s, val = mbox[me][bank][p]["payload"]
            assert s == seq, f"rank {me} read payload of seq {s} while reducing seq {seq}"
            total += val
            yield
        results[me].append(total)


@pytest.mark.parametrize("P", [2, 3, 8])
def test_mailbox_allreduce_two_banks_suffice(P):
    rng = random.Random(P)
    for trial in range(300):
        n_rounds = 12
        values = [[rng.randrange(1000) for _ in range(n_rounds + 1)] for _ in range(P)]
        mbox = [[[{"payload": (0, 0), "flag": 0} for _ in range(P)] for _ in range(2)] for _ in range(P)]
        results = [[] for _ in range(P)]
        gens = [_allreduce_rank(r, P, mbox, n_rounds, values, results) for r in range(P)]
        alive = list(range(P))
        # biased scheduler: one rank is strongly favoured so that it runs as far ahead as the protocol allows
        fav = rng.randrange(P)
        steps = 0
        while alive:
            r = fav if (fav in alive and rng.random() < 0.7) else rng.choice(alive)
            try:
                next(gens[r])
            except StopIteration:
                alive.remove(r)
            steps += 1
            assert steps < 2_000_000, "deadlock in the model"
        want = [sum(values[p][s] for p in range(P)) for s in range(1, n_rounds + 1)]
        for r in range(P):
            assert results[r] == want


def _allreduce_split_rank(me, P, mbox, n_rounds, values, results):
    """CG's r.r reduction with post and wait in DIFFERENT kernels (k_cg_r posts, k_cg_x runs, k_ar_wait collects),
    alternating with the ordinary post+wait reduction of the SpMV epilogue, as in Solve::cg() on several GPUs."""
    for seq in range(1, n_rounds + 1):
        bank = seq & 1
        v = values[me][seq]
        for p in range(P):
            mbox[p][bank][me]["payload"] = (seq, v)
            yield
        for p in range(P):
            mbox[p][bank][me]["flag"] = seq
            yield
        if seq % 2 == 0:                            # the split reduction: independent work between post and wait
            for _ in range(5):
                yield
        for p in range(P):
            while mbox[me][bank][p]["flag"] != seq:
                yield
        total = 0
        for p in range(P):
            s, val = mbox[me][bank][p]["payload"]
            assert s == seq, f"rank {me} read payload of seq {s} while reducing seq {seq}"
            total += val
            yield
        results[me].append(total)


@pytest.mark.parametrize("P", [2, 8])
def test_split_post_wait_allreduce_keeps_the_two_bank_invariant(P):
    rng = random.Random(50 + P)
    for trial in range(200):
        n_rounds = 12
        values = [[rng.randrange(1000) for _ in range(n_rounds + 1)] for _ in range(P)]
        mbox = [[[{"payload": (0, 0), "flag": 0} for _ in range(P)] for _ in range(2)] for _ in range(P)]
        results = [[] for _ in range(P)]
        gens = [_allreduce_split_rank(r, P, mbox, n_rounds, values, results) for r in range(P)]
        alive = list(range(P))
        fav = rng.randrange(P)
        steps = 0
        while alive:
            r = fav if (fav in alive and rng.random() < 0.7) else rng.choice(alive)
            try:
                next(gens[r])
            except StopIteration:
                alive.remove(r)
            steps += 1
            assert steps < 2_000_000, "deadlock in the model"
        want = [sum(values[p][s] for p in range(P)) for s in range(1, n_rounds + 1)]
        for r in range(P):
            assert results[r] == want


def _halo_rank(me, P, recv, n_rounds, data, got, sends, symmetric_flags=True):
    """One rank's SpMV kernels with the exchange fused in (csrc/pk_spmv.cu: k_spmv_tma<HALO>): at its start the kernel
    pushes my entries to the ranks that need them and then raises its flag at EVERY peer (any rank it exchanges with in
    either direction); interior tiles; before the boundary tiles — or, without boundary rows, before the kernel ends — it
    waits for the flag of THIS exchange from every peer, then reads the entries sent to it.  `sends[r]` = ranks r sends
    to (arbitrary, possibly one-directional).  symmetric_flags=False models the flawed variant where flags only follow
    the data (a pure sender is then never throttled by its receiver)."""
    to = sorted(sends[me])
    frm = sorted(q for q in range(P) if me in sends[q])
    peers = sorted(set(to) | set(frm))
    flag_to = peers if symmetric_flags else to
    wait_from = peers if symmetric_flags else frm
    for seq in range(1, n_rounds + 1):
        bank = seq & 1
        for q in to:                                # push: data ...
            recv[q][bank][me]["data"] = (seq, data[me][seq])
            yield
        for q in flag_to:                           # ... fence, flags
            recv[q][bank][me]["flag"] = seq
            yield
        for _ in range(3):                          # interior tiles
            yield
        for q in wait_from:                         # boundary tiles (or kernel exit): wait, then read
            while recv[me][bank][q]["flag"] != seq:
                yield
        for q in frm:
            s, val = recv[me][bank][q]["data"]
            assert s == seq, f"rank {me} read halo of exchange {s} during exchange {seq}"
            got[me].append((seq, q, val))
            yield


def _run_halo_model(P, sends, rng, symmetric_flags=True, trials=300):
    for trial in range(trials):
        n_rounds = 10
        data = [[rng.randrange(1000) for _ in range(n_rounds + 1)] for _ in range(P)]
        recv = [[[{"data": (0, 0), "flag": 0} for _ in range(P)] for _ in range(2)] for _ in range(P)]
        got = [[] for _ in range(P)]
        gens = [_halo_rank(r, P, recv, n_rounds, data, got, sends, symmetric_flags) for r in range(P)]
        alive = list(range(P))
        fav = rng.randrange(P)
        steps = 0
        while alive:
            r = fav if (fav in alive and rng.random() < 0.7) else rng.choice(alive)
            try:
                next(gens[r])
            except StopIteration:
                alive.remove(r)
            steps += 1
            assert steps < 2_000_000, "deadlock in the model"
        for r in range(P):
            for seq, q, val in got[r]:
                assert val == data[q][seq]


@pytest.mark.parametrize("P", [2, 4, 8])
def test_halo_push_two_banks_suffice(P):
    """Symmetric neighbour exchange (slabs, bands): a ring."""
    ring = [sorted({(r - 1) % P, (r + 1) % P} - {r}) for r in range(P)]
    _run_halo_model(P, ring, random.Random(100 + P))


@pytest.mark.parametrize("P", [2, 3, 8])
def test_halo_push_one_directional_patterns(P):
    """Structurally nonsymmetric A (block triangular): rank r only SENDS to r+1 and rank P-1 has no boundary rows to feed
    anyone.  Flags in both directions keep a pure sender at most one exchange ahead of its receiver; random digraphs too."""
    chain = [[r + 1] if r + 1 < P else [] for r in range(P)]
    _run_halo_model(P, chain, random.Random(200 + P))
    rng = random.Random(300 + P)
    for _ in range(10):
        sends = [[q for q in range(P) if q != r and rng.random() < 0.4] for r in range(P)]
        _run_halo_model(P, sends, rng, trials=30)


def test_model_detects_the_unthrottled_pure_sender():
    """Sanity of the model: if flags only follow the data, a pure sender runs two exchanges ahead and overwrites a bank
    its receiver is still reading (the hazard ADVICE r01 named)."""
    with pytest.raises(AssertionError):
        _run_halo_model(2, [[1], []], random.Random(7), symmetric_flags=False, trials=200)


def test_model_detects_the_single_bank_hazard():
    """Sanity of the model itself: with ONE bank a fast rank overwrites a payload its peer has not read yet."""
    src = open(__file__).read().split("def test_model_detects_the_single_bank_hazard")[0].replace("bank = seq & 1", "bank = 0")
    ns = {}
    exec(compile(src, "one_bank_model", "exec"), ns)
    with pytest.raises(AssertionError):
        ns["test_mailbox_allreduce_two_banks_suffice"](3)
