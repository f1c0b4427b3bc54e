"""Index bookkeeping of the tcgen05 GEMM's MMA-issue warp (recnn_b200/csrc/tc_gemm.cuh), modelled in Python.

The default kernel derives stage / phase / A slot / chunk / accumulator buffer / wait parities from the
k-block index i with divisions and modulos; the LEAN kernel keeps running counters.  The two must produce the
same sequence of (mbarrier, parity) waits, accumulate flags and commits for every k-block count -- a wrong
parity is a deadlock (or a data race) on the GPU.  This test enumerates both schemes for all the pipeline
shapes the library instantiates."""
import pytest


def formula_scheme(num_kb, STAGES, A_SLOTS, CH, BK=32):
    ev = []
    for i in range(num_kb):
        s, ph = i % STAGES, (i // STAGES) & 1
        chunk = i // CH
        buf = chunk & 1
        if i % CH == 0:
            ev.append(("wait", "acc_empty", buf, ((chunk >> 1) & 1) ^ 1))
        ev.append(("wait", "split", s, ph))
        slot = i % A_SLOTS
        for k in range(BK // 8):
            ev.append(("mma", s, slot, buf, int((i | k) != 0), int(((i % CH) | k) != 0)))
        ev.append(("commit", "empty", s))
        ev.append(("commit", "a_free", slot))
        if i % CH == CH - 1 or i == num_kb - 1:
            ev.append(("commit", "acc_full", buf))
    return ev


def counter_scheme(num_kb, STAGES, A_SLOTS, CH, BK=32):
    """Transcription of the `if constexpr (C::LEAN)` loop."""
    ev = []
    s = ph = slot = kin = buf = 0
    par = [1, 1]
    lo_acc = 0
    for i in range(num_kb):
        if kin == 0:
            ev.append(("wait", "acc_empty", buf, par[buf]))
            par[buf] ^= 1
        ev.append(("wait", "split", s, ph))
        for k in range(BK // 8):
            lo_flag = lo_acc if k == 0 else 1
            hi_flag = (1 if kin != 0 else 0) if k == 0 else 1
            ev.append(("mma", s, slot, buf, lo_flag, hi_flag))
        ev.append(("commit", "empty", s))
        ev.append(("commit", "a_free", slot))
        if kin == CH - 1 or i == num_kb - 1:
            ev.append(("commit", "acc_full", buf))
        lo_acc = 1
        s += 1
        if s == STAGES:
            s, ph = 0, ph ^ 1
        slot += 1
        if slot == A_SLOTS:
            slot = 0
        kin += 1
        if kin == CH:
            kin, buf = 0, buf ^ 1
    return ev


@pytest.mark.parametrize("STAGES,A_SLOTS", [(4, 2), (6, 5), (6, 4)])     # BN=128 | BN=64 | BN=64 with LO2
def test_lean_counters_reproduce_the_index_formulas(STAGES, A_SLOTS):
    for num_kb in range(0, 140):
        assert counter_scheme(num_kb, STAGES, A_SLOTS, 2) == formula_scheme(num_kb, STAGES, A_SLOTS, 2), num_kb


@pytest.mark.parametrize("STAGES,A_SLOTS,groups", [(4, 2, 2), (6, 5, 2), (6, 5, 4), (6, 4, 4), (6, 4, 2)])
def test_worker_groups_cover_every_k_block_once_and_drain_every_chunk(STAGES, A_SLOTS, groups):
    """Worker side (shared by all variants): group g takes k-blocks g, g+groups, ...; after its k-block i it drains
    every chunk below i // CH.  Every k-block must be split exactly once, every chunk drained exactly once by
    every group, a group's A slot must differ from the slots of the other groups' k-blocks in flight, and the
    accumulator a group drains must be complete no later than the MMAs of k-blocks <= i_g + groups."""
    CH = 2
    for num_kb in range(0, 70):
        num_chunks = (num_kb + CH - 1) // CH
        split_by = {}
        for g in range(groups):
            next_drain, drained = 0, []
            i = g
            while i < num_kb:
                assert i not in split_by
                split_by[i] = g
                if (i + groups) // CH != i // CH:
                    while next_drain < i // CH:
                        drained.append(next_drain)
                        next_drain += 1
                i += groups
            while next_drain < num_chunks:
                drained.append(next_drain)
                next_drain += 1
            assert drained == list(range(num_chunks)), (num_kb, g)
        assert sorted(split_by) == list(range(num_kb))
        assert groups <= A_SLOTS
        for i in range(num_kb - groups + 1):          # k-blocks i .. i+groups-1 may be in flight together
            assert len({(i + d) % A_SLOTS for d in range(groups)}) == groups
