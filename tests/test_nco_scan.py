"""The product's exact NCO fast-forward (csrc/nco_scan.cuh, run on the host via
gpsiq_nco_advance) must equal the literal per-sample recurrences of
plutogpssim.c:2709-2713 / 2741-2746 (oracle_*_nco) bit for bit."""
import random
import struct

import numpy as np
import pytest

import oracle_lib as ol
from pluto_gps_sim_b200 import capi


def bits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


FS = [2.6e6, 3.0e6, 10.0e6, 1.0e6]


def test_code_nco_random():
    rng = random.Random(1234)
    for trial in range(120):
        fs = rng.choice(FS)
        f_carr = rng.uniform(-6000, 6000)
        f_code = 1.023e6 + f_carr / 1540.0
        step = f_code * (1.0 / fs)
        x0 = rng.uniform(0, 1023) if trial % 5 else rng.choice([0.0, 1e-9, 0.5, 1022.999999, 511.99999999])
        n = rng.choice([1, 2, 3, 7, 300000, 123457, 1000000])
        want, ww = ol.oracle_code_nco(x0, step, n)
        got, gw = capi.nco_advance(capi.NCO_CODE, x0, step, n)
        assert bits(got) == bits(want) and gw == ww, (trial, x0, step, n)


def test_carrier_nco_random():
    rng = random.Random(99)
    for trial in range(160):
        fs = rng.choice(FS)
        kind = trial % 4
        if kind == 0:
            f = rng.uniform(-6000, 6000)
        elif kind == 1:
            f = rng.uniform(-3, 3)
        elif kind == 2:
            f = rng.choice([-1, 1]) * 2.0 ** rng.randint(-8, 12)
        else:
            f = rng.uniform(-400, 400)
        step = f * (1.0 / fs)
        x0 = rng.random() if trial % 7 else rng.choice([0.0, 0.5, 0.25, 0.999999999999, 2.0 ** -30])
        n = rng.choice([1, 2, 5, 300000, 777777, 3000000])
        want = ol.oracle_carr_nco(x0, step, n)
        got, _ = capi.nco_advance(capi.NCO_CARRIER, x0, step, n)
        assert bits(got) == bits(want), (trial, x0, step, n)


def test_carrier_nco_exact_ties_and_stuck_phase():
    # step = (m + 1/2) ulp of [0.5,1): every add is a rounding tie
    for m in (1, 2, 3, 1001):
        step = (m + 0.5) * 2.0 ** -53
        for x0 in (0.5, 0.5 + 2.0 ** -53, 0.75):
            want = ol.oracle_carr_nco(x0, step, 100000)
            got, _ = capi.nco_advance(capi.NCO_CARRIER, x0, step, 100000)
            assert bits(got) == bits(want)
    # step below half an ulp: the phase stops moving
    got, _ = capi.nco_advance(capi.NCO_CARRIER, 0.75, 2.0 ** -60, 1000000)
    assert got == 0.75 == ol.oracle_carr_nco(0.75, 2.0 ** -60, 1000000)


def test_scan_matches_reference_epoch_chain():
    """Chain the carrier scan over the golden static scenario: each epoch's end
    phase must equal the reference's own post-loop carr_phase."""
    meta = ol.load_golden_meta("static12")
    desc = ol.load_golden_desc("static12")
    n = meta["samples_per_epoch"]
    ph = desc[0]["carr_phase0"].copy()
    for e in range(desc.shape[0]):
        for c in range(desc.shape[1]):
            ph[c], _ = capi.nco_advance(capi.NCO_CARRIER, ph[c], desc[e, c]["carr_step"], n)
            assert ph[c].hex() == meta["carr_phase_end_hex"][e][c]


def _literal_checkpoints(steps, N, T, x0):
    nt = (N + T - 1) // T
    ck = np.zeros((len(steps), nt))
    x = x0
    for e, d in enumerate(steps):
        for t in range(nt):
            ck[e, t] = x
            x = ol.oracle_carr_nco(x, d, min(T, N - t * T))
    return ck, x


@pytest.mark.parametrize("seed", [5, 6])
def test_parallel_carrier_scan_speculate_translate_verify(seed):
    """The parallel carrier scan (speculative epoch scans from estimated start
    phases, translated onto the exact chain, serial fallback) gives the literal
    recurrence's state at every tile boundary -- whatever the estimate error."""
    rng = random.Random(seed)
    translated = 0
    for trial in range(30):
        fs = rng.choice([2.6e6, 3e6, 1e7])
        E = rng.choice([3, 8, 20])
        N = rng.choice([300000, 100000, 4097])
        T = rng.choice([1024, 2048, 512])
        f0 = rng.uniform(-5000, 5000) if trial % 3 else rng.uniform(-30, 30)
        steps = [(f0 + rng.uniform(-2, 2)) * (1.0 / fs) for _ in range(E)]
        x0 = rng.random()
        err = rng.choice([0.0, 0.0, 1e-12, 1e-9, 1e-6, 1e-3])
        want, wx = _literal_checkpoints(steps, N, T, x0)
        got, gx, fb = capi.carrier_chain_host(steps, N, T, x0, err)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (trial, f0, E, N, T, err)
        assert bits(gx) == bits(wx)
        translated += E - fb
    assert translated > 100          # the fast path is actually exercised


def test_parallel_carrier_scan_unspeculable_steps():
    """Steps that are multiples of 2^-53 (ties possible in [1,2)) and huge steps are never speculated."""
    N, T = 50000, 1024
    for d in (3.0 * 2.0 ** -12, 2.0 ** -9, -2.0 ** -9, 0.3, -0.26, 5 * 2.0 ** -53):
        want, wx = _literal_checkpoints([d, d, d], N, T, 0.123456789)
        got, gx, fb = capi.carrier_chain_host([d, d, d], N, T, 0.123456789, 0.0)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), d
        assert bits(gx) == bits(wx)


def test_negative_doppler_uses_the_derived_parity_variant():
    """step < 0: the second parity variant of every chunk run is DERIVED from the first (shifted by 2^-53 until the first
    wrap whose y + 1.0 is an exact tie, then by 0 or 2^-52; nco_scan.cuh: spec_derive_variant1).  Many negative steps,
    random phases (both parities of the true post-wrap state occur), long epochs (dozens of tie-wraps): every tile-start
    phase equals the literal recurrence, and nearly every epoch is translated, not re-scanned."""
    rng = random.Random(2024)
    fb_total, epochs = 0, 0
    for trial in range(24):
        fs = rng.choice([2.6e6, 1e7])
        f0 = -rng.uniform(300, 5000)
        E, N, T = rng.choice([6, 12]), 300000, 1024
        steps = [(f0 + rng.uniform(-1, 1)) * (1.0 / fs) for _ in range(E)]
        x0 = rng.random()
        err = rng.choice([0.0, 3e-13, 1e-11])
        want, wx = _literal_checkpoints(steps, N, T, x0)
        got, gx, fb = capi.carrier_chain_host(steps, N, T, x0, err)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (trial, f0, err)
        assert bits(gx) == bits(wx)
        fb_total += fb
        epochs += E
    assert fb_total <= epochs // 10, (fb_total, epochs)

    # a step whose ordinary additions in [0.5, 1) can tie (lowest set bit exactly 2^-54): scanned for real, still exact
    d = -(1234567 * 2.0 ** -30 + 2.0 ** -54)
    want, wx = _literal_checkpoints([d] * 4, 100000, 1024, 0.3141592653589793)
    got, gx, fb = capi.carrier_chain_host([d] * 4, 100000, 1024, 0.3141592653589793, 1e-12)
    assert np.array_equal(got.view(np.int64), want.view(np.int64)) and bits(gx) == bits(wx)


@pytest.mark.parametrize("seed", [11, 12])
def test_slice_level_speculation_one_head_scan_per_batch(seed):
    """Level 5 (nco_scan.cuh: slice_chain_group / slice_verify): the groups of a batch are chained speculatively from
    the ESTIMATED batch start, one exact head scan is matched against that trajectory, and every group is then chained
    on its own from its translated start phase.  Whatever the estimate error, every tile-start phase and the end phase
    equal the literal recurrence; the groups' ends always agree with the translation (how != -2); with a good estimate
    the batch is passed by translation (how == 1), with a poor one it is chained serially (how == 0)."""
    rng = random.Random(seed)
    passed = serial = 0
    for trial in range(36):
        fs = rng.choice([2.6e6, 1e7])
        E = rng.choice([5, 12, 17])                      # host groups hold 4 epochs: 2 to 5 groups, ragged last one
        N = rng.choice([260000, 100000, 4097])
        T = rng.choice([1024, 2048])
        f0 = rng.uniform(-5000, 5000) if trial % 4 else rng.uniform(-40, 40)   # incl. epochs that never wrap
        steps = [(f0 + rng.uniform(-2, 2)) * (1.0 / fs) for _ in range(E)]
        x0 = rng.random()
        err = rng.choice([0.0, 0.0, 2e-14, 1e-12, 1e-10, 1e-6, 1e-3])
        want, wx = _literal_checkpoints(steps, N, T, x0)
        got, gx, fb, how = capi.carrier_slice_host(steps, N, T, x0, err)
        assert how in (0, 1), (trial, how)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (trial, f0, E, N, T, err, how)
        assert bits(gx) == bits(wx)
        passed += how == 1
        serial += how == 0
        if err <= 1e-12 and abs(f0) > 300:
            assert how == 1, (trial, f0, err)            # a good estimate must take the fast path
    assert passed >= 12 and serial >= 3, (passed, serial)


def test_slice_level_speculation_negative_steps_both_parities():
    """Negative Doppler: the exact post-wrap state may sit on either parity of the 2^-53 grid; the slice-level chain is
    speculated for both and the matching one is used."""
    rng = random.Random(77)
    hows = []
    for trial in range(24):
        f0 = -rng.uniform(300, 5000)
        steps = [(f0 + rng.uniform(-1, 1)) / 2.6e6 for _ in range(9)]
        x0 = rng.random()
        want, wx = _literal_checkpoints(steps, 260000, 1024, x0)
        got, gx, fb, how = capi.carrier_slice_host(steps, 260000, 1024, x0, rng.choice([0.0, 5e-14, 3e-13]))
        assert np.array_equal(got.view(np.int64), want.view(np.int64)) and bits(gx) == bits(wx) and how in (0, 1)
        hows.append(how)
    assert sum(hows) >= 20, hows


def test_tie_capable_steps_are_translated_through_their_tie_event():
    """A positive step that is a multiple of 2^-53 (1 epoch in ~2^9 of a real run) can tie in [1, 2), where a translation
    by an odd multiple of 2^-52 does not commute with rounding.  Levels 3-5 translate THROUGH such epochs: up to the
    trajectory's first tie-wrap by D, after it by D +- 2^-52 if D is odd (nco_scan.cuh: TieEvent).  Epoch runs with
    one or two tie-capable epochs, random phases and estimate errors (so that D takes both parities): every tile-start
    phase equals the literal recurrence, group trajectories holding a tie event stay usable, and the changed shift is
    actually applied at the group level and at the slice level."""
    rng = random.Random(4242)
    seen = applied_g = applied_s = passed = 0
    ties = np.zeros(2, np.int32)
    for trial in range(60):
        fs = rng.choice([2.6e6, 1e7])
        f0 = rng.uniform(300, 5000)
        E, N, T = rng.choice([8, 12, 16]), rng.choice([260000, 100000]), 1024
        steps = [(f0 + rng.uniform(-1, 1)) / fs for _ in range(E)]
        for e in rng.sample(range(E), rng.choice([1, 2])):
            steps[e] = np.ldexp(float(round(np.ldexp(steps[e], 53))), -53)      # tie-capable: a multiple of 2^-53
        x0 = rng.random()
        err = rng.choice([0.0, 2.0 ** -52, 3 * 2.0 ** -52, 1e-13, 7e-13, 1e-11])
        want, wx = _literal_checkpoints(steps, N, T, x0)
        got, gx, fb, how = capi.carrier_slice_host(steps, N, T, x0, err, ties)
        assert how in (0, 1), (trial, how)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (trial, f0, E, N, err, how)
        assert bits(gx) == bits(wx)
        seen += int(ties[0])
        applied_g += int(ties[1]) % 1000
        applied_s += int(ties[1]) // 1000
        passed += how == 1
        # the plain chain (levels 1-4 only) goes through the same group code
        got2, gx2, _ = capi.carrier_chain_host(steps, N, T, x0, err)
        assert np.array_equal(got2.view(np.int64), want.view(np.int64)) and bits(gx2) == bits(wx)
    assert seen >= 40 and applied_g >= 5 and applied_s >= 3 and passed >= 30, (seen, applied_g, applied_s, passed)


def _literal_with_flags(steps, flags, phase0, N, T, x0):
    nt = (N + T - 1) // T
    ck = np.zeros((len(steps), nt))
    x = x0
    for e, d in enumerate(steps):
        if flags[e] & 1:                       # inactive: the phase passes through, no tile phases
            continue
        if flags[e] & 2:
            x = phase0[e]
        for t in range(nt):
            ck[e, t] = x
            x = ol.oracle_carr_nco(x, d, min(T, N - t * T))
    return ck, x


def test_slice_level_chain_with_inactive_and_reseeded_epochs():
    """A slot may be inactive for some epochs of a batch (the phase passes through: plutogpssim.c:2690 skips the
    channel) and may be re-seeded inside it (allocation, plutogpssim.c:1964).  Inactive epochs -- at the start of the
    batch, at the start of a group, whole groups -- are translated through; a re-seed makes the slice untranslatable
    and it is chained serially.  Every active tile-start phase and the end phase equal the literal recurrence."""
    rng = random.Random(909)
    translated = serial = 0
    for trial in range(40):
        fs = rng.choice([2.6e6, 1e7])
        f0 = rng.uniform(-5000, 5000)
        E, N, T = rng.choice([9, 12, 16]), rng.choice([260000, 100000]), 1024
        steps = [(f0 + rng.uniform(-1, 1)) / fs for _ in range(E)]
        flags = np.zeros(E, np.int32)
        phase0 = np.zeros(E)
        kind = trial % 4
        if kind == 0:                          # inactive at the start of the batch (incl. its whole first group)
            flags[: rng.choice([1, 3, 4, 5])] = 1
        elif kind == 1:                        # scattered inactive epochs, incl. first epochs of groups
            for e in rng.sample(range(E), 3):
                flags[e] = 1
            flags[4] = 1
        elif kind == 2:                        # a whole group in the middle inactive
            flags[4:8] = 1
        else:                                  # a re-seed somewhere after the first epoch
            e = rng.randrange(1, E)
            flags[e] = 2
            phase0[e] = rng.random()
        x0 = rng.random()
        err = rng.choice([0.0, 1e-13, 1e-12])
        want, wx = _literal_with_flags(steps, flags, phase0, N, T, x0)
        got, gx, fb, how = capi.carrier_slice_host(steps, N, T, x0, err, flags=flags, phase0=phase0)
        assert how in (0, 1), (trial, how)
        act = (flags & 1) == 0
        assert np.array_equal(got[act].view(np.int64), want[act].view(np.int64)), (trial, kind, how)
        assert bits(gx) == bits(wx), (trial, kind, how)
        if kind == 3:
            assert how == 0, trial             # re-seeded inside the slice: never translated
        translated += how == 1
        serial += how == 0
    assert translated >= 12 and serial >= 10, (translated, serial)


def test_tie_events_adversarial_steps_and_phases():
    """Ties at every wrap (steps that are small multiples of 2^-53 of very different magnitudes), mixed with ordinary and
    negative steps, start phases on grid points (0, 0.5, 1 - 2^-53, small multiples of 2^-52) and estimate errors of a
    few units of 2^-52 of either sign: the translated chains stay exact, with and without the slice level."""
    rng = random.Random(2025)
    ties = np.zeros(2, np.int32)
    applied = 0
    mags = [2.0 ** -9, 3 * 2.0 ** -12, 5 * 2.0 ** -11, 2.0 ** -10 + 2.0 ** -53, 2.0 ** -10 + 3 * 2.0 ** -53,
            7 * 2.0 ** -14 + 2.0 ** -52, 2.0 ** -12 + 2.0 ** -40, 1.1e-3]
    for trial in range(120):
        E, N, T = rng.choice([5, 8, 12]), rng.choice([4097, 30000]), rng.choice([512, 1024])
        base = rng.choice(mags)
        steps = []
        for _ in range(E):
            m = (rng.choice(mags) if rng.random() < 0.5 else base) * (1 + rng.uniform(-1e-3, 1e-3))
            d = float(np.ldexp(float(round(np.ldexp(m, 53))), -53)) if rng.random() < 0.6 else m
            steps.append(-d if rng.random() < 0.15 else d)
        x0 = rng.choice([rng.random(), 0.0, 0.5, 1 - 2.0 ** -53, 2.0 ** -52 * rng.randrange(1, 1000)])
        err = rng.choice([0.0, 2.0 ** -52, -2.0 ** -52, 3 * 2.0 ** -52, -5 * 2.0 ** -52, 2.0 ** -53, 1e-13, -7e-13, 1e-11])
        want, wx = _literal_checkpoints(steps, N, T, x0)
        got, gx, fb, how = capi.carrier_slice_host(steps, N, T, x0, err, ties)
        assert how in (0, 1), (trial, how)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (trial, steps, x0, err, how)
        assert bits(gx) == bits(wx)
        got2, gx2, _ = capi.carrier_chain_host(steps, N, T, x0, err)
        assert np.array_equal(got2.view(np.int64), want.view(np.int64)) and bits(gx2) == bits(wx), trial
        applied += int(ties[1])
    assert applied % 1000 >= 30 and applied // 1000 >= 10, applied
