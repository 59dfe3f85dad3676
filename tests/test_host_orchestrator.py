"""Host orchestrator (libgpshost.so): the descriptors it computes from a
navigation file must be BIT-identical to the reference's own per-epoch channel state -- against the committed
goldens (made from the compiled reference, tools/gen_golden.py) and, where oracle/_ref is built, against
live runs of the reference with other option combinations (-t, -T, -i, -c, ephemeris roll-over, and -3: a
RINEX-3 file of the same constellation with records of other constellations interleaved)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import refdump
from pluto_gps_sim_b200 import capi, hostapi

NAV12 = os.path.join(ol.GOLDEN, "brdc3540_synth.14n.gz")
NAV32 = os.path.join(ol.GOLDEN, "allsky32_synth.14n.gz")
NAV12_V3 = os.path.join(ol.GOLDEN, "brdc3540_synth.14p.gz")      # RINEX 3.02 (tools/gen_rinex_fixture.py --rinex3)
CIRCLE = os.path.join(refdump.REF_DIR, "circle.csv")
LLH = (30.286502, 120.032669, 100)


def assert_same_descriptors(got, want):
    """bit-for-bit; carr_phase0 only where the reset flag is set (elsewhere the golden carries the sample
    loop's running phase, which the context -- not the host -- owns)."""
    assert got.shape == want.shape
    for f in ("prn", "ms0", "navbits", "flags"):
        bad = np.argwhere(got[f] != want[f])
        assert len(bad) == 0, (f, bad[0], got[f][tuple(bad[0])], want[f][tuple(bad[0])])
    for f in ("code_phase0", "code_step", "carr_step", "gain"):
        bad = np.argwhere(got[f].view(np.uint64) != want[f].view(np.uint64))
        assert len(bad) == 0, (f, bad[0], got[f][tuple(bad[0])].hex(), want[f][tuple(bad[0])].hex())
    m = (want["flags"] & 1) != 0
    assert np.array_equal(got["carr_phase0"][m].view(np.uint64), want["carr_phase0"][m].view(np.uint64))


def test_static12_descriptors_equal_reference_golden():
    with hostapi.Scenario(NAV12, llh=LLH, sample_rate=2600000) as s:
        got = s.next(10)
        assert "PRN   Az    El     Range     Iono" in s.describe()
    assert_same_descriptors(got, ol.load_golden_desc("static12"))


def test_allsky32_descriptors_equal_reference_golden():
    with hostapi.Scenario(NAV32, llh=LLH, sample_rate=10000000, max_chan=32) as s:
        got = s.next(20)
    assert_same_descriptors(got, ol.load_golden_desc("allsky32"))


@pytest.mark.skipif(not os.path.exists(CIRCLE), reason="needs the reference's circle.csv (oracle/_ref)")
def test_circle12_user_motion_across_the_30s_refresh():
    """310 epochs of user motion: NAV frame rebuild and channel re-allocation at t = 30 s."""
    with hostapi.Scenario(NAV12, motion=CIRCLE, sample_rate=2600000) as s:
        got = np.concatenate([s.next(n) for n in (1, 7, 128, 100, 74)])   # batching must not matter
    assert_same_descriptors(got, ol.load_golden_desc("circle12"))


LIVE = [
    # name, reference argv, Scenario kwargs, epochs
    ("start_time", ["-l", "30.286502,120.032669,100", "-s", "2600000", "-t", "2014/12/20,03:10:05"],
     dict(llh=LLH, sample_rate=2600000, start=(2014, 12, 20, 3, 10, 5.0)), 12),
    ("no_iono_ecef", ["-c", "-2758918.64,4772301.12,3197889.44", "-s", "3000000", "-i"],
     dict(xyz=(-2758918.64, 4772301.12, 3197889.44), sample_rate=3000000, iono=False), 12),
    ("overwrite", ["-l", "30.286502,120.032669,100", "-s", "2600000", "-t", "2021/06/07,08:09:10", "-T", "x"],
     dict(llh=LLH, sample_rate=2600000, start=(2021, 6, 7, 8, 9, 10.0), time_overwrite=True), 12),
    ("ephemeris_rollover", ["-l", "30.286502,120.032669,100", "-s", "2600000", "-t", "2014/12/20,00:59:40"],
     dict(llh=LLH, sample_rate=2600000, start=(2014, 12, 20, 0, 59, 40.0)), 330),
]


@pytest.mark.skipif(not refdump.have_ref(), reason="needs the compiled reference (oracle/_ref)")
@pytest.mark.parametrize("start,epochs", [(None, 12), ((2014, 12, 20, 0, 59, 40.0), 330)])
def test_rinex3_against_live_reference(tmp_path, start, epochs):
    """The reference's -3 path (readRinex3, plutogpssim.c:1241-1610) on a RINEX-3 file, incl. an ephemeris roll-over."""
    argv = ["-3", "x", "-l", "30.286502,120.032669,100", "-s", "2600000"]
    if start:
        argv += ["-t", "%04d/%02d/%02d,%02d:%02d:%02d" % tuple(int(v) for v in start)]
    refdump.SCENARIOS["_live3"] = ("ref_harness_O2", "brdc3540_synth.14p.gz", argv, 12)
    recs, _, _ = refdump.run_reference("_live3", epochs, str(tmp_path), want_iq=False)
    want = refdump.to_descriptors(recs)
    with hostapi.Scenario(NAV12_V3, llh=LLH, sample_rate=2600000, start=start, rinex3=True) as s:
        got = s.next(epochs)
    assert_same_descriptors(got, want)
    # the same constellation as the RINEX-2 fixture: the two readers agree with each other as well
    with hostapi.Scenario(NAV12, llh=LLH, sample_rate=2600000, start=start) as s2:
        assert_same_descriptors(got, s2.next(epochs))


def test_rinex_version_mismatch_is_reported():
    with pytest.raises(hostapi.HostError) as e:
        hostapi.Scenario(NAV12, llh=LLH, rinex3=True)            # a RINEX-2 file read as RINEX 3
    assert e.value.status == hostapi.ERR_NAVFILE
    with pytest.raises(hostapi.HostError) as e:
        hostapi.Scenario(NAV12_V3, llh=LLH)                      # and the other way round
    assert e.value.status == hostapi.ERR_NAVFILE


@pytest.mark.skipif(not refdump.have_ref(), reason="needs the compiled reference (oracle/_ref)")
@pytest.mark.parametrize("name,argv,kw,epochs", LIVE, ids=[c[0] for c in LIVE])
def test_option_combinations_against_live_reference(tmp_path, name, argv, kw, epochs):
    refdump.SCENARIOS["_live"] = ("ref_harness_O2", "brdc3540_synth.14n.gz", argv, 12)
    recs, _, _ = refdump.run_reference("_live", epochs, str(tmp_path), want_iq=False)
    want = refdump.to_descriptors(recs)
    with hostapi.Scenario(NAV12, **kw) as s:
        got = s.next(epochs)
    assert_same_descriptors(got, want)


def test_known_answers():
    # values the reference's own functions produce
    assert hostapi.parity(0x8B0000 << 6, 0) == 0x22C00012
    assert hostapi.date2gps(2014, 12, 20, 0, 0, 0.0) == (1823, 518400.0)
    xyz = hostapi.llh2xyz(30.286502, 120.032669, 100.0)
    assert np.allclose(xyz, [-2758918.635941, 4772301.120089, 3197889.437237], atol=1e-6)
    llh = hostapi.xyz2llh(xyz)
    assert abs(llh[0] * 57.2957795131 - 30.286502) < 1e-9 and abs(llh[2] - 100.0) < 1e-3
    # every word the orchestrator emits passes the parity equations it was built with
    with hostapi.Scenario(NAV12, llh=LLH, sample_rate=2600000) as s:
        d = s.next(1)
    assert (d["prn"] > 0).sum() == 12


def test_error_reporting():
    with pytest.raises(hostapi.HostError) as ei:
        hostapi.Scenario("/nonexistent.14n", llh=LLH)
    assert ei.value.status == hostapi.ERR_NAVFILE
    with pytest.raises(hostapi.HostError) as ei:
        hostapi.Scenario(NAV12, llh=LLH, start=(2016, 1, 1, 0, 0, 0.0))          # outside the file's span
    assert ei.value.status == hostapi.ERR_TIME
    with pytest.raises(hostapi.HostError) as ei:
        hostapi.Scenario(NAV12, motion="/nonexistent.csv")
    assert ei.value.status == hostapi.ERR_MOTION
    with pytest.raises(hostapi.HostError):
        hostapi.Scenario(NAV12, llh=LLH, sample_rate=500000)                     # the reference rejects < 1 MHz too


def test_library_exports_every_declared_symbol():
    for name in hostapi.SYMBOLS:
        assert hasattr(hostapi.lib, name)


@pytest.mark.parametrize("name,kw,batches", [
    ("static12", dict(nav=NAV12, llh=LLH, sample_rate=2600000), [700]),
    ("static12-ragged", dict(nav=NAV12, llh=LLH, sample_rate=2600000), [1, 298, 2, 17, 300, 95]),   # batches that end on / straddle the refresh
    ("circle12", dict(nav=NAV12, motion=CIRCLE, sample_rate=2600000), [650]),
    ("allsky32", dict(nav=NAV32, llh=LLH, sample_rate=10000000, max_chan=32), [400]),
    ("rinex3-int-carrier", dict(nav=NAV12_V3, rinex3=True, llh=LLH, sample_rate=2600000, carrier_mode=capi.CARRIER_INT32), [330]),
])
def test_worker_threads_do_not_change_a_bit(name, kw, batches, monkeypatch):
    """gpshost_next spreads the pseudorange computations of a batch over worker threads;
    the serial pass (threads=1) is the reference's order of evaluation.  Every byte of every descriptor must agree,
    across the 30 s refresh (NAV frame, re-allocation) and whatever the batch boundaries are."""
    if "motion" in kw and not os.path.exists(CIRCLE):
        pytest.skip("oracle/_ref/circle.csv not present")
    monkeypatch.delenv("GPSHOST_THREADS", raising=False)
    with hostapi.Scenario(threads=1, **kw) as s:
        want = s.next(sum(batches))
    for threads in (3, 8, 0):
        with hostapi.Scenario(threads=threads, **kw) as s:
            got = np.concatenate([s.next(n) for n in batches])
            t_end = s.time
        assert got.tobytes() == want.tobytes(), (name, threads)
    monkeypatch.setenv("GPSHOST_THREADS", "5")
    with hostapi.Scenario(threads=0, **kw) as s:
        assert s.next(sum(batches)).tobytes() == want.tobytes()
        assert s.time == t_end


def _mutated_nav(tmp_path, src, edits, name="bad.nav"):
    import gzip
    lines = gzip.open(src, "rt").read().split("\n")
    for edit in edits:
        edit(lines)
    p = tmp_path / name
    p.write_text("\n".join(lines))
    return str(p)


def _first_record(lines):
    return next(i for i, l in enumerate(lines) if "END OF HEADER" in l) + 1


def test_corrupted_navigation_files_fail_with_an_error_not_a_crash(tmp_path):
    """Malformed input the reference answers with out-of-bounds reads or an endless loop (found by mutating the
    fixtures under ASan/UBSan): month 0 in a record line (doy[] index -1, plutogpssim.c:844), an eccentricity the
    Kepler iteration cannot converge for (plutogpssim.c:483), geometry that turns NaN (ant_pat[] index,
    plutogpssim.c:2678), a file cut inside a record, an empty file.  A valid file is untouched by the guards:
    every golden above still matches bit for bit."""
    def month0(lines):
        i = _first_record(lines)
        lines[i] = lines[i][:6] + " 0" + lines[i][8:]

    def ecc4(lines):                       # BROADCAST ORBIT 2, second field: e
        for r in range(0, 8 * 12, 8):
            i = _first_record(lines) + r + 2
            lines[i] = lines[i][:22] + " 4.000000000000D+00" + lines[i][41:]

    def nan_sqrta(lines):                  # sqrt(A) = 0 -> n = inf -> NaN ranges
        for r in range(0, 8 * 40, 8):
            i = _first_record(lines) + r + 2
            lines[i] = lines[i][:60] + " 0.000000000000D+00"

    def cut(lines):
        del lines[_first_record(lines) + 3:]

    for edits, status in (([month0], (hostapi.ERR_NOEPH, hostapi.ERR_NAVFILE)), ([cut], (hostapi.ERR_NOEPH, hostapi.ERR_NAVFILE))):
        with pytest.raises(hostapi.HostError) as ei:
            hostapi.Scenario(_mutated_nav(tmp_path, NAV12, edits), llh=LLH, sample_rate=2600000)
        assert ei.value.status in status
    for edits in ([ecc4], [nan_sqrta]):   # opens or not -- but it returns, with an error at the latest from next()
        try:
            with hostapi.Scenario(_mutated_nav(tmp_path, NAV12, edits), llh=LLH, sample_rate=2600000) as s:
                s.next(320)
        except hostapi.HostError as e:
            assert e.status in (hostapi.ERR_NOEPH, hostapi.ERR_ARG, hostapi.ERR_NAVFILE)
    empty = tmp_path / "empty.nav"
    empty.write_text("")
    with pytest.raises(hostapi.HostError):
        hostapi.Scenario(str(empty), llh=LLH, sample_rate=2600000)
    with pytest.raises(hostapi.HostError):
        def month0_v3(lines):              # "Gnn yyyy mm dd ...": month in columns 9-10
            i = next(k for k in range(_first_record(lines), len(lines)) if lines[k].startswith("G"))   # first GPS record
            lines[i] = lines[i][:9] + " 0" + lines[i][11:]
        hostapi.Scenario(_mutated_nav(tmp_path, NAV12_V3, [month0_v3], "bad3.nav"), llh=LLH, sample_rate=2600000, rinex3=True)


# ---- independent restatement of the IS-GPS-200 parity equations (20.3.5.2, table 20-XIV) -------------------
_PARITY_TERMS = {          # D25..D30: (which previous-word bit, data bits d1..d24 that enter)
    25: (29, [1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23]),
    26: (30, [2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24]),
    27: (29, [1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22]),
    28: (30, [2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23]),
    29: (30, [1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24]),
    30: (29, [3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24]),
}


def _spec_word(d, p29, p30):
    """d: list of 24 source data bits d1..d24; p29/p30: D29*, D30* of the previous word -> the 30 transmitted bits."""
    out = [b ^ p30 for b in d]
    for k in range(25, 31):
        prev, terms = _PARITY_TERMS[k]
        v = p29 if prev == 29 else p30
        for t in terms:
            v ^= d[t - 1]
        out.append(v)
    return out


def test_parity_words_satisfy_the_is_gps_200_equations():
    """gpshost_parity (the reference's computeChecksum, plutogpssim.c:291-372) against the parity equations written
    out from the interface specification: 2000 random words, both with plain data (nib = 0) and with the two
    non-information bits solved so that D29 = D30 = 0 (nib = 1: words 2 and 10 of every subframe)."""
    rng = np.random.default_rng(2024)
    for _ in range(2000):
        src = int(rng.integers(0, 1 << 32)) & 0xFFFFFFC0
        p29, p30 = (src >> 31) & 1, (src >> 30) & 1
        d = [(src >> (29 - i)) & 1 for i in range(24)]
        got = hostapi.parity(src, 0)
        bits = [(got >> (29 - i)) & 1 for i in range(30)]
        assert bits == _spec_word(d, p29, p30), hex(src)
        got = hostapi.parity(src, 1)
        bits = [(got >> (29 - i)) & 1 for i in range(30)]
        assert bits[28] == 0 and bits[29] == 0                      # D29 = D30 = 0
        d_solved = [b ^ p30 for b in bits[:24]]                      # the source bits the word now carries
        assert d_solved[:22] == d[:22]                               # only d23, d24 were touched
        assert bits == _spec_word(d_solved, p29, p30), hex(src)


def test_time_and_position_round_trips():
    rng = np.random.default_rng(7)
    for _ in range(300):
        y, m, d = int(rng.integers(1981, 2099)), int(rng.integers(1, 13)), int(rng.integers(1, 29))
        hh, mm, ss = int(rng.integers(0, 24)), int(rng.integers(0, 60)), float(rng.integers(0, 60))
        import datetime
        week, sow = hostapi.date2gps(y, m, d, hh, mm, ss)
        delta = datetime.datetime(y, m, d, hh, mm, int(ss)) - datetime.datetime(1980, 1, 6)
        assert week * 604800 + sow == delta.total_seconds()          # no leap seconds in GPS time
        lat, lon, h = float(rng.uniform(-89, 89)), float(rng.uniform(-179, 179)), float(rng.uniform(-100, 20000))
        back = hostapi.xyz2llh(hostapi.llh2xyz(lat, lon, h))
        assert abs(back[0] * 57.2957795131 - lat) < 1e-7 and abs(back[1] * 57.2957795131 - lon) < 1e-7 and abs(back[2] - h) < 1e-2


@pytest.mark.parametrize("name,kw", [
    ("static12", dict(nav=NAV12, llh=LLH, sample_rate=2600000)),
    ("circle12", dict(nav=NAV12, motion=CIRCLE, sample_rate=2600000)),
    ("allsky32", dict(nav=NAV32, llh=LLH, sample_rate=10000000, max_chan=32)),
    ("ephemeris-rollover", dict(nav=NAV12, llh=LLH, sample_rate=2600000, start=(2014, 12, 20, 0, 59, 40.0))),   # next set at 01:00:00
])
def test_skip_then_next_equals_the_tail_of_one_long_run(name, kw):
    """gpshost_skip (how the owner of a later time slice reaches its first epoch) evaluates one epoch
    per 30 s refresh interval; what follows must be bit-identical to generating everything -- at every kind of cut:
    inside an interval, on the refresh epoch, right after it, several intervals on, and in two hops."""
    if "motion" in kw and not os.path.exists(CIRCLE):
        pytest.skip("oracle/_ref/circle.csv not present")
    with hostapi.Scenario(**kw) as s:
        full = s.next(1300)
    for n in (0, 1, 7, 298, 299, 300, 301, 599, 600, 911, 1200):
        with hostapi.Scenario(**kw) as s:
            s.skip(n)
            t_skip = s.time
            got = s.next(70)
        assert got.tobytes() == full[n:n + 70].tobytes(), (name, n)
        with hostapi.Scenario(**kw) as s:
            s.next(n)
            assert s.time == t_skip
    with hostapi.Scenario(**kw) as s:     # hops and interleaved generation
        s.skip(250)
        a = s.next(100)
        s.skip(333)
        b = s.next(50)
    assert a.tobytes() == full[250:350].tobytes() and b.tobytes() == full[683:733].tobytes()


def test_skip_is_cheap():
    import time
    with hostapi.Scenario(NAV12, llh=LLH, sample_rate=2600000, threads=1) as s:
        t = time.perf_counter()
        s.skip(30000)                      # 50 minutes of signal: 100 evaluated epochs + 100 refresh passes
        dt_skip = time.perf_counter() - t
    with hostapi.Scenario(NAV12, llh=LLH, sample_rate=2600000, threads=1) as s:
        t = time.perf_counter()
        s.next(30000)
        dt_full = time.perf_counter() - t
    assert dt_skip * 10 < dt_full, (dt_skip, dt_full)


@pytest.mark.parametrize("world,E", [(2, 128), (3, 100), (8, 37)])
def test_slice_feeders_of_all_ranks_reassemble_the_single_stream(world, E):
    """Every rank generates only its own time slices (the others are skipped): put back in stream order they are the
    descriptors of one uninterrupted run, RESET_CARRIER flags included -- each (re)allocation is announced exactly
    once, by the rank whose slice holds the first epoch after it."""
    # static scenario started at 00:17:10: PRN rises and is allocated by the refresh at 00:18:00 -> RESET flag on epoch 500,
    # which is inside a slice for (2, 128) and (8, 37) and the FIRST epoch of a slice -- right after a skipped span -- for (3, 100)
    kw = dict(nav=NAV12, llh=LLH, sample_rate=2600000, start=(2014, 12, 20, 0, 17, 10.0))
    steps = 1300 // (world * E)
    with hostapi.Scenario(**kw) as s:
        full = s.next(steps * world * E)
    assert sorted(set(np.nonzero(full["flags"] & 1)[0].tolist())) == [0, 500]
    feeders = [hostapi.SliceFeeder(r, world, E, **kw) for r in range(world)]
    got = np.concatenate([feeders[r].next_slice() for _ in range(steps) for r in range(world)])
    assert got.tobytes() == full.tobytes()
    for f in feeders:
        f.close()


@pytest.mark.parametrize("args,fixture", [
    (["--kind", "gps"], "brdc3540_synth.14n.gz"),
    (["--kind", "allsky"], "allsky32_synth.14n.gz"),
    (["--kind", "gps", "--rinex3"], "brdc3540_synth.14p.gz"),
])
def test_fixture_generator_reproduces_the_committed_navigation_files(tmp_path, args, fixture):
    """tools/gen_rinex_fixture.py (the synthetic-constellation generator) is deterministic:
    it rewrites the committed fixtures byte for byte, so every golden's `nav_fixture_sha256` can be re-derived."""
    import hashlib
    import subprocess
    import sys
    out = tmp_path / fixture
    subprocess.run([sys.executable, os.path.join(ol.REPO, "tools", "gen_rinex_fixture.py")] + args + ["-o", str(out)],
                   check=True, capture_output=True)
    want = open(os.path.join(ol.GOLDEN, fixture), "rb").read()
    assert out.read_bytes() == want
    if fixture == "brdc3540_synth.14n.gz":
        assert hashlib.sha256(want).hexdigest() == ol.load_golden_meta("static12")["nav_fixture_sha256"]
