#!/usr/bin/env python3
"""Synthesise RINEX-2 GPS navigation files in the fixed-column layout the
reference's ``readRinex2`` accepts (plutogpssim.c:874-1233).

``brdc3540.14n`` -- the file every BASELINE.json config names -- is not part of
the reference and cannot be fetched (no network), so the fixtures are generated
deterministically (no RNG) for the same date: 2014-12-20 00:00:00 = GPS week
1823, second 518400, twelve two-hourly ephemeris sets for 32 SVs.

  --kind gps     6 planes x (5|6) slots, GPS-like geometry: >= 12 SVs above the
                 horizon at the config location (30.286502 N, 120.032669 E);
                 the reference takes the first 12 in ascending PRN order.
  --kind allsky  "synthetic full constellation" of BASELINE configs [3]/[4]:
                 all 32 SVs placed so that they stay above the horizon of the
                 config location for the first ~20 minutes (32 live channels
                 when MAX_CHAN is 32).

The output is gzip (mtime 0, so byte-reproducible); zlib's gzopen -- which the
reference uses -- reads it transparently.
"""
import argparse
import gzip
import hashlib
import io
import math

GM = 3.986005e14
OMEGA_E = 7.2921151467e-5
WEEK = 1823
SOW0 = 518400          # 2014-12-20 00:00:00
A_SMA = 26559710.0
RX_LAT, RX_LON = 30.286502, 120.032669


def d19(x):
    """Fortran D19.12 as written by RINEX producers (19 chars)."""
    s = "% .12E" % x
    return s.replace("E", "D")


def d12(x):
    s = "% .4E" % x
    return s.replace("E", "D").rjust(12)


def header():
    def line(body, label):
        return body.ljust(60)[:60] + label.ljust(20) + "\n"
    h = ""
    h += line("     2.10           N: GPS NAV DATA", "RINEX VERSION / TYPE")
    h += line("b200-gps-iq synth   graft               20141221 000000 UTC", "PGM / RUN BY / DATE")
    h += line("synthetic constellation, see tools/gen_rinex_fixture.py", "COMMENT")
    h += line("  " + d12(1.1180e-08) + d12(2.2350e-08) + d12(-5.9600e-08) + d12(-1.1920e-07), "ION ALPHA")
    h += line("  " + d12(9.0110e+04) + d12(1.1470e+05) + d12(-6.5540e+04) + d12(-5.2430e+05), "ION BETA")
    h += line("   " + d19(0.0) + d19(0.0) + "%9d%9d" % (503808, 1824), "DELTA-UTC: A0,A1,T,W")
    h += line("%6d" % 16, "LEAP SECONDS")
    h += line("", "END OF HEADER")
    return h


def wrap_pi(x):
    return (x + math.pi) % (2.0 * math.pi) - math.pi


def sv_elements_gps(sv):
    plane = (sv - 1) % 6
    slot = (sv - 1) // 6
    return dict(
        ecc=0.002 + 0.0005 * sv,
        inc0=math.radians(55.0 + 0.1 * plane),
        omg0=math.radians(60.0 * plane + 10.0),
        m0=math.radians(60.0 * slot + 13.0 * plane + 7.0 * sv),
        aop=math.radians(30.0 + 5.0 * sv),
    )


def sv_elements_allsky(sv):
    """Place SV `sv` so that its sub-satellite point at SOW0 sits on a 4x8 grid
    around the receiver; half the SVs ascending, half descending, so that the
    Doppler spread is realistic (both signs)."""
    row = (sv - 1) // 8          # 0..3
    col = (sv - 1) % 8           # 0..7
    lat = math.radians(8.0 + 13.0 * row + 1.5 * (col % 3))       # 8..50 deg N
    lon = math.radians(RX_LON - 42.0 + 12.0 * col + 3.0 * (row % 2))
    inc = math.radians(55.0 + 0.05 * sv)
    ecc = 0.001 + 0.0003 * sv
    aop = math.radians(20.0 + 9.0 * sv)
    u = math.asin(min(1.0, math.sin(lat) / math.sin(inc)))       # ascending branch
    if sv % 2 == 0:
        u = math.pi - u                                          # descending branch
    dlon = math.atan2(math.cos(inc) * math.sin(u), math.cos(u))
    node_lon = lon - dlon
    omg0 = node_lon + OMEGA_E * SOW0                             # Omega_k = omg0 - w_e*toe at tk=0
    # true anomaly -> mean anomaly (small e)
    nu = u - aop
    E = 2.0 * math.atan2(math.sqrt(1 - ecc) * math.sin(nu / 2), math.sqrt(1 + ecc) * math.cos(nu / 2))
    m0 = E - ecc * math.sin(E)
    return dict(ecc=ecc, inc0=inc, omg0=wrap_pi(omg0), m0=wrap_pi(m0), aop=wrap_pi(aop))


def header3():
    """RINEX 3.02 header as readRinex3 reads it (plutogpssim.c:1263-1368)."""
    def line(body, label):
        return body.ljust(60)[:60] + label.ljust(20) + "\n"
    d17 = lambda x: ("% .10E" % x).replace("E", "D")
    d16 = lambda x: ("% .9E" % x).replace("E", "D")
    h = ""
    h += line("     3.02           N: GNSS NAV DATA    G: GPS", "RINEX VERSION / TYPE")
    h += line("b200-gps-iq synth   graft               20141221 000000 UTC", "PGM / RUN BY / DATE")
    h += line("synthetic constellation, see tools/gen_rinex_fixture.py", "COMMENT")
    h += line("GPSA " + d12(1.1180e-08) + d12(2.2350e-08) + d12(-5.9600e-08) + d12(-1.1920e-07), "IONOSPHERIC CORR")
    h += line("GPSB " + d12(9.0110e+04) + d12(1.1470e+05) + d12(-6.5540e+04) + d12(-5.2430e+05), "IONOSPHERIC CORR")
    h += line("GAUT " + d17(1.0e-9) + d16(2.0e-15) + "%7d%5d" % (345600, 1824), "TIME SYSTEM CORR")   # not GPS: skipped
    h += line("GPUT " + d17(0.0) + d16(0.0) + "%7d%5d" % (503808, 1824), "TIME SYSTEM CORR")
    h += line("%6d" % 16, "LEAP SECONDS")
    h += line("", "END OF HEADER")
    return h


def build(kind, rinex3=False):
    out = io.StringIO()
    out.write(header3() if rinex3 else header())
    n0 = math.sqrt(GM / A_SMA ** 3)
    omgdot = -8.0e-9
    for k in range(12):                       # two-hourly sets
        hh = 2 * k
        toe = SOW0 + 7200 * k
        for sv in range(1, 33):
            el = sv_elements_gps(sv) if kind == "gps" else sv_elements_allsky(sv)
            deltan = 4.5e-9 + 1.0e-11 * sv
            m0 = wrap_pi(el["m0"] + (n0 + deltan) * 7200.0 * k)
            omg0 = wrap_pi(el["omg0"] + omgdot * 7200.0 * k)
            af0 = 1.0e-5 * sv - 1.0e-4
            af1 = 1.0e-12 * sv
            iode = (sv + 8 * k) % 256
            l0 = "%2d %02d %2d %2d %2d %2d%5.1f" % (sv, 14, 12, 20, hh, 0, 0.0) + d19(af0) + d19(af1) + d19(0.0)
            rows = [
                (float(iode), 20.0 - 3.0 * (sv % 7), deltan, m0),                          # IODE Crs dn M0
                (1.0e-6 * (sv % 5 - 2), el["ecc"], 2.0e-6 + 1.0e-7 * (sv % 4), math.sqrt(A_SMA)),  # Cuc e Cus sqrtA
                (float(toe), 1.0e-7 * (sv % 3 - 1), omg0, -1.0e-7 * (sv % 4 - 1)),          # Toe Cic Omega0 Cis
                (el["inc0"], 200.0 + 5.0 * (sv % 9), el["aop"], omgdot),                   # i0 Crc omega OmegaDot
                (1.0e-10 * (sv % 6 - 3), 1.0, float(WEEK), 0.0),                           # IDOT codeL2 week L2P
                (2.0, 0.0, -1.0e-8 + 5.0e-10 * (sv % 8), float(iode)),                     # acc health TGD IODC
                (float(toe - 7200 + 18), 4.0, 0.0, 0.0),                                   # transmission time, fit
            ]
            if rinex3:
                if sv % 8 == 1:   # records of other constellations in between (3 resp. 7 continuation lines): skipped
                    out.write("R%02d %04d %02d %02d %02d %02d %02d" % (sv, 2014, 12, 20, hh, 15, 0) + d19(1e-5) + d19(0.0) + d19(518400.0) + "\n")
                    for _ in range(3):
                        out.write("    " + "".join(d19(v) for v in (1.0e4, -2.0, 0.0, 0.0)) + "\n")
                    out.write("E%02d %04d %02d %02d %02d %02d %02d" % (sv, 2014, 12, 20, hh, 10, 0) + d19(2e-5) + d19(0.0) + d19(0.0) + "\n")
                    for _ in range(7):
                        out.write("    " + "".join(d19(v) for v in (5.0, 6.0, 7.0, 8.0)) + "\n")
                l0 = "G%02d %04d %02d %02d %02d %02d %02d" % (sv, 2014, 12, 20, hh, 0, 0) + d19(af0) + d19(af1) + d19(0.0)
                out.write(l0 + "\n")
                for r in rows:
                    out.write("    " + "".join(d19(v) for v in r) + "\n")
                continue
            out.write(l0 + "\n")
            for r in rows:
                out.write("   " + "".join(d19(v) for v in r) + "\n")
    return out.getvalue()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", choices=["gps", "allsky"], default="gps")
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--rinex3", action="store_true", help="RINEX 3.02 layout (the reference's -3 / readRinex3), with "
                    "records of other constellations interleaved")
    a = ap.parse_args()
    text = build(a.kind, a.rinex3).encode("ascii")
    raw = io.BytesIO()
    with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, filename="") as g:
        g.write(text)
    data = raw.getvalue()
    with open(a.out, "wb") as f:
        f.write(data)
    print(a.out, len(text), "bytes text,", len(data), "bytes gz, sha256(text) =", hashlib.sha256(text).hexdigest())


if __name__ == "__main__":
    main()
