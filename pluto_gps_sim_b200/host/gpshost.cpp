// gpshost.cpp -- host orchestrator: navigation file -> per-epoch channel descriptors.
//
// Everything here runs once per 0.1 s epoch per satellite (microseconds); it exists so that a
// production run does not need the reference at all.  The numbers it produces feed index
// computations in the sample kernels, where one ulp flips a chip, so every floating-point
// expression is evaluated in the order the reference evaluates it (cited per function), with the
// same glibc libm calls, compiled with -ffp-contract=off.  The code is organised differently
// (one Scenario object, table-driven record parsing, a satellite-state struct), the arithmetic
// is the reference's.  See include/gpshost.h for the map of what replaces what.
#include "../../include/gpshost.h"
#include "../../include/gpsiq_desc.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- constants (plutogpssim.h:40-76; IS-GPS-200 values the reference uses) ---------------
constexpr double kWeek = 604800.0, kHalfWeek = 302400.0, kDay = 86400.0, kHour = 3600.0, kMinute = 60.0;
constexpr double kGM = 3.986005e14, kOmegaE = 7.2921151467e-5, kPi = 3.1415926535898;
constexpr double kEarthA = 6378137.0, kEarthE = 0.0818191908426, kRad2Deg = 57.2957795131;
constexpr double kC = 2.99792458e8, kLambda = 0.190293672798365;
constexpr double kCodeHz = 1.023e6, kCarrToCode = 1.0 / 1540.0;
constexpr double k2m5 = 0.03125, k2m19 = 1.907348632812500e-6, k2m29 = 1.862645149230957e-9,
                 k2m31 = 4.656612873077393e-10, k2m33 = 1.164153218269348e-10, k2m43 = 1.136868377216160e-13,
                 k2m55 = 2.775557561562891e-17, k2m50 = 8.881784197001252e-016, k2m30 = 9.313225746154785e-010,
                 k2m27 = 7.450580596923828e-009, k2m24 = 5.960464477539063e-008;
constexpr int kMaxSv = 32, kSets = 13, kWordsPerSf = 10, kNavWords = 60, kMotionMax = 3000;

// receiver antenna attenuation [dB] per 5 degrees of boresight angle (plutogpssim.c:164-169)
const double kAntennaDb[37] = {0.00,  0.00,  0.22,  0.44,  0.67,  1.11,  1.56,  2.00,  2.44,  2.89,  3.56,  4.22,  4.89,
                               5.56,  6.22,  6.89,  7.56,  8.22,  8.89,  9.78,  10.67, 11.56, 12.44, 13.33, 14.44, 15.56,
                               16.67, 17.78, 18.89, 20.00, 21.33, 22.67, 24.00, 25.56, 27.33, 29.33, 31.56};

struct Tow { int week = -1; double sec = 0.0; };
struct Cal { int y = 0, m = 0, d = 0, hh = 0, mm = 0; double sec = 0.0; };

struct Eph {  // one satellite's broadcast ephemeris + derived constants (plutogpssim.h:97-130)
    bool valid = false;
    Cal t; Tow toc, toe;
    int iodc = 0, iode = 0, health = 0, code_l2 = 0;
    double dn = 0, cuc = 0, cus = 0, cic = 0, cis = 0, crc = 0, crs = 0, ecc = 0, sqrta = 0, m0 = 0, omg0 = 0, inc0 = 0,
           aop = 0, omgdot = 0, idot = 0, af0 = 0, af1 = 0, af2 = 0, tgd = 0;
    double n = 0, sq1e2 = 0, A = 0, omgkdot = 0;
};

struct Klob {  // ionosphere + UTC header data (plutogpssim.h:132-140)
    bool enable = true, valid = false;
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, A0 = 0, A1 = 0;
    int dtls = 0, tot = 0, wnt = 0;
};

struct Sight { Tow g; double range = 0, rate = 0, dist = 0, az = 0, el = 0, iono = 0; };

struct Slot {  // one channel (the part of channel_t the host owns)
    int prn = 0;
    double f_carr = 0, f_code = 0, code_phase = 0, carr_phase0 = 0;
    Tow g0;
    uint32_t sf[5][kWordsPerSf];
    uint64_t words[kNavWords];
    int iword = 0, ibit = 0, icode = 0;
    double az = 0, el = 0;
    Sight rho0;
    bool fresh = false;  // (re)allocated since the last emitted epoch
};

thread_local std::string g_error;  // per calling thread: scenarios may be driven from different threads

// ---- time (plutogpssim.c:250-290, 838-866) ------------------------------------------------
bool cal_ok(const Cal& t) {  // what cal_to_tow can take (a malformed record line must not index the month table)
    return t.m >= 1 && t.m <= 12 && t.d >= 0 && t.d <= 31 && t.hh >= 0 && t.hh < 25 && t.mm >= 0 && t.mm < 61 && t.sec >= 0.0 &&
           t.sec < 62.0 && t.y >= 1980 && t.y < 2200;
}

Tow cal_to_tow(const Cal& t) {
    static const int doy[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    const int ye = t.y - 1980;
    int leap = ye / 4 + 1;
    if ((ye % 4) == 0 && t.m <= 2) leap--;
    const int de = ye * 365 + doy[t.m - 1] + t.d + leap - 6;
    Tow g;
    g.week = de / 7;
    g.sec = (double) (de % 7) * kDay + t.hh * kHour + t.mm * kMinute + t.sec;
    return g;
}

Cal tow_to_cal(const Tow& g) {
    const int c = (int) (7 * g.week + floor(g.sec / 86400.0) + 2444245.0) + 1537;
    const int d = (int) ((c - 122.1) / 365.25);
    const int e = 365 * d + d / 4;
    const int f = (int) ((c - e) / 30.6001);
    Cal t;
    t.d = c - e - (int) (30.6001 * f);
    t.m = f - 1 - 12 * (f / 14);
    t.y = d - 4715 - ((7 + t.m) / 10);
    t.hh = ((int) (g.sec / 3600.0)) % 24;
    t.mm = ((int) (g.sec / 60.0)) % 60;
    t.sec = g.sec - 60.0 * floor(g.sec / 60.0);
    return t;
}

double tow_diff(const Tow& a, const Tow& b) {
    double dt = a.sec - b.sec;
    dt += (double) (a.week - b.week) * kWeek;
    return dt;
}

Tow tow_add(const Tow& g0, double dt) {  // rounds to the millisecond, like the reference
    Tow g;
    g.week = g0.week;
    g.sec = g0.sec + dt;
    g.sec = round(g.sec * 1000.0) / 1000.0;
    while (g.sec >= kWeek) { g.sec -= kWeek; g.week++; }
    while (g.sec < 0.0) { g.sec += kWeek; g.week--; }
    return g;
}

// ---- geodesy (plutogpssim.c:296-434) -------------------------------------------------------
double norm3(const double* x) { return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }

void ecef_to_geodetic(const double* xyz, double* llh) {
    const double a = kEarthA, e = kEarthE, eps = 1.0e-3, e2 = e * e;
    if (norm3(xyz) < eps) { llh[0] = 0.0; llh[1] = 0.0; llh[2] = -a; return; }
    const double x = xyz[0], y = xyz[1], z = xyz[2];
    const double rho2 = x * x + y * y;
    double dz = e2 * z, zdz, nh, n;
    for (;;) {
        zdz = z + dz;
        nh = sqrt(rho2 + zdz * zdz);
        const double slat = zdz / nh;
        n = a / sqrt(1.0 - e2 * slat * slat);
        const double dz_new = n * e2 * slat;
        if (fabs(dz - dz_new) < eps) break;
        dz = dz_new;
    }
    llh[0] = atan2(zdz, sqrt(rho2));
    llh[1] = atan2(y, x);
    llh[2] = nh - n;
}

void geodetic_to_ecef(const double* llh, double* xyz) {
    const double a = kEarthA, e = kEarthE, e2 = e * e;
    const double clat = cos(llh[0]), slat = sin(llh[0]), clon = cos(llh[1]), slon = sin(llh[1]);
    const double d = e * slat;
    const double n = a / sqrt(1.0 - d * d);
    const double nph = n + llh[2];
    const double tmp = nph * clat;
    xyz[0] = tmp * clon;
    xyz[1] = tmp * slon;
    xyz[2] = ((1.0 - e2) * n + llh[2]) * slat;
}

struct LocalFrame {  // rows: north, east, up (plutogpssim.c:374-394)
    double r[3][3];
    explicit LocalFrame(const double* llh) {
        const double slat = sin(llh[0]), clat = cos(llh[0]), slon = sin(llh[1]), clon = cos(llh[1]);
        r[0][0] = -slat * clon; r[0][1] = -slat * slon; r[0][2] = clat;
        r[1][0] = -slon;        r[1][1] = clon;         r[1][2] = 0.0;
        r[2][0] = clat * clon;  r[2][1] = clat * slon;  r[2][2] = slat;
    }
    void az_el(const double* v, double& az, double& el) const {  // ecef2neu + neu2azel
        const double n = r[0][0] * v[0] + r[0][1] * v[1] + r[0][2] * v[2];
        const double e = r[1][0] * v[0] + r[1][1] * v[1] + r[1][2] * v[2];
        const double u = r[2][0] * v[0] + r[2][1] * v[1] + r[2][2] * v[2];
        az = atan2(e, n);
        if (az < 0.0) az += (2.0 * kPi);
        const double ne = sqrt(n * n + e * e);
        el = atan2(u, ne);
    }
};

// ---- satellite state (plutogpssim.c:443-546; IS-GPS-200 table 20-IV) --------------------------
struct SvState { double pos[3], vel[3], clk[2]; };

SvState sv_state(const Eph& eph, const Tow& g) {
    SvState s;
    double tk = g.sec - eph.toe.sec;
    if (tk > kHalfWeek) tk -= kWeek;
    else if (tk < -kHalfWeek) tk += kWeek;

    const double mk = eph.m0 + eph.n * tk;
    double ek = mk, ekold = ek + 1.0, one_m_ecosE = 0;
    // Kepler's equation, Newton steps to the reference's tolerance (plutogpssim.c:470-476).  GPS orbits (e < 0.03) are
    // there after <= 5 steps; the cap only ends the loop for a corrupted eccentricity, where the reference spins forever.
    for (int it = 0; it < 64 && fabs(ek - ekold) > 1.0E-14; it++) {
        ekold = ek;
        one_m_ecosE = 1.0 - eph.ecc * cos(ekold);
        ek = ek + (mk - ekold + eph.ecc * sin(ekold)) / one_m_ecosE;
    }
    const double sek = sin(ek), cek = cos(ek);
    const double ekdot = eph.n / one_m_ecosE;
    const double relativistic = -4.442807633E-10 * eph.ecc * eph.sqrta * sek;

    const double pk = atan2(eph.sq1e2 * sek, cek - eph.ecc) + eph.aop;
    const double pkdot = eph.sq1e2 * ekdot / one_m_ecosE;
    const double s2pk = sin(2.0 * pk), c2pk = cos(2.0 * pk);

    const double uk = pk + eph.cus * s2pk + eph.cuc * c2pk;
    const double suk = sin(uk), cuk = cos(uk);
    const double ukdot = pkdot * (1.0 + 2.0 * (eph.cus * c2pk - eph.cuc * s2pk));

    const double rk = eph.A * one_m_ecosE + eph.crc * c2pk + eph.crs * s2pk;
    const double rkdot = eph.A * eph.ecc * sek * ekdot + 2.0 * pkdot * (eph.crs * c2pk - eph.crc * s2pk);

    const double ik = eph.inc0 + eph.idot * tk + eph.cic * c2pk + eph.cis * s2pk;
    const double sik = sin(ik), cik = cos(ik);
    const double ikdot = eph.idot + 2.0 * pkdot * (eph.cis * c2pk - eph.cic * s2pk);

    const double xpk = rk * cuk, ypk = rk * suk;
    const double xpkdot = rkdot * cuk - ypk * ukdot;
    const double ypkdot = rkdot * suk + xpk * ukdot;

    const double ok = eph.omg0 + tk * eph.omgkdot - kOmegaE * eph.toe.sec;
    const double sok = sin(ok), cok = cos(ok);

    s.pos[0] = xpk * cok - ypk * cik * sok;
    s.pos[1] = xpk * sok + ypk * cik * cok;
    s.pos[2] = ypk * sik;

    const double tmp = ypkdot * cik - ypk * sik * ikdot;
    s.vel[0] = -eph.omgkdot * s.pos[1] + xpkdot * cok - tmp * sok;
    s.vel[1] = eph.omgkdot * s.pos[0] + xpkdot * sok + tmp * cok;
    s.vel[2] = ypk * cik * ikdot + ypkdot * sik;

    tk = g.sec - eph.toc.sec;
    if (tk > kHalfWeek) tk -= kWeek;
    else if (tk < -kHalfWeek) tk += kWeek;
    s.clk[0] = eph.af0 + tk * (eph.af1 + tk * eph.af2) + relativistic - eph.tgd;
    s.clk[1] = eph.af1 + 2.0 * tk * eph.af2;
    return s;
}

// ---- Klobuchar ionospheric delay [m] (plutogpssim.c:1612-1683; IS-GPS-200 20.3.3.5.2.5) ----------
double klobuchar(const Klob& k, const Tow& g, const double* llh, double az, double el) {
    if (!k.enable) return 0.0;
    const double E = el / kPi, phi_u = llh[0] / kPi, lam_u = llh[1] / kPi;
    const double F = 1.0 + 16.0 * pow((0.53 - E), 3.0);
    if (!k.valid) return F * 5.0e-9 * kC;
    const double psi = 0.0137 / (E + 0.11) - 0.022;
    double phi_i = phi_u + psi * cos(az);
    if (phi_i > 0.416) phi_i = 0.416;
    else if (phi_i < -0.416) phi_i = -0.416;
    const double lam_i = lam_u + psi * sin(az) / cos(phi_i * kPi);
    const double phi_m = phi_i + 0.064 * cos((lam_i - 1.617) * kPi);
    const double phi_m2 = phi_m * phi_m, phi_m3 = phi_m2 * phi_m;
    double amp = k.a[0] + k.a[1] * phi_m + k.a[2] * phi_m2 + k.a[3] * phi_m3;
    if (amp < 0.0) amp = 0.0;
    double per = k.b[0] + k.b[1] * phi_m + k.b[2] * phi_m2 + k.b[3] * phi_m3;
    if (per < 72000.0) per = 72000.0;
    double t = kDay / 2.0 * lam_i + g.sec;
    while (t >= kDay) t -= kDay;
    while (t < 0) t += kDay;
    const double X = 2.0 * kPi * (t - 50400.0) / per;
    if (fabs(X) < 1.57) {
        const double X2 = X * X, X4 = X2 * X2;
        return F * (5.0e-9 + amp * (1.0 - X2 / 2.0 + X4 / 24.0)) * kC;
    }
    return F * 5.0e-9 * kC;
}

// ---- pseudorange (plutogpssim.c:1691-1747): light time, Earth rotation, clock, ionosphere --------
Sight line_of_sight(const Eph& eph, const Klob& iono, const Tow& g, const double* xyz) {
    SvState s = sv_state(eph, g);
    double los[3] = {s.pos[0] - xyz[0], s.pos[1] - xyz[1], s.pos[2] - xyz[2]};
    const double tau = norm3(los) / kC;
    s.pos[0] -= s.vel[0] * tau;
    s.pos[1] -= s.vel[1] * tau;
    s.pos[2] -= s.vel[2] * tau;
    const double xrot = s.pos[0] + s.pos[1] * kOmegaE * tau;
    const double yrot = s.pos[1] - s.pos[0] * kOmegaE * tau;
    s.pos[0] = xrot;
    s.pos[1] = yrot;
    los[0] = s.pos[0] - xyz[0]; los[1] = s.pos[1] - xyz[1]; los[2] = s.pos[2] - xyz[2];
    Sight r;
    const double range = norm3(los);
    r.dist = range;
    r.range = range - kC * s.clk[0];
    r.rate = (s.vel[0] * los[0] + s.vel[1] * los[1] + s.vel[2] * los[2]) / range;
    r.g = g;
    double llh[3];
    ecef_to_geodetic(xyz, llh);
    LocalFrame(llh).az_el(los, r.az, r.el);
    r.iono = klobuchar(iono, g, llh, r.az, r.el);
    r.range += r.iono;
    return r;
}

bool above_horizon(const Eph& eph, const Tow& g, const double* xyz, double& az, double& el) {  // plutogpssim.c:1896-1916
    double llh[3];
    ecef_to_geodetic(xyz, llh);
    const SvState s = sv_state(eph, g);
    const double los[3] = {s.pos[0] - xyz[0], s.pos[1] - xyz[1], s.pos[2] - xyz[2]};
    LocalFrame(llh).az_el(los, az, el);
    return el * kRad2Deg > 0.0;  // the reference hard-wires a 0 degree mask (plutogpssim.c:1930)
}

// ---- navigation message ----------------------------------------------------------------------
uint32_t ones(uint32_t v) { return (uint32_t) __builtin_popcount(v); }

// (24,30) Hamming parity of IS-GPS-200 table 20-XIV, as plutogpssim.c:751-814: bits 31,30 in = D29*,D30*
uint32_t parity_word(uint32_t source, int nib) {
    static const uint32_t mask[6] = {0x3B1F3480u, 0x1D8F9A40u, 0x2EC7CD00u, 0x1763E680u, 0x2BB1F340u, 0x0B7A89C0u};
    uint32_t d = source & 0x3FFFFFC0u;
    const uint32_t D29 = (source >> 31) & 1u, D30 = (source >> 30) & 1u;
    if (nib) {  // words 2 and 10: solve bits 23, 24 so that D29 = D30 = 0
        if ((D30 + ones(mask[4] & d)) % 2) d ^= (1u << 6);
        if ((D29 + ones(mask[5] & d)) % 2) d ^= (1u << 7);
    }
    uint32_t D = d;
    if (D30) D ^= 0x3FFFFFC0u;
    D |= ((D29 + ones(mask[0] & d)) % 2) << 5;
    D |= ((D30 + ones(mask[1] & d)) % 2) << 4;
    D |= ((D29 + ones(mask[2] & d)) % 2) << 3;
    D |= ((D30 + ones(mask[3] & d)) % 2) << 2;
    D |= ((D30 + ones(mask[4] & d)) % 2) << 1;
    D |= ((D29 + ones(mask[5] & d)) % 2);
    return D & 0x3FFFFFFFu;
}

// Subframes 1-5 without TOW/WN/parity (plutogpssim.c:552-723).  Scaled integers are 64-bit like the
// reference's (unsigned) long; the masks keep what the message carries.
void build_subframes(const Eph& e, const Klob& k, uint32_t sf[5][kWordsPerSf]) {
    typedef unsigned long UL;
    const UL ura = 0, data_id = 1, sv_sf4_p25 = 63, sv_sf5_p25 = 51, sv_sf4_p18 = 56;
    const UL wn = 0;
    const UL toe = (UL) (e.toe.sec / 16.0), toc = (UL) (e.toc.sec / 16.0);
    const UL iode = (UL) e.iode, iodc = (UL) e.iodc;
    const long deltan = (long) (e.dn / k2m43 / kPi);
    const long cuc = (long) (e.cuc / k2m29), cus = (long) (e.cus / k2m29), cic = (long) (e.cic / k2m29),
               cis = (long) (e.cis / k2m29), crc = (long) (e.crc / k2m5), crs = (long) (e.crs / k2m5);
    const UL ecc = (UL) (e.ecc / k2m33), sqrta = (UL) (e.sqrta / k2m19);
    const long m0 = (long) (e.m0 / k2m31 / kPi), omg0 = (long) (e.omg0 / k2m31 / kPi), inc0 = (long) (e.inc0 / k2m31 / kPi),
               aop = (long) (e.aop / k2m31 / kPi), omgdot = (long) (e.omgdot / k2m43 / kPi),
               idot = (long) (e.idot / k2m43 / kPi);
    const long af0 = (long) (e.af0 / k2m31), af1 = (long) (e.af1 / k2m43), af2 = (long) (e.af2 / k2m55),
               tgd = (long) (e.tgd / k2m31);
    const UL health = (UL) e.health, code_l2 = (UL) e.code_l2;
    const UL wna = (UL) (e.toe.week % 256), toa = (UL) (e.toe.sec / 4096.0);
    const long alpha0 = (long) round(k.a[0] / k2m30), alpha1 = (long) round(k.a[1] / k2m27),
               alpha2 = (long) round(k.a[2] / k2m24), alpha3 = (long) round(k.a[3] / k2m24);
    const long beta0 = (long) round(k.b[0] / 2048.0), beta1 = (long) round(k.b[1] / 16384.0),
               beta2 = (long) round(k.b[2] / 65536.0), beta3 = (long) round(k.b[3] / 65536.0);
    const long A0 = (long) round(k.A0 / k2m30), A1 = (long) round(k.A1 / k2m50);
    const long dtls = (long) k.dtls;
    const UL tot = (UL) (k.tot / 4096), wnt = (UL) (k.wnt % 256);
    const UL wnlsf = 1929 % 256, dn = 7;  // leap-second fields are constants in the reference (plutogpssim.c:643-645)
    const long dtlsf = 18;
    const UL preamble = 0x8B0000UL << 6;
    UL w[5][kWordsPerSf];
    memset(w, 0, sizeof w);
    for (int i = 0; i < 5; i++) { w[i][0] = preamble; w[i][1] = (UL) (i + 1) << 8; }

    w[0][2] = ((wn & 0x3FFUL) << 20) | ((code_l2 & 0x3UL) << 18) | ((ura & 0xFUL) << 14) | ((health & 0x3FUL) << 8) |
              (((iodc >> 8) & 0x3UL) << 6);
    w[0][6] = (tgd & 0xFFUL) << 6;
    w[0][7] = ((iodc & 0xFFUL) << 22) | ((toc & 0xFFFFUL) << 6);
    w[0][8] = ((af2 & 0xFFUL) << 22) | ((af1 & 0xFFFFUL) << 6);
    w[0][9] = (af0 & 0x3FFFFFUL) << 8;

    w[1][2] = ((iode & 0xFFUL) << 22) | ((crs & 0xFFFFUL) << 6);
    w[1][3] = ((deltan & 0xFFFFUL) << 14) | (((m0 >> 24) & 0xFFUL) << 6);
    w[1][4] = (m0 & 0xFFFFFFUL) << 6;
    w[1][5] = ((cuc & 0xFFFFUL) << 14) | (((ecc >> 24) & 0xFFUL) << 6);
    w[1][6] = (ecc & 0xFFFFFFUL) << 6;
    w[1][7] = ((cus & 0xFFFFUL) << 14) | (((sqrta >> 24) & 0xFFUL) << 6);
    w[1][8] = (sqrta & 0xFFFFFFUL) << 6;
    w[1][9] = (toe & 0xFFFFUL) << 14;

    w[2][2] = ((cic & 0xFFFFUL) << 14) | (((omg0 >> 24) & 0xFFUL) << 6);
    w[2][3] = (omg0 & 0xFFFFFFUL) << 6;
    w[2][4] = ((cis & 0xFFFFUL) << 14) | (((inc0 >> 24) & 0xFFUL) << 6);
    w[2][5] = (inc0 & 0xFFFFFFUL) << 6;
    w[2][6] = ((crc & 0xFFFFUL) << 14) | (((aop >> 24) & 0xFFUL) << 6);
    w[2][7] = (aop & 0xFFFFFFUL) << 6;
    w[2][8] = (omgdot & 0xFFFFFFUL) << 6;
    w[2][9] = ((iode & 0xFFUL) << 22) | ((idot & 0x3FFFUL) << 8);

    if (k.valid) {  // subframe 4 page 18: ionosphere + UTC
        w[3][2] = (data_id << 28) | (sv_sf4_p18 << 22) | ((alpha0 & 0xFFUL) << 14) | ((alpha1 & 0xFFUL) << 6);
        w[3][3] = ((alpha2 & 0xFFUL) << 22) | ((alpha3 & 0xFFUL) << 14) | ((beta0 & 0xFFUL) << 6);
        w[3][4] = ((beta1 & 0xFFUL) << 22) | ((beta2 & 0xFFUL) << 14) | ((beta3 & 0xFFUL) << 6);
        w[3][5] = (A1 & 0xFFFFFFUL) << 6;
        w[3][6] = ((A0 >> 8) & 0xFFFFFFUL) << 6;
        w[3][7] = ((A0 & 0xFFUL) << 22) | ((tot & 0xFFUL) << 14) | ((wnt & 0xFFUL) << 6);
        w[3][8] = ((dtls & 0xFFUL) << 22) | ((wnlsf & 0xFFUL) << 14) | ((dn & 0xFFUL) << 6);
        w[3][9] = (dtlsf & 0xFFUL) << 22;
    } else {        // page 25 (empty)
        w[3][2] = (data_id << 28) | (sv_sf4_p25 << 22);
    }
    w[4][2] = (data_id << 28) | (sv_sf5_p25 << 22) | ((toa & 0xFFUL) << 14) | ((wna & 0xFFUL) << 6);
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < kWordsPerSf; j++) sf[i][j] = (uint32_t) w[i][j];  // the reference reads them back as 32-bit `unsigned`
}

// Six-subframe word buffer with TOW, WN and chained parity (plutogpssim.c:1820-1894).
bool build_nav_buffer(const Tow& g, Slot& ch, bool first) {
    Tow g0;
    g0.week = g.week;
    g0.sec = (double) (((unsigned long) (g.sec + 0.5)) / 30UL) * 30.0;  // frame (30 s) boundary
    ch.g0 = g0;
    const unsigned long wn = (unsigned long) (g0.week % 1024);
    unsigned long tow = ((unsigned long) g0.sec) / 6UL;
    unsigned long prev = 0;
    if (first) {  // start with subframe 5 of the previous frame
        for (int i = 0; i < kWordsPerSf; i++) {
            uint32_t w = ch.sf[4][i];
            if (i == 1) w |= (uint32_t) ((tow & 0x1FFFFUL) << 13);
            w |= (uint32_t) ((prev << 30) & 0xC0000000UL);
            ch.words[i] = parity_word(w, (i == 1 || i == 9) ? 1 : 0);
            prev = ch.words[i];
        }
    } else {      // keep the subframe 5 transmitted last
        for (int i = 0; i < kWordsPerSf; i++) {
            ch.words[i] = ch.words[kWordsPerSf * 5 + i];
            prev = ch.words[i];
        }
        if ((ch.words[1] & (0x1FFFFUL << 13)) != ((tow & 0x1FFFFUL) << 13)) return false;  // "Invalid TOW in subframe 5"
    }
    for (int s = 0; s < 5; s++) {
        tow++;
        for (int i = 0; i < kWordsPerSf; i++) {
            uint32_t w = ch.sf[s][i];
            if (s == 0 && i == 2) w |= (uint32_t) ((wn & 0x3FFUL) << 20);
            if (i == 1) w |= (uint32_t) ((tow & 0x1FFFFUL) << 13);
            w |= (uint32_t) ((prev << 30) & 0xC0000000UL);
            ch.words[(s + 1) * kWordsPerSf + i] = parity_word(w, (i == 1 || i == 9) ? 1 : 0);
            prev = ch.words[(s + 1) * kWordsPerSf + i];
        }
    }
    return true;
}

// ---- RINEX-2 navigation file (plutogpssim.c:874-1233) -----------------------------------------
// Fixed-column text; a field is cut out of the line buffer, 'D' exponents become 'E', strtod.  The line
// buffer persists between reads like the reference's (a short line leaves the previous line's tail behind).
struct NavReader {
    gzFile fp = nullptr;
    char line[100];
    NavReader() { memset(line, 0, sizeof line); }
    ~NavReader() { if (fp) gzclose(fp); }
    bool next() {
        if (!gzgets(fp, line, (int) sizeof line)) return false;
        const size_t n = strlen(line);
        memset(line + n, 0, sizeof line - n);  // fixed-column reads past a short line see blanks, not the previous line
        return true;
    }
    bool label(const char* s) const { return strncmp(line + 60, s, strlen(s)) == 0; }
    double num(int col, int width) const {
        char tmp[24];
        strncpy(tmp, line + col, (size_t) width);
        tmp[width] = 0;
        for (int i = 0; i < width && tmp[i]; i++)
            if (tmp[i] == 'D' || tmp[i] == 'd') tmp[i] = 'E';
        return atof(tmp);
    }
    int integer(int col, int width) const {
        char tmp[24];
        strncpy(tmp, line + col, (size_t) width);
        tmp[width] = 0;
        return atoi(tmp);
    }
};

// BROADCAST ORBIT 1-7 of one record (four 19-column fields per line from column `base`: 3 in RINEX 2,
// 4 in RINEX 3) and the derived constants (plutogpssim.c:1098-1222 / 1451-1602).  false: the file ended inside.
static bool read_orbit_lines(NavReader& r, Eph& e, int base) {
    const int c0 = base, c1 = base + 19, c2 = base + 38, c3 = base + 57;
    if (!r.next()) return false;
    e.iode = (int) r.num(c0, 19); e.crs = r.num(c1, 19); e.dn = r.num(c2, 19); e.m0 = r.num(c3, 19);
    if (!r.next()) return false;
    e.cuc = r.num(c0, 19); e.ecc = r.num(c1, 19); e.cus = r.num(c2, 19); e.sqrta = r.num(c3, 19);
    if (!r.next()) return false;
    e.toe.sec = r.num(c0, 19); e.cic = r.num(c1, 19); e.omg0 = r.num(c2, 19); e.cis = r.num(c3, 19);
    if (!r.next()) return false;
    e.inc0 = r.num(c0, 19); e.crc = r.num(c1, 19); e.aop = r.num(c2, 19); e.omgdot = r.num(c3, 19);
    if (!r.next()) return false;
    e.idot = r.num(c0, 19); e.code_l2 = (int) r.num(c1, 19); e.toe.week = (int) r.num(c2, 19);
    if (!r.next()) return false;
    e.health = (int) r.num(c1, 19);
    if (e.health > 0 && e.health < 32) e.health += 32;  // summary bit
    e.tgd = r.num(c2, 19); e.iodc = (int) r.num(c3, 19);
    if (!r.next()) return false;  // transmission time / fit interval: not used
    e.valid = true;
    e.A = e.sqrta * e.sqrta;
    e.n = sqrt(kGM / (e.A * e.A * e.A)) + e.dn;
    e.sq1e2 = sqrt(1.0 - e.ecc * e.ecc);
    e.omgkdot = e.omgdot - kOmegaE;
    return true;
}

// returns the number of ephemeris sets (a new set starts when TOC advances by more than one hour)
int load_rinex2(const char* path, std::vector<std::vector<Eph>>& sets, Klob& k, std::string& date) {
    NavReader r;
    r.fp = gzopen(path, "rt");
    if (!r.fp) return -1;
    sets.assign(kSets + 1, std::vector<Eph>(kMaxSv));  // (+1: the reference peeks at set ieph+1, plutogpssim.c:2777)
    int seen = 0;
    while (r.next()) {
        if (r.label("COMMENT")) continue;
        if (r.label("END OF HEADER")) break;
        if (r.label("RINEX VERSION / TYPE")) {
            if (r.num(0, 9) > 3.0) return -2;
            if (r.line[20] != 'N') return -3;
        } else if (r.label("PGM / RUN BY / DATE")) {
            date.assign(r.line + 40, strnlen(r.line + 40, 20));
        } else if (r.label("ION ALPHA")) {
            for (int i = 0; i < 4; i++) k.a[i] = r.num(2 + 12 * i, 12);
            seen |= 1;
        } else if (r.label("ION BETA")) {
            for (int i = 0; i < 4; i++) k.b[i] = r.num(2 + 12 * i, 12);
            seen |= 2;
        } else if (r.label("DELTA-UTC")) {
            k.A0 = r.num(3, 19);
            k.A1 = r.num(22, 19);
            k.tot = r.integer(41, 9);
            k.wnt = r.integer(50, 9);
            if (k.tot % 4096 == 0) seen |= 4;
        } else if (r.label("LEAP SECONDS")) {
            k.dtls = r.integer(0, 6);
            seen |= 8;
        }
    }
    k.valid = (seen == 0xF);

    Tow first;  // TOC that opened the current set
    int set = 0;
    while (r.next()) {
        const int sv = r.integer(0, 2) - 1;
        Cal t;
        t.y = r.integer(3, 2) + 2000;
        t.m = r.integer(6, 2);
        t.d = r.integer(9, 2);
        t.hh = r.integer(12, 2);
        t.mm = r.integer(15, 2);
        t.sec = r.num(18, 2);
        if (!cal_ok(t)) break;  // not a record line
        const Tow g = cal_to_tow(t);
        if (first.week == -1) first = g;
        if (tow_diff(g, first) > kHour) {
            first = g;
            if (++set >= kSets) break;
        }
        if (sv < 0 || sv >= kMaxSv) break;  // not a record line (the reference would index out of bounds here)
        Eph& e = sets[set][sv];
        e.t = t;
        e.toc = g;
        e.af0 = r.num(22, 19); e.af1 = r.num(41, 19); e.af2 = r.num(60, 19);
        if (!read_orbit_lines(r, e, 3)) break;
    }
    if (first.week >= 0) set += 1;
    return set;
}

// RINEX-3 navigation file (plutogpssim.c:1241-1610; the reference's -3 option): same record content, other
// columns -- a record starts "Gnn yyyy mm dd hh mm ss", fields sit one column further right, the header carries
// IONOSPHERIC CORR (GPSA / GPSB), TIME SYSTEM CORR (GPUT) and LEAP SECONDS; records of other constellations
// (first character not 'G') and their continuation lines are skipped.
int load_rinex3(const char* path, std::vector<std::vector<Eph>>& sets, Klob& k, std::string& date) {
    NavReader r;
    r.fp = gzopen(path, "rt");
    if (!r.fp) return -1;
    sets.assign(kSets + 1, std::vector<Eph>(kMaxSv));
    int seen = 0;
    while (r.next()) {
        if (r.label("COMMENT")) continue;
        if (r.label("END OF HEADER")) break;
        if (r.label("RINEX VERSION / TYPE")) {
            if (r.num(0, 9) < 3.0) return -2;
            if (r.line[20] != 'N' && r.line[40] != 'G') return -3;
        } else if (r.label("PGM / RUN BY / DATE")) {
            date.assign(r.line + 40, strnlen(r.line + 40, 20));
        } else if (r.label("IONOSPHERIC CORR")) {
            if (strncmp(r.line, "GPSA", 4) == 0) {
                for (int i = 0; i < 4; i++) k.a[i] = r.num(5 + 12 * i, 12);
                seen |= 1;
            } else if (strncmp(r.line, "GPSB", 4) == 0) {
                for (int i = 0; i < 4; i++) k.b[i] = r.num(5 + 12 * i, 12);
                seen |= 2;
            }
        } else if (r.label("TIME SYSTEM CORR") && strncmp(r.line, "GPUT", 4) == 0) {
            k.A0 = r.num(5, 17);
            k.A1 = r.num(22, 16);
            k.tot = (int) r.num(38, 7);   // (the reference converts a 'D' here too before atoi: same digits)
            k.wnt = r.integer(45, 6);
            if (k.tot % 4096 == 0) seen |= 4;
        } else if (r.label("LEAP SECONDS")) {
            k.dtls = r.integer(0, 6);
            seen |= 8;
        }
    }
    k.valid = (seen == 0xF);

    Tow first;
    int set = 0;
    while (r.next()) {
        if (r.line[0] != 'G') continue;
        const int sv = r.integer(1, 2) - 1;
        Cal t;
        t.y = r.integer(4, 4);
        t.m = r.integer(9, 2);
        t.d = r.integer(12, 2);
        t.hh = r.integer(15, 2);
        t.mm = r.integer(18, 2);
        t.sec = (double) r.integer(21, 2);
        if (!cal_ok(t)) break;  // not a record line
        const Tow g = cal_to_tow(t);
        if (first.week == -1) first = g;
        if (tow_diff(g, first) > kHour) {
            first = g;
            if (++set >= kSets) break;
        }
        if (sv < 0 || sv >= kMaxSv) break;  // not a GPS PRN (the reference would index out of bounds here)
        Eph& e = sets[set][sv];
        e.t = t;
        e.toc = g;
        e.af0 = r.num(23, 19); e.af1 = r.num(42, 19); e.af2 = r.num(61, 19);
        if (!read_orbit_lines(r, e, 4)) break;
    }
    if (first.week >= 0) set += 1;
    return set;
}

// User motion file: "t,x,y,z" rows at 10 Hz, ECEF metres; at most kMotionMax rows; the time column is ignored
// (readUserMotion, plutogpssim.c:1794-1818).  A row with fewer than four fields keeps the previous row's values for
// the missing ones (the reference's scratch variables live across rows); a blank line ends the file.
int load_motion(const char* path, std::vector<double>& xyz) {
    std::unique_ptr<FILE, int (*)(FILE*)> fp(fopen(path, "rt"), fclose);
    if (!fp) return -1;
    xyz.assign((size_t) kMotionMax * 3, 0.0);
    double row[4] = {0.0, 0.0, 0.0, 0.0};
    char text[100];
    int rows = 0;
    while (rows < kMotionMax && fgets(text, sizeof text, fp.get()) &&
           sscanf(text, "%lf,%lf,%lf,%lf", &row[0], &row[1], &row[2], &row[3]) != EOF) {
        std::copy(row + 1, row + 4, xyz.begin() + (size_t) rows * 3);
        rows++;
    }
    return rows;
}

}  // namespace

// ---- the scenario: what main() does around the sample loop (plutogpssim.c:2476-2806) --------------
struct gpshost_scenario {
    gpshost_config cfg;
    std::vector<std::vector<Eph>> sets;
    int nsets = 0, iset = -1;
    Klob iono;
    Klob iono_as_read;  // header values before a -T overwrite touches tot / wnt: what the reference's -v block shows
    std::string rinex_date;
    std::vector<double> motion;
    int nmotion = 0, imotion = 0;
    double xyz0[3] = {0, 0, 0};
    std::vector<Slot> chan;
    int owner[kMaxSv];   // slot of each allocated satellite, -1 = not allocated
    double ant[37];
    Tow g0, grx;
    Cal t0;
    double delt = 0;

    const double* position() const { return cfg.pos_mode == GPSHOST_POS_MOTION ? &motion[(size_t) imotion * 3] : xyz0; }

    // first free slot for every visible, unallocated satellite; release set satellites (plutogpssim.c:1918-1989)
    void allocate(const std::vector<Eph>& eph, const double* xyz) {
        const double origin[3] = {0.0, 0.0, 0.0};
        for (int sv = 0; sv < kMaxSv; sv++) {
            double az, el;
            if (eph[sv].valid && above_horizon(eph[sv], grx, xyz, az, el)) {
                if (owner[sv] != -1) continue;
                size_t i = 0;
                for (; i < chan.size(); i++) {
                    Slot& ch = chan[i];
                    if (ch.prn != 0) continue;
                    ch.prn = sv + 1;
                    ch.az = az;
                    ch.el = el;
                    build_subframes(eph[sv], iono, ch.sf);
                    build_nav_buffer(grx, ch, true);
                    ch.rho0 = line_of_sight(eph[sv], iono, grx, xyz);
                    const double r_xyz = ch.rho0.range;
                    const double r_ref = line_of_sight(eph[sv], iono, grx, origin).range;
                    const double phase_ini = (2.0 * r_ref - r_xyz) / kLambda;
                    if (cfg.carrier_mode == GPSIQ_CARRIER_FLOAT) {
                        ch.carr_phase0 = phase_ini - floor(phase_ini);
                    } else {
                        const double f = phase_ini - floor(phase_ini);
                        ch.carr_phase0 = (double) (unsigned int) (512.0 * 65536.0 * f);  // plutogpssim.c:1966-1967
                    }
                    ch.fresh = true;
                    break;
                }
                if (i < chan.size()) owner[sv] = (int) i;
            } else if (owner[sv] >= 0) {
                chan[(size_t) owner[sv]].prn = 0;
                owner[sv] = -1;
            }
        }
    }

    // per-epoch NCO set-up of one channel from the range at the epoch's end (plutogpssim.c:1754-1787)
    static void code_setup(Slot& ch, const Sight& rho1, double dt) {
        const double rhorate = (rho1.range - ch.rho0.range) / dt;
        ch.f_carr = -rhorate / kLambda;
        ch.f_code = kCodeHz + ch.f_carr * kCarrToCode;
        const double ms = ((tow_diff(ch.rho0.g, ch.g0) + 6.0) - ch.rho0.range / kC) * 1000.0;
        int ims = (int) ms;
        ch.code_phase = (ms - (double) ims) * 1023;
        ch.iword = ims / 600;
        ims -= ch.iword * 600;
        ch.ibit = ims / 20;
        ims -= ch.ibit * 20;
        ch.icode = ims;
        ch.rho0 = rho1;
    }

    int open() {
        g_error.clear();
        if (cfg.max_chan < 1 || cfg.max_chan > GPSIQ_MAX_CHAN || !cfg.nav_path || cfg.sample_rate < 1000000) {
            g_error = "bad configuration (max_chan 1..32, nav_path, sample_rate >= 1 MHz)";
            return GPSHOST_ERR_ARG;
        }
        delt = 1.0 / (double) cfg.sample_rate;
        iono.enable = !cfg.iono_disable;
        if (cfg.pos_mode == GPSHOST_POS_MOTION) {
            nmotion = cfg.motion_path ? load_motion(cfg.motion_path, motion) : -1;
            if (nmotion <= 0) {  // plutogpssim.c:2407-2413
                g_error = nmotion < 0 ? "Failed to open user motion file." : "Failed to read user motion data.";
                return GPSHOST_ERR_MOTION;
            }
        } else if (cfg.pos_mode == GPSHOST_POS_LLH) {
            double llh[3] = {cfg.pos[0] / kRad2Deg, cfg.pos[1] / kRad2Deg, cfg.pos[2]};
            geodetic_to_ecef(llh, xyz0);
        } else {
            memcpy(xyz0, cfg.pos, sizeof xyz0);
        }
        nsets = cfg.rinex3 ? load_rinex3(cfg.nav_path, sets, iono, rinex_date)
                                : load_rinex2(cfg.nav_path, sets, iono, rinex_date);
        if (nsets < 0) { g_error = cfg.rinex3 ? "cannot read RINEX-3 navigation file" : "cannot read RINEX-2 navigation file"; return GPSHOST_ERR_NAVFILE; }
        if (nsets == 0) { g_error = "No ephemeris available."; return GPSHOST_ERR_NOEPH; }
        iono_as_read = iono;

        // span of the file, start time, optional TOC/TOE overwrite (plutogpssim.c:2497-2574)
        Tow gmin, gmax;
        Cal tmin;
        gmax.week = 0;
        for (int sv = 0; sv < kMaxSv; sv++)
            if (sets[0][sv].valid) { gmin = sets[0][sv].toc; tmin = sets[0][sv].t; break; }
        for (int sv = 0; sv < kMaxSv; sv++)
            if (sets[(size_t) nsets - 1][sv].valid) { gmax = sets[(size_t) nsets - 1][sv].toc; break; }
        if (cfg.have_start) {
            t0.y = cfg.start[0]; t0.m = cfg.start[1]; t0.d = cfg.start[2]; t0.hh = cfg.start[3]; t0.mm = cfg.start[4];
            t0.sec = floor(cfg.start_sec);
            if (t0.y <= 1980 || t0.m < 1 || t0.m > 12 || t0.d < 1 || t0.d > 31 || t0.hh < 0 || t0.hh > 23 || t0.mm < 0 ||
                t0.mm > 59 || cfg.start_sec < 0.0 || cfg.start_sec >= 60.0) {
                g_error = "Invalid date and time.";
                return GPSHOST_ERR_TIME;
            }
            g0 = cal_to_tow(t0);
            if (cfg.time_overwrite) {
                Tow gtmp;
                gtmp.week = g0.week;
                gtmp.sec = (double) (((int) (g0.sec)) / 7200) * 7200.0;
                const double dsec = tow_diff(gtmp, gmin);
                iono.wnt = gtmp.week;
                iono.tot = (int) gtmp.sec;
                for (int sv = 0; sv < kMaxSv; sv++)
                    for (int i = 0; i < nsets; i++) {
                        Eph& e = sets[(size_t) i][sv];
                        if (!e.valid) continue;
                        e.toc = tow_add(e.toc, dsec);
                        e.t = tow_to_cal(e.toc);
                        e.toe = tow_add(e.toe, dsec);
                    }
            } else if (tow_diff(g0, gmin) < 0.0 || tow_diff(gmax, g0) < 0.0) {
                char msg[256];  // plutogpssim.c:2556-2563
                Cal tmax;
                for (int sv = 0; sv < kMaxSv; sv++)
                    if (sets[(size_t) nsets - 1][sv].valid) { tmax = sets[(size_t) nsets - 1][sv].t; break; }
                snprintf(msg, sizeof msg,
                         "Invalid start time.\ntmin = %4d/%02d/%02d,%02d:%02d:%02.0f (%d:%.0f)\ntmax = %4d/%02d/%02d,%02d:%02d:%02.0f (%d:%.0f)",
                         tmin.y, tmin.m, tmin.d, tmin.hh, tmin.mm, tmin.sec, gmin.week, gmin.sec, tmax.y, tmax.m, tmax.d, tmax.hh,
                         tmax.mm, tmax.sec, gmax.week, gmax.sec);
                g_error = msg;
                return GPSHOST_ERR_TIME;
            }
        } else {
            g0 = gmin;
            t0 = tmin;
        }
        for (int i = 0; i < nsets && iset < 0; i++)
            for (int sv = 0; sv < kMaxSv; sv++) {
                if (!sets[(size_t) i][sv].valid) continue;
                const double dt = tow_diff(g0, sets[(size_t) i][sv].toc);
                if (dt >= -kHour && dt < kHour) { iset = i; break; }
            }
        if (iset < 0) { g_error = "No current set of ephemerides has been found."; return GPSHOST_ERR_NOEPH; }

        chan.assign((size_t) cfg.max_chan, Slot());
        for (int sv = 0; sv < kMaxSv; sv++) owner[sv] = -1;
        grx = tow_add(g0, 0.0);
        const double* first_pos = cfg.pos_mode == GPSHOST_POS_MOTION ? &motion[0] : xyz0;
        allocate(sets[(size_t) iset], first_pos);
        for (int i = 0; i < 37; i++) ant[i] = pow(10.0, -kAntennaDb[i] / 20.0);
        grx = tow_add(grx, 0.1);
        return GPSHOST_OK;
    }

    // one pass of the reference's epoch loop minus the sample loop (plutogpssim.c:2655-2687, 2761-2805).
    // `sights`, if given, holds this epoch's pseudoranges per slot, computed ahead by run()'s worker threads.
    int epoch(gpsiq_chan_desc* out, const Sight* sights = nullptr) {
        const std::vector<Eph>& eph = sets[(size_t) iset];
        for (size_t i = 0; i < chan.size(); i++) {
            Slot& ch = chan[i];
            if (ch.prn <= 0) {
                memset(&out[i], 0, sizeof out[i]);
                continue;
            }
            const Sight rho = sights ? sights[i] : line_of_sight(eph[(size_t) ch.prn - 1], iono, grx, position());
            ch.az = rho.az;
            ch.el = rho.el;
            code_setup(ch, rho, 0.1);
            const double path_loss = 20200000.0 / rho.dist;
            const int ibs = (int) ((90.0 - rho.el * kRad2Deg) / 5.0);  // elevation -> boresight angle index
            if (ibs < 0 || ibs >= 37) {  // NaN / absurd geometry from a corrupted ephemeris (the reference reads out of bounds)
                g_error = "satellite geometry not finite (corrupted ephemeris?)";
                return GPSHOST_ERR_NOEPH;
            }
            const double gain = path_loss * ant[ibs];
            const int rc = gpsiq_make_desc_inline(&out[i], cfg.carrier_mode, ch.prn, ch.f_carr, ch.f_code, delt, ch.carr_phase0,
                                                  ch.code_phase, ch.words, ch.iword, ch.ibit, ch.icode, gain, ch.fresh);
            if (rc != GPSIQ_OK) { g_error = "descriptor out of range"; return GPSHOST_ERR_ARG; }
            ch.fresh = false;
        }
        end_of_epoch();
        return GPSHOST_OK;
    }

    // what follows the sample loop: the 30 s refresh when due, then time and motion index move on (plutogpssim.c:2761-2805)
    void end_of_epoch() {
        // every 30 s: next frame of NAV words, ephemeris roll-over, re-allocation
        const int igrx = (int) (grx.sec * 10.0 + 0.5);
        if (igrx % 300 == 0) {
            for (Slot& ch : chan)
                if (ch.prn > 0) build_nav_buffer(grx, ch, false);
            for (int sv = 0; sv < kMaxSv; sv++) {
                const Eph& nxt = sets[(size_t) iset + 1][sv];
                if (!nxt.valid) continue;
                if (tow_diff(nxt.toc, grx) < kHour) {
                    iset++;
                    for (Slot& ch : chan)
                        if (ch.prn != 0) build_subframes(sets[(size_t) iset][(size_t) ch.prn - 1], iono, ch.sf);
                }
                break;
            }
            allocate(sets[(size_t) iset], position());
        }
        grx = tow_add(grx, 0.1);
        if (++imotion >= nmotion) imotion = 0;
    }

    // Move n_epochs ahead without producing descriptors (the owner of a later time slice reaches its
    // first epoch without doing the earlier slices' work).  Between two refreshes the only state an epoch leaves behind
    // is each channel's previous pseudorange (code_setup's rho0), so only the LAST epoch of every refresh-to-refresh
    // segment is evaluated; the refresh passes themselves (NAV frame, ephemeris roll-over, re-allocation) are replayed.
    // A channel allocated inside the skipped span loses its RESET_CARRIER flag: its phase has been running since, and
    // the carrier state at a slice boundary comes from the previous slice's owner, not from the host.
    int skip(int n_epochs) {
        int left = n_epochs;
        while (left > 0) {
            // epochs up to and including the next refresh epoch (or all that is left)
            int seg = 0;
            Tow g = grx;
            Tow last = grx;
            int im = imotion, im_last = imotion;
            while (seg < left) {
                last = g;
                im_last = im;
                seg++;
                const bool refresh = ((int) (g.sec * 10.0 + 0.5)) % 300 == 0;
                g = tow_add(g, 0.1);
                if (++im >= nmotion) im = 0;
                if (refresh) break;
            }
            grx = last;          // stand on the segment's last epoch: its ranges become rho0, its refresh (if due) runs
            imotion = im_last;
            const std::vector<Eph>& eph = sets[(size_t) iset];
            for (Slot& ch : chan) {
                if (ch.prn <= 0) continue;
                const Sight rho = line_of_sight(eph[(size_t) ch.prn - 1], iono, grx, position());
                ch.az = rho.az;
                ch.el = rho.el;
                code_setup(ch, rho, 0.1);
                ch.fresh = false;
            }
            end_of_epoch();
            left -= seg;
        }
        return GPSHOST_OK;
    }

    // n_epochs epochs, multi-threaded over epochs.  The pseudorange of a channel at an epoch
    // (satellite state, light time, Earth rotation, ionosphere: most of the host work) depends only on the epoch's time,
    // the receiver position and the ephemeris -- not on the previous epoch.  Everything that IS sequential stays
    // sequential: the millisecond-rounded time steps, the range differences (code_setup), the NAV word counters and the
    // 30 s refresh.  So the stream is cut at the refresh epochs (the channel table and the ephemeris set are constant in
    // between), the times and positions of a segment are stepped through serially, worker threads fill the segment's
    // pseudoranges with the very same function calls, and one thread finishes the descriptors: bit-identical to the
    // serial path by construction (tests/test_host_orchestrator.py compares them anyway).
    int run(gpsiq_chan_desc* out, int n_epochs) {
        const size_t nch = chan.size();
        int workers = cfg.threads > 0 ? cfg.threads : 0;
        if (workers == 0) {
            const char* env = getenv("GPSHOST_THREADS");
            workers = env ? atoi(env) : (int) std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        }
        std::vector<Tow> when;
        std::vector<int> where;
        std::vector<Sight> sights;
        int done = 0;
        while (done < n_epochs) {
            // the segment: up to and including the next refresh epoch
            when.clear();
            where.clear();
            Tow g = grx;
            int im = imotion;
            while (done + (int) when.size() < n_epochs) {
                when.push_back(g);
                where.push_back(im);
                const bool refresh = ((int) (g.sec * 10.0 + 0.5)) % 300 == 0;
                g = tow_add(g, 0.1);
                if (++im >= nmotion) im = 0;
                if (refresh) break;
            }
            const int seg = (int) when.size();
            const int use = std::min(workers, seg / 16);   // below ~16 epochs per thread the spawn costs more than it saves
            if (use <= 1) {
                for (int j = 0; j < seg; j++)
                    if (int rc = epoch(out + (size_t) (done + j) * nch)) return rc;
                done += seg;
                continue;
            }
            sights.resize((size_t) seg * nch);
            const std::vector<Eph>& eph = sets[(size_t) iset];
            auto fill = [&](int j0, int j1) {
                for (int j = j0; j < j1; j++) {
                    const double* pos = cfg.pos_mode == GPSHOST_POS_MOTION ? &motion[(size_t) where[(size_t) j] * 3] : xyz0;
                    for (size_t i = 0; i < nch; i++)
                        if (chan[i].prn > 0)
                            sights[(size_t) j * nch + i] = line_of_sight(eph[(size_t) chan[i].prn - 1], iono, when[(size_t) j], pos);
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < use; t++) pool.emplace_back(fill, (int) ((long) seg * t / use), (int) ((long) seg * (t + 1) / use));
            fill(0, seg / use);
            for (std::thread& t : pool) t.join();
            for (int j = 0; j < seg; j++)
                if (int rc = epoch(out + (size_t) (done + j) * nch, &sights[(size_t) j * nch])) return rc;
            done += seg;
        }
        return GPSHOST_OK;
    }
};

extern "C" {

const char* gpshost_last_error(void) { return g_error.c_str(); }

int gpshost_open(gpshost_scenario** out, const gpshost_config* cfg) {
    if (!out || !cfg) return GPSHOST_ERR_ARG;
    *out = nullptr;
    gpshost_scenario* s = new gpshost_scenario();
    s->cfg = *cfg;
    const int rc = s->open();
    if (rc != GPSHOST_OK) { delete s; return rc; }
    *out = s;
    return GPSHOST_OK;
}

void gpshost_close(gpshost_scenario* s) { delete s; }

int gpshost_next(gpshost_scenario* s, gpsiq_chan_desc* desc, int n_epochs) {
    if (!s || !desc || n_epochs < 0) return GPSHOST_ERR_ARG;
    return s->run(desc, n_epochs);
}

int gpshost_skip(gpshost_scenario* s, int n_epochs) {
    if (!s || n_epochs < 0) return GPSHOST_ERR_ARG;
    return s->skip(n_epochs);
}

int gpshost_time(gpshost_scenario* s, int* week, double* sec) {
    if (!s) return GPSHOST_ERR_ARG;
    if (week) *week = s->grx.week;
    if (sec) *sec = s->grx.sec;
    return GPSHOST_OK;
}

int gpshost_describe(gpshost_scenario* s, char* buf, int buflen) {
    if (!s || !buf || buflen < 1) return GPSHOST_ERR_ARG;
    std::string o;
    char tmp[160];
    snprintf(tmp, sizeof tmp, "RINEX date = %s\nStart time = %4d/%02d/%02d,%02d:%02d:%02.0f (%d:%.0f)\n", s->rinex_date.c_str(),
             s->t0.y, s->t0.m, s->t0.d, s->t0.hh, s->t0.mm, s->t0.sec, s->g0.week, s->g0.sec);
    o += tmp;
    o += "PRN   Az    El     Range     Iono\n";
    for (const Slot& ch : s->chan)
        if (ch.prn > 0) {
            snprintf(tmp, sizeof tmp, "%02d %6.1f %5.1f %11.1f %5.1f\n", ch.prn, ch.az * kRad2Deg, ch.el * kRad2Deg, ch.rho0.dist,
                     ch.rho0.iono);
            o += tmp;
        }
    snprintf(buf, (size_t) buflen, "%s", o.c_str());
    return GPSHOST_OK;
}

int gpshost_describe_iono(gpshost_scenario* s, char* buf, int buflen) {
    if (!s || !buf || buflen < 1) return GPSHOST_ERR_ARG;
    buf[0] = 0;
    const Klob& k = s->iono_as_read;
    if (k.valid)  // the reference prints the block only for a complete header (plutogpssim.c:2487-2495)
        snprintf(buf, (size_t) buflen, "  %12.3e %12.3e %12.3e %12.3e\n  %12.3e %12.3e %12.3e %12.3e\n   %19.11e %19.11e  %9d %9d\n%6d\n",
                 k.a[0], k.a[1], k.a[2], k.a[3], k.b[0], k.b[1], k.b[2], k.b[3], k.A0, k.A1, k.tot, k.wnt, k.dtls);
    return GPSHOST_OK;
}

uint32_t gpshost_parity(uint32_t source, int nib) { return parity_word(source, nib); }

void gpshost_date2gps(int y, int m, int d, int hh, int mm, double sec, int* week, double* sow) {
    Cal t;
    t.y = y; t.m = m; t.d = d; t.hh = hh; t.mm = mm; t.sec = sec;
    const Tow g = cal_to_tow(t);
    if (week) *week = g.week;
    if (sow) *sow = g.sec;
}

void gpshost_llh2xyz(const double llh_rad[3], double xyz[3]) { geodetic_to_ecef(llh_rad, xyz); }
void gpshost_xyz2llh(const double xyz[3], double llh_rad[3]) { ecef_to_geodetic(xyz, llh_rad); }

}  // extern "C"
