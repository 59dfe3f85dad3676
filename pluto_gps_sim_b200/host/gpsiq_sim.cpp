// gpsiq_sim -- command-line front end with the reference's option surface (plutogpssim.c:1991-2012,
// 2296-2390): navigation file + position/motion + start time -> interleaved int16 I/Q stream.
//
//   host orchestration  libgpshost.so  (descriptors, bit-identical to the reference's per-epoch state)
//   sample synthesis     libgpsiq.so    (sm_100a kernels behind the C-ABI; no CPU fallback)
//   sink                 file / stdout / none.  The reference pushes 300000-sample buffers to an ADALM-Pluto
//                        through libiio (plutogpssim.c:2146-2158); libiio is not available in this build, so
//                        the SDR options are accepted and reported as ignored, and the stream is delivered to
//                        a Sink in the same 300000-sample units.
//
// Differences from the reference, all deliberate: it stops after -d seconds (the reference runs until
// a signal arrives); -o/-b/-n are new; -f (FTP download) is refused (no network code here).
#include <getopt.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gpshost.h"
#include "../../include/gpsiq.h"

namespace {

constexpr int kSamplesPerEpoch = 300000;  // NUM_SAMPLES = TX_SAMPLE_FREQ/10, independent of -s (plutogpssim.c:43-44)

struct Sink {  // receives the stream in the reference's push units (one 0.1 s buffer = 300000 I/Q pairs)
    virtual ~Sink() {}
    virtual bool push(const int16_t* iq, size_t pairs) = 0;
};
struct NullSink : Sink {
    bool push(const int16_t*, size_t) override { return true; }
};
struct FileSink : Sink {
    FILE* fp;
    bool own;
    FileSink(FILE* f, bool o) : fp(f), own(o) {}
    ~FileSink() override { if (own && fp) fclose(fp); }
    bool push(const int16_t* iq, size_t pairs) override { return fwrite(iq, 4, pairs, fp) == pairs; }
};

void usage() {
    fprintf(stderr,
            "Usage: gpsiq_sim [options]\n"
            "Options (as pluto-gps-sim):\n"
            "  -e <gps_nav>     RINEX navigation file for GPS ephemerides (required)\n"
            "  -u <user_motion> User motion file (dynamic mode)\n"
            "  -c <location>    ECEF X,Y,Z in meters (static mode) e.g. 3967283.15,1022538.18,4872414.48\n"
            "  -l <location>    Lat,Lon,Hgt (static mode) e.g. 30.286502,120.032669,100\n"
            "  -t <date,time>   Scenario start time YYYY/MM/DD,hh:mm:ss\n"
            "  -T <date,time>   Overwrite TOC and TOE to scenario start time (use \"now\" for the current time)\n"
            "  -s <frequency>   Sampling frequency [Hz] (default: 3000000)\n"
            "  -i               Disable ionospheric delay for spacecraft scenario\n"
            "  -v               Show details about the simulated channels\n"
            "  -A/-B/-U/-N/-g   SDR options of the reference: accepted, ignored (no libiio in this build)\n"
            "Options of this build:\n"
            "  -d <seconds>     Duration [s] (default 1.0; the reference runs until interrupted)\n"
            "  -o <file>        Write the int16 I/Q stream to <file> (\"-\" = stdout; default: discard)\n"
            "  -b <epochs>      0.1 s epochs per GPU batch (default 128)\n"
            "  -n <channels>    Channel slots (default 12 = MAX_CHAN; up to 32)\n");
}

}  // namespace

int main(int argc, char** argv) {
    gpshost_config hc;
    memset(&hc, 0, sizeof hc);
    hc.pos_mode = GPSHOST_POS_LLH;
    hc.pos[0] = 35.681298; hc.pos[1] = 139.766247; hc.pos[2] = 10.0;  // Tokyo (plutogpssim.c:2266-2268)
    hc.sample_rate = 3000000;                                         // TX_SAMPLE_FREQ (plutogpssim.c:43, 2271)
    hc.max_chan = 12;
    hc.carrier_mode = GPSIQ_CARRIER_FLOAT;
    double duration = 1.0;
    int batch = 128;
    bool verbose = false, have_pos = false;
    const char* out_path = nullptr;
    std::string nav, motion;

    if (argc < 3) { usage(); return 1; }
    int opt;
    while ((opt = getopt(argc, argv, "e:3:u:g:c:l:s:T:t:A:B:U:N:vfi?d:o:b:n:")) != -1) {
        switch (opt) {
            case 'e': nav = optarg; break;
            case 'u': motion = optarg; hc.pos_mode = GPSHOST_POS_MOTION; have_pos = true; break;
            case '3': hc.rinex3 = 1; break;  // (takes and ignores an argument, like the reference's "3:" option string)
            case 'f': fprintf(stderr, "ERROR: FTP download is not available (no network code in this build).\n"); return 1;
            case 'c':
                if (sscanf(optarg, "%lf,%lf,%lf", &hc.pos[0], &hc.pos[1], &hc.pos[2]) != 3) { usage(); return 1; }
                hc.pos_mode = GPSHOST_POS_XYZ; have_pos = true;
                break;
            case 'l':
                if (sscanf(optarg, "%lf,%lf,%lf", &hc.pos[0], &hc.pos[1], &hc.pos[2]) != 3) { usage(); return 1; }
                hc.pos_mode = GPSHOST_POS_LLH; have_pos = true;
                break;
            case 's':
                hc.sample_rate = (long long) atoi(optarg);
                if (hc.sample_rate < 1000000) { fprintf(stderr, "ERROR: Invalid sampling frequency.\n"); return 1; }
                break;
            case 'T':
                hc.time_overwrite = 1;
                if (strncmp(optarg, "now", 3) == 0) {
                    time_t timer;
                    time(&timer);
                    const struct tm* gmt = gmtime(&timer);
                    hc.have_start = 1;
                    hc.start[0] = gmt->tm_year + 1900; hc.start[1] = gmt->tm_mon + 1; hc.start[2] = gmt->tm_mday;
                    hc.start[3] = gmt->tm_hour; hc.start[4] = gmt->tm_min; hc.start_sec = (double) gmt->tm_sec;
                }
                break;
            case 't':
                if (sscanf(optarg, "%d/%d/%d,%d:%d:%lf", &hc.start[0], &hc.start[1], &hc.start[2], &hc.start[3], &hc.start[4],
                           &hc.start_sec) != 6) {
                    fprintf(stderr, "ERROR: Invalid date and time.\n");
                    return 1;
                }
                hc.have_start = 1;
                break;
            case 'i': hc.iono_disable = 1; break;
            case 'v': verbose = true; break;
            case 'A': case 'B': case 'U': case 'N': case 'g':
                fprintf(stderr, "note: -%c %s ignored (SDR transport is not part of this build)\n", opt, optarg);
                break;
            case 'd': duration = atof(optarg); break;
            case 'o': out_path = optarg; break;
            case 'b': batch = atoi(optarg); break;
            case 'n': hc.max_chan = atoi(optarg); break;
            default: usage(); return 1;
        }
    }
    if (nav.empty()) { fprintf(stderr, "ERROR: GPS ephemeris file is not specified.\n"); return 1; }
    if (!have_pos) fprintf(stderr, "note: no -l/-c/-u given; using the default location (the reference leaves it uninitialised)\n");
    if (batch < 1 || duration <= 0.0) { usage(); return 1; }
    hc.nav_path = nav.c_str();
    hc.motion_path = motion.empty() ? nullptr : motion.c_str();
    fprintf(stderr, hc.pos_mode == GPSHOST_POS_MOTION ? "Using user motion mode.\n" : "Using static location mode.\n");

    gpshost_scenario* sc = nullptr;
    if (gpshost_open(&sc, &hc) != GPSHOST_OK) { fprintf(stderr, "ERROR: %s\n", gpshost_last_error()); return 1; }
    {
        char buf[4096];
        gpshost_describe(sc, buf, sizeof buf);
        fputs(buf, stderr);
    }

    std::unique_ptr<Sink> sink;
    if (!out_path) sink.reset(new NullSink());
    else if (strcmp(out_path, "-") == 0) sink.reset(new FileSink(stdout, false));
    else {
        FILE* fp = fopen(out_path, "wb");
        if (!fp) { fprintf(stderr, "ERROR: cannot open %s\n", out_path); return 1; }
        sink.reset(new FileSink(fp, true));
    }

    const long total_epochs = (long) (duration * 10.0 + 0.5);
    if (batch > total_epochs) batch = (int) total_epochs;
    gpsiq_config gc;
    memset(&gc, 0, sizeof gc);
    gc.device = 0; gc.max_chan = hc.max_chan; gc.samples_per_epoch = kSamplesPerEpoch;
    gc.carrier_mode = hc.carrier_mode; gc.max_epochs = batch;
    gpsiq_ctx* gq = nullptr;
    if (gpsiq_create(&gq, &gc) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_last_error(nullptr)); return 1; }

    // two batches in flight: descriptors of batch k+1 are generated and scanned while batch k is rendered
    const size_t desc_count = (size_t) batch * (size_t) hc.max_chan;
    gpsiq_chan_desc* desc = (gpsiq_chan_desc*) gpsiq_host_alloc(desc_count * sizeof(gpsiq_chan_desc));
    int16_t* iq[2] = {(int16_t*) gpsiq_host_alloc((size_t) batch * kSamplesPerEpoch * 4),
                      (int16_t*) gpsiq_host_alloc((size_t) batch * kSamplesPerEpoch * 4)};
    if (!desc || !iq[0] || !iq[1]) { fprintf(stderr, "ERROR: pinned host allocation failed\n"); return 1; }

    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<int> sizes;
    long produced = 0;
    auto submit_next = [&]() -> bool {
        const int n = (int) std::min<long>(batch, total_epochs - produced);
        if (n <= 0) return false;
        if (gpshost_next(sc, desc, n) != GPSHOST_OK) { fprintf(stderr, "ERROR: %s\n", gpshost_last_error()); exit(1); }
        if (gpsiq_submit(gq, desc, n) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_last_error(gq)); exit(1); }
        sizes.push_back(n);
        produced += n;
        return true;
    };
    submit_next();
    size_t k = 0;
    bool ok = true;
    while (k < sizes.size() && ok) {
        submit_next();                         // (no-op at the end of the stream)
        int16_t* buf = iq[k & 1];
        if (gpsiq_fetch(gq, buf) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_last_error(gq)); return 1; }
        for (int e = 0; e < sizes[k] && ok; e++)   // the reference's unit: one 300000-sample buffer per push
            ok = sink->push(buf + (size_t) e * kSamplesPerEpoch * 2, kSamplesPerEpoch);
        if (verbose) {
            int week; double sec;
            gpshost_time(sc, &week, &sec);
            fprintf(stderr, "\rTime into run = %4.1f", (double) (k + 1) * batch / 10.0);
        }
        k++;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    if (!ok) fprintf(stderr, "\nERROR: sink refused data\n");
    fprintf(stderr, "%s%ld epochs (%.1f s of signal, %.0f samples) in %.3f s: %.1f Msamples/s, %lld kernel launches\n",
            verbose ? "\n" : "", produced, produced / 10.0, (double) produced * kSamplesPerEpoch, secs,
            (double) produced * kSamplesPerEpoch / secs / 1e6, (long long) gpsiq_launch_count(gq));
    gpsiq_host_free(desc); gpsiq_host_free(iq[0]); gpsiq_host_free(iq[1]);
    gpsiq_destroy(gq);
    gpshost_close(sc);
    return ok ? 0 : 1;
}
