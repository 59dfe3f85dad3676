// gpsiq_sim -- command-line front end with the reference's option surface (plutogpssim.c:1991-2012,
// 2296-2390): navigation file + position/motion + start time -> interleaved int16 I/Q stream.
//
//   host orchestration  libgpshost.so  (descriptors, bit-identical to the reference's per-epoch state)
//   sample synthesis     libgpsiq.so    (sm_100a kernels behind the C-ABI; no CPU fallback)
//   sink                 libgpshost.so  (include/gpssink.h): none / file / stdout / ADALM-Pluto.  The reference pushes
//                        300000-sample buffers to the SDR through libiio (plutogpssim.c:2146-2158); -r selects
//                        the same transport here (libiio is dlopen()ed at that point; without it the open
//                        fails loudly).  Every sink gets the stream in the same 300000-sample units, from
//                        its own writer thread: batch k is drained while batch k+1 is fetched from the GPU.
//
// Differences from the reference, all deliberate: it stops after -d seconds by default (-d 0 runs until
// a signal arrives, as the reference does); -o/-b/-n/-j/-r are new (the radio is one sink among others, not the only one); -f (FTP
// download) is refused (no network code here).
#include <getopt.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <ctime>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gpshost.h"
#include "../../include/gpsiq.h"
#include "../../include/gpssink.h"

namespace {

constexpr int kSamplesPerEpoch = 300000;  // NUM_SAMPLES = TX_SAMPLE_FREQ/10, independent of -s (plutogpssim.c:43-44)

// SIGINT / SIGTERM / SIGQUIT (the three the reference handles, plutogpssim.c:2282-2284) end the run: no new batch is
// submitted, the sink is told to stop after the push unit it is in (0.1 s of signal) and to discard what is queued,
// and it is then closed in order -- for the radio: TX LO off before the context goes away (the reference's handle_sig +
// exit path, plutogpssim.c:2014-2022, 2160-2178).  A second signal skips the orderly GPU drain as well.
volatile sig_atomic_t g_stop = 0;
void on_signal(int) { g_stop = g_stop + 1; }

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
            "  -A <attenuation> Set TX attenuation [dB] (default -20.0)          (used with -r)\n"
            "  -B <bw>          Set RF bandwidth [MHz] (default 5.0)             (used with -r)\n"
            "  -U <uri>         ADALM-Pluto URI                                  (used with -r)\n"
            "  -N <network>     ADALM-Pluto network IP or hostname (default pluto.local)\n"
            "Options of this build:\n"
            "  -d <seconds>     Duration [s] (default 1.0); 0 = until SIGINT/SIGTERM, as the reference runs\n"
            "  -o <file>        Write the int16 I/Q stream to <file> (\"-\" = stdout; default: discard)\n"
            "  -r               Transmit through an ADALM-Pluto (libiio), as the reference does\n"
            "  -b <epochs>      0.1 s epochs per GPU batch (default 128)\n"
            "  -n <channels>    Channel slots (default 12 = MAX_CHAN; up to 32)\n"
            "  -j <threads>     Host threads for the per-epoch orbit/range work (default min(8, cores); 1 = serial)\n"
            "  -g <gpus>        GPUs of this node to spread the stream over, in time slices of one batch (default 1)\n");
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
    int batch = 128, gpus = 1;
    bool verbose = false, have_pos = false, use_radio = false, batch_given = false, moving = false;
    gpssink_radio_config radio;
    gpssink_radio_defaults(&radio);
    const char* out_path = nullptr;
    std::string nav, motion;

    if (argc < 3) { usage(); return 1; }
    int opt;
    while ((opt = getopt(argc, argv, "e:3:u:g:c:l:s:T:t:A:B:U:N:vfi?d:o:b:n:rj:")) != -1) {
        switch (opt) {
            case 'e': nav = optarg; break;
            // -u clears the reference's staticLocationMode and -l / -c never set it again (plutogpssim.c:2301-2318, 2403):
            // user motion wins whatever the option order; -l / -c then only store coordinates nobody reads
            case 'u': motion = optarg; moving = true; have_pos = true; break;
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
                gpssink_radio_option(&radio, 's', optarg);
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
                           &hc.start_sec) != 6 ||
                    hc.start[0] <= 1980 || hc.start[1] < 1 || hc.start[1] > 12 || hc.start[2] < 1 || hc.start[2] > 31 ||
                    hc.start[3] < 0 || hc.start[3] > 23 || hc.start[4] < 0 || hc.start[4] > 59 || hc.start_sec < 0.0 ||
                    hc.start_sec >= 60.0) {  // refused while the options are read, like the reference (plutogpssim.c:2352-2357)
                    fprintf(stderr, "ERROR: Invalid date and time.\n");
                    return 1;
                }
                hc.have_start = 1;
                break;
            case 'i': hc.iono_disable = 1; break;
            case 'v': verbose = true; break;
            case 'A': case 'B': case 'U': case 'N': gpssink_radio_option(&radio, opt, optarg); break;
            // 'g' is in the reference's option string and handled nowhere there (plutogpssim.c:2296); here: GPU count
            case 'g': gpus = atoi(optarg); break;
            case 'r': use_radio = true; break;
            case 'd': duration = atof(optarg); break;
            case 'o': out_path = optarg; break;
            case 'b': batch = atoi(optarg); batch_given = true; break;
            case 'n': hc.max_chan = atoi(optarg); break;
            case 'j': hc.threads = atoi(optarg); break;
            default: usage(); return 1;
        }
    }
    if (nav.empty()) { fprintf(stderr, "ERROR: GPS ephemeris file is not specified.\n"); return 1; }
    if (!have_pos) fprintf(stderr, "note: no -l/-c/-u given; using the default location (the reference leaves it uninitialised)\n");
    if (moving) hc.pos_mode = GPSHOST_POS_MOTION;
    // a transmitting run is paced by the radio (0.1 s per epoch): small batches keep what is in flight -- and so the time
    // a stop request takes -- short; a batch the user asked for is respected up to 10 epochs (1 s)
    if (use_radio) batch = batch_given ? std::min(batch, 10) : 5;
    if (batch < 1 || duration < 0.0 || gpus < 1 || gpus > 16) { usage(); return 1; }
    hc.nav_path = nav.c_str();
    hc.motion_path = motion.empty() ? nullptr : motion.c_str();

    // the mode line comes after the motion file has been read and before the navigation file is (plutogpssim.c:2402-2420)
    if (!moving) fprintf(stderr, "Using static location mode.\n");
    gpshost_scenario* sc = nullptr;
    const int opened = gpshost_open(&sc, &hc);
    if (moving && opened != GPSHOST_ERR_MOTION) fprintf(stderr, "Using user motion mode.\n");
    if (opened != GPSHOST_OK) { fprintf(stderr, "ERROR: %s\n", gpshost_last_error()); return 1; }
    char text[4096];
    if (verbose) {  // ionosphere / UTC parameters of the header (plutogpssim.c:2487-2495)
        gpshost_describe_iono(sc, text, sizeof text);
        fputs(text, stderr);
    }
    if (use_radio) fprintf(stderr, "Gain: %.1fdB\n", radio.gain_db);  // plutogpssim.c:2571
    gpshost_describe(sc, text, sizeof text);   // RINEX date, start time, channel table (plutogpssim.c:2572-2574, 2634-2639)
    fputs(text, stderr);

    // From here on every exit goes through the cleanup block at the end: an open radio sink has its TX LO on, and only
    // gpssink_close powers it down and releases the device (plutogpssim.c:2160-2178).
    gpssink* sink = nullptr;
    gpsiq_multi* gq = nullptr;
    gpsiq_chan_desc* desc = nullptr;
    std::vector<int16_t*> iq;
    bool ok = true, sink_refused = false;
    std::string sink_err;
    long produced = 0;
    int64_t sunk_pairs = 0, pushes = 0;
    const auto t_begin = std::chrono::steady_clock::now();

    const int src = use_radio ? gpssink_open_radio(&sink, &radio) : out_path ? gpssink_open_file(&sink, out_path) : gpssink_open_null(&sink);
    if (src != GPSSINK_OK) { fprintf(stderr, "ERROR: %s\n", gpssink_last_error()); gpshost_close(sc); return 1; }

    const long total_epochs = duration == 0.0 ? LONG_MAX : (long) (duration * 10.0 + 0.5);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_handler = on_signal;
    sigaction(SIGINT, &sa, nullptr);
    sigaction(SIGTERM, &sa, nullptr);
    sigaction(SIGQUIT, &sa, nullptr);
    if (batch > total_epochs) batch = (int) total_epochs;
    do {  // (one pass; `break` = go to the cleanup)
        gpsiq_config gc;
        memset(&gc, 0, sizeof gc);
        gc.device = 0; gc.max_chan = hc.max_chan; gc.samples_per_epoch = kSamplesPerEpoch;
        gc.carrier_mode = hc.carrier_mode; gc.max_epochs = batch;
        if (gpsiq_multi_create(&gq, &gc, gpus) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_last_error(nullptr)); ok = false; break; }

        // In flight at any time: up to `ahead` batches submitted (descriptors generated, scans running) beyond the ones whose
        // rendering has begun, up to `gpus` batches being rendered / copied to the host (one per device), and the sink
        // draining older ones.  One pinned host buffer per batch that can be between fetch_begin and the sink's release.
        const int ahead = gpus + 1, inflight = gpus, nbuf = inflight + 2;
        const size_t desc_count = (size_t) batch * (size_t) hc.max_chan;
        desc = (gpsiq_chan_desc*) gpsiq_host_alloc(desc_count * sizeof(gpsiq_chan_desc));
        for (int i = 0; i < nbuf; i++) iq.push_back((int16_t*) gpsiq_host_alloc((size_t) batch * kSamplesPerEpoch * 4));
        if (!desc || std::find(iq.begin(), iq.end(), nullptr) != iq.end()) { fprintf(stderr, "ERROR: pinned host allocation failed\n"); ok = false; break; }

        std::deque<int> submitted;               // epochs of the batches submitted and not yet begun
        std::deque<std::pair<int, int>> begun;   // (buffer, epochs) of the batches being rendered
        std::vector<int64_t> ticket(nbuf, 0);    // the sink's claim on each pinned buffer
        size_t k = 0;                            // batches begun so far (buffer = k % nbuf)
        auto submit_next = [&]() -> bool {
            const int n = (int) std::min<long>(batch, total_epochs - produced);
            if (n <= 0 || g_stop || !ok) return false;
            if (gpshost_next(sc, desc, n) != GPSHOST_OK) { fprintf(stderr, "ERROR: %s\n", gpshost_last_error()); ok = false; return false; }
            if (gpsiq_multi_submit(gq, desc, n) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_multi_last_error(gq)); ok = false; return false; }
            submitted.push_back(n);
            produced += n;
            return true;
        };
        while (ok) {
            while ((int) (submitted.size() + begun.size()) < ahead + inflight && (int) submitted.size() < 2 * gpus && submit_next()) {}
            if (g_stop) break;                       // stop request: nothing more is fetched or queued
            while (ok && !submitted.empty() && (int) begun.size() < inflight) {   // start rendering on every idle device
                const int b = (int) (k % nbuf);
                if (ticket[b] > 0 && gpssink_wait(sink, ticket[b]) != GPSSINK_OK) { ok = false; sink_refused = true; break; }
                ticket[b] = 0;
                if (gpsiq_multi_fetch_begin(gq, iq[b]) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_multi_last_error(gq)); ok = false; break; }
                begun.push_back({b, submitted.front()});
                submitted.pop_front();
                k++;
            }
            if (!ok || begun.empty()) break;
            if (gpsiq_multi_fetch_end(gq) != GPSIQ_OK) { fprintf(stderr, "ERROR: %s\n", gpsiq_multi_last_error(gq)); ok = false; break; }
            const int b = begun.front().first, n_now = begun.front().second;
            begun.pop_front();
            // n_now push units (one 300000-sample buffer each, plutogpssim.c:2146-2158), written while later batches are fetched
            ticket[b] = gpssink_submit(sink, iq[b], (size_t) n_now * kSamplesPerEpoch);
            if (ticket[b] < 0) { ok = false; sink_refused = true; }
            if (verbose) fprintf(stderr, "\rTime into run = %4.1f", (double) (produced - (long) submitted.size() * batch) / 10.0);
        }
        if (g_stop) gpssink_abort(sink);             // finish the push unit in progress, discard the rest
        if (g_stop < 2)                              // orderly: let the GPUs finish what was begun (the buffers stay valid)
            while (!begun.empty()) { gpsiq_multi_fetch_end(gq); begun.pop_front(); }
    } while (false);

    // ---- cleanup: always the sink first (radio: TX LO off, buffer, channels, context), then the GPUs, then the host side
    if (sink) {
        if (!ok) gpssink_abort(sink);                // after a failure nothing queued is worth transmitting
        gpssink_stats(sink, &sunk_pairs, &pushes);   // waits for the writer
        if (sink_refused) sink_err = gpssink_last_error();
        if (gpssink_close(sink) != GPSSINK_OK) { if (ok) sink_err = gpssink_last_error(); ok = false; sink_refused = true; }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    if (sink_refused) fprintf(stderr, "\nERROR: sink refused data: %s\n", sink_err.c_str());
    fprintf(stderr, "%s%ld epochs (%.1f s of signal, %.0f samples) in %.3f s: %.1f Msamples/s, %lld kernel launches, %lld buffers to the sink\n",
            verbose ? "\n" : "", produced, produced / 10.0, (double) produced * kSamplesPerEpoch, secs,
            (double) produced * kSamplesPerEpoch / secs / 1e6, (long long) gpsiq_multi_launch_count(gq), (long long) (sunk_pairs / kSamplesPerEpoch));
    gpsiq_host_free(desc);
    for (int16_t* p : iq) gpsiq_host_free(p);
    gpsiq_multi_destroy(gq);
    gpshost_close(sc);
    return ok ? 0 : 1;
}
