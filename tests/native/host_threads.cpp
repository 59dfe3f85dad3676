// TEST DRIVER (tests/test_sanitizers.py builds it with -fsanitize=thread together with the host sources).
// gpshost_next with worker threads over several batches that straddle 30 s refreshes: the workers read the channel
// table and ephemerides while filling pseudoranges; the serial pass then mutates them.  Any overlap of the two is a
// ThreadSanitizer report.  Then two scenarios at once, from two threads.
//   host_threads <nav file> <threads>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "gpshost.h"

static int one_scenario(const char* nav, int threads) {
    gpshost_config hc;
    memset(&hc, 0, sizeof hc);
    hc.nav_path = nav;
    hc.pos_mode = GPSHOST_POS_LLH;
    hc.pos[0] = 30.286502; hc.pos[1] = 120.032669; hc.pos[2] = 100.0;
    hc.sample_rate = 2600000;
    hc.max_chan = 12;
    hc.threads = threads;
    gpshost_scenario* sc = nullptr;
    if (gpshost_open(&sc, &hc) != GPSHOST_OK) return 1;
    std::vector<gpsiq_chan_desc> d(12 * 1000);
    for (int r = 0; r < 6; r++)
        if (gpshost_next(sc, d.data(), 1000) != GPSHOST_OK) return 3;
    gpshost_close(sc);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    if (int rc = one_scenario(argv[1], atoi(argv[2]))) return rc;
    // two scenarios driven from two threads at once (one process feeding several GPUs): nothing is shared between them
    int rc[2] = {0, 0};
    std::thread a([&] { rc[0] = one_scenario(argv[1], 3); }), b([&] { rc[1] = one_scenario(argv[1], 1); });
    a.join();
    b.join();
    if (rc[0] || rc[1]) return 10 + rc[0] + rc[1];
    puts("host_threads: ok");
    return 0;
}
