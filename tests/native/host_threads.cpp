// TEST DRIVER (tests/test_sanitizers.py builds it with -fsanitize=thread together with the host sources).
// gpshost_next with worker threads over several batches that straddle 30 s refreshes: the workers read the channel
// table and ephemerides while filling pseudoranges; the serial pass then mutates them.  Any overlap of the two is a
// ThreadSanitizer report.
//   host_threads <nav file> <threads>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gpshost.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    gpshost_config hc;
    memset(&hc, 0, sizeof hc);
    hc.nav_path = argv[1];
    hc.pos_mode = GPSHOST_POS_LLH;
    hc.pos[0] = 30.286502; hc.pos[1] = 120.032669; hc.pos[2] = 100.0;
    hc.sample_rate = 2600000;
    hc.max_chan = 12;
    hc.threads = atoi(argv[2]);
    gpshost_scenario* sc = nullptr;
    if (gpshost_open(&sc, &hc) != GPSHOST_OK) return 1;
    std::vector<gpsiq_chan_desc> d(12 * 1000);
    for (int r = 0; r < 6; r++)
        if (gpshost_next(sc, d.data(), 1000) != GPSHOST_OK) return 3;
    gpshost_close(sc);
    puts("host_threads: ok");
    return 0;
}
