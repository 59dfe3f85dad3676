// TEST DRIVER (tests/test_sanitizers.py builds it with -fsanitize=address,undefined together with the host sources).
// Mutates a navigation file -- truncation, byte noise in the header and first records, deleted spans, inserted digit
// runs -- and runs every mutant through gpshost_open + 320 epochs of gpshost_next (across a 30 s refresh).  Any
// outcome is fine (descriptors or an error code) except a sanitizer report, a crash or a hang.
//   fuzz_nav <plain-text nav file> <rinex3: 0|1> <iterations> <scratch file>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gpshost.h"

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const int v3 = atoi(argv[2]), iters = atoi(argv[3]);
    const char* scratch = argv[4];
    std::string data;
    {
        FILE* f = fopen(argv[1], "rb");
        if (!f) return 2;
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
        fclose(f);
    }
    srand(7);
    int opened = 0, refused = 0;
    static const char kAlphabet[] = "0123456789 .D-+E\n";
    for (int it = 0; it < iters; it++) {
        std::string d = data;
        switch (it % 4) {
            case 0: d.resize((size_t) rand() % d.size()); break;
            case 1:
                for (int k = 0, n = 1 + rand() % 20; k < n; k++) d[(size_t) rand() % (d.size() / 12)] = kAlphabet[rand() % 17];
                break;
            case 2: d.erase((size_t) rand() % d.size(), (size_t) rand() % 2000); break;
            default: d.insert((size_t) rand() % d.size(), std::string((size_t) rand() % 300, kAlphabet[rand() % 16])); break;
        }
        FILE* o = fopen(scratch, "wb");
        fwrite(d.data(), 1, d.size(), o);
        fclose(o);
        gpshost_config hc;
        memset(&hc, 0, sizeof hc);
        hc.nav_path = scratch;
        hc.pos_mode = GPSHOST_POS_LLH;
        hc.pos[0] = 30.2; hc.pos[1] = 120.0; hc.pos[2] = 100.0;
        hc.sample_rate = 2600000;
        hc.max_chan = 12;
        hc.rinex3 = v3;
        hc.threads = (it % 8 == 3) ? 4 : 1;
        gpshost_scenario* sc = nullptr;
        if (gpshost_open(&sc, &hc) == GPSHOST_OK) {
            std::vector<gpsiq_chan_desc> de(12 * 320);
            if (gpshost_next(sc, de.data(), 320) == GPSHOST_OK) opened++; else refused++;
            gpshost_skip(sc, 700);
            gpshost_close(sc);
        } else {
            refused++;
        }
    }
    printf("fuzz_nav: %d mutants produced descriptors, %d were refused\n", opened, refused);
    return 0;
}
