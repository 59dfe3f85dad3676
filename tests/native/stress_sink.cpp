// TEST DRIVER (tests/test_sanitizers.py builds it with -fsanitize=thread together with host/gpssink.cpp).
// A producer that reuses a small ring of buffers the way gpsiq_sim does -- wait for the ticket of the buffer, refill,
// submit -- mixed with synchronous pushes and stats calls, against the file sink; the file must hold every batch
// once, in order.  Any data race between producer and writer thread is a ThreadSanitizer report.
//   stress_sink <scratch directory>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "gpssink.h"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    for (int round = 0; round < 20; round++) {
        const std::string path = std::string(argv[1]) + "/stress_" + std::to_string(round) + ".bin";
        gpssink* s = nullptr;
        if (gpssink_open_file(&s, path.c_str())) return 1;
        std::vector<std::vector<int16_t>> bufs(8, std::vector<int16_t>(2 * 1000));
        int64_t ticket[8] = {0};
        long total = 0;
        for (int k = 0; k < 200; k++) {
            const int b = k % 8;
            if (ticket[b] > 0 && gpssink_wait(s, ticket[b])) return 2;
            for (size_t i = 0; i < bufs[b].size(); i++) bufs[b][i] = (int16_t) (k + (int) i);
            if (k % 17 == 5) {
                if (gpssink_push(s, bufs[b].data(), 1000)) return 3;
                ticket[b] = 0;
            } else {
                ticket[b] = gpssink_submit(s, bufs[b].data(), 1000);
                if (ticket[b] < 0) return 4;
            }
            total += 1000;
            if (k % 50 == 0) {
                int64_t pairs = 0, pushes = 0;
                gpssink_stats(s, &pairs, &pushes);
                if (pairs != total) return 5;
            }
        }
        if (gpssink_close(s)) return 6;
        FILE* f = fopen(path.c_str(), "rb");
        std::vector<int16_t> all(200 * 2000);
        const size_t got = fread(all.data(), 2, all.size(), f);
        fclose(f);
        remove(path.c_str());
        if (got != all.size()) return 7;
        for (int k = 0; k < 200; k++)
            for (int i = 0; i < 2000; i++)
                if (all[(size_t) k * 2000 + i] != (int16_t) (k + i)) return 8;
    }
    puts("stress_sink: ok");
    return 0;
}
