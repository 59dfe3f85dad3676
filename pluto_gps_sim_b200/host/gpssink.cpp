// gpssink.cpp -- output transport behind include/gpssink.h (compiled into libgpshost.so).
//
// The reference's transport is pluto_tx_thread_ep (plutogpssim.c:2058-2190).  This file keeps its device
// contract -- what is written to which IIO attribute, one non-cyclic 300000-pair buffer on 12 kernel buffers,
// copy + push per 0.1 s epoch, LO off / destroy on the way out -- and changes the plumbing around it:
//   * libiio and libad9361 are resolved with dlopen/dlsym when a radio sink is opened (a symbol table, not
//     link-time dependencies), so the library loads and its other sinks work on boxes without them;
//   * the producer/consumer hand-off is a ticketed job queue drained by one writer thread per sink: every push
//     unit goes out exactly once and in order (the reference's timing-based condvar ping-pong can drop or repeat
//     a buffer).
#include "../../include/gpssink.h"

#include <dlfcn.h>

#include <cerrno>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

struct Backend {
    int64_t pushes = 0;
    std::atomic<bool> aborted{false};   // gpssink_abort: stop at the next push unit, discard the rest
    virtual ~Backend() {}
    virtual int write(const int16_t* iq, size_t pairs) = 0;
    virtual int finish() { return GPSSINK_OK; }
};

struct NullBackend : Backend {
    int write(const int16_t*, size_t) override { pushes++; return GPSSINK_OK; }
};

struct FileBackend : Backend {
    FILE* fp = nullptr;
    bool own = false;
    std::string path;
    int write(const int16_t* iq, size_t pairs) override {
        if (fwrite(iq, 4, pairs, fp) != pairs) return fail(GPSSINK_ERR_IO, "short write to %s: %s", path.c_str(), strerror(errno));
        pushes++;
        return GPSSINK_OK;
    }
    int finish() override {
        int rc = GPSSINK_OK;
        if (fp) {
            if (fflush(fp) != 0) rc = fail(GPSSINK_ERR_IO, "flush of %s failed: %s", path.c_str(), strerror(errno));
            if (own) fclose(fp);
            fp = nullptr;
        }
        return rc;
    }
    ~FileBackend() override { finish(); }
};

// ---------------------------------------------------------------------------------------------------------
// libiio, resolved at run time.  Opaque handles only; the signatures are libiio 0.x's (iio.h) for exactly the
// entry points the reference's TX thread uses.
struct IioApi {
    void* lib = nullptr;
    void* lib_ad = nullptr;
    void* (*create_default_context)() = nullptr;
    void* (*create_network_context)(const char*) = nullptr;
    void* (*create_context_from_uri)(const char*) = nullptr;
    void (*context_destroy)(void*) = nullptr;
    void (*strerror_)(int, char*, size_t) = nullptr;
    unsigned (*context_get_devices_count)(const void*) = nullptr;
    void* (*context_find_device)(const void*, const char*) = nullptr;
    int (*device_set_kernel_buffers_count)(const void*, unsigned) = nullptr;
    void* (*device_find_channel)(const void*, const char*, bool) = nullptr;
    ssize_t (*channel_attr_write)(const void*, const char*, const char*) = nullptr;
    int (*channel_attr_write_longlong)(const void*, const char*, long long) = nullptr;
    int (*channel_attr_write_double)(const void*, const char*, double) = nullptr;
    int (*channel_attr_write_bool)(const void*, const char*, bool) = nullptr;
    void (*channel_enable)(void*) = nullptr;
    void (*channel_disable)(void*) = nullptr;
    void* (*device_create_buffer)(const void*, size_t, bool) = nullptr;
    void (*buffer_destroy)(void*) = nullptr;
    void* (*buffer_start)(const void*) = nullptr;
    ssize_t (*buffer_push)(void*) = nullptr;
    int (*ad9361_set_bb_rate)(void*, unsigned long) = nullptr;  // libad9361 (plutogpssim.c:2131)

    ~IioApi() {
        if (lib_ad && lib_ad != lib) dlclose(lib_ad);
        if (lib) dlclose(lib);
    }

    static void* open_first(const char* explicit_path, const char* env, const char* const* fallbacks, std::string* tried) {
        const char* from_env = env ? getenv(env) : nullptr;
        const char* one[2] = {explicit_path ? explicit_path : from_env, nullptr};
        const char* const* names = one[0] ? one : fallbacks;
        for (; *names; names++) {
            if (void* h = dlopen(*names, RTLD_NOW | RTLD_LOCAL)) return h;
            *tried += *tried == "" ? "" : ", ";
            *tried += *names;
        }
        return nullptr;
    }

    int load(const gpssink_radio_config& cfg) {
        static const char* const iio_names[] = {"libiio.so.0", "libiio.so", nullptr};
        static const char* const ad_names[] = {"libad9361.so.0", "libad9361.so", nullptr};
        std::string tried;
        lib = open_first(cfg.iio_lib, "GPSSINK_IIO_LIB", iio_names, &tried);
        if (!lib) return fail(GPSSINK_ERR_BACKEND, "libiio is not available (tried %s): %s", tried.c_str(), dlerror());
        struct { const char* name; void** slot; } table[] = {
            {"iio_create_default_context", (void**) &create_default_context},
            {"iio_create_network_context", (void**) &create_network_context},
            {"iio_create_context_from_uri", (void**) &create_context_from_uri},
            {"iio_context_destroy", (void**) &context_destroy},
            {"iio_strerror", (void**) &strerror_},
            {"iio_context_get_devices_count", (void**) &context_get_devices_count},
            {"iio_context_find_device", (void**) &context_find_device},
            {"iio_device_set_kernel_buffers_count", (void**) &device_set_kernel_buffers_count},
            {"iio_device_find_channel", (void**) &device_find_channel},
            {"iio_channel_attr_write", (void**) &channel_attr_write},
            {"iio_channel_attr_write_longlong", (void**) &channel_attr_write_longlong},
            {"iio_channel_attr_write_double", (void**) &channel_attr_write_double},
            {"iio_channel_attr_write_bool", (void**) &channel_attr_write_bool},
            {"iio_channel_enable", (void**) &channel_enable},
            {"iio_channel_disable", (void**) &channel_disable},
            {"iio_device_create_buffer", (void**) &device_create_buffer},
            {"iio_buffer_destroy", (void**) &buffer_destroy},
            {"iio_buffer_start", (void**) &buffer_start},
            {"iio_buffer_push", (void**) &buffer_push},
        };
        for (auto& e : table) {
            *e.slot = dlsym(lib, e.name);
            if (!*e.slot) return fail(GPSSINK_ERR_BACKEND, "libiio lacks %s", e.name);
        }
        std::string tried_ad;
        lib_ad = open_first(cfg.ad9361_lib, "GPSSINK_AD9361_LIB", ad_names, &tried_ad);
        if (!lib_ad) lib_ad = lib;  // a combined build (and the capture backend of the tests) exports it from libiio
        *(void**) &ad9361_set_bb_rate = dlsym(lib_ad, "ad9361_set_bb_rate");
        return GPSSINK_OK;
    }
};

struct RadioBackend : Backend {
    IioApi api;
    gpssink_radio_config cfg;
    void* ctx = nullptr;
    void* phy = nullptr;
    void* tx_i = nullptr;
    void* tx_q = nullptr;
    void* buffer = nullptr;
    char* buffer_mem = nullptr;
    bool closed = false;

    void* phy_channel(const char* name) { return api.device_find_channel(phy, name, true); }

    std::string iio_error() {
        char buf[256];
        api.strerror_(errno, buf, sizeof buf);
        return buf;
    }

    int open() {
        if (int rc = api.load(cfg)) return rc;
        // context: the default one (local / $IIOD_REMOTE) first, then -N, -U, pluto.local (plutogpssim.c:2071-2081)
        ctx = api.create_default_context();
        if (!ctx) {
            if (cfg.hostname) ctx = api.create_network_context(cfg.hostname);
            else if (cfg.uri) ctx = api.create_context_from_uri(cfg.uri);
            else ctx = api.create_network_context("pluto.local");
        }
        if (!ctx) return fail(GPSSINK_ERR_DEVICE, "Failed creating IIO context: %s", iio_error().c_str());
        if (api.context_get_devices_count(ctx) == 0) return fail(GPSSINK_ERR_DEVICE, "No supported PLUTOSDR devices found.");
        void* tx = api.context_find_device(ctx, "cf-ad9361-dds-core-lpc");
        if (!tx) return fail(GPSSINK_ERR_DEVICE, "Error opening PLUTOSDR TX device: %s", iio_error().c_str());
        api.device_set_kernel_buffers_count(tx, (unsigned) cfg.kernel_buffers);  // default is 4 (plutogpssim.c:2101-2102)

        phy = api.context_find_device(ctx, "ad9361-phy");
        if (!phy) return fail(GPSSINK_ERR_DEVICE, "ad9361-phy not found in the IIO context");
        // transmit chain of the AD9361 (plutogpssim.c:2105-2110), then RX LO off and TX LO frequency (2112-2118)
        void* chain = phy_channel("voltage0");
        void* rx_lo = phy_channel("altvoltage0");
        void* tx_lo = phy_channel("altvoltage1");
        if (!chain || !rx_lo || !tx_lo)  // (the reference would hand libiio a null channel here)
            return fail(GPSSINK_ERR_DEVICE, "ad9361-phy lacks the %s channel", !chain ? "voltage0" : !rx_lo ? "altvoltage0" : "altvoltage1");
        api.channel_attr_write(chain, "rf_port_select", cfg.rfport);
        api.channel_attr_write_longlong(chain, "rf_bandwidth", cfg.bw_hz);
        api.channel_attr_write_longlong(chain, "sampling_frequency", cfg.fs_hz);
        api.channel_attr_write_double(chain, "hardwaregain", cfg.gain_db);
        api.channel_attr_write_bool(rx_lo, "powerdown", true);
        api.channel_attr_write_longlong(tx_lo, "frequency", cfg.lo_hz);

        // streaming channels of the DAC core: I then Q, with the altvoltage names as fall-back (plutogpssim.c:2120-2129)
        static const char* const names[2][2] = {{"voltage0", "altvoltage0"}, {"voltage1", "altvoltage1"}};
        void** slots[2] = {&tx_i, &tx_q};
        for (int k = 0; k < 2; k++) {
            *slots[k] = api.device_find_channel(tx, names[k][0], true);
            if (!*slots[k]) *slots[k] = api.device_find_channel(tx, names[k][1], true);
            if (!*slots[k]) return fail(GPSSINK_ERR_DEVICE, "TX streaming channel %s not found", names[k][0]);
        }
        api.channel_enable(tx_i);
        api.channel_enable(tx_q);

        // baseband rate incl. the FIR the AD9361 needs below 25/12 MS/s (plutogpssim.c:2131)
        if (api.ad9361_set_bb_rate) api.ad9361_set_bb_rate(api.context_find_device(ctx, "ad9361-phy"), (unsigned long) cfg.fs_hz);
        else if (cfg.fs_hz * 12 < 25000000LL) return fail(GPSSINK_ERR_BACKEND, "libad9361 (ad9361_set_bb_rate) is needed for %lld S/s", cfg.fs_hz);
        else fprintf(stderr, "note: libad9361 not found; sampling_frequency was set directly\n");

        buffer = api.device_create_buffer(tx, (size_t) cfg.pairs_per_push, false);
        if (!buffer) return fail(GPSSINK_ERR_DEVICE, "Could not create TX buffer.");
        api.channel_attr_write_bool(tx_lo, "powerdown", false);  // TX LO on (plutogpssim.c:2139-2141)
        buffer_mem = (char*) api.buffer_start(buffer);
        return GPSSINK_OK;
    }

    int write(const int16_t* iq, size_t pairs) override {
        const size_t unit = (size_t) cfg.pairs_per_push;
        if (pairs % unit) return fail(GPSSINK_ERR_ARG, "radio sink takes whole %zu-pair buffers (got %zu pairs)", unit, pairs);
        for (size_t done = 0; done < pairs; done += unit) {
            if (aborted.load(std::memory_order_relaxed)) return GPSSINK_OK;   // stop within one 0.1 s push, like the reference
            memcpy(buffer_mem, iq + 2 * done, unit * 4);           // plutogpssim.c:2148
            const ssize_t n = api.buffer_push(buffer);              // plutogpssim.c:2152
            if (n < 0) return fail(GPSSINK_ERR_PUSH, "Error pushing buf %d", (int) n);
            pushes++;
        }
        return GPSSINK_OK;
    }

    int finish() override {  // plutogpssim.c:2160-2178
        if (closed) return GPSSINK_OK;
        closed = true;
        if (ctx) {  // TX LO off -- also after a failed start, like the reference, but never through a null handle
            void* dev = api.context_find_device(ctx, "ad9361-phy");
            void* lo = dev ? api.device_find_channel(dev, "altvoltage1", true) : nullptr;
            if (lo) api.channel_attr_write_bool(lo, "powerdown", true);
        }
        if (buffer) api.buffer_destroy(buffer);
        if (tx_i) api.channel_disable(tx_i);
        if (tx_q) api.channel_disable(tx_q);
        if (ctx) api.context_destroy(ctx);
        buffer = nullptr; ctx = nullptr;
        return GPSSINK_OK;
    }
    ~RadioBackend() override { if (api.lib) finish(); }
};

}  // namespace

struct gpssink {
    std::unique_ptr<Backend> be;
    int64_t pairs = 0;
    // writer thread state
    struct Job { int64_t ticket; const int16_t* iq; size_t pairs; };
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> jobs;
    std::thread writer;
    int64_t next_ticket = 1, done_ticket = 0;
    int status = GPSSINK_OK;     // sticky: the first failure
    std::string status_msg;
    bool stop = false;

    int write_now(const int16_t* iq, size_t n) {
        const int rc = be->write(iq, n);
        if (rc == GPSSINK_OK) pairs += (int64_t) n;
        return rc;
    }

    void run() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
            if (jobs.empty()) return;  // stop requested and drained
            const Job j = jobs.front();
            jobs.pop_front();
            int rc = status;
            if (rc == GPSSINK_OK && !be->aborted.load()) {   // (aborted: queued batches are discarded, not written)
                lk.unlock();
                rc = write_now(j.iq, j.pairs);
                std::string msg = rc == GPSSINK_OK ? "" : g_err;  // g_err is per thread: carry it over to the waiter
                lk.lock();
                if (rc != GPSSINK_OK) { status = rc; status_msg = msg; }
            }
            done_ticket = j.ticket;
            cv_done.notify_all();
        }
    }

    void drain_and_join() {
        if (!writer.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_job.notify_all();
        writer.join();
    }
};

extern "C" {

void gpssink_radio_defaults(gpssink_radio_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->fs_hz = 3000000;              // TX_SAMPLE_FREQ (plutogpssim.c:43, 2271)
    cfg->bw_hz = 2 * cfg->fs_hz;       // plutogpssim.c:2270
    cfg->lo_hz = 1575420000LL;         // plutogpssim.c:2272
    cfg->rfport = "A";
    cfg->gain_db = -20.0;
    cfg->kernel_buffers = 12;
    cfg->pairs_per_push = GPSSINK_PUSH_PAIRS;
}

int gpssink_radio_option(gpssink_radio_config* cfg, int letter, const char* arg) {
    if (!cfg || !arg) return fail(GPSSINK_ERR_ARG, "null argument");
    switch (letter) {
        case 'A':  // plutogpssim.c:2367-2371
            cfg->gain_db = atof(arg);
            if (cfg->gain_db > 0.0) cfg->gain_db = 0.0;
            if (cfg->gain_db < -80.0) cfg->gain_db = -80.0;
            return GPSSINK_OK;
        case 'B': {  // plutogpssim.c:2372-2376; MHz, rounded to Hz the way the reference's MHZ() does
            long long hz = (long long) (atof(arg) * 1000000.0 + .5);
            if (hz > 5000000LL) hz = 5000000LL;
            if (hz < 1000000LL) hz = 1000000LL;
            cfg->bw_hz = hz;
            return GPSSINK_OK;
        }
        case 'U': cfg->uri = arg; return GPSSINK_OK;
        case 'N': cfg->hostname = arg; return GPSSINK_OK;
        case 's': cfg->fs_hz = (long long) atoi(arg); return GPSSINK_OK;  // plutogpssim.c:2325 (bw_hz is NOT rescaled there)
        default: return fail(GPSSINK_ERR_ARG, "not a radio option: -%c", letter);
    }
}

static int adopt(gpssink** out, Backend* be) {
    gpssink* s = new gpssink();
    s->be.reset(be);
    *out = s;
    return GPSSINK_OK;
}

int gpssink_open_null(gpssink** out) {
    if (!out) return fail(GPSSINK_ERR_ARG, "null argument");
    return adopt(out, new NullBackend());
}

int gpssink_open_file(gpssink** out, const char* path) {
    if (!out || !path || !*path) return fail(GPSSINK_ERR_ARG, "null argument");
    std::unique_ptr<FileBackend> f(new FileBackend());
    f->path = path;
    if (strcmp(path, "-") == 0) {
        f->fp = stdout;
    } else {
        f->fp = fopen(path, "wb");
        f->own = true;
        if (!f->fp) return fail(GPSSINK_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    }
    return adopt(out, f.release());
}

int gpssink_open_radio(gpssink** out, const gpssink_radio_config* cfg) {
    if (!out || !cfg) return fail(GPSSINK_ERR_ARG, "null argument");
    if (cfg->pairs_per_push <= 0 || cfg->kernel_buffers <= 0 || !cfg->rfport || cfg->fs_hz <= 0)
        return fail(GPSSINK_ERR_ARG, "radio configuration not initialised (use gpssink_radio_defaults)");
    std::unique_ptr<RadioBackend> r(new RadioBackend());
    r->cfg = *cfg;
    if (int rc = r->open()) {
        const std::string keep = g_err;
        if (r->api.lib) r->finish();  // the reference's exit path also runs after a failed start (plutogpssim.c:2160)
        g_err = keep;
        return rc;
    }
    return adopt(out, r.release());
}

int gpssink_push(gpssink* s, const int16_t* iq, size_t pairs) {
    if (!s || (!iq && pairs)) return fail(GPSSINK_ERR_ARG, "null argument");
    {
        std::unique_lock<std::mutex> lk(s->mu);   // keep order with queued batches
        s->cv_done.wait(lk, [&] { return s->done_ticket == s->next_ticket - 1; });
        if (s->status != GPSSINK_OK) return fail(s->status, "%s", s->status_msg.c_str());
    }
    const int rc = s->write_now(iq, pairs);
    if (rc != GPSSINK_OK) {
        std::lock_guard<std::mutex> lk(s->mu);
        s->status = rc;
        s->status_msg = g_err;
    }
    return rc;
}

int64_t gpssink_submit(gpssink* s, const int16_t* iq, size_t pairs) {
    if (!s || (!iq && pairs)) return fail(GPSSINK_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->writer.joinable()) s->writer = std::thread([s] { s->run(); });
    const int64_t t = s->next_ticket++;
    s->jobs.push_back({t, iq, pairs});
    s->cv_job.notify_one();
    return t;
}

int gpssink_wait(gpssink* s, int64_t ticket) {
    if (!s) return fail(GPSSINK_ERR_ARG, "null argument");
    std::unique_lock<std::mutex> lk(s->mu);
    if (ticket <= 0 || ticket >= s->next_ticket) return fail(GPSSINK_ERR_ARG, "unknown ticket %lld", (long long) ticket);
    s->cv_done.wait(lk, [&] { return s->done_ticket >= ticket; });
    if (s->status != GPSSINK_OK) return fail(s->status, "%s", s->status_msg.c_str());
    return GPSSINK_OK;
}

int gpssink_abort(gpssink* s) {
    if (!s) return fail(GPSSINK_ERR_ARG, "null argument");
    s->be->aborted.store(true);
    return GPSSINK_OK;
}

int gpssink_stats(gpssink* s, int64_t* pairs, int64_t* pushes) {
    if (!s) return fail(GPSSINK_ERR_ARG, "null argument");
    std::unique_lock<std::mutex> lk(s->mu);
    s->cv_done.wait(lk, [&] { return s->done_ticket == s->next_ticket - 1; });
    if (pairs) *pairs = s->pairs;
    if (pushes) *pushes = s->be->pushes;
    return GPSSINK_OK;
}

int gpssink_close(gpssink* s) {
    if (!s) return GPSSINK_OK;
    s->drain_and_join();
    int rc = s->status;
    std::string msg = s->status_msg;
    const int rc2 = s->be->finish();
    if (rc == GPSSINK_OK && rc2 != GPSSINK_OK) { rc = rc2; msg = g_err; }
    delete s;
    if (rc != GPSSINK_OK) g_err = msg;
    return rc;
}

const char* gpssink_last_error(void) { return g_err.c_str(); }

}  // extern "C"
