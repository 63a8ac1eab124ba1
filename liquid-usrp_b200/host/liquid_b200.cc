// liquid_b200.cc -- the slice of the liquid-dsp C API (include/liquid/liquid.h) that
// lib/ofdmtxrx.cc and the src/ programs call directly, implemented over the B200 library:
//   ofdmflexframegen_*   -> b2_ofdmgen_*   (lib/ofdmtxrx.cc:79-84,293,314-328,377-387)
//   ofdmflexframesync_*  -> b2_ofdmsync_*  (lib/ofdmtxrx.cc:91,242-247,482,518-525,625)
//   msresamp_crcf_*      -> b2_msresamp_*  (src/flexframe_rx.cc:179,240,275)
//   liquid_getopt_str2*, liquid_print_*_schemes (src/multichannel_tx.cc:46-49,92-94)
// Sample-at-a-time calls are legal and are batched internally: ofdmflexframesync_execute()
// queues samples and runs the GPU when `batch` samples are pending (or on
// ofdmflexframesync_flush / reset / destroy); callbacks fire in stream order.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <time.h>
#include <vector>

#include <liquid/liquid.h>
#include "b200_ofdm.h"

extern "C" void b2_set_callback_indices(uint64_t detect_index, uint64_t complete_index);

static int env_device()
{
    const char * e = getenv("B2_DEVICE");
    return e ? atoi(e) : 0;
}
static void die(const char * where)
{
    // liquid-dsp aborts on invalid configuration; keep that contract for the C handles
    fprintf(stderr, "error: %s, %s\n", where, b2_last_error());
    exit(1);
}

extern "C" {

const char * liquid_libversion(void) { return LIQUID_VERSION; }

// ------------------------------------------------------------------ scheme names
static const struct { const char * name; const char * fullname; unsigned int id; } k_mods[] = {
    {"bpsk", "binary phase-shift keying", LIQUID_MODEM_BPSK}, {"qpsk", "quaternary phase-shift keying", LIQUID_MODEM_QPSK},
    {"qam4", "quadrature amplitude-shift keying (4)", LIQUID_MODEM_QAM4}, {"qam16", "quadrature amplitude-shift keying (16)", LIQUID_MODEM_QAM16},
    {"qam64", "quadrature amplitude-shift keying (64)", LIQUID_MODEM_QAM64}, {"qam256", "quadrature amplitude-shift keying (256)", LIQUID_MODEM_QAM256}};
static const struct { const char * name; unsigned int id; } k_fecs[] = {
    {"none", LIQUID_FEC_NONE}, {"h128", LIQUID_FEC_HAMMING128}, {"g2412", LIQUID_FEC_GOLAY2412}, {"v27", LIQUID_FEC_CONV_V27}};
static const struct { const char * name; unsigned int id; } k_crcs[] = {{"none", LIQUID_CRC_NONE}, {"crc32", LIQUID_CRC_32}};

modulation_scheme liquid_getopt_str2mod(const char * s)
{
    for (size_t i = 0; i < sizeof(k_mods) / sizeof(k_mods[0]); i++)
        if (strcmp(s, k_mods[i].name) == 0) return (modulation_scheme)k_mods[i].id;
    fprintf(stderr, "warning: liquid_getopt_str2mod(), unknown/unsupported mod scheme : %s\n", s);
    return LIQUID_MODEM_UNKNOWN;
}
fec_scheme liquid_getopt_str2fec(const char * s)
{
    for (size_t i = 0; i < sizeof(k_fecs) / sizeof(k_fecs[0]); i++)
        if (strcmp(s, k_fecs[i].name) == 0) return (fec_scheme)k_fecs[i].id;
    fprintf(stderr, "warning: liquid_getopt_str2fec(), unknown/unsupported fec scheme : %s\n", s);
    return LIQUID_FEC_UNKNOWN;
}
crc_scheme liquid_getopt_str2crc(const char * s)
{
    for (size_t i = 0; i < sizeof(k_crcs) / sizeof(k_crcs[0]); i++)
        if (strcmp(s, k_crcs[i].name) == 0) return (crc_scheme)k_crcs[i].id;
    fprintf(stderr, "warning: liquid_getopt_str2crc(), unknown/unsupported crc scheme : %s\n", s);
    return LIQUID_CRC_UNKNOWN;
}
void liquid_print_modulation_schemes(void)
{
    printf("          ");
    for (size_t i = 0; i < sizeof(k_mods) / sizeof(k_mods[0]); i++) printf("%s%s", i ? ", " : "", k_mods[i].name);
    printf("\n");
}
void liquid_print_fec_schemes(void)
{
    printf("          ");
    for (size_t i = 0; i < sizeof(k_fecs) / sizeof(k_fecs[0]); i++) printf("%s%s", i ? ", " : "", k_fecs[i].name);
    printf("\n");
}
void liquid_print_crc_schemes(void)
{
    printf("          ");
    for (size_t i = 0; i < sizeof(k_crcs) / sizeof(k_crcs[0]); i++) printf("%s%s", i ? ", " : "", k_crcs[i].name);
    printf("\n");
}

// ------------------------------------------------------------------ ofdmflexframegen
struct ofdmflexframegen_s {
    b2_ofdmgen * g;
    unsigned int M, cp, taper, W;
    ofdmflexframegenprops_s props;
    std::vector<std::complex<float> > frame;     // all symbols of the assembled frame (incl. tail)
    unsigned int n_symbols, symbol;              // symbols in `frame`, next symbol to hand out
    unsigned int buf_index;                      // position inside the current symbol (write())
    int assembled, complete;
    unsigned int payload_len;
};

static const ofdmflexframegenprops_s k_props_default = {LIQUID_CRC_32, LIQUID_FEC_NONE, LIQUID_FEC_HAMMING128, LIQUID_MODEM_QPSK};
void ofdmflexframegenprops_init_default(ofdmflexframegenprops_s * p) { *p = k_props_default; }

ofdmflexframegen ofdmflexframegen_create(unsigned int M, unsigned int cp, unsigned int taper, unsigned char * p, ofdmflexframegenprops_s * props)
{
    ofdmflexframegen q = new ofdmflexframegen_s;
    q->g = NULL;
    if (b2_ofdmgen_create(M, cp, taper, p, env_device(), &q->g) != B2_OK) die("ofdmflexframegen_create()");
    q->M = M; q->cp = cp; q->taper = taper; q->W = M + cp;
    q->props = props ? *props : k_props_default;
    q->n_symbols = 0; q->symbol = 0; q->buf_index = q->W; q->assembled = 0; q->complete = 0; q->payload_len = 0;
    return q;
}
void ofdmflexframegen_destroy(ofdmflexframegen q)
{
    b2_ofdmgen_destroy(q->g);
    delete q;
}
void ofdmflexframegen_reset(ofdmflexframegen q)
{
    b2_ofdmgen_reset(q->g);
    q->frame.clear();
    q->n_symbols = 0; q->symbol = 0; q->buf_index = q->W; q->assembled = 0; q->complete = 0;
}
int ofdmflexframegen_is_assembled(ofdmflexframegen q) { return q->assembled; }
void ofdmflexframegen_getprops(ofdmflexframegen q, ofdmflexframegenprops_s * p) { *p = q->props; }
void ofdmflexframegen_setprops(ofdmflexframegen q, ofdmflexframegenprops_s * p) { q->props = p ? *p : k_props_default; }
unsigned int ofdmflexframegen_getframelen(ofdmflexframegen q) { return q->n_symbols ? q->n_symbols - 1 : 0; }
void ofdmflexframegen_print(ofdmflexframegen q)
{
    printf("ofdmflexframegen (b200): M=%u cp=%u taper=%u payload=%u B symbols=%u\n", q->M, q->cp, q->taper, q->payload_len, q->n_symbols);
}
// the whole frame is produced on the device at assemble time and streamed out by write()
void ofdmflexframegen_assemble(ofdmflexframegen q, const unsigned char * header, const unsigned char * payload, unsigned int len)
{
    unsigned int n = 0;
    if (b2_ofdmgen_assemble(q->g, header, payload, len, (int)q->props.check, (int)q->props.fec0, (int)q->props.fec1,
                            (int)q->props.mod_scheme, &n) != B2_OK) die("ofdmflexframegen_assemble()");
    q->frame.resize((size_t)n * q->W);
    int last = 0;
    if (b2_ofdmgen_write(q->g, (float *)q->frame.data(), n, &last) != B2_OK) die("ofdmflexframegen_assemble()");
    q->n_symbols = n; q->symbol = 0; q->buf_index = q->W; q->assembled = 1; q->complete = 0; q->payload_len = len;
}
int ofdmflexframegen_write(ofdmflexframegen q, liquid_float_complex * buf, unsigned int n)
{
    static const std::complex<float> zero(0.0f, 0.0f);
    for (unsigned int i = 0; i < n; i++) {
        if (q->buf_index >= q->W) {
            // next symbol; the tail buffer is the last one and clears "assembled"
            if (q->symbol < q->n_symbols) {
                q->symbol++;
                if (q->symbol == q->n_symbols) { q->complete = 1; q->assembled = 0; }
            } else {
                q->symbol = q->n_symbols + 1;        // past the frame: silence
            }
            q->buf_index = 0;
        }
        buf[i] = (q->symbol >= 1 && q->symbol <= q->n_symbols) ? q->frame[(size_t)(q->symbol - 1) * q->W + q->buf_index] : zero;
        q->buf_index++;
    }
    return q->complete;
}
int ofdmflexframegen_writesymbol(ofdmflexframegen q, liquid_float_complex * buffer)
{
    if (!q->assembled || q->n_symbols < 2) {
        memset((void *)buffer, 0, sizeof(std::complex<float>) * q->W);
        return 1;
    }
    memcpy((void *)buffer, &q->frame[(size_t)q->symbol * q->W], sizeof(std::complex<float>) * q->W);
    q->symbol++;
    if (q->symbol == q->n_symbols - 1) {         // that was the last payload symbol; no tail in this API
        ofdmflexframegen_reset(q);
        return 1;
    }
    return 0;
}

// ------------------------------------------------------------------ ofdmflexframesync
struct ofdmflexframesync_s {
    b2_ofdmsync * s;
    framesync_callback callback;
    void * userdata;
    std::vector<std::complex<float> > pending;
    size_t batch;
    long long first_ns, max_wait_ns;       // oldest staged sample / $B2_SYNC_MAX_LATENCY_MS (default 5 ms)
    float rssi, cfo;
};

ofdmflexframesync ofdmflexframesync_create(unsigned int M, unsigned int cp, unsigned int taper, unsigned char * p,
                                           framesync_callback callback, void * userdata)
{
    ofdmflexframesync q = new ofdmflexframesync_s;
    q->s = NULL;
    q->batch = 1u << 16;
    if (const char * e = getenv("B2_SYNC_BATCH")) {
        unsigned long v = strtoul(e, NULL, 10);
        if (v >= 1 && v <= (1ul << 24)) q->batch = v;
    }
    if (b2_ofdmsync_create(M, cp, taper, p, 1, env_device(), q->batch, &q->s) != B2_OK) die("ofdmflexframesync_create()");
    q->callback = callback; q->userdata = userdata;
    q->first_ns = 0; q->max_wait_ns = 5000000ll;
    if (const char * e = getenv("B2_SYNC_MAX_LATENCY_MS")) q->max_wait_ns = (long long)(atof(e) * 1e6);
    q->pending.reserve(q->batch);
    q->rssi = 0.0f; q->cfo = 0.0f;
    return q;
}
void ofdmflexframesync_flush(ofdmflexframesync q)
{
    if (!q->pending.empty()) {
        if (b2_ofdmsync_execute(q->s, (const float *)q->pending.data(), q->pending.size()) != B2_OK) die("ofdmflexframesync_execute()");
        q->pending.clear();
    }
    size_t n = 0, nb = 0;
    if (b2_ofdmsync_poll(q->s, NULL, 0, &n, NULL, 0, &nb) != B2_OK || n == 0) return;
    std::vector<b2_frame_rec> recs(n);
    std::vector<uint8_t> payloads(nb ? nb : 1);
    if (b2_ofdmsync_poll(q->s, recs.data(), n, &n, payloads.data(), payloads.size(), &nb) != B2_OK) die("ofdmflexframesync_execute()");
    for (size_t i = 0; i < n; i++) {
        const b2_frame_rec & r = recs[i];
        q->rssi = r.rssi; q->cfo = r.cfo;
        if (!q->callback) continue;
        framesyncstats_s st;
        st.evm = r.evm; st.rssi = r.rssi; st.cfo = r.cfo; st.framesyms = NULL; st.num_framesyms = 0;
        st.mod_scheme = r.mod_scheme; st.mod_bps = r.mod_bps; st.check = r.check; st.fec0 = r.fec0; st.fec1 = r.fec1;
        unsigned char header[8];
        memcpy(header, r.header, 8);
        unsigned char * payload = (r.header_valid && r.payload_len) ? payloads.data() + r.payload_offset : NULL;
        b2_set_callback_indices(r.detect_index, r.complete_index);
        q->callback(header, r.header_valid, payload, r.payload_len, r.payload_valid, st, q->userdata);
    }
}
void ofdmflexframesync_destroy(ofdmflexframesync q)
{
    ofdmflexframesync_flush(q);
    b2_ofdmsync_destroy(q->s);
    delete q;
}
void ofdmflexframesync_reset(ofdmflexframesync q)
{
    ofdmflexframesync_flush(q);
    b2_ofdmsync_reset(q->s);
}
static inline long long b2_now_ns()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}
void ofdmflexframesync_execute(ofdmflexframesync q, liquid_float_complex * x, unsigned int n)
{
    if (n == 0) return;
    if (q->pending.empty() && q->max_wait_ns > 0) q->first_ns = b2_now_ns();
    unsigned int i = 0;
    while (i < n) {
        size_t c = q->batch - q->pending.size();
        if (c > n - i) c = n - i;
        q->pending.insert(q->pending.end(), x + i, x + i + c);
        i += (unsigned int)c;
        if (q->pending.size() >= q->batch) { ofdmflexframesync_flush(q); if (q->max_wait_ns > 0) q->first_ns = b2_now_ns(); }
    }
    // liquid fires the callback inside this call; here samples are staged, but never for longer than the bound
    // (a receive worker that feeds one sample per call, lib/ofdmtxrx.cc:620-626, still gets its batching)
    if (!q->pending.empty() && q->max_wait_ns > 0 && b2_now_ns() - q->first_ns >= q->max_wait_ns) ofdmflexframesync_flush(q);
}
void ofdmflexframesync_print(ofdmflexframesync q) { (void)q; printf("ofdmflexframesync (b200)\n"); }
float ofdmflexframesync_get_rssi(ofdmflexframesync q) { return q->rssi; }
float ofdmflexframesync_get_cfo(ofdmflexframesync q) { return q->cfo; }
void ofdmflexframesync_debug_enable(ofdmflexframesync q) { (void)q; }
void ofdmflexframesync_debug_disable(ofdmflexframesync q) { (void)q; }
void ofdmflexframesync_debug_print(ofdmflexframesync q, const char * filename) { (void)q; (void)filename; }

// ------------------------------------------------------------------ msresamp_crcf
struct msresamp_crcf_s { b2_msresamp * r; float rate; };
msresamp_crcf msresamp_crcf_create(float rate, float As)
{
    msresamp_crcf q = new msresamp_crcf_s;
    q->r = NULL; q->rate = rate;
    if (b2_msresamp_create(rate, As, env_device(), &q->r) != B2_OK) die("msresamp_crcf_create()");
    return q;
}
void msresamp_crcf_destroy(msresamp_crcf q) { b2_msresamp_destroy(q->r); delete q; }
void msresamp_crcf_reset(msresamp_crcf q) { b2_msresamp_reset(q->r); }
float msresamp_crcf_get_delay(msresamp_crcf q) { (void)q; return 7.0f; }
void msresamp_crcf_execute(msresamp_crcf q, liquid_float_complex * x, unsigned int nx, liquid_float_complex * y, unsigned int * ny)
{
    size_t n = 0;
    // the caller sizes y as in src/flexframe_rx.cc:198 ((int)(2*rate) + 64 per input sample block)
    if (b2_msresamp_execute(q->r, (const float *)x, nx, (float *)y, (size_t)(2.0f * q->rate * nx) + 64, &n) != B2_OK) die("msresamp_crcf_execute()");
    *ny = (unsigned int)n;
}

} // extern "C"
