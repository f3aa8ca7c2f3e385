/*
 * oracle/ref_harness_ws.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Runs the UNMODIFIED reference websocket callback (main.c:39-181) without libwebsockets:
 * main.c is textually #included (from where it lies under the reference tree, via -I) against
 * the API stand-in oracle/lws_shim/libwebsockets.h, its main() renamed, and lws_write() below
 * records every (bytes, length, write mode) the callback hands to the socket.  What the
 * tests read back is therefore the reference's own wire stream:
 *   spectrum message  "t s;f %u;b %u;s %d;d" + 1024 payload bytes           main.c:80-84
 *   audio message     "FF;t a;d" + 8 fragments of <= 2048 bytes of float32   main.c:86-110
 * including audio_get_audio_payload's drain behaviour at pool-buffer boundaries
 * (audio_main.c:53-63).  The consumer schedule is the harness's: after every replayed USB
 * buffer the callback is pumped until it writes nothing (ref_harness_cbb.c calls the hook).
 *
 * Two properties of main.c matter to whoever builds this file:
 *   - tmpbuffer is char[30] (main.c:46) and the spectrum header for an 8- or 9-digit
 *     frequency is 30 or 31 characters + NUL (main.c:81): the reference overruns it by one or
 *     two bytes.  This unit is compiled with -fno-stack-protector so that the overrun lands
 *     where it lands in the reference's own build instead of tripping a canary.
 *   - the callback sleeps 5 ms whenever it wrote nothing (main.c:129-132); usleep is made a
 *     no-op here, the virtual clock of ref_harness_cbb.c is what paces the run.
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define usleep(us) ((void) (us))
#define main ref_rtlws_main
#include "main.c"                 /* resolved through -I<reference>/src */
#undef main
#undef usleep

/* ---- the lws stand-in: record writes, ignore everything else ---- */
static uint8_t* g_log_bytes = NULL;
static int64_t g_log_cap = 0;
static int64_t g_log_n = 0;
static int32_t* g_log_rec = NULL;      /* [max_rec][3]: byte offset (low 31 bits), length, write mode */
static int g_log_max_rec = 0;
static int g_log_n_rec = 0;
static int g_log_overflow = 0;
static int g_writes_this_call = 0;

int lws_write(struct lws* wsi, unsigned char* buf, size_t len, int mode)
{
    (void) wsi;
    g_writes_this_call++;
    if (g_log_bytes == NULL || g_log_n + (int64_t) len > g_log_cap || g_log_n_rec >= g_log_max_rec)
    {
        g_log_overflow = 1;
        return (int) len;
    }
    memcpy(g_log_bytes + g_log_n, buf, len);
    g_log_rec[3 * g_log_n_rec + 0] = (int32_t) g_log_n;
    g_log_rec[3 * g_log_n_rec + 1] = (int32_t) len;
    g_log_rec[3 * g_log_n_rec + 2] = (int32_t) mode;
    g_log_n += (int64_t) len;
    g_log_n_rec++;
    return (int) len;
}
int lws_callback_on_writable(struct lws* wsi) { (void) wsi; return 0; }
int lws_callback_on_writable_all_protocol(const struct lws_context* c, const struct lws_protocols* p)
{
    (void) c;
    (void) p;
    return 0;
}
struct lws_context* lws_get_context(const struct lws* wsi) { (void) wsi; return NULL; }
const struct lws_protocols* lws_get_protocol(struct lws* wsi) { (void) wsi; return NULL; }
struct lws_context* lws_create_context(struct lws_context_creation_info* info) { (void) info; return NULL; }
int lws_service(struct lws_context* context, int timeout_ms) { (void) context; (void) timeout_ms; return -1; }
void lws_context_destroy(struct lws_context* context) { (void) context; }
void lwsl_notice(const char* format, ...) { (void) format; }

/* http_handler.c (static file serving) is outside the path; main() only copies its protocol entry */
struct lws_protocols* get_http_protocol()
{
    static struct lws_protocols none = { "http-only", NULL, 0, 0, 0, NULL };
    return &none;
}

/* ---- driving the callback ---- */
static struct per_session_data__rtl_ws g_pss;

void ref_ws_set_log(uint8_t* bytes, int64_t cap, int32_t* rec, int max_rec)
{
    g_log_bytes = bytes;
    g_log_cap = cap;
    g_log_rec = rec;
    g_log_max_rec = max_rec;
    g_log_n = 0;
    g_log_n_rec = 0;
    g_log_overflow = 0;
}
int ref_ws_n_records(void) { return g_log_n_rec; }
int64_t ref_ws_n_bytes(void) { return g_log_n; }
int ref_ws_overflowed(void) { return g_log_overflow; }

/* a client connects (main.c:55-65) and the send buffer main() would have allocated (main.c:194) */
int ref_ws_connect(void)
{
    if (send_buffer == NULL)
        send_buffer = calloc(LWS_SEND_BUFFER_PRE_PADDING + SEND_BUFFER_SIZE + LWS_SEND_BUFFER_POST_PADDING, 1);
    memset(&g_pss, 0, sizeof(g_pss));
    return callback_rtl_ws(NULL, LWS_CALLBACK_ESTABLISHED, &g_pss, NULL, 0);
}

/* a text command from the client: "start", "stop", "freq <kHz>", "bw <kHz>", "spectrumgain <dB>" (main.c:139-176) */
int ref_ws_command(const char* cmd)
{
    return callback_rtl_ws(NULL, LWS_CALLBACK_RECEIVE, &g_pss, (void*) cmd, strlen(cmd));
}

/* LWS_CALLBACK_SERVER_WRITEABLE until one call writes nothing (or max_calls); returns the writes made */
int ref_ws_pump(int max_calls)
{
    int total = 0;
    int i;
    for (i = 0; i < max_calls; ++i)
    {
        g_writes_this_call = 0;
        callback_rtl_ws(NULL, LWS_CALLBACK_SERVER_WRITEABLE, &g_pss, NULL, 0);
        if (g_writes_this_call == 0)
            break;
        total += g_writes_this_call;
    }
    return total;
}

/* commands queued by the test; delivered when the "client" connects, i.e. at the first consumer poll:
 * by then cbb_init() has run, as it has in the real server before any connection (main.c:199 vs :222) */
#define WS_MAX_CMDS 8
static char g_cmds[WS_MAX_CMDS][64];
static int g_n_cmds = 0;
static int g_connected = 0;

int ref_ws_queue_command(const char* cmd)
{
    if (g_n_cmds >= WS_MAX_CMDS || strlen(cmd) >= sizeof(g_cmds[0]))
        return -1;
    strcpy(g_cmds[g_n_cmds++], cmd);
    return 0;
}

static void ws_consumer_hook(void)
{
    int i;
    if (!g_connected)
    {
        ref_ws_connect();
        for (i = 0; i < g_n_cmds; ++i)
            ref_ws_command(g_cmds[i]);
        g_connected = 1;
    }
    ref_ws_pump(4096);
}

extern void ref_cbb_set_consumer(void (*hook)(void));

/* route ref_cbb_run's per-buffer consumer poll through the websocket callback */
void ref_ws_attach(int on)
{
    ref_cbb_set_consumer(on ? ws_consumer_hook : NULL);
    g_connected = 0;
    if (!on)
        g_n_cmds = 0;
}

int ref_ws_sent_audio_fragments(void) { return g_pss.sent_audio_fragments; }
