/*
 * oracle/lws_shim/libwebsockets.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A from-scratch stand-in for the slice of the libwebsockets API that the reference's
 * src/main.c touches (main.c:6 include; :39-40 callback signature; :53-178 callback
 * reasons; :80-110 LWS_SEND_BUFFER_*_PADDING, lws_write and its LWS_WRITE_* modes;
 * :135,169 lws_callback_on_writable*; :189-228 protocol table / context calls).
 * libwebsockets itself is absent from this image (system -lwebsockets, reference
 * Makefile:21, unpinned).  The unmodified main.c is #included by
 * oracle/ref_harness_ws.c against this header so that its own
 * LWS_CALLBACK_SERVER_WRITEABLE branch produces the wire bytes the tests pin the product's
 * emitters to; lws_write() here just records what it is handed.
 * It is NOT libwebsockets: no sockets, no framing, no service loop.
 */
#ifndef ORACLE_LWS_SHIM_H
#define ORACLE_LWS_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct lws;
struct lws_context;
struct lws_extension;

enum lws_callback_reasons {
    LWS_CALLBACK_ESTABLISHED = 0,
    LWS_CALLBACK_CLOSED = 4,
    LWS_CALLBACK_RECEIVE = 6,
    LWS_CALLBACK_SERVER_WRITEABLE = 11,
    LWS_CALLBACK_HTTP = 12,
    LWS_CALLBACK_PROTOCOL_DESTROY = 28
};

/* write modes: low bits select the frame type, LWS_WRITE_NO_FIN is a flag OR-ed in */
enum lws_write_protocol {
    LWS_WRITE_TEXT = 0,
    LWS_WRITE_BINARY = 1,
    LWS_WRITE_CONTINUATION = 2,
    LWS_WRITE_HTTP = 3,
    LWS_WRITE_NO_FIN = 0x40
};

#define LWS_SEND_BUFFER_PRE_PADDING 16
#define LWS_SEND_BUFFER_POST_PADDING 4

typedef int (*lws_callback_function)(struct lws* wsi, enum lws_callback_reasons reason, void* user, void* in,
                                     size_t len);

struct lws_protocols {
    const char* name;
    lws_callback_function callback;
    size_t per_session_data_size;
    size_t rx_buffer_size;
    unsigned int id;
    void* user;
};

struct lws_context_creation_info {
    int port;
    const char* iface;
    const struct lws_protocols* protocols;
    const struct lws_extension* extensions;
    int gid;
    int uid;
    unsigned int options;
    void* user;
};

int lws_write(struct lws* wsi, unsigned char* buf, size_t len, int mode);
int lws_callback_on_writable(struct lws* wsi);
int lws_callback_on_writable_all_protocol(const struct lws_context* context, const struct lws_protocols* protocol);
struct lws_context* lws_get_context(const struct lws* wsi);
const struct lws_protocols* lws_get_protocol(struct lws* wsi);
struct lws_context* lws_create_context(struct lws_context_creation_info* info);
int lws_service(struct lws_context* context, int timeout_ms);
void lws_context_destroy(struct lws_context* context);
void lwsl_notice(const char* format, ...);

#ifdef __cplusplus
}
#endif
#endif
