"""Wire formats: the product's host-side emitters against the byte
stream the UNMODIFIED main.c callback produces (oracle/ref_harness_ws.c over the libwebsockets API
stand-in), and against the committed fixture tests/golden/ws_stream.npz made from the same run.
No GPU needed: these exercise b200_wire_spectrum_header / _message / _audio_fragment /
_reference_drain_index only.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as _po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ws_stream.npz")
LWS_BINARY, LWS_CONTINUATION, LWS_NO_FIN = 1, 2, 0x40       # oracle/lws_shim/libwebsockets.h


def split_stream(records):
    """-> (spectrum messages, audio fragments) in order."""
    spec = [b for m, b in records if b.startswith(b"t s;")]
    audio = [(m, b) for m, b in records if not b.startswith(b"t s;")]
    return spec, audio


def expected_audio_writes(wire, audio: np.ndarray, n_fragments: int, drain: bool):
    """The lws_write calls for the first n_fragments fragments of the audio stream `audio`."""
    frags = wire.audio_fragments()
    out = []
    w = 0
    for i in range(n_fragments):
        off, ln, flags = frags[i % wire.AUDIO_FRAGMENTS]
        idx = np.arange(w, w + 512)
        if drain:
            idx = np.array([wire.reference_drain_index(int(j)) for j in idx])
        body = audio[idx].astype("<f4").tobytes()
        out.append((flags, (b"FF;t a;d" if i % wire.AUDIO_FRAGMENTS == 0 else b"") + body))
        assert ln == len(out[-1][1])
        w += 512
    return out


def check_stream(wire, records, payload_rows, audio, freq, rate, gain):
    spec, frags = split_stream(records)
    assert len(spec) >= 2 and len(frags) >= 16
    # spectrum messages: header + one of the driver's payload rows, in order
    hdr = wire.spectrum_header(freq, rate, gain)
    assert hdr == f"t s;f {freq};b {rate};s {gain};d".encode()
    row = -1
    for msg in spec:
        assert msg[:len(hdr)] == hdr and len(msg) == len(hdr) + 1024
        body = np.frombuffer(msg[len(hdr):], dtype=np.uint8)
        hits = [r for r in range(row + 1, len(payload_rows)) if np.array_equal(payload_rows[r], body)]
        assert hits, "a spectrum message carries a payload the driver never published"
        row = hits[0]
        assert wire.spectrum_message(freq, rate, gain, payload_rows[row]) == msg
    # audio fragments: flags, header placement and the drain order, byte for byte
    exp = expected_audio_writes(wire, audio, len(frags), drain=True)
    for i, ((mode, got), (flags, want)) in enumerate(zip(frags, exp)):
        lws = (LWS_BINARY if flags & wire.BINARY else LWS_CONTINUATION) | (LWS_NO_FIN if flags & wire.NO_FIN else 0)
        assert mode == lws, f"fragment {i}: write mode {mode:#x} vs {lws:#x}"
        assert got == want, f"fragment {i} differs"


@pytest.fixture(scope="module")
def wire(pkg):
    pkg.lib()
    return pkg.wire


def test_header_and_fragment_table(wire):
    assert wire.spectrum_header(100000000, 2048000, 0) == b"t s;f 100000000;b 2048000;s 0;d"
    assert wire.spectrum_header(4294967295, 4294967295, -2147483648) == b"t s;f 4294967295;b 4294967295;s -2147483648;d"
    f = wire.audio_fragments()
    assert f[0] == (0, 2056, wire.BINARY | wire.NO_FIN)
    assert f[1] == (2056, 2048, wire.CONTINUATION | wire.NO_FIN)
    assert f[6] == (8 + 6 * 2048, 2048, wire.CONTINUATION | wire.NO_FIN)
    assert f[7] == (8 + 7 * 2048, 2048, wire.CONTINUATION)
    assert sum(x[1] for x in f) == wire.AUDIO_MESSAGE_BYTES


def test_header_too_small_is_an_error(pkg, wire):
    import ctypes as C
    buf = C.create_string_buffer(16)
    assert pkg.lib().b200_wire_spectrum_header(buf, 16, 100000000, 2048000, 0) == -1
    assert b"do not fit" in pkg.lib().b200_last_error()


def test_drain_index_map(wire):
    L = wire.POOL_BUFFER_LEN
    assert [wire.reference_drain_index(w) for w in (0, 511, 512, L - 1)] == [0, 511, 512, L - 1]
    # every buffer after the first: its first 512 samples are the previous buffer's first 512
    assert wire.reference_drain_index(L) == 0 and wire.reference_drain_index(L + 511) == 511
    assert wire.reference_drain_index(L + 512) == L + 512
    assert wire.reference_drain_index(3 * L + 100) == 2 * L + 100
    assert wire.reference_drain_index(-1) == -1


@pytest.mark.skipif(not _po.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_emitters_match_unmodified_main_c(po, synth, wire):
    iq = synth.s3_fm(131072 * 12, seed=77)
    records = po.Ref().ws_run(iq, commands=("spectrumgain 17", "freq 99900", "start"))
    drv = po.Ref().cbb_run(iq, gain_db=17)
    check_stream(wire, records, drv["payload"], drv["audio"], 99900000, 2048000, 17)


@pytest.mark.skipif(not _po.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_nothing_is_sent_before_start_and_after_stop(po, synth):
    iq = synth.s3_fm(131072 * 4, seed=78)
    assert po.Ref().ws_run(iq, commands=()) == []
    assert po.Ref().ws_run(iq, commands=("start", "stop")) == []


def test_emitters_match_golden_stream(synth, wire):
    g = np.load(GOLDEN)
    records = [(int(m), g["bytes"][o:o + n].tobytes()) for o, n, m in g["records"]]
    check_stream(wire, records, g["payload"], g["audio"], int(g["freq"]), int(g["rate"]), int(g["gain"]))
