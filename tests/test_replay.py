"""The replay sensor: libb200replay.so implements the reference's rtl_sensor.h
over a capture.  No GPU here: chunking, ring behaviour, cancel, loops, pacing, several virtual dongles at
once, and the reference's UNMODIFIED signal_source.c running on top of it (oracle/_ref/libreplay_rtlws.so)."""
import ctypes as C
import os
import re
import threading
import time

import numpy as np
import pytest

from oracle import pyoracle as _po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUF = 262144


@pytest.fixture(scope="module")
def rp(pkg):
    return pkg.replay


def test_exports_the_reference_sensor_interface(rp):
    lib = rp.lib()
    text = open(os.path.join(ROOT, "include", "rtl_sensor_replay.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    declared = set(re.findall(r"\b((?:rtl|b200_replay)_[a-z_0-9]+)\s*\(", text))
    assert declared == set(rp.EXPORTED_SYMBOLS)
    assert not [n for n in declared if not hasattr(lib, n)]


def test_defaults_and_setters_follow_rtl_sensor_c(rp):
    d = rp.VirtualDongle(0, np.zeros(BUF, dtype=np.uint8))
    L = rp.lib()
    assert (L.rtl_sample_rate(d.dev), L.rtl_freq(d.dev), L.rtl_gain(d.dev)) == (2048000, 100000000, 25.4)
    assert L.rtl_set_frequency(d.dev, 99900000) == 0 and L.rtl_freq(d.dev) == 99900000
    assert L.rtl_set_sample_rate(d.dev, 1024000) == 0 and L.rtl_sample_rate(d.dev) == 1024000
    assert L.rtl_set_gain(d.dev, 12.5) == 0 and L.rtl_gain(d.dev) == 12.5
    d.close()
    dev = C.c_void_p()
    assert L.rtl_init(C.byref(dev), -1) == -1 and L.rtl_init(C.byref(dev), 1 << 20) == -1


def test_whole_buffers_in_order_tail_dropped_and_loops(rp):
    rng = np.random.default_rng(1)
    iq = rng.integers(0, 256, BUF * 4 + 12345, dtype=np.uint8)
    d = rp.VirtualDongle(1, iq, loops=2)
    got, addrs = [], []

    def on(b):
        got.append(b.copy())
        addrs.append(b.ctypes.data)
    assert d.read_async(on) == 0
    assert len(got) == 8 and all(len(b) == BUF for b in got)
    for i, b in enumerate(got):
        assert np.array_equal(b, iq[BUF * (i % 4):BUF * (i % 4 + 1)])
    assert d.delivered_bytes() == 8 * BUF
    assert len(set(addrs)) == 8          # eight different ring buffers of the fifteen
    d.close()
    # shorter than one buffer: nothing is delivered (the reference's own stub returns at once too)
    d = rp.VirtualDongle(1, iq[:1000])
    assert d.read_async(lambda b: got.append(b)) == 0 and len(got) == 8
    d.close()


def test_ring_wraps_after_fifteen_buffers(rp):
    iq = np.arange(BUF * 17, dtype=np.uint32).astype(np.uint8)
    d = rp.VirtualDongle(2, iq)
    addrs = []
    d.read_async(lambda b: addrs.append(b.ctypes.data))
    assert len(addrs) == 17 and len(set(addrs)) == 15 and addrs[15] == addrs[0] and addrs[16] == addrs[1]
    d.close()


def test_cancel_stops_an_endless_replay_and_pacing_is_real_time(rp):
    iq = np.zeros(BUF * 2, dtype=np.uint8)
    d = rp.VirtualDongle(3, iq, loops=0)
    n = [0]

    def on(b):
        n[0] += 1
        if n[0] == 40:
            d.cancel()
    assert d.read_async(on) == 0 and n[0] == 40
    d.close()
    d = rp.VirtualDongle(3, iq, loops=3, realtime=True)       # 6 buffers of 64 ms at 2.048 MS/s
    t0 = time.perf_counter()
    d.read_async(lambda b: None)
    dt = time.perf_counter() - t0
    assert 0.38 <= dt < 1.0, dt
    d.close()


def test_virtual_dongles_run_side_by_side(rp):
    rng = np.random.default_rng(2)
    caps = [rng.integers(0, 256, BUF * 3, dtype=np.uint8) for _ in range(6)]
    dongles = [rp.VirtualDongle(10 + i, c) for i, c in enumerate(caps)]
    got = [[] for _ in caps]
    threads = [threading.Thread(target=d.read_async, args=(lambda b, g=g: g.append(b.copy()),)) for d, g in zip(dongles, got)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for c, g in zip(caps, got):
        assert np.array_equal(np.concatenate(g), c)
    [d.close() for d in dongles]


@pytest.mark.skipif(not os.path.exists(_po.REPLAY_SO), reason="oracle/_ref/libreplay_rtlws.so not built")
def test_unmodified_signal_source_over_the_replay_sensor(rp):
    """signal_source.c's worker thread, callback list and fan-out (signal_source.c:29-70), unmodified, reading
    from the replay sensor: every registered callback sees every buffer as cmplx_u8 samples, in order."""
    rng = np.random.default_rng(3)
    iq = rng.integers(0, 256, BUF * 5, dtype=np.uint8)
    L = rp.lib(mode=C.RTLD_GLOBAL)
    ss = C.CDLL(_po.REPLAY_SO)
    CB = C.CFUNCTYPE(None, C.POINTER(C.c_ubyte), C.c_int)
    seen_a, seen_b = [], []
    cb_a = CB(lambda p, n: seen_a.append(np.ctypeslib.as_array(p, shape=(2 * n,)).copy()))
    cb_b = CB(lambda p, n: seen_b.append(n))
    d = rp.VirtualDongle(20, iq)
    L.b200_replay_gate(20, 0)
    ss.signal_source_start.argtypes = [C.c_void_p]
    ss.signal_source_start(d.dev)
    ss.signal_source_add_callback(cb_a)
    ss.signal_source_add_callback(cb_b)
    L.b200_replay_gate(20, 1)
    deadline = time.time() + 10
    while d.delivered_bytes() < iq.size and time.time() < deadline:
        time.sleep(0.005)
    ss.signal_source_stop()
    assert seen_b == [BUF // 2] * 5
    assert np.array_equal(np.concatenate(seen_a), iq)
    d.close()
