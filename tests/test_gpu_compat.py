"""GPU parity of the reference-named interface (include/rtlws_compat.h) and of the
host-buffer session: the calls an unmodified cbb_main.c / audio_main.c would make."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_spectrum_add_semantics(pkg, cuda, po, synth):
    iq = synth.s2_tones(1024 * 6, seed=3)
    s = pkg.spectrum_alloc(1024)
    ps = np.zeros(1024)
    for f in range(6):                                   # cbb_main.c:50-59
        assert pkg.spectrum_add_cmplx_u8(s, iq[1024 * f:1024 * (f + 1)], ps, 1024) == 0
    want = po.Spectrum(1024).rows(iq, K=6)[0]
    np.testing.assert_allclose(ps, want, rtol=1e-4, atol=1e-4 * want.mean())
    # accumulate into a dirty caller buffer, then the wrong length
    rng = np.random.default_rng(4)
    init = rng.uniform(0, 50, 1024)
    a = init.copy()
    assert pkg.spectrum_add_cmplx_u8(s, iq[:1024], a, 1024) == 0
    b = init.copy()
    po.Spectrum(1024).add_cmplx_u8(iq[:1024], b)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4 * b.mean())
    assert np.isclose(a[512], init[512] + a[511])
    keep = a.copy()
    assert pkg.spectrum_add_cmplx_u8(s, iq[:512], a, 512) == -1        # spectrum.c:51-52
    assert np.array_equal(a, keep)
    # the two caller-less input types
    x32 = iq[:1024].astype(np.int32) - 128
    c = np.zeros(1024)
    assert pkg.spectrum_add_cmplx_s32(s, x32, c, 1024) == 0
    d = np.zeros(1024)
    po.Spectrum(1024).add_cmplx_s32(x32, d)
    np.testing.assert_allclose(c, d, rtol=1e-4, atol=1e-4 * d.mean())
    xr = rng.standard_normal(1024).astype(np.float32)
    e = np.zeros(1024)
    assert pkg.spectrum_add_real_f32(s, xr, e, 1024) == 0
    f = np.zeros(1024)
    po.Spectrum(1024).add_real_f32(xr, f)
    np.testing.assert_allclose(e, f, rtol=1e-4, atol=1e-4 * f.mean())
    pkg.spectrum_free(s)
    s4 = pkg.spectrum_alloc(4096)
    iq4 = synth.s2_tones(4096, N=4096)
    g = np.zeros(4096)
    assert pkg.spectrum_add_cmplx_u8(s4, iq4, g, 4096) == 0
    np.testing.assert_allclose(g, po.Spectrum(4096).rows(iq4)[0], rtol=1e-4, atol=1e-4 * g.mean())
    pkg.spectrum_free(s4)


@pytest.mark.parametrize("R", [1, 5, 10, 16])
def test_cic_decimate_bit_exact_with_state(pkg, cuda, po, synth, R):
    iq = synth.s1_noise(R * 700, seed=R)
    cut = R * 300
    delay = pkg.CicDelayLine()
    st = po.CicState()
    for part in (iq[:cut], iq[cut:]):
        dst = np.zeros((len(part) // R, 2), np.int32)
        assert pkg.cic_decimate(R, part, len(part), dst, len(dst), delay) == 0
        _, want, st = po.cic_decimate(R, part, st)
        assert np.array_equal(dst, want)
        assert [delay.integrator_prev_out.re, delay.integrator_prev_out.im] == list(st.integrator_prev_out)
        assert [delay.comb_prev_in.re, delay.comb_prev_in.im] == list(st.comb_prev_in)
    dst = np.zeros((10, 2), np.int32)
    assert pkg.cic_decimate(R, iq[:R * 10], R * 10, dst, 9, delay) == -1      # resample.c:18-19
    # a caller-set state with integrator != comb input only shifts the first output
    odd = pkg.CicDelayLine()
    odd.integrator_prev_out.re, odd.comb_prev_in.re = 1000, 400
    ost = po.CicState()
    ost.integrator_prev_out[0], ost.comb_prev_in[0] = 1000, 400
    assert pkg.cic_decimate(R, iq[:R * 10], R * 10, dst, 10, odd) == 0
    _, want, ost = po.cic_decimate(R, iq[:R * 10], ost)
    assert np.array_equal(dst, want) and odd.integrator_prev_out.re == ost.integrator_prev_out[0]


def test_halfband_decimate_with_delay_line(pkg, cuda, po):
    rng = np.random.default_rng(5)
    x = rng.standard_normal(4096).astype(np.float32)
    d_gpu = np.zeros(10, np.float32)
    d_cpu = np.zeros(10, np.float32)
    got, want = [], []
    for lo, hi in ((0, 1000), (1000, 1020), (1020, 4096)):
        out = np.zeros((hi - lo) // 2, np.float32)
        pkg.halfband_decimate(x[lo:hi], out, len(out), d_gpu)
        got.append(out)
        want.append(po.halfband_decimate(x[lo:hi], d_cpu))
        assert np.array_equal(d_gpu, d_cpu)
    assert np.abs(np.concatenate(got) - np.concatenate(want)).max() <= 1e-6
    imp = np.zeros(64, np.float32)
    imp[0] = 1
    out = np.zeros(32, np.float32)
    pkg.halfband_decimate(imp, out, 32, np.zeros(10, np.float32))
    np.testing.assert_allclose(out[:6], [0.01824, -0.11614, 0.34790, 0.34790, -0.11614, 0.01824], rtol=1e-6)


def test_rf_decimator_reblocking_and_callbacks(pkg, cuda, po, synth):
    iq = synth.s3_fm(2 * 204800 + 5000, seed=8)
    d = pkg.RfDecimator()
    seen = []
    order = []
    d.add_callback(lambda sig, n: (seen.append(sig.copy()), order.append("a")))
    d.add_callback(lambda sig, n: order.append("b"))
    assert d.decimate_cmplx_u8(iq[:100]) == -1                     # rf_decimator.c:90-91
    assert d.set_parameters(0.0, 10) == -1 and d.set_parameters(2048000.0, 0) == -1
    assert d.set_parameters(2048000.0, 10) == 0
    for pos in range(0, len(iq), 131072):                          # source-buffer sized pushes
        assert d.decimate_cmplx_u8(iq[pos:pos + 131072]) == 0
    assert [len(s) for s in seen] == [20480, 20480] and order == ["a", "b", "a", "b"]
    want, _ = po.chain_run(iq)
    assert np.array_equal(np.concatenate(seen), want)
    d.remove_callbacks()
    assert d.decimate_cmplx_u8(iq[:204800]) == 0 and len(seen) == 2
    # main.c:149-155: changing the rate re-derives the block sizes and forgets the surplus
    assert d.set_parameters(1024000.0, 5) == 0
    d.add_callback(lambda sig, n: seen.append(sig.copy()))
    assert d.decimate_cmplx_u8(iq[:102400]) == 0
    assert len(seen) == 3 and len(seen[2]) == 20480
    d.free()


def test_session_host_buffers(pkg, cuda, po, synth):
    torch = cuda
    n_streams, n = 6, 5120 * 6
    iq = np.stack([synth.s3_fm(2 * n, seed=90 + s) for s in range(n_streams)])
    sess = pkg.Session(n_streams, n)
    h_iq = torch.empty((n_streams, n, 2), dtype=torch.uint8).pin_memory()
    h_db = torch.empty((n_streams, n // 1024, 1024), dtype=torch.float32).pin_memory()
    h_audio = torch.empty((n_streams, n // 40), dtype=torch.float32).pin_memory()
    audio_parts = []
    for b in range(2):                                             # two batches: history carries inside
        h_iq.copy_(torch.as_tensor(iq[:, b * n:(b + 1) * n]))
        sess.chain(h_iq, n, h_db, h_audio)
        audio_parts.append(h_audio.numpy().copy())
        for s in range(n_streams):
            rows = po.Spectrum(1024).rows(iq[s, b * n:(b + 1) * n])
            ok = rows > 1e-6 * rows.mean(axis=1, keepdims=True)
            assert np.abs(h_db.numpy()[s][ok] - 10 * np.log10(rows[ok])).max() <= 0.01
    got = np.concatenate(audio_parts, axis=1)
    for s in range(n_streams):
        _, dec, _ = po.cic_decimate(10, iq[s])
        _, _, want, _ = po.fm_demodulate(dec)
        assert np.abs(got[s] - want).max() <= 1e-4
    # pageable host memory works too
    sess.reset()
    db2 = np.empty((n_streams, n // 1024, 1024), np.float32)
    au2 = np.empty((n_streams, n // 40), np.float32)
    sess.chain(np.ascontiguousarray(iq[:, :n]), n, db2, au2)
    assert np.array_equal(au2, audio_parts[0])
    with pytest.raises(pkg.B200Error):
        sess.chain(h_iq, 5000, h_db, h_audio)
    sess.close()


def test_host_alloc_is_pinned_and_sm_count(pkg, cuda):
    """b200_host_alloc hands out page-locked memory (what makes the session's copies asynchronous); b200_sm_count is the
    device's SM count (148 on a B200: the persistent kernels size their grids with it)."""
    import ctypes as C
    torch = cuda
    L = pkg.lib()
    ptr = L.b200_host_alloc(1 << 20)
    assert ptr
    buf = (C.c_ubyte * (1 << 20)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.uint8)
    arr[:] = 7
    t = torch.from_numpy(arr)
    d = t.cuda(non_blocking=True)
    torch.cuda.synchronize()
    assert int(d.sum()) == 7 << 20
    del t, d, arr, buf
    L.b200_host_free(ptr)
    assert L.b200_sm_count() == torch.cuda.get_device_properties(0).multi_processor_count


def test_session_products_host_buffers(pkg, cuda, po, synth):
    """b200_session_products (what bench.py times as e2e_products): host IQ in, the reference's OWN products back --
    the FM audio and, per stream, the payload bytes of the 6-frame average at the start of the batch
    (cbb_main.c:40-70 + 106-135) -- two batches with the history carried inside, a gain, and argument errors."""
    torch = cuda
    n_streams, n, gain = 5, 5120 * 8, 17
    iq = np.stack([synth.s2_tones(2 * n, seed=400 + s) for s in range(n_streams)])
    sess = pkg.Session(n_streams, n)
    h_iq = torch.empty((n_streams, n, 2), dtype=torch.uint8).pin_memory()
    h_audio = torch.empty((n_streams, n // 40), dtype=torch.float32).pin_memory()
    h_pay = torch.empty((n_streams, 1024), dtype=torch.uint8).pin_memory()
    audio_parts = []
    for b in range(2):
        h_iq.copy_(torch.as_tensor(iq[:, b * n:(b + 1) * n]))
        sess.products(h_iq, n, h_audio, h_pay, gain_db=gain, K_avg=6)
        audio_parts.append(h_audio.numpy().copy())
        for s in range(n_streams):
            rows = po.Spectrum(1024).rows(iq[s, b * n:b * n + 6 * 1024], K=6)
            want, dbf = po.db_payload(rows[0], 6, gain)
            diff = h_pay.numpy()[s] != want
            assert diff.mean() < 0.01 and (np.abs(dbf[diff] - np.rint(dbf[diff])) <= 0.01).all()
    got = np.concatenate(audio_parts, axis=1)
    for s in range(n_streams):
        _, dec, _ = po.cic_decimate(10, iq[s])
        _, _, want, _ = po.fm_demodulate(dec)
        assert np.abs(got[s] - want).max() <= 1e-4
    # the same batches through b200_session_chain give the same audio, bit for bit
    sess.reset()
    h_db = torch.empty((n_streams, n // 1024, 1024), dtype=torch.float32).pin_memory()
    h_iq.copy_(torch.as_tensor(iq[:, :n]))
    sess.chain(h_iq, n, h_db, h_audio, gain_db=gain)
    assert np.array_equal(h_audio.numpy(), audio_parts[0])
    with pytest.raises(pkg.B200Error):
        sess.products(h_iq, n, h_audio, None)                  # the payload buffer is what the call is for
    with pytest.raises(pkg.B200Error):
        sess.products(h_iq, n, h_audio, h_pay, K_avg=0)
    sess.close()


@pytest.mark.parametrize("R", [12, 5, 16])
def test_session_and_stream_other_down_factors(pkg, cuda, po, synth, R):
    """The reference re-derives R = fs / 192000 when the client sends `bw` (main.c:152-155): the fast-path
    entry points take it too.  Two batches through a session and chunked pushes through a stream."""
    torch = cuda
    tile = int(pkg.lib().b200_chain_tile_samples(R))
    assert tile % 1024 == 0 and tile % (4 * R) == 0
    n_streams, n = 3, tile * 4
    iq = np.stack([synth.s3_fm(2 * n, seed=120 + s) for s in range(n_streams)])
    want_audio = []
    for s in range(n_streams):
        _, dec, _ = po.cic_decimate(R, iq[s])
        _, _, a, _ = po.fm_demodulate(dec)
        want_audio.append(a)
    sess = pkg.Session(n_streams, n, R=R)
    h_db = np.empty((n_streams, n // 1024, 1024), np.float32)
    h_audio = np.empty((n_streams, n // (4 * R)), np.float32)
    parts = []
    for b in range(2):
        sess.chain(np.ascontiguousarray(iq[:, b * n:(b + 1) * n]), n, h_db, h_audio)
        parts.append(h_audio.copy())
        for s in range(n_streams):
            rows = po.Spectrum(1024).rows(iq[s, b * n:(b + 1) * n])
            ok = rows > 1e-6 * rows.mean(axis=1, keepdims=True)
            assert np.abs(h_db[s][ok] - 10 * np.log10(rows[ok])).max() <= 0.01
    got = np.concatenate(parts, axis=1)
    for s in range(n_streams):
        assert np.abs(got[s] - want_audio[s]).max() <= 1e-4
    with pytest.raises(pkg.B200Error):
        sess.chain(np.ascontiguousarray(iq[:, :tile + 8]), tile + 8, h_db, h_audio)
    sess.close()
    ps = pkg.PushStream(n_streams, tile * 2, R=R)
    pos = [0] * n_streams
    rng = np.random.default_rng(R)
    while any(p < 2 * n for p in pos):
        s = int(rng.integers(n_streams))
        k = min(int(rng.integers(1, 40000)), 2 * n - pos[s])
        ps.push(s, iq[s, pos[s]:pos[s] + k])
        pos[s] += k
    ps.flush()
    for s in range(n_streams):
        a = np.concatenate([x for _, x in sorted(ps.audio[s], key=lambda t: t[0])])
        assert a.shape == want_audio[s].shape and np.abs(a - want_audio[s]).max() <= 1e-4
        f = np.concatenate([x for _, x in sorted(ps.spectra[s], key=lambda t: t[0])])
        assert f.shape == (2 * n // 1024, 1024)
    ps.close()
    with pytest.raises(pkg.B200Error):
        pkg.Session(1, 5120, R=300)


def test_push_stream_groups_and_stragglers(pkg, cuda, po, synth):
    """Full batches are submitted in runs when a whole group of (up to 16) streams is ready; a stream that runs
    ahead of its group goes out on its own, and poll() submits what is waiting.  20 streams = two groups; stream 5
    pushes three batches before anybody else starts, stream 18 never finishes its last batch."""
    n_streams, batch = 20, 5120 * 4
    n = batch * 3
    iq = np.stack([synth.s3_fm(n, seed=160 + s) for s in range(n_streams)])
    ps = pkg.PushStream(n_streams, batch)
    before = pkg.launch_count()
    ps.push(5, iq[5])                                    # three batches: the third forces the first out alone
    assert ps.pending(5) == 0
    for b in range(3):
        for s in range(n_streams):
            if s == 5:
                continue
            hi = (b + 1) * batch if not (s == 18 and b == 2) else (b + 1) * batch - 1000
            ps.push(s, iq[s, b * batch:hi])
        ps.poll()
    ps.flush()
    launches = pkg.launch_count() - before
    assert launches < 2 * 3 * n_streams, launches          # far fewer than two launches per stream and batch
    assert ps.pending(18) == batch - 1000
    for s in range(n_streams):
        n_done = 2 * batch if s == 18 else n
        _, dec, _ = po.cic_decimate(10, iq[s, :n_done])
        _, _, want, _ = po.fm_demodulate(dec)
        a = np.concatenate([x for _, x in sorted(ps.audio[s], key=lambda t: t[0])])
        assert a.shape == want.shape and np.abs(a - want).max() <= 1e-4
        firsts = [f for f, _ in ps.spectra[s]]
        assert firsts == sorted(firsts) and len(firsts) == n_done // batch      # delivered in order, once each
    ps.close()


def test_push_stream_payload_sink(pkg, cuda, po, synth):
    """What the reference sends its client: the 6-frame average at the start of every batch as payload bytes
    (cbb_main.c:48-59,121-130), computed on the device; the per-frame rows are neither computed nor copied."""
    n_streams, batch, n_batches, gain = 2, 204800, 3, 17
    iq = np.stack([synth.s2_tones(batch * n_batches, seed=140 + s) for s in range(n_streams)])
    ps = pkg.PushStream(n_streams, batch, gain_db=gain, payload_K=6, frames=False)
    for s in range(n_streams):
        for pos in range(0, batch * n_batches, 131072):
            ps.push(s, iq[s, pos:pos + 131072])
    ps.flush()
    for s in range(n_streams):
        assert ps.spectra[s] == [] and len(ps.payloads[s]) == n_batches
        for b, (first, k, payload) in enumerate(sorted(ps.payloads[s], key=lambda t: t[0])):
            assert first == b * (batch // 1024) and k == 6
            rows = po.Spectrum(1024).rows(iq[s, b * batch:b * batch + 6 * 1024], K=6)
            want, dbf = po.db_payload(rows[0], 6, gain)
            diff = payload != want
            assert diff.mean() < 0.01 and (np.abs(dbf[diff] - np.rint(dbf[diff])) <= 0.01).all()
        _, dec, _ = po.cic_decimate(10, iq[s])
        _, _, want_audio, _ = po.fm_demodulate(dec)
        a = np.concatenate([x for _, x in sorted(ps.audio[s], key=lambda t: t[0])])
        assert np.abs(a - want_audio).max() <= 1e-4
    ps.close()


def test_push_stream_spectrum_sink_installed_late(pkg, cuda, po, synth):
    """The per-frame dB buffers are allocated when a spectrum sink first arrives.  A stream that ran payload-only and
    gets its sink after two batches delivers rows for the later batches only -- never uninitialised memory for the
    earlier ones -- and audio and payloads are unaffected."""
    batch, n_batches = 204800, 4
    iq = synth.s2_tones(batch * n_batches, seed=171)
    ps = pkg.PushStream(1, batch, payload_K=6, frames=False)
    ps.push(0, iq[:2 * batch])
    ps.flush()
    assert ps.spectra[0] == [] and len(ps.payloads[0]) == 2
    ps.enable_frames()
    ps.push(0, iq[2 * batch:])
    ps.flush()
    assert [f for f, _ in ps.spectra[0]] == [2 * (batch // 1024), 3 * (batch // 1024)]
    assert len(ps.payloads[0]) == 4 and len(ps.audio[0]) == 4
    db = np.concatenate([r for _, r in ps.spectra[0]])
    rows = po.Spectrum(1024).rows(iq[2 * batch:])
    ok = rows > 1e-6 * rows.mean(axis=1, keepdims=True)
    assert np.abs(db[ok] - 10 * np.log10(rows[ok])).max() <= 0.01
    ps.close()


def test_push_stream_like_a_signal_source_callback(pkg, cuda, po, synth):
    """Three dongles pushing 131072-sample source buffers (signal_source.c:31) in round robin;
    batches of one reference block (204800 samples); sinks receive everything in order."""
    n_streams, batch = 3, 204800
    n = 3 * batch + 50_000                      # the last 50 000 samples stay pending
    iq = [synth.s3_fm(n, seed=300 + s) for s in range(n_streams)]
    ps = pkg.PushStream(n_streams, batch)
    for pos in range(0, n, 131072):
        for s in range(n_streams):
            ps.push(s, iq[s][pos:pos + 131072])
    ps.flush()
    for s in range(n_streams):
        assert ps.pending(s) == 50_000
        assert [f for f, _ in ps.spectra[s]] == [0, 200, 400]
        assert [f for f, _ in ps.audio[s]] == [0, 5120, 10240]
        db = np.concatenate([r for _, r in ps.spectra[s]])
        audio = np.concatenate([a for _, a in ps.audio[s]])
        rows = po.Spectrum(1024).rows(iq[s][:3 * batch])
        ok = rows > 1e-6 * rows.mean(axis=1, keepdims=True)
        assert np.abs(db[ok] - 10 * np.log10(rows[ok])).max() <= 0.01
        _, want = po.chain_run(iq[s][:3 * batch])
        assert np.abs(audio - want).max() <= 1e-4
    # chunking independence: one stream, odd pushes
    ps2 = pkg.PushStream(1, batch)
    for pos in range(0, n, 7777):
        ps2.push(0, iq[0][pos:pos + 7777])
    ps2.flush()
    a2 = np.concatenate([a for _, a in ps2.audio[0]])
    assert np.array_equal(a2, np.concatenate([a for _, a in ps.audio[0]]))
    with pytest.raises(pkg.B200Error):
        pkg.PushStream(1, 1000)
    ps.close()
    ps2.close()
