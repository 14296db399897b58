"""The reference GUI with its 'GPU' receiver mode.

In the build container /root/reference is mounted: gui_patch edits a copy of the REAL
scripts/fft_analyzer_gui.py, the copy is imported with stand-ins for flask / flask_socketio / PyQt5 /
serial (tests/stubs/gui_stubs.py), and its own ReceiverController, UdpReceiver,
MultiPacketAssembler, decode and rate-limit code run against frames produced by libfra.  Without a
GPU the library behind FraContext is the host-emulated build of the same sources
(tests/emul) - test infrastructure, swapped in by the test, never by the product.  The -m gpu
tests at the bottom do the transport half on real GPU frames (the GUI source does not travel to
the GPU box)."""
import os
import random
import socket

import numpy as np
import pytest

from oracle import golden as g

GUI_PATH = "/root/reference/scripts/fft_analyzer_gui.py"
HTML_PATH = "/root/reference/scripts/templates/index.html"
needs_reference = pytest.mark.skipif(not os.path.exists(GUI_PATH), reason="reference GUI not mounted (GPU box)")


@pytest.fixture()
def emulated_lib(monkeypatch):
    """FraContext on the host-emulated libfra (CPU container only)."""
    import torch
    from fpga_real_time_fft_analyzer_b200 import _lib, context
    from tests.emul import emul_lib
    monkeypatch.setattr(_lib, "_lib", emul_lib.lib())
    monkeypatch.setattr(context.FraContext, "pinned",
                        lambda self, name, shape, dtype: self._pinned.setdefault((name, tuple(shape), dtype),
                                                                                 torch.empty(shape, dtype=dtype)))
    yield
    monkeypatch.setattr(_lib, "_lib", None)


@pytest.fixture()
def gui(emulated_lib):
    from fpga_real_time_fft_analyzer_b200 import gui_patch
    from tests.stubs import gui_stubs
    src = gui_patch.patch_gui_source(open(GUI_PATH).read())
    gui_stubs.EMITTED.clear()
    gui_stubs.TIMERS.clear()
    mod = gui_stubs.load_gui(src)
    mod._stubs = gui_stubs
    return mod


def batches(channels, n_batches, seed=0):
    return [g.tone_noise(range(channels), n=16384, seed=seed + i) for i in range(n_batches)]


@needs_reference
def test_patch_anchors_and_html():
    from fpga_real_time_fft_analyzer_b200 import gui_patch
    src = open(GUI_PATH).read()
    out = gui_patch.patch_gui_source(src)
    assert "class GpuQtReceiver(QtCore.QObject)" in out and 'elif mode == "GPU":' in out
    assert out.count("(UartReceiver, GpuQtReceiver)") >= 4 and "isinstance(self.current_receiver, UartReceiver)" not in out
    # nothing but insertions and the isinstance / transport edits: every original line is still there
    kept = [ln for ln in src.splitlines() if "isinstance(" not in ln and "command = ETHERNET_MODE_CMD" not in ln]
    patched_lines = set(out.splitlines())
    assert all(ln in patched_lines for ln in kept)
    with pytest.raises(gui_patch.PatchError):
        gui_patch.patch_gui_source(out)                       # already patched
    with pytest.raises(gui_patch.PatchError):
        gui_patch.patch_gui_source("x = 1\n")
    html = gui_patch.patch_index_html(open(HTML_PATH).read())
    assert '<option value="GPU">' in html and html.count("<option value=") == open(HTML_PATH).read().count("<option value=") + 1


@needs_reference
def test_receiver_controller_gpu_mode_frames_stats_and_commands(gui, rom):
    xs = batches(2, 6)
    it = iter(xs)
    gui.web_config.update({"gpu_source": lambda: next(it), "gpu_channels": 2, "gpu_display_channel": 1, "comm_mode": "GPU"})
    ctrl = gui.ReceiverController()
    ctrl.start_receiver("GPU")
    rx = ctrl.current_receiver
    assert type(rx).__name__ == "GpuQtReceiver" and gui.receiver_state["is_active"]
    status = [p for e, p in gui._stubs.EMITTED if e == "receiver_status"]
    assert status[-1] == {"active": True, "mode": "GPU", "message": "GPU receiver started successfully"}
    # the QTimer the receiver started: one tick = one batch = one frame_data event (first frame is always shown)
    timer = rx.read_timer
    timer.fire()
    frames = [p for e, p in gui._stubs.EMITTED if e == "frame_data"]
    assert len(frames) == 1
    p = frames[0]
    for key in ("frequency", "data", "incoming_fps", "display_fps", "frames_received", "frames_displayed", "frames_dropped",
                "packet_count", "peak_magnitude", "peak_frequency", "peak_bin", "timestamp", "receiver_active",
                "freq_range_start", "freq_range_end"):              # the payload of GUI:439-455
        assert key in p, key
    assert p["frames_received"] == 1 and p["frames_displayed"] == 1 and p["receiver_active"] is True
    # what is plotted = the GUI's own decode of the frame libfra made: window (mode 0xB1 after reset) -> FFT / N
    w = g.window(xs[0][1][None], rom)
    rq, iq = g.quantize_bins(np.fft.fft(w.astype(np.float64), axis=-1), -14)
    want = np.sqrt(rq[0].astype(np.float32) ** 2 + iq[0].astype(np.float32) ** 2)
    got = np.array(p["data"]["magnitude"], dtype=np.float32)
    assert got.shape == (16384,) and np.abs(got - want).max() <= 2.0          # bins within 1 LSB each
    assert p["peak_bin"] == int(np.argmax(got))
    # the 30 FPS limiter of the GUI (should_display_frame, GUI:281-292) applies to GPU frames: a frame that arrives
    # less than 1/30 s after the last displayed one is received but dropped (the emulated kernels take longer
    # than that, so the last display time is moved instead of racing the clock)
    gui.receiver_state["last_display_time"] = gui.time.time() + 60.0
    timer.fire()
    assert gui.receiver_state["frames_received"] == 2 and gui.receiver_state["frames_dropped"] == 1
    assert len([1 for e, _ in gui._stubs.EMITTED if e == "frame_data"]) == 1
    # commands go to the GPU backend through the controller's own slots (the isinstance sites)
    ctrl.send_fpga_command(gui.FILTER_DEFAULT_CMD)
    assert rx.rx.ctx.mode == 0x00
    assert [p for e, p in gui._stubs.EMITTED if e == "receiver_status"][-1]["message"] == "Default Filter command sent successfully (0x00)"
    sections = [[0, 1, 0, 64, -67, 19], [64, 127, 64, 64, -85, 40]]
    ctrl.send_filter_coeff_signal.emit(sections)                              # the Qt signal the socket handler uses
    assert list(rx.rx.ctx.bank(1)) == [v for s in sections for v in s]
    assert "uploaded successfully" in [p for e, p in gui._stubs.EMITTED if e == "receiver_status"][-1]["message"]
    ctrl.send_fpga_command(gui.FILTER_CUSTOM_CMD)
    assert rx.rx.ctx.mode == 0xA1
    ctrl.send_start_commands("UART")                                          # 0x55 then 0xA5 through the backend
    c = rx.rx.ctx.counters()
    assert c["start"] >= 2 and c["request"] == 1
    # the socket.io 'set_mode' handler: stop, three resets, transport byte, restart - all on the GPU backend
    gui.receiver_controller = ctrl
    import types
    gui.time = types.SimpleNamespace(time=gui.time.time, sleep=lambda s: None)    # the handler's settle delays, not the process's clock
    gui.socketio.handlers["set_mode"]()
    assert type(ctrl.current_receiver).__name__ == "GpuQtReceiver" and ctrl.current_receiver is not rx and not rx.active
    assert ctrl.current_receiver.rx.ctx.transport == gui.ETHERNET_MODE_CMD
    ctrl.stop_receiver()
    assert ctrl.current_receiver is None and not gui.receiver_state["is_active"]


@needs_reference
def test_udp_loopback_into_the_reference_udp_receiver(gui, rom):
    """GPU frames -> UdpFrameSender -> a real UDP socket -> the reference's UdpReceiver.process_payload ->
    MultiPacketAssembler -> decode -> frame_data: the unmodified receive side rebuilds and plots the frame."""
    from fpga_real_time_fft_analyzer_b200 import GpuReceiver
    from fpga_real_time_fft_analyzer_b200.udp_emitter import FPGA_SRC_PORT, UdpFrameSender
    assert gui.ETHERNET_PAYLOAD_SIZE == 1025 and gui.PACKETS_PER_FRAME == 64
    xs = batches(1, 2, seed=7)
    it = iter(xs)
    rx = GpuReceiver(lambda: next(it), channels=1)
    rx.send_command(0x00)
    rx.send_ethernet_start()
    sink = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    sink.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 20)
    sink.bind(("127.0.0.1", 0))
    sink.settimeout(2.0)
    try:
        try:
            sender = UdpFrameSender(dst=sink.getsockname(), src_ip="127.0.0.1", src_port=FPGA_SRC_PORT)
        except OSError:
            sender = UdpFrameSender(dst=sink.getsockname(), src_ip="127.0.0.1", src_port=0)
        udp = gui.UdpReceiver("127.0.0.1", 6006)                   # the reference class (its Qt socket is a stand-in)
        sent = []
        for _ in range(2):
            frame = rx.poll()[0]
            sent.append(frame)
            sender.send_frame(frame)
            datagrams = []
            for _ in range(64):
                data, addr = sink.recvfrom(2048)
                assert addr[1] == sender.src_port
                datagrams.append(data)
            random.Random(5).shuffle(datagrams)                    # the assembler orders by count byte
            gui.receiver_state["last_display_time"] = 0.0          # let the limiter show both
            for d in datagrams:
                udp.process_payload(d)
        plotted = [p for e, p in gui._stubs.EMITTED if e == "frame_data"]
        assert len(plotted) == 2 and sender.frames_sent == 2 and sender.packets_sent == 128
        for frame, p in zip(sent, plotted):
            assert np.array_equal(np.array(p["data"]["magnitude"], dtype=np.float32), gui.decode_mag_16iq_le(frame))
        y, _ = __import__("oracle.cgolden", fromlist=["x"]).window_iir(xs[1], rom, 0x00, g.BANK0_COEFF, g.BANK0_COEFF,
                                                                       __import__("oracle.cgolden", fromlist=["x"]).window_iir(xs[0], rom, 0x00, g.BANK0_COEFF, g.BANK0_COEFF)[1])
        rq, iq = g.quantize_bins(np.fft.fft(y.astype(np.float64), axis=-1), -14)
        re, im = gui.decode_iq_components(sent[1])
        assert np.abs(re - rq[0]).max() <= 1 and np.abs(im - iq[0]).max() <= 1
        sender.close()
    finally:
        sink.close()
        rx.stop()


def test_frame_gate_drops_and_backlog(emulated_lib):
    """a9: frames between hand-shakes are dropped (IMP/sequencer_dsp.vhd:50-82) - a free-running receiver
    keeps filtering while the gate is closed and counts what nobody took; a slow consumer's backlog is trimmed."""
    from fpga_real_time_fft_analyzer_b200 import GpuReceiver
    from oracle import cgolden as cg
    rom = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "hann_rom.i16"), dtype="<i2")
    xs = batches(2, 5, seed=3)
    it = iter(xs)
    rx = GpuReceiver(lambda: next(it), channels=2, free_running=True, max_backlog=2)
    try:
        rx.send_command(0x00)
        assert rx.poll() == [] and rx.stats["frames_dropped"] == 0          # not started: the sequencer is idle
        rx.send_command(0xFE)                                                # UART transport: one batch per 0xA5
        rx.send_start_sequence()
        assert len(rx.poll()) == 2 and rx.stats["frames_received"] == 2      # batch 0 shown
        assert rx.poll() == [] and rx.poll() == []                           # batches 1, 2: filtered, dropped
        assert rx.stats["frames_dropped"] == 4 and rx.stats["batches"] == 3
        rx.send_data_request()
        frames = rx.poll()                                                   # batch 3, history carried through the dropped ones
        st = None
        for x in xs[:4]:
            y, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, g.BANK0_COEFF, st)
        rq, iq = g.quantize_bins(np.fft.fft(y.astype(np.float64), axis=-1), -14)
        re, im, _ = g.decode_frame(np.frombuffer(frames[1], np.uint8))
        assert np.abs(re - rq[1]).max() <= 1 and np.abs(im - iq[1]).max() <= 1
    finally:
        rx.stop()
    it2 = iter(batches(1, 5))
    rx = GpuReceiver(lambda: next(it2), channels=1, max_backlog=2)
    try:
        rx.send_ethernet_start()
        for _ in range(5):
            assert rx.poll_into_buffer() == 1
        assert len(rx.frame_buffer) == 2 and rx.stats["frames_dropped"] == 3 and rx.stats["frames_received"] == 5
    finally:
        rx.stop()


@pytest.mark.gpu
def test_gpu_frames_over_udp_loopback(rom):
    """the transport half on the GPU box: real GPU frames through UdpFrameSender and a UDP socket; the 64
    datagrams carry count bytes 0..63 and rebuild the frame in any arrival order"""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device; there is no CPU fallback")
    from fpga_real_time_fft_analyzer_b200 import GpuReceiver
    from fpga_real_time_fft_analyzer_b200.udp_emitter import UdpFrameSender, serve
    xs = batches(3, 3, seed=11)
    it = iter(xs)
    rx = GpuReceiver(lambda: next(it), channels=3)
    sink = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    sink.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 1 << 21)
    sink.bind(("127.0.0.1", 0))
    sink.settimeout(5.0)
    try:
        rx.send_command(0x00)
        rx.send_ethernet_start()
        with UdpFrameSender(dst=sink.getsockname(), src_ip="127.0.0.1", src_port=0) as sender:
            assert serve(rx, sender, channel=2, max_batches=3, period_s=0.0) == 3
            st = None
            from oracle import cgolden as cg
            for x in xs:
                parts = {}
                for _ in range(64):
                    d, _ = sink.recvfrom(2048)
                    assert len(d) == 1025
                    parts[d[0]] = d[1:]
                frame = b"".join(parts[i] for i in range(64))
                y, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, g.BANK0_COEFF, st)
                rq, iq = g.quantize_bins(np.fft.fft(y.astype(np.float64), axis=-1), -14)
                re, im, _ = g.decode_frame(np.frombuffer(frame, np.uint8))
                assert np.abs(re - rq[2]).max() <= 1 and np.abs(im - iq[2]).max() <= 1
    finally:
        sink.close()
        rx.stop()
