"""GpuReceiver: the reference GUI's receiver contract (command set, 65536-byte frames,
UDP wire format) on top of the CUDA path."""
import numpy as np
import pytest

from oracle import golden as g


def test_udp_payload_format_reassembles_like_the_reference(gui_vectors):
    """64 payloads of count byte + 1024 data bytes; the fixture recorded that the
    reference's MultiPacketAssembler rebuilds the frame from exactly this format."""
    from fpga_real_time_fft_analyzer_b200.receiver import frame_to_udp_payloads
    rng = np.random.default_rng(gui_vectors["decode"]["seed"])
    frame = rng.integers(0, 256, size=65536, dtype=np.uint8).tobytes()
    payloads = frame_to_udp_payloads(frame)
    assert len(payloads) == 64 and all(len(p) == 1025 for p in payloads)
    assert gui_vectors["assembler"]["reassembled_equals_frame"] is True
    rebuilt = {}
    for i in gui_vectors["assembler"]["order"]:          # the order the fixture fed the reference assembler
        p = payloads[i]
        rebuilt[p[0]] = p[1:]
    assert b"".join(rebuilt[i] for i in range(64)) == frame
    with pytest.raises(ValueError):
        frame_to_udp_payloads(frame[:-1])


@pytest.mark.gpu
def test_gpu_receiver_protocol_and_frames(rom, gui_vectors):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device; there is no CPU fallback")
    from fpga_real_time_fft_analyzer_b200 import GpuReceiver
    c, n = 3, 16384
    batches = [g.tone_noise(range(c), n=n, seed=s) for s in range(4)]
    it = iter(batches)
    rx = GpuReceiver(lambda: next(it), channels=c, fft_size=n)
    try:
        assert rx.poll() == []                                   # nothing before the start command
        assert rx.send_command(0xFE)                             # UART transport: one batch per 0xA5
        assert rx.send_start_sequence()                          # 0x55 then 0xA5
        frames = rx.poll()
        assert len(frames) == c and all(len(f) == 65536 for f in frames)
        assert rx.poll() == []                                   # no second request yet
        w = g.window(batches[0], rom)                            # default mode 0xB1: window only
        rq, iq_ = g.quantize_bins(np.fft.fft(w.astype(np.float64), axis=-1), -14)
        re, im, _ = g.decode_frame(np.frombuffer(frames[1], np.uint8))
        assert np.abs(re - rq[1]).max() <= 1 and np.abs(im - iq_[1]).max() <= 1
        # the GUI's quantised sections go out as 0xF1 + 12 bytes, section-major (GUI:603)
        sections = gui_vectors["quantize"][0]["sections"]
        assert rx.send_filter_coefficients(sections) and rx.send_command(0xA1)
        assert list(rx.ctx.bank(1)) == [v for sec in sections for v in sec]
        assert rx.send_command(0xEF)                             # Ethernet transport: frames stream
        assert len(rx.poll()) == c and len(rx.poll()) == c
        assert rx.send_command(0xFF) and not rx.send_command(0xFF)   # 2 s reset cooldown (GUI:571-576)
        assert rx.ctx.mode == 0xB1
    finally:
        rx.stop()
    assert not rx.active and rx.send_command(0x00) is False


@pytest.mark.gpu
def test_gpu_receiver_on_a_pipelined_context(rom):
    """A FRA_PIPELINE context holds the FFT of a call back until the next one; the receiver joins
    before it hands frames out, so poll() returns finished frames (not uninitialised memory) for
    CUDA sources too."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device; there is no CPU fallback")
    from fpga_real_time_fft_analyzer_b200 import GpuReceiver, _abi
    c, n = 6, 16384
    batches = [g.tone_noise(range(c), n=n, seed=s) for s in range(3)]
    it = iter(torch.from_numpy(b).cuda() for b in batches)
    ref = GpuReceiver(lambda: None, channels=c, fft_size=n)
    rx = GpuReceiver(lambda: next(it), channels=c, fft_size=n, flags=_abi.FRA_PIPELINE)
    try:
        for r in (ref, rx):
            assert r.send_command(0x00) and r.send_ethernet_start()
        for b in batches:
            got = rx.poll()
            want = ref.process_batch(torch.from_numpy(b).cuda())["frames"].cpu().numpy()
            assert len(got) == c
            for ch in range(c):
                assert got[ch] == want[ch].tobytes()
    finally:
        rx.stop()
        ref.stop()
