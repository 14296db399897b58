"""Parity of the CUDA path (through the C ABI, via FraContext) with the oracle on a
real B200.  Bars: window + IIR12 bit-exact; fp32 bins within 1e-4 relative L2 of a
numpy float64 FFT of the same filtered frames (BASELINE.json north_star); int16 bins
= floor(own fp32 bins * scale) exactly and within 1 LSB of the float64 oracle;
magnitude bit-identical to the GUI decode of the same frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import cgolden as cg          # noqa: E402  (checker only)
from oracle import golden as g            # noqa: E402

B1 = np.array([32, 10, -33, 119, 35, 0, 52, -16, 11, 84, -10, 0], dtype=np.int8)
FFT_TOL = 1e-4


@pytest.fixture(scope="module")
def fra():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    import fpga_real_time_fft_analyzer_b200 as pkg
    from fpga_real_time_fft_analyzer_b200 import _abi, synth
    pkg._abi = _abi
    pkg.synth = synth
    return pkg


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def adversarial(rng, c, n):
    x = rng.integers(-32768, 32768, size=(c, n)).astype(np.int16)
    x[0, :40] = -32768
    x[-1, -40:] = -32768
    return x


def rel_l2(got, ref):
    return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("variant", ["lane", "split", "duo"])
@pytest.mark.parametrize("channels", [1, 5, 33, 257])
def test_k1_bit_exact_state_reload(fra, rom, variant, channels):
    flags = {"lane": fra._abi.FRA_K1_FORCE_LANE, "split": fra._abi.FRA_K1_FORCE_SPLIT,
             "duo": fra._abi.FRA_K1_FORCE_DUO}[variant]
    rng = np.random.default_rng(channels)
    n = 16384
    with fra.FraContext(channels, n, flags=flags) as ctx:
        ctx.command(0x00)
        st = None
        for frame in range(3):                       # state carried over >= 3 frames
            x = adversarial(rng, channels, n)
            y, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1, st)
            out = ctx.process(dev(x), continuous=frame > 0, want=("filtered",))
            assert np.array_equal(out["filtered"].cpu().numpy(), y), (variant, channels, frame)
            assert np.array_equal(ctx.get_state().cpu().numpy(), st)
        # mid-stream 0xF1 reload + bank switch at a frame boundary, history retained
        assert ctx.command(bytes([0xF1]) + B1.tobytes()) and ctx.command(0xA1)
        x = adversarial(rng, channels, n)
        y, st = cg.window_iir(x, rom, 0xA1, g.BANK0_COEFF, B1, st)
        out = ctx.process(dev(x), continuous=True, want=("filtered",))
        assert np.array_equal(out["filtered"].cpu().numpy(), y)
        # burst after a gap: history from zero
        y, st = cg.window_iir(x, rom, 0xA1, g.BANK0_COEFF, B1, None)
        out = ctx.process(dev(x), continuous=False, want=("filtered",))
        assert np.array_equal(out["filtered"].cpu().numpy(), y)
        assert np.array_equal(ctx.get_state().cpu().numpy(), st)


@pytest.mark.parametrize("seed", range(1, 9))
def test_k1_random_int8_coefficients_fuzz(fra, rom, seed):
    rng = np.random.default_rng(seed)
    n, c = 2048, 70
    coef = rng.integers(-128, 128, 12).astype(np.int8)
    if seed == 1:
        coef[:] = [-128, 127, -128, -128, 127, 0, 127, -128, 127, 127, -128, 0]
    if seed == 5:           # |A1| at the limit of k1_duo's two-instruction recurrence, the rest extreme
        coef[:] = [-128, 127, -128, -128, 60, 0, 127, -128, 127, 127, -60, 0]
    if seed == 6:
        coef[:] = [127, -128, 127, 127, -60, 0, -128, 127, -128, -128, 60, 0]
    if seed >= 7:
        coef[4], coef[10] = rng.integers(-60, 61, 2)
    for flags in (fra._abi.FRA_K1_FORCE_LANE, fra._abi.FRA_K1_FORCE_SPLIT,
                  fra._abi.FRA_K1_FORCE_DUO,
                  fra._abi.FRA_K1_FORCE_SPLIT | fra._abi.FRA_K1_SPECULATE):
        with fra.FraContext(c, n, flags=flags) as ctx:
            ctx.load_bank1(coef)
            ctx.set_mode(0xA1)
            st0 = rng.integers(-32768, 32768, (c, 6, 4)).astype(np.int16)
            ctx.set_state(dev(st0))
            x = adversarial(rng, c, n)
            y, st = cg.window_iir(x, rom, 0xA1, g.BANK0_COEFF, coef, st0)
            out = ctx.process(dev(x), continuous=True, want=("filtered",))
            assert np.array_equal(out["filtered"].cpu().numpy(), y)
            assert np.array_equal(ctx.get_state().cpu().numpy(), st)


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_fft_sizes_tolerance_and_framing(fra, rom, n):
    rng = np.random.default_rng(n)
    b = 3 * max(1, 16384 // n) + 1
    x = adversarial(rng, b, n)
    log2n = int(np.log2(n))
    with fra.FraContext(b, n) as ctx:
        got = ctx.fft_only(dev(x)).cpu().numpy()
        ref = np.fft.fft(x.astype(np.float64), axis=-1)
        assert rel_l2(got, ref) < FFT_TOL
        assert max(rel_l2(got[i], ref[i]) for i in range(b)) < FFT_TOL          # per frame, too
        out = ctx.process(dev(x), want=("filtered", "frames", "iq", "mag", "phase"))     # bypass: window fused
        out = {k: v.cpu().numpy() for k, v in out.items()}
        w = g.window(x, rom)
        assert np.array_equal(out["filtered"], w)
        ref = np.fft.fft(w.astype(np.float64), axis=-1)
        got = out["iq"][..., 0] + 1j * out["iq"][..., 1]
        assert rel_l2(got, ref) < FFT_TOL
        assert np.array_equal(cg.quantize_pack(got.astype(np.complex128), -log2n, 0), out["frames"])
        re, im, mag = g.decode_frame(out["frames"])
        rq, iq_ = g.quantize_bins(ref, -log2n)
        assert np.abs(re - rq).max() <= 1 and np.abs(im - iq_).max() <= 1
        assert ((re != rq) | (im != iq_)).mean() < 1e-3
        assert np.array_equal(mag.view(np.uint32), out["mag"].view(np.uint32))
        assert np.abs(out["phase"] - np.arctan2(im, re)).max() < 2e-6


def test_full_chain_tone_noise_and_host_path(fra, rom):
    n, c = 16384, 96
    x = g.tone_noise(range(c), n=n, seed=4)
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        a = {k: v.cpu().numpy() for k, v in ctx.process(dev(x), want=("filtered", "frames", "iq", "mag")).items()}
        pinned = torch.from_numpy(x).pin_memory()
        b = {k: v.numpy().copy() for k, v in ctx.process_host(pinned, want=("filtered", "frames", "iq", "mag")).items()}
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        y, _ = cg.window_iir(x, rom, 0, g.BANK0_COEFF, B1)
        assert np.array_equal(a["filtered"], y)
        ref = np.fft.fft(y.astype(np.float64), axis=-1)
        got = a["iq"][..., 0] + 1j * a["iq"][..., 1]
        assert rel_l2(got, ref) < FFT_TOL
        # the tone of channel c sits at bin round(f_c / fs * N) of the filtered spectrum's peak neighbourhood
        _, _, mag = g.decode_frame(a["frames"])
        assert np.array_equal(mag.view(np.uint32), a["mag"].view(np.uint32))


def test_scale_rounding_saturation(fra, rom):
    n, c = 4096, 5
    x = adversarial(np.random.default_rng(5), c, n)
    for flags, rounding in ((0, 0), (fra._abi.FRA_ROUND_NEAREST, 1)):
        with fra.FraContext(c, n, flags=flags) as ctx:
            for ls in (-12, -8, 0):
                out = ctx.process(dev(x), log2_scale=ls, want=("frames", "iq"))
                iq = out["iq"].cpu().numpy()
                got = (iq[..., 0] + 1j * iq[..., 1]).astype(np.complex128)
                assert np.array_equal(cg.quantize_pack(got, ls, rounding), out["frames"].cpu().numpy()), (flags, ls)


def test_reset_semantics(fra, rom):
    n, c = 2048, 12
    x = adversarial(np.random.default_rng(2), c, n)
    with fra.FraContext(c, n) as ctx:
        ctx.command(bytes([0xF1]) + B1.tobytes() + bytes([0xA1]))
        ctx.process(dev(x), want=("filtered",))
        assert ctx.get_state().any().item()
        ctx.command(0xFF)
        assert ctx.mode == 0xB1 and not ctx.bank(1).any() and not ctx.get_state().any().item()
        out = ctx.process(dev(x), continuous=True, want=("filtered",))
        assert np.array_equal(out["filtered"].cpu().numpy(), g.window(x, rom))
        ctx.command(0xA1)
        assert not ctx.process(dev(x), want=("filtered",))["filtered"].any().item()


def test_config2_full_size_subset_and_properties(fra, rom):
    """BASELINE config 2: 4096 independent channels x 16384, state reset per frame.
    Subset parity on channels {0, 1, C/2, C-1} + 60 seeded random ones; size-independent
    properties on all of them: Parseval per channel, Hermitian symmetry of the frame."""
    c, n = 4096, 16384
    x = fra.synth.tone_noise(c, n, "cuda")
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        out = ctx.process(x, want=("filtered", "frames", "iq"))
        assert ctx.last_kernel_count == 2
        rng = np.random.default_rng(0)
        pick = np.unique(np.concatenate([[0, 1, c // 2, c - 1], rng.integers(0, c, 60)]))
        xs = x[pick].cpu().numpy()
        y, _ = cg.window_iir(xs, rom, 0, g.BANK0_COEFF, B1)
        assert np.array_equal(out["filtered"][pick].cpu().numpy(), y)
        iq = out["iq"][pick].cpu().numpy()
        ref = np.fft.fft(y.astype(np.float64), axis=-1)
        assert rel_l2(iq[..., 0] + 1j * iq[..., 1], ref) < FFT_TOL
        # Parseval on every channel: sum |X|^2 = N sum y^2
        e_t = out["filtered"].to(torch.float64).pow(2).sum(dim=1) * n
        e_f = out["iq"].to(torch.float64).pow(2).sum(dim=(1, 2))
        assert torch.allclose(e_t, e_f, rtol=1e-5)
        # Hermitian symmetry of the fp32 bins: X[N-k] = conj(X[k]) exactly (emitted from the same registers)
        iqa = out["iq"]
        assert torch.equal(iqa[:, 1:, 0], iqa[:, 1:, 0].flip(1)) and torch.equal(iqa[:, 1:, 1], -iqa[:, 1:, 1].flip(1))
        # int16 frame decodes to floor(bins / N)
        fr = out["frames"].view(torch.int16).view(c, n, 2)
        want = torch.floor(out["iq"].to(torch.float64) / n).to(torch.int16)
        assert torch.equal(fr, want)


def test_config2_every_sample_of_the_full_size(fra, rom):
    """BASELINE config 2 in full - all 4096 channels x 16384 samples, not a subset: the window + IIR12 output and the
    final history of EVERY channel bit-exact against the C golden model (67 M samples, a few seconds of one host core),
    with full-range int16 on every 16th channel so that the wrap-around paths are exercised at size; and every
    channel's spectrum within tolerance of a float64 FFT of its filtered frame."""
    c, n = 4096, 16384
    x = fra.synth.tone_noise(c, n, "cuda")
    x[::16] = fra.synth.full_range(c // 16, n, "cuda", seed=9)
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        out = ctx.process(x, want=("filtered", "iq"))
        y, st = cg.window_iir(x.cpu().numpy(), rom, 0, g.BANK0_COEFF, B1)
        got = out["filtered"].cpu().numpy()
        assert np.array_equal(got, y)
        assert np.array_equal(ctx.get_state().cpu().numpy(), st)
        for c0 in range(0, c, 512):
            iq = out["iq"][c0:c0 + 512].cpu().numpy()
            ref = np.fft.fft(y[c0:c0 + 512].astype(np.float64), axis=-1)
            spec = iq[..., 0] + 1j * iq[..., 1]
            assert rel_l2(spec, ref) < FFT_TOL, c0
            worst = np.linalg.norm(spec - ref, axis=1) / np.linalg.norm(ref, axis=1)
            assert worst.max() < FFT_TOL, (c0, worst.argmax())


def test_config3_continuous_sharded_channels(fra, rom):
    """BASELINE config 3 at one GPU's share for 8 GPUs (8192 channels): IIR history
    carried across frames by the lane-per-channel kernel; splitting the stream in
    frames must equal the golden model run over the concatenated stream."""
    c, n, frames = 8192, 16384, 3
    with fra.FraContext(c, n, flags=fra._abi.FRA_K1_FORCE_LANE) as ctx:
        ctx.command(0x00)
        pick = np.array([0, 1, c // 2, c - 1, 777, 4242])
        st = None
        for f in range(frames):
            x = fra.synth.tone_noise(c, n, "cuda", frame=f)
            out = ctx.process(x, continuous=f > 0, want=("filtered",))
            y, st = cg.window_iir(x[pick].cpu().numpy(), rom, 0, g.BANK0_COEFF, B1, st)
            assert np.array_equal(out["filtered"][pick].cpu().numpy(), y)
        assert np.array_equal(ctx.get_state()[pick].cpu().numpy(), st)


@pytest.mark.parametrize("channels,frames", [(65536, 8), (32768, 3), (16384, 3), (8192, 3)])
def test_config3_default_dispatch_continuous(fra, rom, channels, frames):
    """BASELINE config 3 on the kernel the product picks by itself (no FORCE_* flag): 65536
    continuous channels over 8 frames (the one-GPU case), and the shares of 2, 4 and 8 GPUs
    (32768 = the window where the lane-per-channel kernel is chosen).  A seeded subset of the
    channels against the golden model run over the concatenated stream, final history equal,
    and the last frame's bins = floor(FFT(filtered) / N) within 1 LSB."""
    c, n = channels, 16384
    rng = np.random.default_rng(c)
    pick = np.unique(np.concatenate([[0, 1, 31, 32, c // 2 - 1, c // 2, c - 33, c - 32, c - 1], rng.integers(0, c, 40)]))
    pick_t = torch.from_numpy(pick).cuda()
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        out = {"filtered": torch.empty((c, n), dtype=torch.int16, device="cuda"),
               "frames": torch.empty((c, 4 * n), dtype=torch.uint8, device="cuda")}
        st = None
        for f in range(frames):
            x = fra.synth.tone_noise(c, n, "cuda", frame=f)
            if f == 1:
                x[pick_t[:8]] = fra.synth.full_range(8, n, "cuda", seed=f)      # full-range int16 on a few of them
            ctx.process(x, continuous=f > 0, want=("filtered", "frames"), out=out)
            assert ctx.last_kernel_count == 2
            y, st = cg.window_iir(x[pick_t].cpu().numpy(), rom, 0, g.BANK0_COEFF, B1, st)
            assert np.array_equal(out["filtered"][pick_t].cpu().numpy(), y), f
            del x
        assert np.array_equal(ctx.get_state()[pick_t].cpu().numpy(), st)
        re, im, _ = g.decode_frame(out["frames"][pick_t].cpu().numpy())
        rq, iq_ = g.quantize_bins(np.fft.fft(y.astype(np.float64), axis=-1), -14)
        assert np.abs(re - rq).max() <= 1 and np.abs(im - iq_).max() <= 1
        # every channel: the int16 frame is Hermitian (X[N-k] = conj(X[k]) up to the floor of -im)
        fr = out["frames"].view(torch.int16).view(c, n, 2)
        assert torch.equal(fr[:, 1:, 0], fr[:, 1:, 0].flip(1))


def test_config3_every_sample_of_the_full_size(fra, rom):
    """BASELINE config 3 in full - all 65536 continuous channels, two consecutive 16384-sample frames with the history
    carried (2.1 G samples): window + IIR12 output of EVERY channel and the final history bit-exact against the C golden
    model, which runs on all host cores (ctypes releases the GIL; ~25 core-seconds per frame).  Full-range int16 on
    every 64th channel in the second frame."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    c, n, frames = 65536, 16384, 2
    workers = max(1, min(32, os.cpu_count() or 1))
    chunk = 1024
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        out = {"filtered": torch.empty((c, n), dtype=torch.int16, device="cuda")}
        st = [None] * (c // chunk)
        for f in range(frames):
            x = fra.synth.tone_noise(c, n, "cuda", frame=f)
            if f == 1:
                x[::64] = fra.synth.full_range(c // 64, n, "cuda", seed=11)
            ctx.process(x, continuous=f > 0, want=("filtered",), out=out)
            xh = x.cpu().numpy()
            del x
            got = out["filtered"].cpu().numpy()

            def check(i):
                y, s_new = cg.window_iir(xh[i * chunk:(i + 1) * chunk], rom, 0, g.BANK0_COEFF, B1, st[i])
                return i, s_new, bool(np.array_equal(got[i * chunk:(i + 1) * chunk], y))
            with ThreadPoolExecutor(workers) as pool:
                for i, s_new, ok in pool.map(check, range(c // chunk)):
                    assert ok, (f, i)
                    st[i] = s_new
        assert np.array_equal(ctx.get_state().cpu().numpy(), np.concatenate(st, axis=0))


def test_pipeline_mode_config2_size(fra, rom):
    """FRA_PIPELINE at the 4096-channel configuration (one window+IIR CTA per SM beside two FFT
    CTAs - the co-residency the mode exists for): same bytes as the sequential context."""
    c, n, frames = 4096, 16384, 4
    xs = [fra.synth.tone_noise(c, n, "cuda", frame=f) for f in range(frames)]
    xs[2][:64] = fra.synth.full_range(64, n, "cuda", seed=3)
    want = []
    with fra.FraContext(c, n) as ref:
        ref.command(0x00)
        for i, x in enumerate(xs):
            want.append(ref.process(x, continuous=i > 0, want=("frames",))["frames"].clone())
        st_ref = ref.get_state().clone()
    with fra.FraContext(c, n, flags=fra._abi.FRA_PIPELINE) as ctx:
        ctx.command(0x00)
        got = [ctx.process(x, continuous=i > 0, want=("frames",))["frames"] for i, x in enumerate(xs)]
        ctx.join()
        torch.cuda.synchronize()
        for i in range(frames):
            assert torch.equal(got[i], want[i]), i
        assert torch.equal(ctx.get_state(), st_ref)
    pick = np.array([0, 63, 64, 2047, 4095])
    y, st = None, None
    for x in xs:
        y, st = cg.window_iir(x[pick].cpu().numpy(), rom, 0, g.BANK0_COEFF, B1, st)
    assert np.array_equal(st_ref[pick].cpu().numpy(), st)


@pytest.mark.parametrize("c,frames", [(16384, 3), (65536, 2)])
def test_pipeline_mode_config3_shares(fra, rom, c, frames):
    """FRA_PIPELINE - the mode bench.py's headline runs in - at 16384 channels (config 3's share of four GPUs: the
    pipelined context picks the lane-per-channel kernel there, the sequential one the stage-pair kernel) and at all
    65536: same frame bytes and history as the sequential context, and the filter history of a seeded subset equal to
    the golden model's over the concatenated stream."""
    n = 16384
    xs = [fra.synth.tone_noise(c, n, "cuda", frame=f) for f in range(frames)]
    xs[1][:48] = fra.synth.full_range(48, n, "cuda", seed=5)
    want = []
    with fra.FraContext(c, n) as ref:
        ref.command(0x00)
        for i, x in enumerate(xs):
            want.append(ref.process(x, continuous=i > 0, want=("frames",))["frames"].clone())
        st_ref = ref.get_state().clone()
    with fra.FraContext(c, n, flags=fra._abi.FRA_PIPELINE) as ctx:
        ctx.command(0x00)
        got = [ctx.process(x, continuous=i > 0, want=("frames",))["frames"] for i, x in enumerate(xs)]
        ctx.join()
        torch.cuda.synchronize()
        for i in range(frames):
            assert torch.equal(got[i], want[i]), i
        assert torch.equal(ctx.get_state(), st_ref)
    pick = np.array([0, 31, 47, 48, 8191, 8192, c - 1])
    st = None
    for x in xs:
        _, st = cg.window_iir(x[pick].cpu().numpy(), rom, 0, g.BANK0_COEFF, B1, st)
    assert np.array_equal(st_ref[pick].cpu().numpy(), st)


def test_command_bytes_inside_an_upload_are_data(fra, rom):
    """rx_filter_coeff's busy masking (NEW/command_control.vhd:51) on the real library: payload
    bytes that look like commands (0xFF, 0x00, 0xF1, 0xB1) are coefficients, arbitrary
    fragmentation of the byte stream, and the filter then runs with exactly those bytes."""
    n, c = 2048, 40
    rng = np.random.default_rng(77)
    payload = np.array([0xFF, 0x00, 0xF1, 0xB1, 0x15, 0x00, 0xA1, 0x55, 0xA5, 0xEF, 0xFE, 0xFF], dtype=np.uint8)
    stream = bytes([0x00, 0xF1]) + payload.tobytes() + bytes([0xA1, 0x42])
    with fra.FraContext(c, n) as ctx:
        dec = g.CommandDecoder()
        pos = 0
        while pos < len(stream):
            k = int(rng.integers(1, 5))
            done = ctx.command(stream[pos:pos + k])
            dec.feed(stream[pos:pos + k])
            assert done == (not dec.busy)
            pos += k
        assert ctx.mode == dec.mode == 0xA1
        assert np.array_equal(ctx.bank(1), payload.view(np.int8)) and np.array_equal(dec.bank1, payload.view(np.int8))
        assert ctx.counters() == {"start": 0, "request": 0, "reset": 0, "upload": 1}
        x = adversarial(rng, c, n)
        y, _ = cg.window_iir(x, rom, 0xA1, g.BANK0_COEFF, payload.view(np.int8))
        assert np.array_equal(ctx.process(dev(x), want=("filtered",))["filtered"].cpu().numpy(), y)
        # random streams: the library's decoder state equals the oracle decoder's
        interesting = [0x00, 0xA1, 0xB1, 0x55, 0xA5, 0xEF, 0xFE, 0xF1, 0x42]
        for trial in range(10):
            s2 = bytes(int(rng.choice(interesting)) if rng.random() < 0.7 else int(rng.integers(0, 255))
                       for _ in range(int(rng.integers(1, 80))))
            ctx.command(s2)
            dec.feed(s2)
            assert ctx.mode == dec.mode and ctx.transport == dec.transport
            assert np.array_equal(ctx.bank(1), dec.bank1)
            while dec.busy:                               # finish the upload so that the next trial starts idle
                ctx.command(b"\x01")
                dec.feed(b"\x01")


def test_config4_dual_banks_reload_and_all_outputs(fra, rom):
    """BASELINE config 4: bytes A1, then F1 + 12 at a frame boundary mid-run, then 00;
    int16 I/Q frame + fp32 magnitude + fp32 phase."""
    c, n = 64, 16384
    script = [bytes([0xA1]), bytes([0xF1]) + B1.tobytes(), bytes([0x00])]
    with fra.FraContext(c, n) as ctx:
        dec = g.CommandDecoder()
        st = None
        for f, cmd in enumerate(script):
            ctx.command(cmd)
            dec.feed(cmd)
            x = g.tone_noise(range(c), n=n, seed=10 + f)
            out = {k: v.cpu().numpy() for k, v in
                   ctx.process(dev(x), continuous=f > 0, want=("filtered", "frames", "mag", "phase")).items()}
            y, st = cg.window_iir(x, rom, dec.mode, g.BANK0_COEFF, dec.bank1, st)
            assert np.array_equal(out["filtered"], y), f
            re, im, mag = g.decode_frame(out["frames"])
            assert np.array_equal(mag.view(np.uint32), out["mag"].view(np.uint32))
            assert np.abs(out["phase"] - np.arctan2(im, re)).max() < 2e-6
            rq, iq_ = g.quantize_bins(np.fft.fft(y.astype(np.float64), axis=-1), -14)
            assert np.abs(re - rq).max() <= 1 and np.abs(im - iq_).max() <= 1


def test_config5_single_stream(fra, rom):
    """BASELINE config 5: one long stream.  exact=True is bit-exact (six-lane systolic
    chain); the time-parallel chunked scan stays within the dead band of the truncating
    sections and says by how much (it is NOT bit-exact and does not claim to be)."""
    n = 1 << 20
    x = g.tone_noise([5], n=n, seed=2)[0]
    yr, st = cg.window_iir(x[None], rom, 0x00, g.BANK0_COEFF, B1)
    with fra.FraContext(1, 16384) as ctx:
        ctx.command(0x00)
        y, stats = ctx.iir_stream(dev(x), exact=True)
        assert stats["exact"] == 1 and np.array_equal(y.cpu().numpy(), yr[0])
        assert np.array_equal(ctx.get_state().cpu().numpy(), st)
        y, stats = ctx.iir_stream(dev(x), exact=False)
        assert stats["exact"] == 0 and stats["n_chunks"] > 8 and stats["max_state_dev"] <= 32
        assert np.abs(y.cpu().numpy().astype(int) - yr[0].astype(int)).max() <= 64
        # an overflowing (wrapping) cascade is chaotic: the scan detects it and recomputes exactly
        ctx.load_bank1(B1)
        ctx.set_mode(0xA1)
        xr = np.random.default_rng(4).integers(-32768, 32768, 1 << 18).astype(np.int16)
        yw, _ = cg.window_iir(xr[None], rom, 0xA1, g.BANK0_COEFF, B1)
        y, stats = ctx.iir_stream(dev(xr), exact=False)
        assert stats["exact"] == 1 and np.array_equal(y.cpu().numpy(), yw[0])


def test_pipeline_mode_matches_sequential(fra, rom):
    """FRA_PIPELINE: FFT of call i beside window+IIR of call i+1 on internal streams.  Same
    bits as the sequential context over several continuous frames, state included."""
    rng = np.random.default_rng(11)
    c, n, frames = 70, 16384, 5
    xs = [dev(adversarial(rng, c, n)) for _ in range(frames)]
    want = []
    with fra.FraContext(c, n) as ref:
        ref.command(0x00)
        for i, x in enumerate(xs):
            o = ref.process(x, continuous=i > 0, want=("frames", "mag"))
            want.append((o["frames"].clone(), o["mag"].clone()))
        st_ref = ref.get_state().clone()
    with fra.FraContext(c, n, flags=fra._abi.FRA_PIPELINE) as ctx:
        ctx.command(0x00)
        got = [ctx.process(x, continuous=i > 0, want=("frames", "mag")) for i, x in enumerate(xs)]
        ctx.join()
        torch.cuda.synchronize()
        for i in range(frames):
            assert torch.equal(got[i]["frames"], want[i][0]), i
            assert torch.equal(got[i]["mag"], want[i][1]), i
        assert torch.equal(ctx.get_state(), st_ref)
        # caller-owned outputs rewritten every call (filtered + frames): the library orders the
        # reuse itself; the last frame must equal the oracle continued from the state above
        y, st = cg.window_iir(xs[0].cpu().numpy(), rom, 0x00, g.BANK0_COEFF, B1, st_ref.cpu().numpy())
        y2, st2 = cg.window_iir(xs[1].cpu().numpy(), rom, 0x00, g.BANK0_COEFF, B1, st)
        out = {"filtered": torch.empty((c, n), dtype=torch.int16, device="cuda"),
               "frames": torch.empty((c, 4 * n), dtype=torch.uint8, device="cuda")}
        ctx.process(xs[0], continuous=True, want=("filtered", "frames"), out=out)
        ctx.process(xs[1], continuous=True, want=("filtered", "frames"), out=out)
        ctx.join()
        torch.cuda.synchronize()
        assert np.array_equal(out["filtered"].cpu().numpy(), y2)
        assert np.array_equal(ctx.get_state().cpu().numpy(), st2)


def test_process_host_async_matches_sync(fra, rom):
    """Two host calls in flight (upload of frame i+1 beside the download of frame i): same
    bytes as the synchronous call, history carried across the calls."""
    rng = np.random.default_rng(12)
    c, n, frames = 1024, 16384, 4            # 2^24 samples: the sliced, three-stream path
    xs = [torch.from_numpy(adversarial(rng, c, n)).pin_memory() for _ in range(frames)]
    with fra.FraContext(c, n) as ref, fra.FraContext(c, n) as ctx:
        ref.command(0x00)
        ctx.command(0x00)
        pending = None
        for i, x in enumerate(xs):
            w = {k: v.clone() for k, v in ref.process_host(x, continuous=i > 0, want=("frames", "filtered")).items()}
            cur = ctx.process_host_async(x, continuous=i > 0, want=("frames", "filtered")) + (w,)
            if pending is not None:
                ctx.host_wait(pending[1])
                for k in ("frames", "filtered"):
                    assert torch.equal(pending[0][k], pending[2][k]), (i - 1, k)
            pending = cur
        ctx.host_wait(pending[1])
        for k in ("frames", "filtered"):
            assert torch.equal(pending[0][k], pending[2][k]), ("last", k)
        assert torch.equal(ctx.get_state(), ref.get_state())
    y, st = None, None
    for x in xs:
        y, st = cg.window_iir(x.numpy()[:8], rom, 0x00, g.BANK0_COEFF, B1, st)
    assert np.array_equal(pending[0]["filtered"].numpy()[:8], y)


@pytest.mark.parametrize("variant", ["lane", "split", "duo", "auto"])
def test_six_independent_sections(fra, rom, variant):
    """fra_load_sections: every stage its own coefficients, bit-exact."""
    flags = {"lane": fra._abi.FRA_K1_FORCE_LANE, "split": fra._abi.FRA_K1_FORCE_SPLIT,
             "duo": fra._abi.FRA_K1_FORCE_DUO, "auto": 0}[variant]
    rng = np.random.default_rng(31)
    c, n = 45, 16384
    with fra.FraContext(c, n, flags=flags) as ctx:
        for trial in range(3):
            sec = rng.integers(-128, 128, (6, 6)).astype(np.int8)
            if trial >= 1:
                sec[:, 4] = rng.integers(-60, 61, 6)
            if trial == 2:
                sec[:, 1] = 0
            ctx.load_sections(sec)
            ctx.set_mode(0xA1)
            assert np.array_equal(ctx.sections(), sec)
            st0 = rng.integers(-32768, 32768, (c, 6, 4)).astype(np.int16)
            ctx.set_state(dev(st0))
            x = adversarial(rng, c, n)
            y, st = cg.window_iir_sections(x, rom, sec, st0)
            out = ctx.process(dev(x), continuous=True, want=("filtered",))
            assert np.array_equal(out["filtered"].cpu().numpy(), y), (variant, trial)
            assert np.array_equal(ctx.get_state().cpu().numpy(), st)
        ctx.load_bank1(B1)                                   # back to the RTL's two alternating sets
        assert np.array_equal(ctx.sections(), np.stack([B1[:6], B1[6:]] * 3))
        ctx.set_mode(0x00)
        assert np.array_equal(ctx.sections(), np.stack([g.BANK0_COEFF[:6], g.BANK0_COEFF[6:]] * 3))


def test_64k_frames_full_chain(fra, rom):
    """N = 65536 (BASELINE config 5's upper end): window + IIR12 bit-exact, FFT by the
    split / 2 x 32K / join path within tolerance, host path identical."""
    rng = np.random.default_rng(64)
    c, n = 5, 65536
    x = adversarial(rng, c, n)
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        out = {k: v.cpu().numpy() for k, v in ctx.process(dev(x), want=("filtered", "frames", "iq", "mag")).items()}
        y, st = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1)
        assert np.array_equal(out["filtered"], y)
        assert np.array_equal(ctx.get_state().cpu().numpy(), st)
        ref = np.fft.fft(y.astype(np.float64), axis=-1)
        got = out["iq"][..., 0] + 1j * out["iq"][..., 1]
        assert rel_l2(got, ref) < FFT_TOL
        assert np.array_equal(cg.quantize_pack(got.astype(np.complex128), -16, 0), out["frames"])
        re, im, mag = g.decode_frame(out["frames"])
        assert np.array_equal(mag.view(np.uint32), out["mag"].view(np.uint32))
        host = ctx.process_host(torch.from_numpy(x).pin_memory(), want=("frames",))
        assert np.array_equal(host["frames"].numpy(), out["frames"])


@pytest.mark.parametrize("n,other", [(32768, "wide"), (65536, "split")])
def test_cluster_and_wide_cta_frames_agree(fra, rom, n, other):
    """32K / 64K frames: the default (64 KiB CTAs in a cluster of two / four, sub-sequences exchanged through
    distributed shared memory) against the other path of that size - FRA_K2_WIDE_CTA (32K in one 128 KiB CTA),
    FRA_K2_64K_SPLIT (64K as two 32K transforms through HBM).
    Both within tolerance of float64 and of each other, frames equal up to the last bit of a few bins."""
    rng = np.random.default_rng(n + 7)
    c = 7
    x = adversarial(rng, c, n)
    outs = []
    for flags in (0, fra._abi.FRA_K2_WIDE_CTA if other == "wide" else fra._abi.FRA_K2_64K_SPLIT):
        with fra.FraContext(c, n, flags=flags) as ctx:
            spec = ctx.fft_only(dev(x)).cpu().numpy()
            o = {k: v.cpu().numpy() for k, v in ctx.process(dev(x), want=("frames", "iq", "mag", "phase")).items()}
            only = ctx.process(dev(x), want=("frames",))["frames"].cpu().numpy()          # the frames-only instantiation
            assert np.array_equal(only, o["frames"])
            outs.append((spec, o))
    ref = np.fft.fft(x.astype(np.float64), axis=-1)
    for spec, o in outs:
        assert rel_l2(spec, ref) < FFT_TOL
    # the same butterflies in the same order; only the compiler's FMA contraction may differ between the instantiations
    assert rel_l2(outs[0][0], outs[1][0]) < 1e-6
    ra, ia, _ = g.decode_frame(outs[0][1]["frames"])
    rb, ib, _ = g.decode_frame(outs[1][1]["frames"])
    assert np.abs(ra - rb).max() <= 1 and np.abs(ia - ib).max() <= 1
    assert ((ra != rb) | (ia != ib)).mean() < 1e-3
    for spec, o in outs:                                                               # each variant is self-consistent
        got = o["iq"][..., 0] + 1j * o["iq"][..., 1]
        assert np.array_equal(cg.quantize_pack(got.astype(np.complex128), -int(np.log2(n)), 0), o["frames"])
        _, _, mag = g.decode_frame(o["frames"])
        assert np.array_equal(mag.view(np.uint32), o["mag"].view(np.uint32))


def test_magnitude_averaging_in_the_pack_stage(fra, rom):
    rng = np.random.default_rng(19)
    c, n, alpha = 9, 16384, 0.125
    with fra.FraContext(c, n) as ctx:
        ctx.command(0x00)
        ctx.set_mag_average(alpha)
        out = {"mag": torch.zeros((c, n), dtype=torch.float32, device="cuda"),
               "frames": torch.empty((c, 4 * n), dtype=torch.uint8, device="cuda")}
        avg = np.zeros((c, n), np.float32)
        for frame in range(3):
            ctx.process(dev(adversarial(rng, c, n)), continuous=frame > 0, out=out)
            _, _, mag = g.decode_frame(out["frames"].cpu().numpy())
            avg = (avg + np.float32(alpha) * (mag.astype(np.float32) - avg)).astype(np.float32)
            assert np.allclose(out["mag"].cpu().numpy(), avg, rtol=1e-6, atol=1e-4)
        with pytest.raises(fra.FraError):
            ctx.set_mag_average(0.0)


@pytest.mark.parametrize("variant", ["lane", "duo", "auto"])
def test_all_biased_step_at_the_edge_of_its_condition(fra, rom, variant):
    """biquad_step_biased (five FFMAs + one PRMT per stage, every operand carrying the PRMT bias):
    coefficient sets on the edge of its eligibility condition, random history, full-range input,
    16K frames over two continuous calls; the general step (FRA_K1_NO_BIASED) gives the same bits."""
    from tests.test_emul_kernels import biased_edge_sections
    flags = {"lane": fra._abi.FRA_K1_FORCE_LANE, "duo": fra._abi.FRA_K1_FORCE_DUO, "auto": 0}[variant]
    rng = np.random.default_rng(78)
    c, n = 70, 16384
    for trial in range(3):
        sec = biased_edge_sections(rng)
        if trial == 2:
            sec[:, 1] = 0
            sec[:, 2] = np.clip(sec[:, 2], -60, 60)
            sec[:, 4] = np.clip(sec[:, 4], -30, 30)
            sec[:, 0] = -sec[:, 2]
        st0 = rng.integers(-32768, 32768, (c, 6, 4)).astype(np.int16)
        xa, xb = adversarial(rng, c, n), adversarial(rng, c, n)
        ya, st = cg.window_iir_sections(xa, rom, sec, st0)
        yb, st = cg.window_iir_sections(xb, rom, sec, st)
        for fl in (flags, flags | fra._abi.FRA_K1_NO_BIASED):
            with fra.FraContext(c, n, flags=fl) as ctx:
                ctx.load_sections(sec)
                ctx.set_mode(0xA1)
                ctx.set_state(dev(st0))
                assert np.array_equal(ctx.process(dev(xa), continuous=True, want=("filtered",))["filtered"].cpu().numpy(), ya)
                assert np.array_equal(ctx.process(dev(xb), continuous=True, want=("filtered",))["filtered"].cpu().numpy(), yb)
                assert np.array_equal(ctx.get_state().cpu().numpy(), st)


@pytest.mark.parametrize("n", [1024, 4096, 16384, 32768])
def test_fixed_point_fft_mode(fra, rom, n):
    """FRA_FFT_FIXED16: 16-bit data, 16-bit phase factors, scaled 1/N, truncation - the
    configuration of xfft_0.xci - as a radix-2^2 integer pipeline.  Bit-exact against the integer oracle
    (oracle/fixed_fft.py); within the quantisation noise expected of such a pipeline against the float64
    FFT (a few LSB peak, ~1 LSB rms).  Bit-level parity with the proprietary Xilinx core: UNPINNED."""
    from oracle.fixed_fft import fixed_fft
    rng = np.random.default_rng(n)
    c = 37
    x = g.tone_noise(range(c), n=n, seed=5)
    x[-1] = rng.integers(-32768, 32768, n)
    with fra.FraContext(c, n, flags=fra._abi.FRA_FFT_FIXED16) as ctx:
        ctx.command(0x00)
        out = {k: v.cpu().numpy() for k, v in ctx.process(dev(x), want=("filtered", "frames", "iq", "mag", "phase")).items()}
        y, _ = cg.window_iir(x, rom, 0x00, g.BANK0_COEFF, B1)
        assert np.array_equal(out["filtered"], y)
        re, im = fixed_fft(y)
        gre, gim, gmag = g.decode_frame(out["frames"])
        assert np.array_equal(gre, re) and np.array_equal(gim, im)
        assert np.array_equal(out["mag"].view(np.uint32), gmag.astype(np.float32).view(np.uint32))
        assert np.abs(out["phase"] - np.arctan2(gim, gre)).max() < 2e-6
        assert np.array_equal(out["iq"][..., 0], re.astype(np.float32) * n)
        ref = np.fft.fft(y.astype(np.float64), axis=-1) / n
        err = (re + 1j * im) - ref
        assert np.abs(err).max() < 8.0 and np.sqrt((np.abs(err) ** 2).mean()) < 1.5
        # the float path on the same input: its int16 bins agree with the fixed-point ones to the same few LSB
        with fra.FraContext(c, n) as fl:
            fl.command(0x00)
            fre, fim, _ = g.decode_frame(fl.process(dev(x), want=("frames",))["frames"].cpu().numpy())
        assert np.abs(fre.astype(int) - re).max() <= 8 and np.abs(fim.astype(int) - im).max() <= 8
        ctx.command(0xB1)                                     # bypass: window fused into the load
        o2 = ctx.process(dev(x), want=("frames",))["frames"].cpu().numpy()
        re, im = fixed_fft(g.window(x, rom))
        gre, gim, _ = g.decode_frame(o2)
        assert np.array_equal(gre, re) and np.array_equal(gim, im)
        with pytest.raises(fra.FraError):
            ctx.process(dev(x), log2_scale=-3, want=("frames",))
    with pytest.raises(fra.FraError):
        fra.FraContext(1, 65536, flags=fra._abi.FRA_FFT_FIXED16)


@pytest.mark.parametrize("n,c", [(1024, 70), (16384, 1100), (65536, 9)])
def test_half_spectrum_host_transfer(fra, rom, n, c):
    """FRA_HOST_HALF_SPECTRUM: the frames cross PCIe as bins 0..N/2 + one bit per bin and the host's cores
    complete the Hermitian half (AVX2 path, several threads, the sliced three-stream path at 1100 x 16K):
    byte-identical to the full transfer, synchronous and with two calls in flight."""
    rng = np.random.default_rng(n)
    xs = []
    for i in range(3):
        x = adversarial(rng, c, n)
        x[1] = 0
        x[2] = -32768
        x[3 % c] = g.tone_noise([3], n=n, seed=i)[0]
        xs.append(torch.from_numpy(x).pin_memory())
    with fra.FraContext(c, n) as full, fra.FraContext(c, n, flags=fra._abi.FRA_HOST_HALF_SPECTRUM) as half:
        full.command(0x00); half.command(0x00)
        want = [full.process_host(x, continuous=i > 0, want=("frames",))["frames"].clone() for i, x in enumerate(xs)]
        pending = None
        for i, x in enumerate(xs):
            cur = half.process_host_async(x, continuous=i > 0, want=("frames",))
            if pending is not None:
                half.host_wait(pending[1])
                assert torch.equal(pending[0]["frames"], want[i - 1]), i - 1
            pending = cur
        half.host_wait(pending[1])
        assert torch.equal(pending[0]["frames"], want[-1])
        assert torch.equal(half.get_state(), full.get_state())
        # a saturating scale takes the full transfer
        a = full.process_host(xs[0], log2_scale=-6, want=("frames",))["frames"].clone()
        b = half.process_host(xs[0], log2_scale=-6, want=("frames",))["frames"]
        assert torch.equal(a, b)
        # mixed transfer: a fixed share of every channel slice as half spectra, the rest whole - same bytes out,
        # and the bytes that crossed the link are what the share says
        ref = full.process_host(xs[1], continuous=True, want=("frames",))["frames"].clone()
        st = full.get_state().clone()
        state0 = half.get_state().clone()
        full_bytes = c * n * 4
        for share in (0.0, 0.3, 0.77, 1.0):
            half.set_host_half_share(share)
            half.set_state(state0)
            got = half.process_host(xs[1], continuous=True, want=("frames",))["frames"]
            assert torch.equal(got, ref), share
            assert torch.equal(half.get_state(), st)
            h2d, d2h, sh = half.host_transfer()
            assert h2d == c * n * 2 and sh == share
            assert d2h <= full_bytes and (share > 0.0 or d2h == full_bytes) and (share < 1.0 or d2h < 0.53 * full_bytes)
        # adaptive again (the default): whatever share the controller picks, the frames stay the same
        half.set_host_half_share(-1.0)
        ref0 = full.process_host(xs[2], continuous=False, want=("frames",))["frames"].clone()
        pending = None
        for i in range(19):                             # the share search takes at most 17 calls
            cur = half.process_host_async(xs[2], continuous=False, want=("frames",))
            if pending is not None:
                half.host_wait(pending[1])
                assert torch.equal(pending[0]["frames"], ref0), i
            pending = cur
        half.host_wait(pending[1])
        assert torch.equal(pending[0]["frames"], ref0)
        assert half.host_transfer()[2] in (1.0, 0.875, 0.75, 0.625, 0.5, 0.375)
