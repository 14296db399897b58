"""GpuReceiver - a 'GPU' receiver backend beside the reference GUI's UartReceiver and
UdpReceiver (scripts/fft_analyzer_gui.py:355-747 of the reference; "GUI" below).

Same command set, same 65536-byte frame format: the frames this class hands out go
straight into the GUI's own decode_mag_16iq_le / decode_iq_components (GUI:250-270).
The method names, argument meaning and error behaviour (swallow, print, return
False - GUI:567-613) follow UartReceiver so that ReceiverController can treat both
alike (INTEGRATION.md shows the four-line patch).  It is a plain object: under
PyQt the GUI wraps `poll` in a QTimer exactly as UartReceiver does with read_data
(GUI:496-498)."""
from __future__ import annotations

import time

import numpy as np

from . import _abi
from .context import FraContext

FRAME_SIZE_BYTES = 65536      # GUI:41-44
SAMPLES_PER_FRAME = 16384
PACKETS_PER_FRAME = 64        # GUI:48-50
PACKET_DATA_SIZE = FRAME_SIZE_BYTES // PACKETS_PER_FRAME


class GpuReceiver:
    """source(): returns the next batch of samples, int16 [channels, fft_size], as a
    numpy array / CPU torch tensor (e2e path) or a CUDA torch tensor (resident path)."""

    def __init__(self, source, channels: int = 1, fft_size: int = SAMPLES_PER_FRAME, device: int = 0,
                 continuous: bool = True, want=("frames",), flags: int = 0, free_running: bool = False,
                 max_backlog: int = 10):
        self.ctx = FraContext(channels, fft_size, device, flags)      # raises like UartReceiver's open (GUI:482-484)
        self.source = source
        self.channels, self.fft_size = channels, fft_size
        self.continuous = continuous
        self.want = tuple(want)
        self.active = True
        self.frame_buffer = []            # frames not yet handed to the GUI (bytes objects)
        self._start_base = 0              # start commands seen before the last reset
        self._seen_request = 0
        self._first = True
        self.last_reset_time = 0.0
        # free_running: the ADC and the filter never stop (IMP/dsp_system_top.vhd:427-449) - a poll() that finds
        # the frame gate closed still takes the next batch from the source, filters it (history carried)
        # and drops its frames, as `sequencer` discards the FFT output between hand-shakes
        # (IMP/sequencer_dsp.vhd:50-82).  Off: the source is only asked for samples that will be shown.
        self.free_running = free_running
        self.max_backlog = max_backlog    # frames kept for a slow consumer (GUI:687-689 trims its backlog the same way)
        self.stats = {"frames_received": 0, "frames_dropped": 0, "batches": 0, "samples": 0}
        print(f"GPU receiver opened: cuda:{device}, {channels} channel(s) x {fft_size} samples")

    # -------------------------------------------------- command surface (UartReceiver's)
    def send_command(self, command) -> bool:                           # GUI:567-585
        try:
            if not self.active:
                return False
            if command == _abi.FPGA_RESET_CMD:
                now = time.time()
                if (now - self.last_reset_time) < 2.0:
                    print("FPGA reset ignored - cooldown active")
                    return False
                self.last_reset_time = now
            self.ctx.command(int(command) & 0xFF)
            if command == _abi.FPGA_RESET_CMD:
                self._after_reset()
            print(f"GPU command sent: 0x{int(command) & 0xFF:02X}")
            return True
        except Exception as e:
            print(f"GPU command error: {e}")
            return False

    @staticmethod
    def _byte(val):                                                    # GUI:587-589
        return int(val) & 0xFF

    def send_filter_coefficients(self, coefficients) -> bool:          # GUI:591-613
        """Two sections x [b0,b1,b2,a0,a1,a2], section-major: 0xF1 then 12 bytes."""
        try:
            if not self.active:
                return False
            payload = bytes(self._byte(c) for sec in coefficients for c in sec)
            self.ctx.command(bytes([_abi.FILTER_UPDATE_CMD]))
            done = self.ctx.command(payload)
            print(f"Sent {len(payload)} coefficient bytes" + ("" if done else " (upload incomplete)"))
            return True
        except Exception as e:
            print(f"Filter coefficient send error: {e}")
            return False

    def send_filter_sections(self, sections) -> bool:
        """Superset of send_filter_coefficients: 1..6 independent
        sections of int8 bytes in the RTL's register order B0,B1,B2,A0,A1,A2, e.g. from
        filter_design.quantize_sections(sos, rtl_compatible=True).  Goes through the 0xF1 byte
        protocol when the design fits the RTL's two alternating sets."""
        try:
            if not self.active:
                return False
            from . import filter_design
            how = filter_design.upload(self.ctx, sections, select=False)
            print(f"Sent filter sections ({how})")
            return True
        except Exception as e:
            print(f"Filter section send error: {e}")
            return False

    def send_start_sequence(self) -> bool:                             # GUI:529-541: 0x55 then 0xA5
        return self._send_raw(_abi.START_COMMAND) and self.send_data_request()

    def send_data_request(self) -> bool:                               # GUI:543-553
        return self._send_raw(_abi.UART_REQUEST_CMD)

    def send_ethernet_start(self) -> bool:                             # GUI:555-565
        return self._send_raw(_abi.START_COMMAND)

    def force_mode_reset(self) -> bool:                                # GUI:500-527 (three resets)
        try:
            for _ in range(3):
                self.ctx.command(_abi.FPGA_RESET_CMD)
            self._after_reset()
            self.frame_buffer.clear()
            print("Complete GPU receiver reset performed")
            return True
        except Exception as e:
            print(f"Mode reset error: {e}")
            return False

    def stop(self):                                                    # GUI:742-747
        self.active = False
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    def _send_raw(self, byte) -> bool:
        try:
            if not self.active:
                return False
            self.ctx.command(byte)
            return True
        except Exception as e:
            print(f"GPU command error: {e}")
            return False

    def _after_reset(self):
        self._first = True
        c = self.ctx.counters()
        self._start_base, self._seen_request = c["start"], c["request"]     # sequencer back in IDLE

    # ------------------------------------------------------------------ data
    def _armed(self) -> bool:
        """Frames flow after a start command (0x55).  In UART transport one batch is
        released per 0xA5 request (IMP/sequ2.vhd:214-218); in Ethernet transport they
        stream (IMP/sequ2.vhd:121-172)."""
        c = self.ctx.counters()
        if c["start"] <= self._start_base:
            return False
        if self.ctx.transport == _abi.UART_MODE_CMD:
            if c["request"] > self._seen_request:
                self._seen_request += 1
                return True
            return False
        return True

    def process_batch(self, x=None):
        """Run one batch through the GPU chain; returns the dict of outputs
        (CPU pinned tensors for host input, CUDA tensors for CUDA input)."""
        if x is None:
            x = self.source()
        carry = self.continuous and not self._first
        self._first = False
        if hasattr(x, "is_cuda") and x.is_cuda:
            out = self.ctx.process(x, continuous=carry, want=self.want)
            if self.ctx.flags & _abi.FRA_PIPELINE:
                # a pipelined context holds this call's FFT back until the next call: a receiver hands
                # its frames out now, so the current stream joins the internal streams first
                self.ctx.join()
        else:
            out = self.ctx.process_host(x, continuous=carry, want=self.want)
        self.stats["batches"] += 1
        self.stats["samples"] += self.channels * self.fft_size
        return out

    def poll(self):
        """What UartReceiver.read_data + process_buffer do on the timer (GUI:616-689):
        returns the list of complete frames (bytes, 4 * fft_size each) now available."""
        if not self.active:
            return []
        if not self._armed():
            if self.free_running and self.ctx.counters()["start"] > self._start_base:
                self.process_batch()                               # the chain runs on; nobody takes the frames
                self.stats["frames_dropped"] += self.channels
            return []
        out = self.process_batch()
        frames = out["frames"]
        if hasattr(frames, "is_cuda") and frames.is_cuda:
            frames = frames.cpu()
        arr = frames.numpy() if hasattr(frames, "numpy") else np.asarray(frames)
        got = [arr[c].tobytes() for c in range(self.channels)]
        self.stats["frames_received"] += len(got)
        return got

    def poll_into_buffer(self):
        """poll() for a consumer that drains `frame_buffer` at its own pace (the GUI's display loop): new
        frames are appended; beyond `max_backlog` batches the oldest are dropped and counted, as
        UartReceiver.process_buffer trims an overflowing buffer (GUI:687-689)."""
        got = self.poll()
        if got:
            self.frame_buffer.append(got)
            while len(self.frame_buffer) > self.max_backlog:
                self.stats["frames_dropped"] += len(self.frame_buffer.pop(0))
        return len(got)


def frame_to_udp_payloads(frame: bytes):
    """The reference's Ethernet wire format for one frame: 64 payloads of 1 count byte
    + 1024 data bytes (IMP/phy_rmii_if.vhd:173-175,322-323; consumer
    MultiPacketAssembler, GUI:308-352).  Lets the UNMODIFIED UdpReceiver consume GPU
    frames when they are sent from 169.254.252.255:5005 to port 6006."""
    if len(frame) != FRAME_SIZE_BYTES:
        raise ValueError(f"Invalid frame size: {len(frame)} (expected {FRAME_SIZE_BYTES})")
    return [bytes([i]) + frame[i * PACKET_DATA_SIZE:(i + 1) * PACKET_DATA_SIZE] for i in range(PACKETS_PER_FRAME)]
