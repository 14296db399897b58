"""udp_emitter - sends GPU-produced frames in the reference's Ethernet wire format, so that the
UNMODIFIED UdpReceiver / MultiPacketAssembler of the reference GUI (scripts/fft_analyzer_gui.py:
308-460) can display them.

Wire format (IMP/phy_rmii_if.vhd:173-175, 322-323; IMP/head_data.mif): one frame = 64 UDP
datagrams of 1025 bytes = one count byte (0..63, the 6-bit mark_cnt) + 1024 frame bytes; source
port 5005, destination port 6006 (the FPGA broadcasts from 169.254.252.255; the GUI filters on
web_config['expect_src_ip'] / ['expect_src_port'], GUI:382-385).  The MAC / IP / UDP headers, CRC
and checksum of the RTL are the operating system's job here."""
from __future__ import annotations

import socket
import time

from .receiver import FRAME_SIZE_BYTES, PACKETS_PER_FRAME, frame_to_udp_payloads

FPGA_SRC_PORT = 5005          # head_data.mif bytes 34-35 / GUI:21
GUI_DST_PORT = 6006           # head_data.mif bytes 36-37 / GUI:19


class UdpFrameSender:
    """Sends 65536-byte frames as 64 x 1025-byte datagrams from `src_port` to `dst`."""

    def __init__(self, dst=("127.0.0.1", GUI_DST_PORT), src_ip="", src_port=FPGA_SRC_PORT, broadcast=False,
                 pace_s=0.0):
        self.dst = dst
        self.pace_s = pace_s                      # pause between datagrams (the RTL needs ~86 us per packet at 100 Mb/s)
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        if broadcast:
            self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_BROADCAST, 1)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 4 * FRAME_SIZE_BYTES)
        self.sock.bind((src_ip, src_port))
        self.frames_sent = 0
        self.packets_sent = 0

    @property
    def src_port(self):
        return self.sock.getsockname()[1]

    def send_frame(self, frame: bytes):
        for payload in frame_to_udp_payloads(frame):
            self.sock.sendto(payload, self.dst)
            self.packets_sent += 1
            if self.pace_s:
                time.sleep(self.pace_s)
        self.frames_sent += 1

    def send_frames(self, frames):
        for f in frames:
            self.send_frame(f)

    def close(self):
        if self.sock is not None:
            self.sock.close()
            self.sock = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def serve(receiver, sender: UdpFrameSender, channel: int = 0, max_batches=None, period_s: float = 1.0 / 30.0):
    """The FPGA's Ethernet mode as a loop: every period one batch goes through the GPU chain and
    channel `channel`'s frame goes out (30 frames per second is what the reference's link and GUI
    sustain, README.md:168, GUI:53).  Returns the number of frames sent."""
    n = 0
    while receiver.active and (max_batches is None or n < max_batches):
        t0 = time.time()
        frames = receiver.poll()
        if frames:
            sender.send_frame(frames[min(channel, len(frames) - 1)])
            n += 1
        dt = period_s - (time.time() - t0)
        if dt > 0:
            time.sleep(dt)
    return n


__all__ = ["UdpFrameSender", "serve", "FPGA_SRC_PORT", "GUI_DST_PORT", "PACKETS_PER_FRAME"]
