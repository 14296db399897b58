"""gui_patch - gives the reference GUI (scripts/fft_analyzer_gui.py of
mfkiwl/fpga-real-time-fft-analyzer) its third receiver backend, 'GPU', next to UART and ETHERNET.

The reference GUI is not vendored here; this module edits a copy of it:

    python -m fpga_real_time_fft_analyzer_b200.gui_patch  path/to/scripts/fft_analyzer_gui.py  out.py
    python -m fpga_real_time_fft_analyzer_b200.gui_patch  --html path/to/templates/index.html  out.html

`patch_gui_source` finds its anchors with `ast` (class and function names, not line numbers) and
makes the five edits of INTEGRATION.md section 3:

  1. class GpuQtReceiver (below) is inserted in front of `class ReceiverController` (GUI:749);
  2. `ReceiverController.start_receiver` (GUI:817-849) gains an `elif mode == "GPU":` branch;
  3. every `isinstance(<x>.current_receiver, UartReceiver)` (GUI:763, 854, 908, 1017) also accepts
     GpuQtReceiver, so commands and 0xF1 uploads go to the GPU backend instead of a temporary
     serial port;
  4. the transport byte sent on a mode switch (GUI:1035, 1246) is UART only for "UART": the GPU
     backend streams frames like the Ethernet one;
  5. web_config gains 'gpu_channels' / 'gpu_display_channel' / 'gpu_device'.

Everything else - decode_mag_16iq_le, the 30 FPS limiter, the frame_data payload, the filter
designer - stays the reference's own code and is what the GPU frames go through.
"""
from __future__ import annotations

import ast
import re
import sys

# The Qt shell around GpuReceiver.  It lives in the GUI module's namespace and uses that
# module's own helpers (receiver_state, update_fps, should_display_frame, decode_mag_16iq_le,
# UdpReceiver.emit_plot_data - which is what builds the frame_data payload of GUI:439-455).
GPU_QT_RECEIVER = '''
class GpuQtReceiver(QtCore.QObject):
    """'GPU' receiver backend (fpga_real_time_fft_analyzer_b200): same command surface as
    UartReceiver (GUI:500-613), frames through the same decode + emit path as UdpReceiver."""

    def __init__(self, source=None, channels=1, device=0, display_channel=0, fft_size=SAMPLES_PER_FRAME):
        super().__init__()
        from fpga_real_time_fft_analyzer_b200 import GpuReceiver
        if source is None:
            source = self._demo_source(channels, fft_size, device)
        self.rx = GpuReceiver(source, channels=channels, fft_size=fft_size, device=device)   # raises without a CUDA device
        self.display_channel = display_channel
        self.active = True
        self.read_timer = QtCore.QTimer(self)
        self.read_timer.timeout.connect(self.read_data)
        self.read_timer.start(1)
        print(f"GPU receiver started: {channels} channel(s), displaying channel {display_channel}")

    @staticmethod
    def _demo_source(channels, fft_size, device):
        """synthetic 12-bit tone + noise, consecutive frames of one continuous stream, made on the GPU"""
        from fpga_real_time_fft_analyzer_b200 import synth
        state = {"frame": 0}

        def next_batch():
            x = synth.tone_noise(channels, fft_size, f"cuda:{device}", frame=state["frame"])
            state["frame"] += 1
            return x
        return next_batch

    def read_data(self):
        if not self.active:
            return
        try:
            frames = self.rx.poll()
        except Exception as e:
            print(f"GPU read error: {e}")
            return
        if not frames:
            return
        frame = frames[min(self.display_channel, len(frames) - 1)]
        receiver_state['frames_received'] += 1
        receiver_state['fps_counters']['incoming'] += 1
        update_fps()
        if should_display_frame("GPU"):
            receiver_state['frames_displayed'] += 1
            receiver_state['fps_counters']['display'] += 1
            try:
                UdpReceiver.emit_plot_data(self, decode_mag_16iq_le(frame), frame)
            except Exception as e:
                print(f"FFT decode error: {e}")

    # command surface: the names, arguments and return values of UartReceiver
    def send_command(self, command):
        return self.rx.send_command(command)

    def send_filter_coefficients(self, coefficients):
        return self.rx.send_filter_coefficients(coefficients)

    def send_start_sequence(self):
        return self.rx.send_start_sequence()

    def send_data_request(self):
        return self.rx.send_data_request()

    def send_ethernet_start(self):
        return self.rx.send_ethernet_start()

    def force_mode_reset(self):
        return self.rx.force_mode_reset()

    def stop(self):
        self.active = False
        if hasattr(self, 'read_timer'):
            self.read_timer.stop()
        self.rx.stop()

'''

START_BRANCH = '''            elif mode == "GPU":
                self.current_receiver = GpuQtReceiver(
                    web_config.get('gpu_source'), channels=web_config.get('gpu_channels', 1),
                    device=web_config.get('gpu_device', 0), display_channel=web_config.get('gpu_display_channel', 0))
                self.current_receiver.send_ethernet_start()
                print(f"Receiver started in {mode} mode")
                socketio.emit('receiver_status', {
                    'active': True,
                    'mode': mode,
                    'message': f'{mode} receiver started successfully'
                })
'''


class PatchError(RuntimeError):
    pass


def patch_gui_source(src: str) -> str:
    """The reference GUI's source with the 'GPU' receiver mode added.  Raises PatchError if an
    anchor is missing (a different GUI version): nothing is guessed."""
    tree = ast.parse(src)
    lines = src.splitlines(keepends=True)
    classes = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}
    for need in ("UartReceiver", "UdpReceiver", "ReceiverController"):
        if need not in classes:
            raise PatchError(f"class {need} not found")
    if "GpuQtReceiver" in classes:
        raise PatchError("already patched")
    ctrl = classes["ReceiverController"]
    start = next((f for f in ctrl.body if isinstance(f, ast.FunctionDef) and f.name == "start_receiver"), None)
    if start is None:
        raise PatchError("ReceiverController.start_receiver not found")
    # (2) the UART branch of start_receiver: `elif mode == "UART":` inside the try block
    uart_if = None
    for node in ast.walk(start):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Name) \
                and node.test.left.id == "mode" and isinstance(node.test.comparators[0], ast.Constant) \
                and node.test.comparators[0].value == "UART":
            uart_if = node
    if uart_if is None or uart_if.orelse:
        raise PatchError('start_receiver: `elif mode == "UART":` (last branch) not found')
    insert_branch_at = uart_if.end_lineno                     # 1-based line after which the new branch goes
    # (1) class text in front of ReceiverController (decorators excluded: it has none)
    insert_class_at = ctrl.lineno - 1
    out = lines[:insert_class_at] + [GPU_QT_RECEIVER] + lines[insert_class_at:insert_branch_at] + [START_BRANCH] \
        + lines[insert_branch_at:]
    text = "".join(out)
    # (3) isinstance(..., UartReceiver) -> also GpuQtReceiver
    text, n_inst = re.subn(r"isinstance\(([\w\.]*current_receiver), UartReceiver\)",
                           r"isinstance(\1, (UartReceiver, GpuQtReceiver))", text)
    if n_inst < 4:
        raise PatchError(f"expected at least four isinstance(..., UartReceiver) sites, found {n_inst}")
    # (4) transport byte on a mode switch
    text, n_mode = re.subn(r'command = ETHERNET_MODE_CMD if mode == "ETHERNET" else UART_MODE_CMD',
                           'command = UART_MODE_CMD if mode == "UART" else ETHERNET_MODE_CMD', text)
    if n_mode < 1:
        raise PatchError("mode-switch transport command not found")
    # (5) configuration keys
    text, n_cfg = re.subn(r"(web_config = \{\n)", r"\1    'gpu_channels': 1,\n    'gpu_display_channel': 0,\n    'gpu_device': 0,\n", text, count=1)
    if n_cfg != 1:
        raise PatchError("web_config not found")
    ast.parse(text)                                           # still valid Python
    return text


def patch_index_html(html: str) -> str:
    """templates/index.html:295-298: a third <option> in the communication-mode <select>."""
    m = re.search(r'(<option value="ETHERNET"[^>]*>[^<]*</option>)', html)
    if not m:
        raise PatchError('<option value="ETHERNET"> not found')
    return html[:m.end()] + '\n                    <option value="GPU">GPU (B200)</option>' + html[m.end():]


def main(argv):
    if len(argv) == 4 and argv[1] == "--html":
        open(argv[3], "w").write(patch_index_html(open(argv[2]).read()))
    elif len(argv) == 3:
        open(argv[2], "w").write(patch_gui_source(open(argv[1]).read()))
    else:
        print(__doc__)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
