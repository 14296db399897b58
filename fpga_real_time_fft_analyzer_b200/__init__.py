"""fpga_real_time_fft_analyzer_b200 - B200-native receive chain (window ROM -> IIR12
-> 16K FFT -> bin framing) of mfkiwl/fpga-real-time-fft-analyzer behind a C ABI.

Only the hot path lives here: csrc/ (sm_100a CUDA kernels + the C ABI of
include/fra.h, built in-tree into libfra.so) and the host-side mirror of the
reference GUI's receiver interface.  Importing the package loads nothing; the
first FraContext loads libfra.so and fails loudly if it or a CUDA device is missing."""
from . import _abi, filter_design
from ._abi import (FILTER_CUSTOM_CMD, FILTER_DEFAULT_CMD, FILTER_NONE_CMD, FILTER_UPDATE_CMD, FPGA_RESET_CMD,
                   ETHERNET_MODE_CMD, START_COMMAND, UART_MODE_CMD, UART_REQUEST_CMD)
from ._lib import FraError
from .context import FraContext
from .receiver import GpuReceiver, frame_to_udp_payloads
from .udp_emitter import UdpFrameSender

__all__ = ["FraContext", "FraError", "GpuReceiver", "UdpFrameSender", "frame_to_udp_payloads", "filter_design", "_abi"]
