import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fpga_real_time_fft_analyzer_b200 import FraContext
n = int(sys.argv[1]); b = (1 << 26) // n
ctx = FraContext(b, n)
x = torch.randint(-32768, 32767, (b, n), dtype=torch.int16, device="cuda")
for i in range(3):
    y = ctx.fft_only(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    y = ctx.fft_only(x)
e1.record(); torch.cuda.synchronize()
print(n, b, e0.elapsed_time(e1) / 10, "ms")
