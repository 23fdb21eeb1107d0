import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ckfft_b200 as ck
n, batch, inv = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = ck.Context(n, ck.BOTH)
x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1))
y = ctx.complex_inverse(x) if inv else ctx.complex_forward(x)
torch.cuda.synchronize()
ref = torch.fft.ifft(x, dim=1) * n if inv else torch.fft.fft(x, dim=1)
print(n, batch, inv, "max err", float((y - ref).abs().max()))
