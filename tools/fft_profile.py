#!/usr/bin/env python
"""Run the FFT+argmax kernel a few times for one length (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, qpsk_b200
n = int(sys.argv[1]); nb = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
dev = torch.device("cuda", 0)
x = torch.randn((nb, n, 2), device=dev)
bins = torch.empty(nb, dtype=torch.int32, device=dev); mag = torch.empty(nb, device=dev)
f = qpsk_b200.Fft(n)
for _ in range(5):
    f.argmax_device(x.data_ptr(), nb, bins.data_ptr(), mag.data_ptr())
torch.cuda.synchronize()
print(n, nb, f.kernel_ms(), "ms", nb * n * 8 / f.kernel_ms() / 1e6, "GB/s")
