#!/usr/bin/env python
"""Run the batched rrc_fir kernel a few times for one tap count (ncu target).
usage: python tools/fir_profile.py 127|256 [nchan] [nsamples]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, qpsk_b200
ntaps = int(sys.argv[1]); nchan = int(sys.argv[2]) if len(sys.argv) > 2 else 16384; T = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
dev = torch.device("cuda", 0)
x = torch.randn((nchan, T, 2), device=dev)
f = qpsk_b200.Fir(qpsk_b200.rrc_make(ntaps, 9600.0, 2400.0 if ntaps == 127 else 1200.0, 0.35), nchan)
for _ in range(3):
    f.filter_device(x.data_ptr(), T)
torch.cuda.synchronize()
print(ntaps, nchan, T, f.kernel_ms(), "ms", nchan * T / f.kernel_ms() / 1e3, "Msamples/s")
