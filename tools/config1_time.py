"""configs[1] alone (1,024 x 1200-baud channels x 256 frames, device-resident): ms per call and the K1 / K3 sums.
usage: python tools/config1_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, qpsk_b200
dev = torch.device("cuda", 0)
C, F = 1024, 256
pcm = bench.synth_pcm_gpu(torch, qpsk_b200, C, F * 512, dev, 0, seed=5, rs=1200.0, doppler_hz_per_s=5.0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for div in (sys.argv[1:] or [""]):
  if div: os.environ["QPSK_B200_CHUNK_DIV"] = div
  rx = qpsk_b200.Receiver(C, F, rs=1200.0, device=0, decode_frames=True)
  for _ in range(5):
      rx.process_device(pcm.data_ptr(), F, st.cuda_stream)
  torch.cuda.synchronize()
  ms = []
  for _ in range(30):
      flush.add_(1)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); rx.process_device(pcm.data_ptr(), F, st.cuda_stream); e1.record(); torch.cuda.synchronize()
      ms.append(e0.elapsed_time(e1))
  print("div", div or "default", "config1 ms median %.3f min %.3f  kernels (front sum, loop sum) %s  dibit checksum %d" % (
      float(np.median(ms)), min(ms), rx.kernel_ms(), int(rx.dibits().astype(np.int64).sum())), flush=True)
  rx.close()
