"""Pinned host<->device copy bandwidth of the box (bounds the e2e bench)."""
import torch
n = 16 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
  fn(); torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps
ms = t(lambda: d.copy_(h, non_blocking=True)); print('H2D 16 MiB: %.3f ms  %.1f GB/s' % (ms, n / ms / 1e6))
ms = t(lambda: h.copy_(d, non_blocking=True)); print('D2H 16 MiB: %.3f ms  %.1f GB/s' % (ms, n / ms / 1e6))
def both():
  e = torch.cuda.Event(); e.record()
  with torch.cuda.stream(s1):
    s1.wait_event(e); d.copy_(h, non_blocking=True)
  with torch.cuda.stream(s2):
    s2.wait_event(e); h2.copy_(d2, non_blocking=True)
  torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
ms = t(both); print('both directions 16 MiB each: %.3f ms  %.1f GB/s per direction' % (ms, n / ms / 1e6))
