"""Device memory and streams: PyTorch is the allocator and the stream provider, nothing more."""
import numpy as np


class CudaBackend:
  """fp64 / int32 buffers on one CUDA device, handed to the C ABI as raw pointers."""

  def __init__(self, device=None):
    import torch
    if not torch.cuda.is_available():
      raise RuntimeError('pymoc_b200: no CUDA device visible; the engine has no CPU path')
    self.torch = torch
    self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    from . import _lib
    self.lib = _lib.lib()

  def upload(self, arr):
    t = self.torch.from_numpy(np.ascontiguousarray(arr))
    return t.to(self.device, non_blocking=False)

  def zeros(self, shape, dtype=np.float64):
    tdt = {np.float64: self.torch.float64, np.int32: self.torch.int32, np.uint32: self.torch.int32}[dtype]
    return self.torch.zeros(shape, dtype=tdt, device=self.device)

  @staticmethod
  def ptr(buf):
    return None if buf is None else buf.data_ptr()

  def download(self, buf):
    return buf.cpu().numpy()

  def assign(self, buf, arr):
    buf.copy_(self.torch.from_numpy(np.ascontiguousarray(arr)).view(buf.dtype).reshape(buf.shape))

  def stream(self):
    return self.torch.cuda.current_stream(self.device).cuda_stream

  def sync(self):
    self.torch.cuda.synchronize(self.device)
