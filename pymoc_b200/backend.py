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

  def guard(self):
    """Context in which this backend's device is the current CUDA device: the library launches on the current
    device, so a backend built for another device must switch to it around every call."""
    return self.torch.cuda.device(self.device)


class PinnedHostBackend:
  """Pinned HOST buffers for the host-pointer entry points of the C ABI (``pmoc_model_run_host``,
  ``pmoc_host_open/step/close``), which do their own host<->device copies.  Same interface as
  :class:`CudaBackend`; the buffers are torch tensors in page-locked memory, viewed as numpy arrays."""

  def __init__(self):
    import torch
    if not torch.cuda.is_available():
      raise RuntimeError('pymoc_b200: no CUDA device visible; the engine has no CPU path')
    from . import _lib
    self.torch, self.lib, self.bytes_in, self.bytes_out = torch, _lib.lib(), 0, 0

  def upload(self, arr):
    a = np.ascontiguousarray(arr)
    t = self.torch.empty(a.shape, dtype=self.torch.from_numpy(a[:0]).dtype, pin_memory=True)
    t.numpy()[...] = a
    self.bytes_in += a.nbytes
    return t

  def zeros(self, shape, dtype=np.float64):
    tdt = {np.float64: self.torch.float64, np.int32: self.torch.int32, np.uint32: self.torch.int32}[dtype]
    t = self.torch.zeros(shape, dtype=tdt, pin_memory=True)
    self.bytes_out += t.numel() * t.element_size()
    return t

  @staticmethod
  def ptr(buf):
    return None if buf is None else buf.data_ptr()

  def download(self, buf):
    return buf.numpy().copy()

  def assign(self, buf, arr):
    buf.numpy()[...] = np.asarray(arr).reshape(buf.shape)

  def stream(self):
    return None

  def sync(self):
    pass

  def guard(self):
    import contextlib
    return contextlib.nullcontext()
