"""pymoc_b200 -- B200-native batched-ensemble engine for PyMOC's time-stepping hot path.

Drop-in surface: :mod:`pymoc_b200.modules` (``Column``, ``Psi_Thermwind``, ``Psi_SO``,
``SO_ML`` with the reference's signatures) and :mod:`pymoc_b200.utils`; batched front
end: :class:`pymoc_b200.ensemble.Ensemble` over a :class:`pymoc_b200.spec.ModelSpec`.
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/pymoc_b200.h``; there is no CPU fallback.
"""
__version__ = '0.1.0'
