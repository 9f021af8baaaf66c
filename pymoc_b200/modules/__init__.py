"""Drop-in module classes with the reference's names and signatures
(src/pymoc/modules/__init__.py:2-6); every method runs a CUDA kernel through the C ABI."""
from .column import Column
from .psi_SO import Psi_SO
from .psi_thermwind import Psi_Thermwind
from .SO_ML import SO_ML
