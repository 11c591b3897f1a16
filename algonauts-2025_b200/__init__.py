"""algonauts-2025_b200 — B200-native (sm_100a) implementation of TRIBE's FmriEncoder hot path.

The directory name is not an importable identifier; import it as ``algonauts2025_b200`` (see the loader module of that
name at the repository root).  Layout:

* ``csrc/``       hand-written CUDA kernels + the C ABI (``include/tribe_b200.h``) built into ``libtribe_b200.so``
* ``_lib.py``     nvcc build + ctypes binding of the C ABI
* ``ops.py``      torch-tensor wrappers of the ABI calls
* ``model.py``    ``FmriEncoder`` / ``FmriEncoderConfig``   (mirror of reference ``algonauts2025/model.py``)
* ``pl_module.py````BrainModule``                           (mirror of reference ``algonauts2025/pl_module.py``)
* ``metrics.py``  Pearson metrics / evaluation              (mirror of ``modeling_utils/metrics/base.py``, ``main.py:459-477``)
* ``parallel.py`` data-parallel gradient all-reduce, ensemble and parcel-sharded evaluation over NCCL
"""
from ._lib import TribeError, build, launch_count, load  # noqa: F401

__all__ = ["TribeError", "build", "load", "launch_count"]
