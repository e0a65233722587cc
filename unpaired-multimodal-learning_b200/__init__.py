"""uml_b200 - B200-native hot path of Unpaired Multimodal Learning (UML).

Host side: Python/PyTorch for device memory, streams and torch.distributed.
Device side: hand-written sm_100a kernels behind a C ABI (``include/uml_b200.h``,
``lib/libuml_b200.so``).  There is no CPU or eager-PyTorch fallback: importing
``uml_b200.ops`` raises if the library has not been built, and every op raises if its
tensors are not on a CUDA device.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the .so until first use)
