"""TEST INFRASTRUCTURE: huff_encoding_b200.engine.Engine over the CPU model of the library (tests/emu).  Same methods --
only the constructor differs: no CUDA device, host tensors, whose data_ptr() the model treats as device pointers."""
import contextlib

import torch

from huff_encoding_b200 import _lib as L
from huff_encoding_b200.api import Context
from huff_encoding_b200.engine import Engine


class ModelEngine(Engine):
    def __init__(self):
        if "emu" not in L.SO_PATH:
            raise RuntimeError("ModelEngine needs HUFFB200_SO to point at tests/emu/_build/libhuffb200_emu.so")
        self.device_index = 0
        self.device = torch.device("cpu")
        self.ctx = Context(0)
        self.lib = L.load()
        self._hist = torch.zeros(256, dtype=torch.int64)
        self._stream = None
        self.comm_world = None

    @contextlib.contextmanager
    def _ordered(self):
        yield
