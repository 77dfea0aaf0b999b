"""huff_encoding_b200 -- B200-native (sm_100a) drop-in for huff_coding's u8 hot path.

    from huff_encoding_b200 import compress, decompress, compress_with_tree, HuffTree, ByteWeights, build_weights_map

The compute lives in libhuffb200.so (hand-written CUDA kernels behind the C ABI of include/huffb200.h).
Importing this package never falls back to a CPU implementation: the compute entry points raise when the library
or a CUDA device is missing.
"""
from . import _lib
from .api import (ByteWeights, CompressData, CompressError, CompressedDataFromBytesError, Context, FromBinError,
                  HuffCudaError, HuffPanic, HuffTree, build_weights_map, compress, compress_with_tree, decompress,
                  default_context)

__all__ = [
    "ByteWeights", "CompressData", "CompressError", "CompressedDataFromBytesError", "Context", "FromBinError",
    "HuffCudaError", "HuffPanic", "HuffTree", "build_weights_map", "compress", "compress_with_tree", "decompress",
    "default_context",
]
