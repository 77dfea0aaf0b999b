"""Where the `-m gpu` tests put their tensors and which engine they drive.

On the B200 box: cuda:0 and huff_encoding_b200.engine.Engine (the product).  With HB_EMU=1 (no GPU; tests/emu, the CPU
execution model of the library, loaded through HUFFB200_SO): host tensors and an engine whose "device" pointers are host
pointers.  The model is test infrastructure; the product never loads it."""
import os

MODEL = os.environ.get("HB_EMU") == "1"


def dev():
    import torch
    return torch.device("cpu") if MODEL else torch.device("cuda", 0)


def dev_sync():
    if not MODEL:
        import torch
        torch.cuda.synchronize()


def make_engine():
    if MODEL:
        from tests.emu.model_engine import ModelEngine
        return ModelEngine()
    from huff_encoding_b200.engine import Engine
    return Engine(0)
