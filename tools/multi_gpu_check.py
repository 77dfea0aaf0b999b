"""torchrun --nproc-per-node N tools/multi_gpu_check.py [workload] [bytes_total]
Checks on real GPUs that the sharded path reproduces the single-GPU stream bit for bit and round-trips."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from huff_encoding_b200 import datagen as G
from huff_encoding_b200.engine import Engine
from huff_encoding_b200.sharded import ShardedCodec

kind = sys.argv[1] if len(sys.argv) > 1 else "english"
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else (256 << 20) + 12345
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
codec = ShardedCodec(eng, world, rank, dist)
per = n_total // world
lo, hi = rank * per, (n_total if rank == world - 1 else (rank + 1) * per)
shard = getattr(G, kind)(hi - lo, offset=lo, device=dev)
comp_buf = torch.zeros(shard.numel() + shard.numel() // 4 + 4096, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()        # shard and comp_buf are written on torch's stream; the engine stream must not race it
with torch.cuda.stream(eng.stream):
    info = codec.compress(shard, comp_buf)
    out_buf = torch.zeros(shard.numel() + 64, dtype=torch.uint8, device=dev)
    n = codec.decompress(comp_buf, info, out_buf)
    eng.sync()
assert n == shard.numel() and torch.equal(out_buf[:n], shard), "shard round trip failed"
del out_buf
torch.cuda.synchronize()
# concatenate the shard streams ON THE DEVICE of rank 0 (OR-merge of the byte two shards share) and compare with the
# stream one GPU produces for the whole input
total_bits = info["total_bits"]
n_bytes = (total_bits + 7) // 8
cap = max((b + 7) // 8 + 2 for b in info["all_bits"])
send = torch.zeros(cap, dtype=torch.uint8, device=dev)
send[: info["comp_len"]] = comp_buf[: info["comp_len"]]
recv = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
dist.gather(send, recv, dst=0)
pad4 = torch.zeros(((n_bytes + 15) // 16) * 16 + 4096, dtype=torch.uint8, device=dev)
if rank == 0:
    off = 0
    for g in range(world):
        ln = (off % 8 + info["all_bits"][g] + 7) // 8
        pad4[off // 8: off // 8 + ln] |= recv[g][:ln]
        off += info["all_bits"][g]
    del recv
    full = getattr(G, kind)(n_total, device=dev)
    one, clen, pad1, tree1 = Engine(local).compress(full)
    del full
    assert clen == n_bytes and pad1 == info["padding_bits"], (clen, n_bytes, pad1, info["padding_bits"])
    assert torch.equal(one[:clen], pad4[:n_bytes]), "concatenated shards differ from the single-GPU stream"
    assert tree1.read_codes() == info["tree"].read_codes()
    del one
torch.cuda.synchronize()
# byte-sharded decode of the gathered stream (speculative entries + neighbour verification)
dist.broadcast(pad4, src=0)
cut = [n_bytes * g // world // 16 * 16 for g in range(world)] + [n_bytes]
b0 = max(cut[rank] - 4096, 0)
b1 = min(cut[rank + 1] + 4096, n_bytes)
buf = pad4[b0: ((b1 + 15) // 16) * 16].contiguous()
torch.cuda.synchronize()        # buf was produced on the default stream; the engine stream must not race it
with torch.cuda.stream(eng.stream):
    out, cnt, letter_off = codec.decompress_byte_sharded(buf, b0, cut[rank], cut[rank + 1], total_bits, info["tree"],
                                                         lambda k: torch.zeros(k + 64, dtype=torch.uint8, device=dev))
    eng.sync()
expect = getattr(G, kind)(cnt, offset=letter_off, device=dev)
assert torch.equal(out[:cnt], expect), "byte-sharded decode mismatch"
tot = torch.tensor([cnt], dtype=torch.int64, device=dev)
dist.all_reduce(tot)
assert int(tot.item()) == n_total
if rank == 0:
    print(f"multi_gpu_check ok: world={world} {kind} n={n_total} stream={n_bytes} B launches={eng.kernel_launches()}")
dist.destroy_process_group()
