"""torchrun worker: one stream lives on rank 1, is broadcast once over NCCL, every rank decodes its page
range with the CUDA path; shards are checked against the source bytes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brotli_g_sdk_b200 as bg
from brotli_g_sdk_b200 import datagen
from brotli_g_sdk_b200.multi_gpu import StreamGeometry, cuda_decode_fn, decode_sharded_stream

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
data = datagen.mixed(37 * 65536 + 4321, seed=5)          # every rank can regenerate the source to check its shard
owner = 1 % world
stream = None
if rank == owner:
    s = bg.Encode(data)
    stream = torch.zeros(((len(s) + 15) // 16) * 16 + 64, dtype=torch.uint8, device="cuda")
    stream[: len(s)] = torch.from_numpy(s).cuda()
    stream = stream[: len(s)]
dec = bg.BrotligDecoder(local)
timing = {}
shard, (lo, hi), nbytes = decode_sharded_stream(stream, owner, cuda_decode_fn(dec, timing), device=torch.device("cuda", local),
                                                capacity=len(data) + 4096, timing=timing)   # one collective: [size | stream]
geo = StreamGeometry(38, 65536, 4321, len(data))
want = data[lo * 65536: lo * 65536 + geo.range_bytes(lo, hi)]
ok = torch.tensor([int(np.array_equal(shard.cpu().numpy(), want))], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED OK" if int(ok.item()) == 1 else "SHARDED MISMATCH", "world", world, "stream bytes", nbytes, timing)
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
