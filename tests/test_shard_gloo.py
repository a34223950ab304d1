"""CPU-only, world_size 2 over gloo: GOP-shard assignment and the NAL-unit gather used for multi-GPU runs
(ks265codec_b200/shard.py).  Each rank 'encodes' its shards with the CPU model; rank 0's gathered stream
must equal the single-process stream byte for byte."""
import os
import subprocess
import sys

from katlib import ROOT

from ks265codec_b200 import shard as ksh

WORKER = r'''
import os, sys, ctypes as C, hashlib
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, os.path.join(%(root)r, "tools"))
import torch.distributed as dist
import gen_yuv
from katlib import SeqCfg, oracle, ptr
from ks265codec_b200 import shard as ksh
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
w, h, n, iper = 96, 64, 10, 3
yuv = np.frombuffer(gen_yuv.make(w, h, n, seed=9), np.uint8)
O = oracle(); O.ora_encode_sequence.restype = C.c_long
def enc(first, cnt):
    cfg = SeqCfg(w, h, cnt, 30, iper, 0, 64, 16, 2, 1, 1, 3)
    bs = np.zeros(w * h * 3 * cnt + 100000, np.uint8)
    fs = w * h * 3 // 2
    part = np.ascontiguousarray(yuv[first * fs:(first + cnt) * fs])
    nb = O.ora_encode_sequence(C.byref(cfg), ptr(part), ptr(bs), C.c_size_t(bs.size), None)
    return bs[:nb].tobytes()
shards = ksh.shard_frames(n, iper)
local = {s: enc(*shards[s]) for s in ksh.assign_shards(len(shards), rank, world)}
out = ksh.gather_bitstreams(local, len(shards))
if rank == 0:
    whole = enc(0, n)          # closed GOPs: the single-process stream is the concatenation of the shard streams
    assert out == whole, (len(out), len(whole))
    print("GATHER_OK", hashlib.md5(out).hexdigest())
else:
    assert out is None
dist.destroy_process_group()
'''


def test_assignment_covers_every_shard_once():
    for n in (1, 2, 7, 16):
        for world in (1, 2, 4, 8):
            got = sorted(s for r in range(world) for s in ksh.assign_shards(n, r, world))
            assert got == list(range(n))
    assert ksh.shard_frames(10, 3) == [(0, 3), (3, 3), (6, 3), (9, 1)]


def test_gather_world2_gloo(tmp_path):
    port = 29000 + os.getpid() % 2000
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-800:]
    assert "GATHER_OK" in outs[0][0]
