"""The N>1 path on CPU: two processes over gloo shard a batch exactly like bench.py / a multi-process driver does
(static contiguous split, no data-path collective), run their shard through the C ABI (emulated kernels -- test
infrastructure) and reduce times / cell counts; rank 0 checks the union against the oracle."""
import os
import subprocess
import sys
import textwrap

from common import ROOT


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from common import capi, orc, scoring_to_params, EMU_SO
    from seqalib_b200 import shard, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 37
    lo, hi = shard.shard_range(n_total, rank, world)
    owned = torch.zeros(n_total, dtype=torch.int32); owned[lo:hi] = 1
    dist.all_reduce(owned)
    assert bool((owned == 1).all()), "shards must tile the batch exactly once"
    lib = capi.Lib(EMU_SO)
    sc = orc.Scoring.linear(-1, 1, -1)
    bases, off1, off2, l1, l2 = synth.batch(synth.SEED, lo, hi - lo, 1)
    res = lib.align_batch(scoring_to_params("sw", sc), bases, off1, off2, l1, l2)
    cells = float((l1.astype(np.float64) * l2).sum())
    total_cells = shard.reduce_sum(cells)
    tmax = shard.reduce_max(1.0 + rank)
    assert tmax == float(world)
    scores = torch.zeros(n_total, dtype=torch.int64); scores[lo:hi] = torch.from_numpy(res.score[:hi - lo].astype(np.int64))
    dist.all_reduce(scores)
    shard.barrier()
    if rank == 0:
        fb, fo1, fo2, fl1, fl2 = synth.batch(synth.SEED, 0, n_total, 1)
        assert total_cells == float((fl1.astype(np.float64) * fl2).sum())
        for p in range(n_total):
            a = bytes(fb[int(fo1[p]):int(fo1[p]) + int(fl1[p])]).decode(); b = bytes(fb[int(fo2[p]):int(fo2[p]) + int(fl2[p])]).decode()
            assert orc.oracle_align("sw", sc, a, b)["score"] == int(scores[p]), p
        print("GLOO-OK")
    dist.destroy_process_group()
""")


def test_world_size_2_gloo(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "GLOO-OK" in out.stdout, out.stdout[-3000:]
