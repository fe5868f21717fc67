"""Multi-rank GPU test of the point-sharded dense grid (chore_b200.dist.query_grid_sharded, SURVEY.md 8e / BASELINE config 2):
2 ranks over NCCL evaluate disjoint slabs of one image's grid and all-gather them; the result equals the single-GPU grid
bit for bit.  Needs >= 2 CUDA devices (skipped otherwise); spawned with torch.multiprocessing."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import chore_b200
    from chore_b200 import dist as cdist
    from oracle import chore_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    net = chore_b200.CHORE(device=str(dev))
    net.load_state_dict(O.make_state_dict(0, "unit"))
    feat, tmpx = O.synth_features(5, B=1, hw=32)
    net.im_feat_list, net.tmpx = [feat.to(dev)], tmpx.to(dev)
    cc = torch.tensor([[1008., 995.]], device=dev)
    res, bmin, bmax = (24, 20, 29), [-3.0, -0.9, 0.2], [3.0, 1.8, 4.0]          # 13 920 points: ragged 128-aligned shards
    outs = cdist.query_grid_sharded(net, res, bmin, bmax, cc, head_mask=1 | 8)
    if rank == 0:
        single = net.query_grid(res, bmin, bmax, cc, 0, head_mask=1 | 8)
        ok = all(torch.equal(a, b) for a, b in zip((outs[0], outs[3]), (single[0], single[3]))) and outs[1] is None
        torch.save({"ok": ok, "shape": tuple(outs[0].shape)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_query_grid_sharded_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["shape"] == (2, 24 * 20 * 29)
