"""2-GPU (NCCL) check of the distributed index build, launched by torchrun:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_build_check.py
train_distributed / add_distributed on the CUDA path must reproduce the single-process oracle
(k-means within fp32 tie noise and identical on all ranks; list contents, ids and search bit-exact)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{torch.cuda.current_device()}"))
    P = importlib.import_module("abstracts-search_b200")
    from oracle import ivf as oivf
    from oracle import synth as osynth

    d, nlist, n, nq, k, nprobe = 1024, 64, 20000, 16, 10, 8
    x = osynth.corpus(11, 0, n, d, nlist)
    cut = np.linspace(0, n, world + 1).astype(int)
    cut[1] += 333  # uneven slices
    mine = x[cut[rank]:cut[rank + 1]]
    local = P.IndexIVFFlat(d, nlist, device=torch.cuda.current_device())
    local.set_shard(rank, world)
    local.cp.max_points_per_centroid = 200  # 64 * 200 < 20000: subsampling branch
    sh = P.ShardedIndexIVFFlat(local)
    sh.train_distributed(mine)
    cent = local.get_centroids()
    ref_cent = oivf.kmeans_train(x, nlist, max_points_per_centroid=200)
    # fp32 scores computed in a different summation order may flip a near-tie between two centroids,
    # so demand near-identity (same bar as test_train_matches_oracle_on_lattice) ...
    diff = np.abs(cent - ref_cent).max(axis=1)
    # (a flipped point moves the two centroids involved by O(1/cluster size); most centroids are identical)
    assert np.median(diff) < 1e-5 and (diff < 1e-4).mean() >= 0.8 and diff.max() < 0.5, \
        f"rank {rank}: distributed k-means far from the oracle: median {np.median(diff)}, max {diff.max()}, " \
        f"identical {(diff < 1e-4).mean()}"
    # ... and bit-identical centroids on every rank (the coarse step must be replicated exactly)
    c_all = [torch.empty_like(torch.from_numpy(cent).cuda()) for _ in range(world)]
    dist.all_gather(c_all, torch.from_numpy(cent).cuda())
    assert all(torch.equal(c_all[0], c) for c in c_all), "centroids differ between ranks"
    ref_cent = cent
    h = len(mine) // 2
    sh.add_distributed(mine[:h])
    sh.add_distributed(torch.from_numpy(mine[h:]).cuda())
    assert sh.ntotal == n
    order = np.concatenate([np.arange(cut[r], cut[r] + (cut[r + 1] - cut[r]) // 2) for r in range(world)] +
                           [np.arange(cut[r] + (cut[r + 1] - cut[r]) // 2, cut[r + 1]) for r in range(world)])
    ref = oivf.IVFFlat(d, nlist)
    ref.set_centroids(ref_cent)
    ref.add(x[order])
    sizes = torch.from_numpy(local.list_sizes()).cuda()
    dist.all_reduce(sizes)
    assert np.array_equal(sizes.cpu().numpy(), ref.list_sizes())
    for l in range(rank, nlist, world):
        codes, ids = local.get_list(l)
        assert np.array_equal(ids, ref.ids[l]) and np.array_equal(codes, ref.codes[l]), (rank, l)
    sh.nprobe = nprobe
    q = osynth.queries(11, 0, nq, d, nlist, n)
    D, I = sh.search(torch.from_numpy(q).cuda(), k)
    Dr, Ir = ref.search(q, k, nprobe=nprobe)
    assert np.array_equal(I.cpu().numpy(), Ir) and np.array_equal(D.cpu().numpy(), Dr)
    # the same search with the exchange over NVLink peer memory (CUDA IPC between the two processes)
    # instead of the NCCL all-gather: identical results, repeated to cycle the two-deep ring
    sh.use_peer_exchange(max_results=nq * k)
    qd = torch.from_numpy(q).cuda()
    for _ in range(5):
        Dp, Ip = sh.search(qd, k)
        assert np.array_equal(Ip.cpu().numpy(), Ir) and np.array_equal(Dp.cpu().numpy(), Dr), f"rank {rank}: peer exchange"
    emb = torch.full((nq // world, d), float(rank + 1), device="cuda")
    px = P.PeerExchange.over_group(torch.cuda.current_device(), emb.numel() * 4)
    for it in range(3):
        got = px.allgather(emb + it)
        want = torch.stack([torch.full_like(emb, float(r + 1 + it)) for r in range(world)])
        assert torch.equal(got, want), f"rank {rank}: peer all-gather"
    # two-stage shards (fp16 shortlist + exact re-score) behind both exchanges: same result
    loc2 = P.IndexIVFFlat(d, nlist, device=torch.cuda.current_device())
    loc2.set_two_stage(32)
    loc2.set_shard(rank, world)
    loc2.set_centroids(ref_cent)
    loc2.add(x[order])
    sh2 = P.ShardedIndexIVFFlat(loc2)
    sh2.nprobe = nprobe
    D2, I2 = sh2.search(qd, k)
    assert np.array_equal(I2.cpu().numpy(), Ir) and np.array_equal(D2.cpu().numpy(), Dr), f"rank {rank}: two-stage + nccl"
    sh2.use_peer_exchange(max_results=nq * k)
    for _ in range(3):
        D2, I2 = sh2.search(qd, k)
        assert np.array_equal(I2.cpu().numpy(), Ir) and np.array_equal(D2.cpu().numpy(), Dr), f"rank {rank}: two-stage + peer"
    # the batch SPREAD over the ranks: coarse computed where the slice lives, {queries | coarse ids} in ONE
    # all-gather (NVLink peer stores, then NCCL), two-stage shards, fused push: same result
    per = nq // world
    qloc = qd[rank * per:(rank + 1) * per].contiguous()
    pxq = P.PeerExchange.over_group(torch.cuda.current_device(), per * (d * 4 + 8 * nprobe))
    for _ in range(3):
        Ds, Is = sh2.search_spread(qloc, k, px_queries=pxq)
        assert np.array_equal(Is.cpu().numpy(), Ir[: per * world]) and np.array_equal(Ds.cpu().numpy(), Dr[: per * world]), \
            f"rank {rank}: search_spread + peer"
        assert torch.equal(sh2.last_queries, qd[: per * world])
    Ds, Is = sh.search_spread(qloc, k)  # NCCL all-gather of the record
    assert np.array_equal(Is.cpu().numpy(), Ir[: per * world]) and np.array_equal(Ds.cpu().numpy(), Dr[: per * world]), \
        f"rank {rank}: search_spread + nccl"
    torch.cuda.synchronize()
    assert sh._px.status() == 0 and px.status() == 0 and sh2._px.status() == 0 and pxq.status() == 0
    dist.barrier()
    if rank == 0:
        print(f"dist_build_check ok: world={world}, k-means matches, lists and search bit-exact, NVLink peer exchange bit-exact")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
