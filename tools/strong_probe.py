"""Only the `strong` object of bench.py (configs[4] sharded over the ranks + rank 0's single-GPU solve of the same window).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/strong_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from ppo_pkg import ppo
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uid = [ppo.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = ppo.nccl_init(uid[0], rank, world, local_rank)
    out = bench.strong_scaling_config4(ppo, torch, dist, rank, world, local_rank, comm)
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    ppo.nccl_destroy(comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
