"""P-GPU run == 1-GPU run on the same input (launched under torchrun on P GPUs; also used by the -m gpu test-suite
through subprocess when >= 2 GPUs are visible).  The check itself lives in pis_b200/multigpu_check.py (bench.py --gpus N
runs it too, before its timed region).  Prints one JSON line on rank 0.

  torchrun --nproc-per-node P tools/multi_check.py [ncell] [steps] [T0]
  env: PISB_HALO_MODE (0 auto, 1 NCCL, 2 peer memory or fail), PISB_FORCE_VARIANT (3 / 5: thread-per-atom / pair-list step
       kernels at any brick size), PISB_FUSE_VV"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pis_b200.distributed import init_process_group  # noqa: E402
from pis_b200.multigpu_check import run_check  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    T0 = float(sys.argv[3]) if len(sys.argv) > 3 else 60.0
    rank, world = init_process_group()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    out = run_check(rank, world, local, ncell=ncell, steps=steps, T0=T0, halo_mode=int(os.environ.get("PISB_HALO_MODE", "0")),
                    force_variant=int(os.environ.get("PISB_FORCE_VARIANT", "0")),
                    fuse_vv=int(os.environ["PISB_FUSE_VV"]) if "PISB_FUSE_VV" in os.environ else None)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
