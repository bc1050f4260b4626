"""utils/dist.py of the reference -> lavender_b200.dist (same function names; NCCL over NVLink, gloo on CPU)."""
from lavender_b200.dist import (NoOp, all_gather, dist_init, get_local_rank, get_local_size, get_rank,  # noqa: F401
                                get_world_size, is_main_process, iter_tqdm, reduce_dict, set_seed, synchronize)
