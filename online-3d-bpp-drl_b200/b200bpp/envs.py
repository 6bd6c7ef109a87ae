"""`make_vec_envs` with the reference's signature (acktr/envs.py:77-118), returning the GPU env."""
from .vec_env import BinPackingVecEnv


def make_vec_envs(env_name, seed, num_processes, gamma, log_dir, device, allow_early_resets,
                  num_frame_stack=None, args=None, **kwargs):
    """env_name / gamma / log_dir / allow_early_resets / num_frame_stack are accepted for signature
    compatibility (the reference ignores gamma too: VecNormalize(ob=False, ret=False), envs.py:112).
    `args` is the namespace of acktr/arguments.py: container_size, enable_rotation, box_size_set, data_type."""
    if num_frame_stack is not None:
        raise NotImplementedError("frame stacking is not part of the bin-packing path")
    return BinPackingVecEnv(
        num_envs=num_processes,
        container_size=getattr(args, "container_size", (10, 10, 10)),
        box_set=getattr(args, "box_size_set", None),
        enable_rotation=getattr(args, "enable_rotation", False),
        data_type=getattr(args, "data_type", "rs"),
        seed=seed, device=device, **kwargs)
