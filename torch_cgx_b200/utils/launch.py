"""Launcher glue. The reference must be started with ``mpirun`` and maps
``OMPI_COMM_WORLD_*`` to the env:// variables by hand
(/root/reference/examples/cifar_train.py:61-66, README.md:67-76). Here any
launcher works (torchrun, mp.spawn, mpirun, srun); this helper normalises the
environment."""
from __future__ import annotations

import os


def map_launcher_env(default_addr: str = "127.0.0.1", default_port: int = 4040) -> tuple[int, int, int]:
    """Fill RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT from
    OpenMPI or SLURM variables when torchrun did not set them.
    Returns (rank, world_size, local_rank)."""
    env = os.environ
    if "RANK" not in env:
        if "OMPI_COMM_WORLD_RANK" in env:
            env["RANK"] = env["OMPI_COMM_WORLD_RANK"]
            env["WORLD_SIZE"] = env["OMPI_COMM_WORLD_SIZE"]
            env.setdefault("LOCAL_RANK", env.get("OMPI_COMM_WORLD_LOCAL_RANK", "0"))
        elif "SLURM_PROCID" in env:
            env["RANK"] = env["SLURM_PROCID"]
            env["WORLD_SIZE"] = env.get("SLURM_NTASKS", "1")
            env.setdefault("LOCAL_RANK", env.get("SLURM_LOCALID", "0"))
        else:
            env["RANK"] = "0"
            env["WORLD_SIZE"] = "1"
    env.setdefault("WORLD_SIZE", "1")
    env.setdefault("LOCAL_RANK", "0")
    env.setdefault("MASTER_ADDR", default_addr)
    env.setdefault("MASTER_PORT", str(default_port))
    return int(env["RANK"]), int(env["WORLD_SIZE"]), int(env["LOCAL_RANK"])
