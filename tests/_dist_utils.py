"""Helpers to run a function on N ranks with torch.multiprocessing (no mpirun)."""
import os
import traceback

import torch
import torch.multiprocessing as mp


def _entry(rank, world, port, fn, args, env, errq):
    try:
        os.environ.update({k: str(v) for k, v in env.items()})
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        os.environ["RANK"] = str(rank)
        os.environ["WORLD_SIZE"] = str(world)
        os.environ["LOCAL_RANK"] = str(rank)
        fn(rank, world, *args)
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
        raise


def spawn(fn, world, args=(), env=None, timeout=240):
    from conftest import free_port

    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = free_port()
    procs = []
    for r in range(world):
        p = ctx.Process(target=_entry, args=(r, world, port, fn, args, env or {}, errq))
        p.start()
        procs.append(p)
    failed = False
    for p in procs:
        p.join(timeout)
        if p.is_alive():
            p.terminate()
            failed = True
        elif p.exitcode != 0:
            failed = True
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    if failed or msgs:
        raise AssertionError("distributed test failed:\n" + "\n".join(f"[rank {r}] {m}" for r, m in msgs))
