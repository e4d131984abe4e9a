"""File-descriptor passing used by the VMM / multicast heap (comm/fd_channel): SCM_RIGHTS over
abstract unix datagram sockets, names exchanged out of band. CPU only."""
import os

import torch.multiprocessing as mp

import torch_cgx_b200 as cgx

C = cgx._C


def _child(q_names, q_done, idx):
    ch = C.FdChannel(f"t{idx}")
    q_names.put((idx, ch.name()))
    # receive one descriptor from the parent and prove it is the same open file
    fd, kind, src = ch.recv(10_000)
    assert kind == 7 and src == 42
    # (all receivers share ONE open file description, hence one file offset: read positionally)
    q_done.put((idx, os.pread(fd, 100, 0).decode()))
    os.close(fd)


def test_descriptor_travels_between_processes(tmp_path):
    ctx = mp.get_context("spawn")
    q_names, q_done = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_child, args=(q_names, q_done, i)) for i in range(3)]
    for p in procs:
        p.start()
    names = dict(q_names.get(timeout=60) for _ in procs)
    path = tmp_path / "payload.txt"
    path.write_text("hello over SCM_RIGHTS")
    me = C.FdChannel("parent")
    fd = os.open(path, os.O_RDONLY)
    try:
        for i in range(3):
            assert me.send(names[i], fd, 7, 42)
    finally:
        os.close(fd)
    got = dict(q_done.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert got == {0: "hello over SCM_RIGHTS", 1: "hello over SCM_RIGHTS", 2: "hello over SCM_RIGHTS"}


def test_recv_times_out_quietly():
    ch = C.FdChannel("lonely")
    assert ch.recv(50) is None
