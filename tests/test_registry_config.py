"""Layer registry + env config surface."""
import os

import pytest
import torch

import torch_cgx_b200 as cgx
from torch_cgx_b200 import CGXState

C = cgx._C


def test_env_defaults(clean_cgx_env):
    cfg = C.engine_config()
    assert cfg["bits"] == 32 and cfg["bucket_size"] == 512
    assert cfg["fusion_bytes"] == 64 << 20
    assert cfg["min_compress_elems"] == 16
    assert cfg["inner_comm"] == "P2P" and cfg["cross_comm"] == "NCCL"
    assert cfg["inner_reduction"] == "SRA" and cfg["cross_reduction"] == "RING"
    assert cfg["intra_broadcast"] is True and cfg["intra_compress"] is True
    assert cfg["skip_incomplete"] is False and cfg["stochastic"] is False


def test_env_overrides(clean_cgx_env, monkeypatch):
    monkeypatch.setenv("CGX_COMPRESSION_QUANTIZATION_BITS", "4")
    monkeypatch.setenv("CGX_COMPRESSION_BUCKET_SIZE", "128")
    monkeypatch.setenv("CGX_COMPRESSION_SKIP_INCOMPLETE_BUCKETS", "1")
    monkeypatch.setenv("CGX_FUSION_BUFFER_SIZE_MB", "8")
    monkeypatch.setenv("CGX_INNER_COMMUNICATOR_TYPE", "SHM")   # reference default -> P2P here
    monkeypatch.setenv("CGX_CROSS_COMMUNICATOR_TYPE", "MPI")   # no MPI on this stack -> NCCL
    monkeypatch.setenv("CGX_INNER_REDUCTION_TYPE", "Ring")
    monkeypatch.setenv("CGX_COMPRESSION_MINIMAL_SIZE", "100")
    monkeypatch.setenv("CGX_STOCHASTIC_ROUNDING", "1")
    cfg = C.engine_config()
    assert cfg["bits"] == 4 and cfg["bucket_size"] == 128 and cfg["skip_incomplete"] is True
    assert cfg["fusion_bytes"] == 8 << 20
    assert cfg["inner_comm"] == "P2P" and cfg["cross_comm"] == "NCCL"
    assert cfg["inner_reduction"] == "RING"
    assert cfg["min_compress_elems"] == 100
    assert cfg["stochastic"] is True
    # values are re-read on every call (the reference's tests depend on it)
    monkeypatch.setenv("CGX_COMPRESSION_QUANTIZATION_BITS", "8")
    assert C.engine_config()["bits"] == 8
    monkeypatch.setenv("CGX_COMPRESSION_QUANTIZATION_BITS", "12")  # out of range == off
    assert C.engine_config()["bits"] == 32


def test_unregistered_buffer_uses_env(clean_cgx_env, monkeypatch):
    monkeypatch.setenv("CGX_COMPRESSION_QUANTIZATION_BITS", "4")
    layers, bucket = C.extract_layers(10_000)
    assert bucket == -1 and layers == [(0, 10_000, 4, 512)]
    layers, _ = C.extract_layers(8)           # < 16 elements: never compressed
    assert layers == [(0, 8, 32, 512)]
    layers, _ = C.extract_layers(16)          # numel > min_elems is required
    assert layers[0][2] == 32


def test_registered_layers_and_cursor(clean_cgx_env):
    cgx.register_layer(0, 0, 4096, 4, 512)
    cgx.register_layer(0, 1, 64, 32, 512)
    cgx.register_layer(1, 0, 1000, 8, 64)
    assert C.num_registered_buckets() == 2
    # cursor cycles 0,1,0,... like the reference
    l0, b0 = C.extract_layers(4160)
    l1, b1 = C.extract_layers(1000)
    l2, b2 = C.extract_layers(4160)
    assert (b0, b1, b2) == (0, 1, 0)
    assert l0 == [(0, 4096, 4, 512), (4096, 64, 32, 512)]
    assert l1 == [(0, 1000, 8, 64)]
    # a size that matches no bucket falls back to one env-configured layer instead of throwing
    lx, bx = C.extract_layers(777)
    assert bx == -1 and lx == [(0, 777, 32, 512)]
    # out-of-order call is resolved by size
    ly, by = C.extract_layers(1000)
    assert by == 1
    # explicit bucket index
    lz, bz = C.extract_layers(4160, 0)
    assert bz == 0 and lz == l0


def test_setters(clean_cgx_env):
    cgx.register_layer(0, 0, 4096, 4, 512)
    cgx.set_quantization_bits(0, 0, 2)
    cgx.set_quantization_bucket_size(0, 0, 128)   # really sets the bucket size (reference bug §2.8 #1)
    assert C.registered_bucket(0) == [(4096, 2, 128)]
    layers, _ = C.extract_layers(4096, 0)
    assert layers == [(0, 4096, 2, 128)]
    cgx.set_quantization_bits(0, 0, 32)
    assert C.extract_layers(4096, 0)[0] == [(0, 4096, 32, 128)]
    cgx.reset_layers()
    assert C.num_registered_buckets() == 0


def test_cgx_state_defaults(clean_cgx_env, monkeypatch):
    class FakePG:
        pass

    s = CGXState(FakePG())
    assert s.layer_min_size == 1024 and s.quantization_bits == 32 and s.quantization_bucket_size == 1024
    monkeypatch.setenv("CGX_COMPRESSION_QUANTIZATION_BITS", "4")
    monkeypatch.setenv("CGX_COMPRESSION_BUCKET_SIZE", "256")
    monkeypatch.setenv("CGX_COMPRESSION_MINIMAL_SIZE", "4096")
    s = CGXState(FakePG(), layer_min_size=100)
    assert s.layer_min_size == 4096 and s.quantization_bits == 4 and s.quantization_bucket_size == 256
    s = CGXState(FakePG(), compression_params={"bits": 2, "bucket_size": 64})
    assert s.quantization_bits == 2 and s.quantization_bucket_size == 64
    assert not s.should_compress_(torch.zeros(10_000))          # 1-D
    assert not s.should_compress_(torch.zeros(10, 10))          # too small
    assert s.should_compress_(torch.zeros(100, 100))
    sd = s.state_dict()
    s2 = CGXState(FakePG())
    s2.load_state_dict(sd)
    assert s2.quantization_bits == 2 and s2.step == 0


def test_launcher_env_mapping(monkeypatch):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("OMPI_COMM_WORLD_RANK", "3")
    monkeypatch.setenv("OMPI_COMM_WORLD_SIZE", "8")
    monkeypatch.setenv("OMPI_COMM_WORLD_LOCAL_RANK", "3")
    assert cgx.map_launcher_env() == (3, 8, 3)
    assert os.environ["MASTER_ADDR"] == "127.0.0.1" and os.environ["MASTER_PORT"] == "4040"


def test_bench_reference_arm_reports_unavailable():
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference"], capture_output=True, text=True)
    assert out.returncode == 0
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and "unavailable" in d


def test_compression_context_manager(clean_cgx_env, monkeypatch):
    monkeypatch.setenv("CGX_COMPRESSION_BUCKET_SIZE", "128")
    assert C.engine_config()["bits"] == 32
    with cgx.compression(bits=2, bucket_size=256, stochastic=True, seed=9):
        cfg = C.engine_config()
        assert (cfg["bits"], cfg["bucket_size"], cfg["stochastic"], cfg["seed"]) == (2, 256, True, 9)
        with cgx.compression(bits=8):
            assert C.engine_config()["bits"] == 8 and C.engine_config()["bucket_size"] == 256
        assert C.engine_config()["bits"] == 2
    cfg = C.engine_config()
    assert cfg["bits"] == 32 and cfg["bucket_size"] == 128 and cfg["stochastic"] is False
    assert "CGX_SEED" not in os.environ
