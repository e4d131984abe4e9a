"""Model zoo (ResNet / GPT-2 / ViT used by bench.py and the examples) and small utilities."""
import torch

from torch_cgx_b200 import models
from torch_cgx_b200.models.gpt2 import GPT2, GPT2Config
from torch_cgx_b200.models.vit import ViT, ViTConfig
from torch_cgx_b200.utils.data import SyntheticHostDataset


def _nparams(m):
    return sum(p.numel() for p in m.parameters())


def test_parameter_counts_match_the_named_architectures():
    assert _nparams(models.resnet50()) == 25_557_032      # torchvision resnet50
    assert _nparams(models.resnet18()) == 11_689_512      # torchvision resnet18
    with torch.device("meta"):
        assert abs(_nparams(models.gpt2_medium()) / 1e6 - 354.9) < 0.5   # GPT-2 medium (vocab padded to 50304)
        assert abs(_nparams(models.vit_l16()) / 1e6 - 304.3) < 0.5       # ViT-L/16


def test_tiny_models_train_one_step():
    torch.manual_seed(0)
    net = models.resnet18(num_classes=10, cifar_stem=True)
    x, y = torch.randn(4, 3, 32, 32), torch.randint(0, 10, (4,))
    loss = torch.nn.functional.cross_entropy(net(x), y)
    loss.backward()
    assert all(p.grad is not None for p in net.parameters())
    g = GPT2(GPT2Config.tiny())
    idx = torch.randint(0, 512, (2, 16))
    l2 = g(idx, idx)
    l2.backward()
    assert torch.isfinite(l2)
    v = ViT(ViTConfig.tiny())
    out = v(torch.randn(2, 3, 32, 32))
    assert out.shape == (2, 10)


def test_synthetic_dataset_shapes():
    ds = SyntheticHostDataset.images(4, 3, 8, 8, 10, n_batches=2, seed=1)
    x, y = ds[3]
    assert x.shape == (4, 3, 8, 8) and y.shape == (4,) and len(ds) == 2
    assert ds.bytes_per_batch() == 4 * 3 * 8 * 8 * 4 + 4 * 8
    dt = SyntheticHostDataset.tokens(2, 16, 100, n_batches=1)
    a, b = dt[0]
    assert a.shape == (2, 16) and torch.equal(a[:, 1:], b[:, :-1])
