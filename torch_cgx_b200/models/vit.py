"""Vision Transformer (Dosovitskiy et al. 2020) for the DDP benchmarks
(BASELINE.json configs[3]: "ViT-L/16 DDP bits=2 QSGD stochastic rounding")."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class ViTConfig:
    image_size: int = 224
    patch: int = 16
    dim: int = 1024
    depth: int = 24
    heads: int = 16
    mlp_dim: int = 4096
    num_classes: int = 1000

    @staticmethod
    def large16() -> "ViTConfig":
        return ViTConfig()

    @staticmethod
    def base16() -> "ViTConfig":
        return ViTConfig(dim=768, depth=12, heads=12, mlp_dim=3072)

    @staticmethod
    def tiny() -> "ViTConfig":
        return ViTConfig(image_size=32, patch=8, dim=64, depth=2, heads=4, mlp_dim=128, num_classes=10)


class EncoderBlock(nn.Module):
    def __init__(self, c: ViTConfig):
        super().__init__()
        self.ln1 = nn.LayerNorm(c.dim, eps=1e-6)
        self.qkv = nn.Linear(c.dim, 3 * c.dim)
        self.proj = nn.Linear(c.dim, c.dim)
        self.ln2 = nn.LayerNorm(c.dim, eps=1e-6)
        self.fc1 = nn.Linear(c.dim, c.mlp_dim)
        self.fc2 = nn.Linear(c.mlp_dim, c.dim)
        self.heads = c.heads

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = self.qkv(self.ln1(x)).view(B, T, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v)
        x = x + self.proj(a.transpose(1, 2).reshape(B, T, C))
        return x + self.fc2(F.gelu(self.fc1(self.ln2(x))))


class ViT(nn.Module):
    def __init__(self, c: ViTConfig):
        super().__init__()
        self.config = c
        n = (c.image_size // c.patch) ** 2
        self.patch_embed = nn.Conv2d(3, c.dim, c.patch, c.patch)
        self.cls = nn.Parameter(torch.zeros(1, 1, c.dim))
        self.pos = nn.Parameter(torch.randn(1, n + 1, c.dim) * 0.02)
        self.blocks = nn.ModuleList(EncoderBlock(c) for _ in range(c.depth))
        self.ln = nn.LayerNorm(c.dim, eps=1e-6)
        self.head = nn.Linear(c.dim, c.num_classes)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.patch_embed(x).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls.expand(x.size(0), -1, -1), x], dim=1) + self.pos
        for b in self.blocks:
            x = b(x)
        return self.head(self.ln(x[:, 0]))


def vit_l16() -> ViT:
    return ViT(ViTConfig.large16())


def vit_b16() -> ViT:
    return ViT(ViTConfig.base16())
