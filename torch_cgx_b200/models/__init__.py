from .gpt2 import GPT2, GPT2Config, gpt2_medium, gpt2_small  # noqa: F401
from .resnet import ResNet, resnet18, resnet50  # noqa: F401
from .vit import ViT, ViTConfig, vit_b16, vit_l16  # noqa: F401
