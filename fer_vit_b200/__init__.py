"""fer_vit_b200 — B200-native (sm_100a) train-step path of FER-ViT's LatentViT / HybridLatentViT / ImageViT."""
from .native_module import set_default_precision, get_default_precision
from .runtime import CrossEntropyLoss, cross_entropy, mixup_cross_entropy
from .models_fer_vit import (LatentViT, LatentViTv2, HybridLatentViT, AdapterModule, create_hybrid_latent_vit,
                             RECOMMENDED_STRATEGIES, ImageViT, PatchEmbedding, create_vit_tiny, create_vit_small,
                             create_vit_base, LatentDecomposer, ExpressionAwareViT)
from .modules import LEAM, SemanticPE, LayerWiseNorm
from .graph import GraphedTrainStep, GraphedMixupTrainStep
from .optim import FusedAdamW
from .data import (LatentAugment, PackedLatentCache, latent_batch, mixup, get_latent_train_transforms,
                   get_latent_val_transforms)

__all__ = ["set_default_precision", "get_default_precision", "CrossEntropyLoss", "cross_entropy", "LatentViT",
           "LatentViTv2", "HybridLatentViT", "AdapterModule", "create_hybrid_latent_vit", "RECOMMENDED_STRATEGIES",
           "ImageViT", "PatchEmbedding", "create_vit_tiny", "create_vit_small", "create_vit_base", "LEAM",
           "SemanticPE", "LayerWiseNorm", "GraphedTrainStep", "GraphedMixupTrainStep", "FusedAdamW", "mixup_cross_entropy", "LatentAugment",
           "PackedLatentCache", "latent_batch", "mixup", "get_latent_train_transforms", "get_latent_val_transforms", "LatentDecomposer", "ExpressionAwareViT"]
